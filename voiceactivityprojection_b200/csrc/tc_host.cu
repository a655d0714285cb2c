// Host-side helpers shared by the tensor-core kernels: tensor-map encoding (cuTensorMapEncodeTiled through the
// runtime's driver entry point: no link-time dependency on libcuda), the per-thread 16-bit format switch, and the
// folding of conv0 + ChannelNorm into centred, pre-scaled taps with closed-form per-frame statistics
// (vap/encoder_components.py:62-70,83-84).
#include <cstdio>
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

thread_local int g_fp16 = 0;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool make_tmap(CUtensorMap* map, const void* base, int elem_bytes, int rank, const uint64_t* dims,
               const uint64_t* strides_elems, const uint32_t* box, int swizzle, std::string* err) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    if (err) *err = "cuTensorMapEncodeTiled not available";
    return false;
  }
  cuuint64_t gdim[5], gstride[5];
  cuuint32_t bdim[5], estride[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estride[i] = 1;
    if (i > 0) gstride[i - 1] = strides_elems[i - 1] * elem_bytes;  // bytes
  }
  const CUresult r = fn(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bdim, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[320];
      snprintf(buf, sizeof buf,
               "cuTensorMapEncodeTiled failed (%d): rank %d elem %d dims [%llu,%llu,%llu,%llu] strides [%llu,%llu,%llu] box "
               "[%u,%u,%u,%u]",
               (int)r, rank, elem_bytes, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
               (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
               (unsigned long long)(rank > 1 ? strides_elems[0] : 0), (unsigned long long)(rank > 2 ? strides_elems[1] : 0),
               (unsigned long long)(rank > 3 ? strides_elems[2] : 0), box[0], rank > 1 ? box[1] : 0,
               rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
      *err = buf;
    }
    return false;
  }
  return true;
}

bool make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box, std::string* err) {
  return make_tmap(map, base, 2, rank, dims, strides_elems, box, 128, err);
}

// Host folding of the conv0 + ChannelNorm parameters (double precision).
// w: (256, 1, 10) conv weight, bias (256), g / beta: ChannelNorm affine (256).
void conv0_v2_fold(const float* w, const float* bias, const float* g, float* u /*[10][256]*/, float* d /*[256]*/,
                   Conv0Stats* cs) {
  double wbar[10] = {0}, bbar = 0;
  for (int c = 0; c < kDim; ++c) {
    for (int k = 0; k < 10; ++k) wbar[k] += w[c * 10 + k];
    bbar += bias[c];
  }
  for (int k = 0; k < 10; ++k) wbar[k] /= kDim;
  bbar /= kDim;
  double G[10][10] = {{0}}, h[10] = {0}, s = 0;
  for (int c = 0; c < kDim; ++c) {
    double uc[10];
    const double dc = bias[c] - bbar;
    for (int k = 0; k < 10; ++k) uc[k] = w[c * 10 + k] - wbar[k];
    for (int k = 0; k < 10; ++k) {
      for (int l = 0; l < 10; ++l) G[k][l] += uc[k] * uc[l];
      h[k] += dc * uc[k];
      u[k * kDim + c] = (float)(g[c] * uc[k]);
    }
    s += dc * dc;
    d[c] = (float)(g[c] * dc);
  }
  for (int k = 0; k < 10; ++k) {
    for (int l = 0; l < 10; ++l) cs->G[k][l] = (float)G[k][l];
    cs->h2[k] = (float)(2.0 * h[k]);
  }
  cs->s = (float)s;
}

}  // namespace vapb
