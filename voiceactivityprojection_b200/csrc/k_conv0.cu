// conv0 of the CPC gEncoder fused with its ChannelNorm and ReLU.
// Reference: vap/encoder_components.py:83-84,99 (Conv1d(1,256,k=10,s=5,p=3)),
// :62-70 (ChannelNorm: unbiased variance over the 256 channels of one time step).
//
// Bandwidth-bound on its output: one frame = 10 input samples -> 256 channels.
// One warp owns a frame at a time; a lane owns 8 channels, keeps their 80 taps in
// registers and the frame's mean/variance is two warp-shuffle reductions.
// Output is channels-last (seq, pad + frame, 256) so that conv1's implicit-GEMM
// rows are contiguous spans of it.
#include <cuda_fp16.h>

#include "common.cuh"

namespace vapb {

constexpr int C0_FRAMES = 64;  // frames per CTA
constexpr int C0_THREADS = 256;
constexpr int C0_SAMPLES = 5 * C0_FRAMES + 5;

template <bool OUT_BF16>
__global__ void __launch_bounds__(C0_THREADS)
conv0_cn_relu_kernel(const float* __restrict__ wav, int batch, long long n_samples, int seq0, long long L0,
                     const float* __restrict__ w, const float* __restrict__ bias,
                     const float* __restrict__ g, const float* __restrict__ b, void* __restrict__ out,
                     long long out_seq_stride, int out_pad_rows) {
  __shared__ float xs[C0_SAMPLES];
  const int lseq = blockIdx.y;
  const int seq = seq0 + lseq;       // channel-major sequence id: c * batch + item
  const int ch = seq / batch, item = seq % batch;
  const float* x = wav + ((long long)item * 2 + ch) * n_samples;
  const long long f0 = (long long)blockIdx.x * C0_FRAMES;
  const long long s0 = 5 * f0 - 3;
  for (int i = threadIdx.x; i < C0_SAMPLES; i += C0_THREADS) {
    long long s = s0 + i;
    xs[i] = (s >= 0 && s < n_samples) ? x[s] : 0.0f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // channels lane*4..+3 and 128+lane*4..+3
  float wr[8][10], br[8], gr[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = (j < 4 ? 0 : 128) + lane * 4 + (j & 3);
#pragma unroll
    for (int k = 0; k < 10; ++k) wr[j][k] = w[k * kDim + c];
    br[j] = bias[c];
    gr[j] = g[c];
    be[j] = b[c];
  }
  __syncthreads();
  for (int fi = warp; fi < C0_FRAMES; fi += C0_THREADS / 32) {
    const long long f = f0 + fi;
    if (f >= L0) break;
    float xv[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) xv[k] = xs[5 * fi + k];
    float a[8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) v = fmaf(wr[j][k], xv[k], v);
      v += br[j];
      a[j] = v;
      s += v;
    }
    const float mean = warp_sum(s) * (1.0f / kDim);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] -= mean;
      q = fmaf(a[j], a[j], q);
    }
    const float var = warp_sum(q) * (1.0f / (kDim - 1));
    const float rstd = 1.0f / sqrtf(var + kEps);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaf(a[j] * rstd, gr[j], be[j]), 0.f);
    const long long row = (long long)lseq * out_seq_stride + (out_pad_rows + f) * kDim;
    if (OUT_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a[0], a[1]), p1 = __floats2bfloat162_rn(a[2], a[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(a[4], a[5]), p3 = __floats2bfloat162_rn(a[6], a[7]);
      uint2 u0, u1;
      u0.x = *reinterpret_cast<uint32_t*>(&p0);
      u0.y = *reinterpret_cast<uint32_t*>(&p1);
      u1.x = *reinterpret_cast<uint32_t*>(&p2);
      u1.y = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint2*>(o + lane * 4) = u0;
      *reinterpret_cast<uint2*>(o + 128 + lane * 4) = u1;
    } else {
      float* o = reinterpret_cast<float*>(out) + row;
      *reinterpret_cast<float4*>(o + lane * 4) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(o + 128 + lane * 4) = make_float4(a[4], a[5], a[6], a[7]);
    }
  }
}

int launch_conv0(cudaStream_t st, const float* wav, int batch, long long n_samples, int seq0, int nseq,
                 long long L0, const float* w, const float* bias, const float* g, const float* b, void* out,
                 int out_bf16, long long out_seq_stride, int out_pad_rows) {
  dim3 grid((unsigned)((L0 + C0_FRAMES - 1) / C0_FRAMES), (unsigned)nseq);
  if (out_bf16)
    conv0_cn_relu_kernel<true><<<grid, C0_THREADS, 0, st>>>(wav, batch, n_samples, seq0, L0, w, bias, g, b, out,
                                                            out_seq_stride, out_pad_rows);
  else
    conv0_cn_relu_kernel<false><<<grid, C0_THREADS, 0, st>>>(wav, batch, n_samples, seq0, L0, w, bias, g, b, out,
                                                             out_seq_stride, out_pad_rows);
  return 1;
}

// Zero rows [row0, row0+nrows) of every sequence of a channels-last buffer (the
// physical zero padding the implicit-GEMM convolutions read).
__global__ void zero_rows_kernel(uint4* buf, long long seq_stride_v, long long row0_v, long long n_v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_v) buf[(long long)blockIdx.y * seq_stride_v + row0_v + i] = make_uint4(0, 0, 0, 0);
}

int launch_zero_rows(cudaStream_t st, void* buf, int elem_bytes, int nseq, long long seq_stride_elems,
                     long long row0, long long nrows) {
  if (nrows <= 0 || nseq <= 0) return 0;
  const long long vec_per_row = (long long)kDim * elem_bytes / 16;
  const long long n_v = nrows * vec_per_row;
  dim3 grid((unsigned)((n_v + 255) / 256), (unsigned)nseq);
  zero_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<uint4*>(buf), seq_stride_elems * elem_bytes / 16,
                                         row0 * vec_per_row, n_v);
  return 1;
}

// Stage export for diagnostics: any activation -> dense fp32 (nseq, rows, 256).
template <int SRC_FMT>  // 0 fp32, 1 bf16, 2 fp16
__global__ void to_f32_kernel(const void* __restrict__ src, RowMap map, int rows_per_seq, float* __restrict__ dst,
                              long long total_rows, int blocked) {
  const long long r = (long long)blockIdx.x * (blockDim.x / 64) + threadIdx.x / 64;
  if (r >= total_rows) return;
  const int c4 = (threadIdx.x & 63) * 4;
  const long long off = blocked ? (((r >> 7) * 64 + (c4 >> 2)) * 128 + (r & 127)) * 4
                                : (r / rows_per_seq) * map.seq_stride + (r % rows_per_seq) * map.row_stride + c4;
  float4 v;
  if (SRC_FMT == 1) {
    const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(src) + off;
    v = make_float4(__bfloat162float(s[0]), __bfloat162float(s[1]), __bfloat162float(s[2]), __bfloat162float(s[3]));
  } else if (SRC_FMT == 2) {
    const __half* s = reinterpret_cast<const __half*>(src) + off;
    v = make_float4(__half2float(s[0]), __half2float(s[1]), __half2float(s[2]), __half2float(s[3]));
  } else {
    v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + off);
  }
  *reinterpret_cast<float4*>(dst + r * kDim + c4) = v;
}

int launch_to_f32(cudaStream_t st, const void* src, int src_bf16, RowMap src_map, int nseq, int rows_per_seq,
                  float* dst, int src_blocked) {
  const long long total = (long long)nseq * rows_per_seq;
  const unsigned grid = (unsigned)((total + 3) / 4);
  if (src_bf16 == 2)
    to_f32_kernel<2><<<grid, 256, 0, st>>>(src, src_map, rows_per_seq, dst, total, 0);
  else if (src_bf16)
    to_f32_kernel<1><<<grid, 256, 0, st>>>(src, src_map, rows_per_seq, dst, total, 0);
  else
    to_f32_kernel<0><<<grid, 256, 0, st>>>(src, src_map, rows_per_seq, dst, total, src_blocked);
  return 1;
}

}  // namespace vapb
