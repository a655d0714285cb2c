// BF16 implicit-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), with the
// same fused row epilogues as the fp32 kernel (k_gemm_f32.cu).
//
//   C[m, n] = sum_k A[row(m), k] * W[n, k]        A, W bf16 (K-major), C fp32 in TMEM
//
// * A tiles (128 rows x 64 k) come through a 3-D TMA tensor map (k, t, seq) whose
//   row stride is smaller than the row length for a strided conv: output frame t
//   of a Conv1d(k, s) over a channels-last activation is the contiguous span of
//   k*256 elements starting at padded input frame s*t, so consecutive TMA rows
//   overlap in memory and no im2col buffer exists
//   (vap/encoder_components.py:85-92,100-103; vap/encoder.py:24-30).
// * W tiles (256 n x 64 k) through a 2-D map over the [N][K] weight.
// * One CTA per SM, persistent over tiles; warp 0 = TMA producer, warp 1 = MMA
//   issuer (one thread, tcgen05.mma cta_group::1, M=128 N=256 K=16), warp 2
//   owns the TMEM allocation, warps 4-7 = epilogue. 4-stage smem ring; two
//   256-column fp32 accumulators in TMEM so the epilogue of tile i overlaps the
//   MMAs of tile i+1.
// * Epilogue: thread r of the 128 owns accumulator row r (tcgen05.ld 32x32b), so
//   the channel norm that follows every conv / precedes every transformer op is a
//   thread-local reduction over the 256 columns:
//     v = acc + bias -> ChannelNorm|LayerNorm -> ReLU|GELU -> + residual
//       (-> += previous out) -> out1 (fp32 and/or bf16) ; out2 = LayerNorm2(v) bf16
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

constexpr int TBM = 128, TBN = 256, TBK = 64, TSTAGES = 4;
constexpr int TA_BYTES = TBM * TBK * 2, TB_BYTES = TBN * TBK * 2;
constexpr int TSTAGE_BYTES = TA_BYTES + TB_BYTES;
constexpr int TC_SMEM = TSTAGES * TSTAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + (1024 + 4 * 256 + 512) * 4 /*EpiVecs*/;

struct alignas(64) TcGemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  int nseq, rows_per_seq, tiles_per_seq, n_tiles_n, num_k_blocks;
  // epilogue
  const float* bias;
  int norm1;
  const float *g1, *b1;
  int act;
  const float* resid;
  RowMap resid_map;
  int accumulate;            // v += out1_f32 (previous contents)
  float* out1_f32;           // optional fp32 copy of v
  __nv_bfloat16* out1_bf16;  // optional bf16 copy of v
  RowMap out1_map;           // shared by both out1 copies (same logical shape, N columns)
  int norm2;
  const float *g2, *b2;
  __nv_bfloat16* out2_bf16;
  RowMap out2_map;
  int N;
};

__device__ __forceinline__ void store_chunk_f32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void store_chunk_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16(v[8 * i], v[8 * i + 1]);
    u.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(dst + 8 * i) = u;
  }
}

constexpr int TC_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int TC_EPI_THREADS = 256;

// Per-column epilogue vectors staged once per CTA (they do not change per tile).
struct EpiVecs {
  float bias[1024];
  float g1[256], b1[256], g2[256], b2[256];
  float part[2][128][2];  // per-row partial (sum, sumsq) of the two column halves
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + TSTAGES * TSTAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TSTAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TSTAGES + 4);
  EpiVecs& ev = *reinterpret_cast<EpiVecs*>(smem_gen + TSTAGES * TSTAGE_BYTES + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp >= 4) {
    const int e = threadIdx.x - 128;  // 0..255
    for (int i = e; i < 1024; i += TC_EPI_THREADS) ev.bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    ev.g1[e] = p.norm1 != NORM_NONE ? p.g1[e] : 1.f;
    ev.b1[e] = p.norm1 != NORM_NONE ? p.b1[e] : 0.f;
    ev.g2[e] = p.norm2 != NORM_NONE ? p.g2[e] : 1.f;
    ev.b2[e] = p.norm2 != NORM_NONE ? p.b2[e] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int num_m_tiles = p.nseq * p.tiles_per_seq;
  const int num_tiles = num_m_tiles * p.n_tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
        const int seq = mt / p.tiles_per_seq, t0 = (mt % p.tiles_per_seq) * TBM;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), TSTAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * TSTAGE_BYTES;
          tma_load_3d(a_dst, &p.tma_a, full_bar(stage), kb * TBK, t0, seq);
          tma_load_2d(a_dst + TA_BYTES, &p.tma_b, full_bar(stage), kb * TBK, nt * TBN);
          if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TBM, TBN, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TBN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * TSTAGE_BYTES;
          const uint32_t b_addr = a_addr + TA_BYTES;
#pragma unroll
          for (int k = 0; k < TBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128-byte swizzle atom
            const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // thread = (column half, TMEM lane quadrant, lane): row quad*32+lane, columns [128*half, +128)
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int cbase = half * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
      const int seq = mt / p.tiles_per_seq, t = (mt % p.tiles_per_seq) * TBM + row_in_tile;
      const bool valid = t < p.rows_per_seq;
      const int n0 = nt * TBN + cbase;  // first global column of this thread
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TBN + cbase;

      float mean1 = 0.f, rstd1 = 1.f;
      if (p.norm1 != NORM_NONE) {
        float s = 0.f, ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(r[i]) + ev.bias[n0 + c * 32 + i];
            s += v;
            ss = fmaf(v, v, ss);
          }
        }
        ev.part[half][row_in_tile][0] = s;
        ev.part[half][row_in_tile][1] = ss;
        epi_bar();
        s += ev.part[half ^ 1][row_in_tile][0];
        ss += ev.part[half ^ 1][row_in_tile][1];
        mean1 = s * (1.0f / kDim);
        const float var = fmaxf(ss - s * mean1, 0.f) * (p.norm1 == NORM_CHANNEL ? 1.0f / (kDim - 1) : 1.0f / kDim);
        rstd1 = rsqrtf(var + kEps);
        epi_bar();  // part[] may be rewritten below
      }

      const long long o1 = (long long)seq * p.out1_map.seq_stride + (long long)t * p.out1_map.row_stride + n0;
      const float* resid = p.resid ? p.resid + (long long)seq * p.resid_map.seq_stride +
                                         (long long)t * p.resid_map.row_stride + n0
                                   : nullptr;
      float s2 = 0.f, ss2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(r[i]) + ev.bias[n0 + c * 32 + i];
          if (p.norm1 != NORM_NONE)
            x = fmaf((x - mean1) * rstd1, ev.g1[cbase + c * 32 + i], ev.b1[cbase + c * 32 + i]);
          v[i] = apply_act(x, p.act);
        }
        if (valid) {
          if (resid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = *reinterpret_cast<const float4*>(resid + c * 32 + 4 * i);
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          if (p.accumulate) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = *reinterpret_cast<const float4*>(p.out1_f32 + o1 + c * 32 + 4 * i);
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          if (p.out1_f32) store_chunk_f32(p.out1_f32 + o1 + c * 32, v);
          if (p.out1_bf16) store_chunk_bf16(p.out1_bf16 + o1 + c * 32, v);
        }
        if (p.norm2 != NORM_NONE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s2 += v[i];
            ss2 = fmaf(v[i], v[i], ss2);
            r[i] = __float_as_uint(v[i]);
          }
          tmem_st32(taddr + c * 32, r);  // keep v for the LayerNorm2 pass
        }
      }
      if (p.norm2 != NORM_NONE) {
        tmem_st_wait();
        ev.part[half][row_in_tile][0] = s2;
        ev.part[half][row_in_tile][1] = ss2;
        epi_bar();
        s2 += ev.part[half ^ 1][row_in_tile][0];
        ss2 += ev.part[half ^ 1][row_in_tile][1];
        const float mean2 = s2 * (1.0f / kDim);
        const float var2 = fmaxf(ss2 - s2 * mean2, 0.f) * (1.0f / kDim);
        const float rstd2 = rsqrtf(var2 + kEps);
        epi_bar();
        const long long o2 = (long long)seq * p.out2_map.seq_stride + (long long)t * p.out2_map.row_stride + n0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = fmaf((__uint_as_float(r[i]) - mean2) * rstd2, ev.g2[cbase + c * 32 + i], ev.b2[cbase + c * 32 + i]);
          if (valid) store_chunk_bf16(p.out2_bf16 + o2 + c * 32, v);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host ---------------------------------------------------------------------
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

int launch_gemm_tc(cudaStream_t st, const TcGemmArgs& a, int n_sm, std::string* err) {
  if (a.N % TBN || a.K % TBK) {
    if (err) *err = "gemm_tc: N must be a multiple of 256 and K of 64";
    return -1;
  }
  if ((a.e.norm1 != NORM_NONE || a.e.norm2 != NORM_NONE) && a.N != TBN) {
    if (err) *err = "gemm_tc: row norms need N == 256";
    return -1;
  }
  TcGemmParams p{};
  {
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.rows_per_seq, (uint64_t)a.nseq};
    const uint64_t strides[2] = {(uint64_t)a.a_map.row_stride,
                                 (uint64_t)(a.nseq > 1 ? a.a_map.seq_stride : a.a_map.row_stride * a.rows_per_seq)};
    const uint32_t box[3] = {TBK, TBM, 1};
    if (!make_tmap_bf16(&p.tma_a, a.A, 3, dims, strides, box, err)) return -1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    const uint64_t strides[1] = {(uint64_t)a.K};
    const uint32_t box[2] = {TBK, TBN};
    if (!make_tmap_bf16(&p.tma_b, a.W, 2, dims, strides, box, err)) return -1;
  }
  p.nseq = a.nseq;
  p.rows_per_seq = a.rows_per_seq;
  p.tiles_per_seq = (a.rows_per_seq + TBM - 1) / TBM;
  p.n_tiles_n = a.N / TBN;
  p.num_k_blocks = a.K / TBK;
  p.N = a.N;
  const Epilogue& e = a.e;
  p.bias = e.bias;
  p.norm1 = e.norm1; p.g1 = e.g1; p.b1 = e.b1;
  p.act = e.act;
  p.resid = e.resid; p.resid_map = e.resid_map;
  p.accumulate = e.accumulate;
  p.out1_f32 = a.out1_f32;
  p.out1_bf16 = a.out1_bf16;
  p.out1_map = e.out1_map;
  p.norm2 = e.norm2; p.g2 = e.g2; p.b2 = e.b2;
  p.out2_bf16 = reinterpret_cast<__nv_bfloat16*>(e.out2);
  p.out2_map = e.out2_map;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    configured = true;
  }
  const int tiles = p.nseq * p.tiles_per_seq * p.n_tiles_n;
  const int grid = tiles < n_sm ? tiles : n_sm;
  gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(p);
  return 1;
}

}  // namespace vapb
