// fp32-class GEMM on the tensor cores for the parity mode: fp32 operands in memory, every element split into two fp16
// parts on the fly (x = hi + lo, 22 significant bits), three tcgen05 MMAs per K step
//     A W^T  ~=  A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T          (the dropped lo x lo term is 2^-22 relative)
// accumulated in fp32 in TMEM, and the same fused row epilogue as the CUDA-core kernel it replaces
// (k_gemm_f32.cu: bias -> ChannelNorm | LayerNorm -> ReLU | GELU(erf) -> + residual -> (+= out) -> out1 ; out2 = LayerNorm2).
// Serves every contraction of the fp32 path: the implicit-GEMM convolutions (vap/encoder_components.py:85-92,100-103,
// vap/encoder.py:24-30), the gAR input projection, and the Linear layers of vap/modules.py.
//
// * A tiles (128 rows x 32 k, fp32) come through a 3-D TMA map (k, t, seq) with 128-byte swizzle into a staging buffer;
//   rows may overlap in memory (strided conv over a channels-last activation, no im2col). Four converter warps
//   (thread = row) split them into hi / lo fp16 operand tiles in the un-swizzled K-major core-matrix layout.
// * W is split and laid out on the HOST (x3_pack_weight): per (256-column tile, 32-k block) one contiguous 32 KB block
//   [hi | lo][k chunk of 8][256 rows][8 halves], fetched with one bulk copy.
// * warp 0 = producer (TMA + bulk copy), warp 1 = MMA issuer, warp 2 = TMEM allocation, warps 4-7 = converters,
//   warps 8-15 = epilogue (thread = accumulator row x column half; two-pass statistics like the fp32 kernel).
#include <cuda_fp16.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int X_BM = 128, X_BN = 256, X_BK = 32, X_STAGES = 3;
constexpr int X_RAW = X_BM * X_BK * 4;      // 16 KB fp32 staging (SW128: rows of 128 B)
constexpr int X_APL = X_BM * X_BK * 2;      // 8 KB per A plane: [4 chunks][128 rows][16 B]
constexpr int X_WPL = X_BN * X_BK * 2;      // 16 KB per W plane: [4 chunks][256 rows][16 B]
constexpr int X_STAGE = X_RAW + 2 * X_APL + 2 * X_WPL;  // 64 KB
constexpr int X_OFF_BAR = X_STAGES * X_STAGE;
constexpr int X_OFF_VEC = X_OFF_BAR + 256;
constexpr int X_THREADS = 512;

struct X3Vecs {
  float bias[1024];
  float g1[256], b1[256], g2[256], b2[256];
  float part[2][128][2];
};
constexpr int X_SMEM = X_OFF_VEC + (int)sizeof(X3Vecs) + 1024;
static_assert(X_SMEM <= 232448, "shared memory budget");

struct alignas(64) X3Params {
  CUtensorMap tma_a;   // (K, rows_per_seq, nseq) fp32, box (32, 128, 1), SW128
  const __half* w;     // packed split weight (x3_pack_weight)
  int nseq, rows_per_seq, tiles_per_seq, n_tiles_n, num_k_blocks, N;
  const float* bias;
  int norm1;
  const float *g1, *b1;
  int act;
  const float* resid;
  RowMap resid_map;
  int accumulate;
  float* out1;
  RowMap out1_map;
  int norm2;
  const float *g2, *b2;
  float* out2;
  RowMap out2_map;
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(X_THREADS, 1) gemm_x3_kernel(const __grid_constant__ X3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + X_OFF_BAR;
  auto raw_full = [&](int s) { return bar_base + 8u * s; };            // TMA A + bulk W have landed
  auto op_full = [&](int s) { return bar_base + 8u * (4 + s); };       // converters have written the A planes (4 warps)
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };     // MMA commit: the stage may be refilled
  auto tfull_bar = [&](int a) { return bar_base + 8u * (12 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (14 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 16;
  X3Vecs& ev = *reinterpret_cast<X3Vecs*>(smem_gen + X_OFF_VEC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    for (int s = 0; s < X_STAGES; ++s) {
      mbar_init(raw_full(s), 1);
      mbar_init(op_full(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp >= 8) {
    const int e = threadIdx.x - 256;  // 0..255
    for (int i = e; i < 1024; i += 256) ev.bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
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
    // ===== producer: the fp32 A tile (TMA) and the packed hi | lo W block (one bulk copy) of every k-block
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
        const int seq = mt / p.tiles_per_seq, t0 = (mt % p.tiles_per_seq) * X_BM;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(raw_full(stage), X_RAW + 2 * X_WPL);
          const uint32_t base = smem_base + stage * X_STAGE;
          tma_load_3d(base, &p.tma_a, raw_full(stage), kb * X_BK, t0, seq);
          bulk_load(base + X_RAW + 2 * X_APL, p.w + ((size_t)nt * p.num_k_blocks + kb) * (2 * X_WPL / 2), 2 * X_WPL,
                    raw_full(stage));
          if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: per K = 16 step  hi x hi, lo x hi, hi x lo
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_16(X_BM, X_BN, 0, 0, 1);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * X_BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(op_full(stage), phase);
          tc_fence_after();
          const uint32_t a_hi = smem_base + stage * X_STAGE + X_RAW, a_lo = a_hi + X_APL;
          const uint32_t w_hi = a_lo + X_APL, w_lo = w_hi + X_WPL;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t ah = make_smem_desc_nosw(a_hi + k * 4096, 2048, 128), al = make_smem_desc_nosw(a_lo + k * 4096, 2048, 128);
            const uint64_t wh = make_smem_desc_nosw(w_hi + k * 8192, 4096, 128), wl = make_smem_desc_nosw(w_lo + k * 8192, 4096, 128);
            umma_bf16(d_tmem, ah, wh, idesc, (kb | k) != 0);
            umma_bf16(d_tmem, al, wh, idesc, 1);
            umma_bf16(d_tmem, ah, wl, idesc, 1);
          }
          umma_commit(empty_bar(stage));
          if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== converters: thread = row of the tile; fp32 (SW128 staging) -> fp16 hi / lo operand planes
    const int r = threadIdx.x - 128;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(raw_full(stage), phase);
        const uint8_t* raw = smem_gen + stage * X_STAGE + r * 128;
        uint8_t* hi_pl = smem_gen + stage * X_STAGE + X_RAW + r * 16;
        uint8_t* lo_pl = hi_pl + X_APL;
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // k chunk of 8 = two 16-byte groups of the fp32 row
          uint32_t h[4], l[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(raw + ((((uint32_t)(2 * j + q)) ^ (uint32_t)(r & 7)) << 4));
            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
            h[2 * q] = *reinterpret_cast<const uint32_t*>(&h0);
            h[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            l[2 * q] = *reinterpret_cast<const uint32_t*>(&l0);
            l[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&l1);
          }
          *reinterpret_cast<uint4*>(hi_pl + j * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(lo_pl + j * 2048) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(op_full(stage));
        if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===== epilogue: thread = (accumulator row, column half); statistics in two passes like k_gemm_f32.cu
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int cbase = half * 128;
    auto epi_bar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // sum over the row's 256 columns of f(value): this thread's 128, then the other half's through shared memory
    auto row_total = [&](float mine) {
      ev.part[half][row_in_tile][0] = mine;
      epi_bar();
      const float tot = mine + ev.part[half ^ 1][row_in_tile][0];
      epi_bar();
      return tot;
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
      const int seq = mt / p.tiles_per_seq, t = (mt % p.tiles_per_seq) * X_BM + row_in_tile;
      const bool valid = t < p.rows_per_seq;
      const int n0 = nt * X_BN + cbase;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * X_BN + cbase;
      float mean1 = 0.f, rstd1 = 1.f;
      if (p.norm1 != NORM_NONE) {
        float s = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]) + ev.bias[n0 + c * 32 + i];
        }
        mean1 = row_total(s) * (1.0f / kDim);
        float q = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = __uint_as_float(r[i]) + ev.bias[n0 + c * 32 + i] - mean1;
            q = fmaf(d, d, q);
          }
        }
        const float var = row_total(q) * (p.norm1 == NORM_CHANNEL ? 1.0f / (kDim - 1) : 1.0f / kDim);
        rstd1 = 1.0f / sqrtf(var + kEps);
      }
      const long long o1 = (long long)seq * p.out1_map.seq_stride + (long long)t * p.out1_map.row_stride + n0;
      const float* resid = p.resid ? p.resid + (long long)seq * p.resid_map.seq_stride +
                                         (long long)t * p.resid_map.row_stride + n0
                                   : nullptr;
      float s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(r[i]) + ev.bias[n0 + c * 32 + i];
          if (p.norm1 != NORM_NONE) x = fmaf((x - mean1) * rstd1, ev.g1[cbase + c * 32 + i], ev.b1[cbase + c * 32 + i]);
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
              const float4 q = *reinterpret_cast<const float4*>(p.out1 + o1 + c * 32 + 4 * i);
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(p.out1 + o1 + c * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (p.norm2 != NORM_NONE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s2 += v[i];
            r[i] = __float_as_uint(v[i]);
          }
          tmem_st32(taddr + c * 32, r);  // keep v for the LayerNorm2 passes
        }
      }
      if (p.norm2 != NORM_NONE) {
        tmem_st_wait();
        const float mean2 = row_total(s2) * (1.0f / kDim);
        float q2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = __uint_as_float(r[i]) - mean2;
            q2 = fmaf(d, d, q2);
          }
        }
        const float rstd2 = 1.0f / sqrtf(row_total(q2) * (1.0f / kDim) + kEps);
        const long long o2 = (long long)seq * p.out2_map.seq_stride + (long long)t * p.out2_map.row_stride + n0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o;
              o.x = fmaf((__uint_as_float(r[i]) - mean2) * rstd2, ev.g2[cbase + c * 32 + i], ev.b2[cbase + c * 32 + i]);
              o.y = fmaf((__uint_as_float(r[i + 1]) - mean2) * rstd2, ev.g2[cbase + c * 32 + i + 1], ev.b2[cbase + c * 32 + i + 1]);
              o.z = fmaf((__uint_as_float(r[i + 2]) - mean2) * rstd2, ev.g2[cbase + c * 32 + i + 2], ev.b2[cbase + c * 32 + i + 2]);
              o.w = fmaf((__uint_as_float(r[i + 3]) - mean2) * rstd2, ev.g2[cbase + c * 32 + i + 3], ev.b2[cbase + c * 32 + i + 3]);
              *reinterpret_cast<float4*>(p.out2 + o2 + c * 32 + i) = o;
            }
          }
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

}  // namespace

// Host: split an fp32 weight Wt[K][N] (the fp32 path's packing, K index contiguous per output column is NOT assumed:
// element (k, n) at wt[k * N + n]) into fp16 hi / lo parts in the kernel's block layout:
// [N / 256][K / 32][plane hi, lo][k chunk of 8][256 rows][8].
void x3_pack_weight(const float* wt, int K, int N, std::vector<__half>* out) {
  const int nt = N / X_BN, nk = K / X_BK;
  out->assign((size_t)N * K * 2, __float2half_rn(0.f));
  for (int t = 0; t < nt; ++t)
    for (int kb = 0; kb < nk; ++kb) {
      __half* blk = out->data() + ((size_t)t * nk + kb) * (2 * X_BN * X_BK);
      for (int n = 0; n < X_BN; ++n)
        for (int k = 0; k < X_BK; ++k) {
          const float v = wt[(size_t)(kb * X_BK + k) * N + t * X_BN + n];
          const __half h = __float2half_rn(v);
          const __half l = __float2half_rn(v - __half2float(h));
          const size_t off = (size_t)(k >> 3) * (X_BN * 8) + (size_t)n * 8 + (k & 7);
          blk[off] = h;
          blk[X_BN * X_BK + off] = l;
        }
    }
}

// Same problem / epilogue conventions as launch_gemm_f32 (all buffers fp32), w_packed from x3_pack_weight (device).
// Returns launches or -1 (unsupported shape: the caller falls back to the CUDA-core kernel).
int launch_gemm_x3(cudaStream_t st, const GemmProblem& g, const Epilogue& e, const void* w_packed, int n_sm,
                   std::string* err) {
  if (g.N % X_BN || g.K % X_BK || g.N > 1024 || g.M % g.rows_per_seq ||
      ((e.norm1 != NORM_NONE || e.norm2 != NORM_NONE) && g.N != X_BN) || e.out1_bf16 || e.out2_bf16) {
    if (err) *err = "gemm_x3: unsupported problem";
    return -1;
  }
  X3Params p{};
  const int nseq = g.M / g.rows_per_seq;
  {
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.rows_per_seq, (uint64_t)nseq};
    const uint64_t strides[2] = {(uint64_t)g.a_map.row_stride,
                                 (uint64_t)(nseq > 1 ? g.a_map.seq_stride : g.a_map.row_stride * g.rows_per_seq)};
    const uint32_t box[3] = {X_BK, X_BM, 1};
    if (!make_tmap(&p.tma_a, g.A, 4, 3, dims, strides, box, 128, err)) return -1;
  }
  p.w = static_cast<const __half*>(w_packed);
  p.nseq = nseq;
  p.rows_per_seq = g.rows_per_seq;
  p.tiles_per_seq = (g.rows_per_seq + X_BM - 1) / X_BM;
  p.n_tiles_n = g.N / X_BN;
  p.num_k_blocks = g.K / X_BK;
  p.N = g.N;
  p.bias = e.bias;
  p.norm1 = e.norm1; p.g1 = e.g1; p.b1 = e.b1;
  p.act = e.act;
  p.resid = e.resid; p.resid_map = e.resid_map;
  p.accumulate = e.accumulate;
  p.out1 = static_cast<float*>(e.out1);
  p.out1_map = e.out1_map;
  p.norm2 = e.norm2; p.g2 = e.g2; p.b2 = e.b2;
  p.out2 = static_cast<float*>(e.out2);
  p.out2_map = e.out2_map;
  static bool configured_on[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM) != cudaSuccess) {
      if (err) *err = "gemm_x3: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const int tiles = nseq * p.tiles_per_seq * p.n_tiles_n;
  const int grid = tiles < n_sm ? tiles : n_sm;
  gemm_x3_kernel<<<grid, X_THREADS, X_SMEM, st>>>(p);
  return 1;
}

}  // namespace vapb
