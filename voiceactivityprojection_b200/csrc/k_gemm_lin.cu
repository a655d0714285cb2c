// BF16 tcgen05 GEMM for the row-local layers of the transformer (q/k/v/proj, FFN,
// combinator, vap_head; vap/modules.py:9-21,93-95,109,246-275,434-449,
// vap/model.py:261) and the downsample conv (vap/encoder.py:24-30).
//
// Mainloop of the first-generation GEMM (TMA -> smem ring -> tcgen05.mma M128 N256 K16, two
// TMEM accumulators) but these GEMMs have K = 256..1280, so a 128 x 256 tile spends
// ~1 us on the tensor pipe and the kernel lives or dies by its epilogue and its HBM
// traffic. The epilogue therefore moves data only in full lines:
//   * bf16 outputs (the next GEMM's / attention's TMA-loaded operands) are packed into
//     a 128-byte-swizzled staging tile in shared memory (conflict-free: lane r writes
//     16-byte chunk j at r*128 + ((j ^ (r&7)) << 4)) and leave with ONE TMA store per
//     128 x 64 tile; rows beyond the sequence are clipped by the tensor map;
//   * the fp32 residual stream is private to the library, so it is kept in a
//     row-blocked layout [row/128][col/4][row%128][4]: the thread that owns
//     accumulator row r reads/writes float4s that are contiguous across the warp;
//   * fp32 row-major outputs (logits) go through the same staging tiles (128 x 32).
// With K = 256 (every q/k/v/proj/FFN-in GEMM) the kernel runs "weight-resident": a CTA owns one 256-column
// slice of W for its whole life (128 KB of shared memory, loaded once) and streams only A tiles, because at
// 1 us of tensor time per tile re-fetching 128 KB of W per tile made these GEMMs L2->SM-bandwidth bound
// (measured: 2.36 GB through L2 for a 1 GB qkv GEMM). CTAs that share an M tile but own different W slices
// walk the M tiles in lockstep, so A is fetched from HBM once.
// TMEM loads are double-buffered against the math; GELU is a packed-FMA polynomial (gelu_poly2).
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int LBM = 128, LBN = 256, LBK = 64;
constexpr int LA_BYTES = LBM * LBK * 2, LB_BYTES = LBN * LBK * 2;
constexpr int LSTAGE_BYTES = LA_BYTES + LB_BYTES;   // 48 KB
constexpr int LSTG_HALF = 128 * 128;                // staging per column half: 2 x (128 rows x 64 B) bf16 tiles (SW64),
                                                    // or 1 x (128 rows x 128 B) fp32 tile (SW128)
constexpr int L_THREADS = 384, L_EPI_THREADS = 256;
// shared-memory plan: [operand region][staging 2 x 16 KB][barriers 256 B][LinVecs]
//   streaming:        3 stages x (A 16 KB + W 32 KB)
//   weight-resident:  W 4 x 32 KB, then 3 stages x A 16 KB
//   NST = 4 (the K >= 1024 convolutions: the tensor pipe idled 29 % of the time waiting for operands with 3 stages
//   in flight, ncu) pays for the fourth stage with single-buffered bf16 staging (8 KB per half instead of 16)
template <bool WRES, int NST>
struct LinSmem {
  static constexpr int NSTG = NST == 4 ? 1 : 2;  // bf16 staging buffers per column half
  static constexpr int OPERANDS = WRES ? 4 * LB_BYTES + NST * LA_BYTES : NST * LSTAGE_BYTES;
  static constexpr int OFF_STG = OPERANDS;
  static constexpr int STG_HALF = NSTG == 2 ? LSTG_HALF : LSTG_HALF / 2;
  static constexpr int OFF_BAR = OFF_STG + 2 * STG_HALF;
  static constexpr int OFF_VEC = OFF_BAR + 256;
};

struct LinVecs {
  float bias[256], g1[256], b1[256], g2[256], b2[256];
  float part[2][128][2];
};
template <bool WRES, int NST>
constexpr int lin_smem_bytes() { return LinSmem<WRES, NST>::OFF_VEC + (int)sizeof(LinVecs) + 1024 /*alignment slack*/; }

struct alignas(64) LinParams {
  CUtensorMap tma_a, tma_b;
  CUtensorMap tma_o1;   // out1 bf16  (N, rows, nseq) box (32,128,1) SW64
  CUtensorMap tma_o2;   // out2 bf16
  CUtensorMap tma_of;   // out1 fp32 row-major (N, rows, nseq) box (32,128,1) SW128
  int nseq, rows_per_seq, tiles_per_seq, n_tiles_n, num_k_blocks, N;
  const float* bias;
  int norm1;
  const float *g1, *b1;
  int act;
  const float* resid;     // fp32, blocked layout
  int accumulate;         // v += previous out1_f32 (blocked)
  float* out1_f32;        // blocked layout (f32_mode 1) or row-major through tma_of (f32_mode 2)
  int f32_mode;
  int has_o1_bf16;
  int norm2;
  const float *g2, *b2;
  int fp16;  // 16-bit operand / output format: 0 bf16, 1 fp16
  int prefetch;  // L2-prefetch the A rows of the tile this many tiles ahead (env VAPB_LIN_PREFETCH; 0 = off)
};

// GELU(erf) for two values with packed fp32 FMAs and no MUFU: gelu(x) = relu(x) + g(|x|), where
// g(a) = -a * Phi(-a) is smooth on [0, inf) and below 1.3e-4 beyond a = 4; g is a degree-7 polynomial on [0, 4]
// (Chebyshev fit, max abs error of the whole GELU 2.2e-4 in fp32 Horner form — below the bf16 rounding of the
// values it produces). tanh.approx made the FFN-in epilogue MUFU-bound (ncu: 515 us vs 238 us for the same GEMM
// without the activation).
__device__ __forceinline__ float2 gelu_poly2(float2 x) {
  const float2 a = make_float2(fminf(fabsf(x.x), 4.0f), fminf(fabsf(x.y), 4.0f));
  float2 acc = make_float2(6.604950176551938e-4f, 6.604950176551938e-4f);
  acc = __ffma2_rn(acc, a, make_float2(-1.0183836333453655e-2f, -1.0183836333453655e-2f));
  acc = __ffma2_rn(acc, a, make_float2(5.9112582355737686e-2f, 5.9112582355737686e-2f));
  acc = __ffma2_rn(acc, a, make_float2(-1.445402055978775e-1f, -1.445402055978775e-1f));
  acc = __ffma2_rn(acc, a, make_float2(4.981609806418419e-2f, 4.981609806418419e-2f));
  acc = __ffma2_rn(acc, a, make_float2(3.8568615913391113e-1f, 3.8568615913391113e-1f));
  acc = __ffma2_rn(acc, a, make_float2(-4.9917131662368774e-1f, -4.9917131662368774e-1f));
  acc = __ffma2_rn(acc, a, make_float2(1.4609939171350561e-5f, 1.4609939171350561e-5f));
  return __fadd2_rn(acc, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float2 act_fast2(float2 v, int act) {
  if (act == ACT_RELU) return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
  if (act == ACT_GELU) return gelu_poly2(v);
  return v;
}

// offset (floats) of (row m, column c) in the blocked fp32 layout; c % 4 == 0 for float4 access
__device__ __forceinline__ long long blocked_off(long long m, int c) {
  return (((m >> 7) * 64 + (c >> 2)) * 128 + (m & 127)) * 4 + (c & 3);
}

__device__ __forceinline__ void bar_half(int half) {
  asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
}
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

}  // namespace

// Epilogue features compiled into an instantiation. The do-everything epilogue is ~140 KB of SASS and its warps
// stalled on instruction fetch (ncu: stall_no_instruction was the top stall reason), so each launch picks the
// smallest instantiation that covers what it needs.
enum : int {
  F_PRE = 1,     // bias and/or norm1 on the accumulator
  F_ACT = 2,     // ReLU / GELU
  F_RESID = 4,   // + residual (blocked fp32)
  F_ACC = 8,     // += previous out1_f32 (blocked fp32)
  F_F32B = 16,   // out1_f32, blocked
  F_F32R = 32,   // out1_f32, row-major through TMA
  F_O1B = 64,    // out1 bf16
  F_N2 = 128,    // out2 = LayerNorm2(v) bf16
  F_ALL = 255
};

template <bool WRES, int F, int NST>
__global__ void __launch_bounds__(L_THREADS, 1) gemm_lin_kernel(const __grid_constant__ LinParams p) {
  constexpr int LSTAGES = NST, NSTG = LinSmem<WRES, NST>::NSTG, STG_HALF = LinSmem<WRES, NST>::STG_HALF;
  constexpr bool C_PRE = F & F_PRE, C_ACT = F & F_ACT, C_RESID = F & F_RESID, C_ACC = F & F_ACC, C_F32B = F & F_F32B,
                 C_F32R = F & F_F32R, C_O1B = F & F_O1B, C_N2 = F & F_N2;
  constexpr int L_OFF_STG = LinSmem<WRES, NST>::OFF_STG, L_OFF_BAR = LinSmem<WRES, NST>::OFF_BAR,
                L_OFF_VEC = LinSmem<WRES, NST>::OFF_VEC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L_OFF_BAR;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 12;
  const uint32_t w_full = bar_base + 8u * 14;
  LinVecs& ev = *reinterpret_cast<LinVecs*>(smem_gen + L_OFF_VEC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    for (int s = 0; s < LSTAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), L_EPI_THREADS);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp >= 4) {
    const int e = threadIdx.x - 128;
    ev.bias[e] = p.bias ? p.bias[e] : 0.f;
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
  // PDL: everything above read only weights; the producer thread also requests its resident weight slice first
  if (!(warp == 0 && lane == 0)) pdl_wait();

  // Tile walk. Streaming: (mt, nt) pairs strided over the grid. Weight-resident: the CTA's nt is fixed
  // (blockIdx % n_tiles_n) and it strides over M tiles with the CTAs of its group.
  const int num_m_tiles = p.nseq * p.tiles_per_seq;
  const int groups = WRES ? (int)gridDim.x / p.n_tiles_n : 1;
  const int my_nt = WRES ? (int)blockIdx.x % p.n_tiles_n : 0;
  const int tile_first = WRES ? ((int)blockIdx.x / p.n_tiles_n < groups ? (int)blockIdx.x / p.n_tiles_n : num_m_tiles)
                              : (int)blockIdx.x;
  const int tile_stride = WRES ? groups : (int)gridDim.x;
  const int num_tiles = WRES ? num_m_tiles : num_m_tiles * p.n_tiles_n;
  auto tile_mt = [&](int tile) { return WRES ? tile : tile / p.n_tiles_n; };
  auto tile_nt = [&](int tile) { return WRES ? my_nt : tile % p.n_tiles_n; };
  constexpr uint32_t A_RING = WRES ? 4 * LB_BYTES : 0;            // first byte of the A ring
  constexpr uint32_t A_STRIDE = WRES ? LA_BYTES : LSTAGE_BYTES;   // bytes between ring stages

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (WRES && tile_first < num_tiles) {
        mbar_arrive_expect_tx(w_full, 4 * LB_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem_base + kb * LB_BYTES, &p.tma_b, w_full, kb * LBK, my_nt * LBN);
      }
      pdl_wait();
      // The ring holds 3-4 k-blocks (48-64 KB) per SM; at ~1.5 us of DRAM latency that is ~32 GB/s per SM, which is
      // what the K = 256 GEMMs ran at (ncu: the epilogue warps idle on tfull, the tensor pipe 44 % active). The A rows
      // of the tile PF tiles ahead are therefore pulled into L2 while this tile loads.
      const int PF = p.prefetch;
      auto prefetch_tile = [&](int tile) {
        if (tile >= num_tiles) return;
        const int mt = tile_mt(tile);
        const int seq = mt / p.tiles_per_seq, t0 = (mt % p.tiles_per_seq) * LBM;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) tma_prefetch_3d(&p.tma_a, kb * LBK, t0, seq);
      };
      for (int i = 1; i < PF; ++i) prefetch_tile(tile_first + i * tile_stride);
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        const int mt = tile_mt(tile), nt = tile_nt(tile);
        const int seq = mt / p.tiles_per_seq, t0 = (mt % p.tiles_per_seq) * LBM;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), WRES ? LA_BYTES : LSTAGE_BYTES);
          const uint32_t a_dst = smem_base + A_RING + stage * A_STRIDE;
          tma_load_3d(a_dst, &p.tma_a, full_bar(stage), kb * LBK, t0, seq);
          if (!WRES) tma_load_2d(a_dst + LA_BYTES, &p.tma_b, full_bar(stage), kb * LBK, nt * LBN);
          if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
        }
        if (PF) prefetch_tile(tile + PF * tile_stride);  // behind this tile's own loads in the TMA queue
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_16(LBM, LBN, 0, 0, p.fp16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      if (WRES && tile_first < num_tiles) {
        mbar_wait(w_full, 0);
        tc_fence_after();
      }
      for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * LBN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + A_RING + stage * A_STRIDE;
          const uint32_t b_addr = WRES ? smem_base + kb * LB_BYTES : a_addr + LA_BYTES;
#pragma unroll
          for (int k = 0; k < LBK / 16; ++k)
            umma_bf16(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(b_addr + k * 32, 0, 1024),
                      idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));
          if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // thread = (column half, TMEM lane quadrant, lane): accumulator row quad*32+lane, columns [128*half, +128)
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const bool leader = (threadIdx.x - 128 - half * 128) == 0;  // issues this half's TMA stores
    const int cbase = half * 128;
    const uint32_t stg_addr = smem_base + L_OFF_STG + half * STG_HALF;
    uint8_t* stg_gen = smem_gen + L_OFF_STG + half * STG_HALF;
    const uint32_t sw128 = (uint32_t)(row_in_tile & 7), sw64 = (uint32_t)((row_in_tile >> 1) & 3);
    uint32_t stg_cnt = 0;
    // bf16 staging: 32 columns of the tile = 128 rows x 64 B (SWIZZLE_64B: chunk j of row r at r*64 + ((j ^ (r/2 & 3)) << 4),
    // conflict-free for 8 consecutive rows), two buffers. All 128 threads of the half write their row, then the
    // leader hands the tile to TMA; a buffer is reused once the store that read it two tiles ago has drained.
    auto stage_bf16 = [&](const CUtensorMap* map, const float (&v)[32], int c0, int c1, int c2) {
      if (leader) bulk_wait_read<NSTG - 1>();
      bar_half(half);
      const uint32_t boff = NSTG == 2 ? (stg_cnt & 1u) * 8192u : 0u;
      uint8_t* rowp = stg_gen + boff + (uint32_t)row_in_tile * 64u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack16(v[8 * j], v[8 * j + 1], p.fp16);
        u.y = pack16(v[8 * j + 2], v[8 * j + 3], p.fp16);
        u.z = pack16(v[8 * j + 4], v[8 * j + 5], p.fp16);
        u.w = pack16(v[8 * j + 6], v[8 * j + 7], p.fp16);
        *reinterpret_cast<uint4*>(rowp + (((uint32_t)j ^ sw64) << 4)) = u;
      }
      fence_proxy_async();
      bar_half(half);
      if (leader) {
        tma_store_3d(map, stg_addr + boff, c0, c1, c2);
        bulk_commit();
      }
      ++stg_cnt;
    };
    // fp32 row-major staging (logits): 32 columns = 128 rows x 128 B (SWIZZLE_128B), single buffer
    auto stage_f32 = [&](const CUtensorMap* map, const float (&v)[32], int c0, int c1, int c2) {
      if (leader) bulk_wait_read<0>();
      bar_half(half);
      uint8_t* rowp = stg_gen + (uint32_t)row_in_tile * 128u;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(rowp + (((uint32_t)j ^ sw128) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async();
      bar_half(half);
      if (leader) {
        tma_store_3d(map, stg_addr, c0, c1, c2);
        bulk_commit();
      }
    };

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_first; tile < num_tiles; tile += tile_stride) {
      const int mt = tile_mt(tile), nt = tile_nt(tile);
      const int seq = mt / p.tiles_per_seq, t0 = (mt % p.tiles_per_seq) * LBM;
      const int t = t0 + row_in_tile;
      const bool valid = t < p.rows_per_seq;
      const long long m = (long long)seq * p.rows_per_seq + t;  // dense row (blocked fp32 buffers)
      const int n0 = nt * LBN + cbase;                            // first global column of this thread
      if (C_RESID && p.resid && valid && (lane & 7) == 0) {
        // pull the residual rows of this tile into L2 while its MMAs run (the adds below stalled on DRAM latency)
#pragma unroll
        for (int c = 0; c < 32; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + blocked_off(m, n0 + c * 4)));
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * LBN + cbase;

      float mean1 = 0.f, rstd1 = 1.f;
      if (C_PRE && p.norm1 != NORM_NONE) {
        float s = 0.f, ss = 0.f;
        uint32_t r[2][32];
        tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(r[c & 1][i]) + ev.bias[cbase + c * 32 + i];
            s += v;
            ss = fmaf(v, v, ss);
          }
        }
        ev.part[half][row_in_tile][0] = s;
        ev.part[half][row_in_tile][1] = ss;
        bar_epi();
        s += ev.part[half ^ 1][row_in_tile][0];
        ss += ev.part[half ^ 1][row_in_tile][1];
        mean1 = s * (1.0f / kDim);
        const float var = fmaxf(ss - s * mean1, 0.f) * (p.norm1 == NORM_CHANNEL ? 1.0f / (kDim - 1) : 1.0f / kDim);
        rstd1 = rsqrtf(var + kEps);
        bar_epi();
      }

      float s2 = 0.f, ss2 = 0.f;
      {
        uint32_t r[2][32];
        tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
          float v[32];
          if (C_PRE && (p.bias || p.norm1 != NORM_NONE)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bi = *reinterpret_cast<const float4*>(&ev.bias[cbase + c * 32 + i]);
              const float4 g = *reinterpret_cast<const float4*>(&ev.g1[cbase + c * 32 + i]);
              const float4 b = *reinterpret_cast<const float4*>(&ev.b1[cbase + c * 32 + i]);
              const float bb[4] = {bi.x, bi.y, bi.z, bi.w}, gg[4] = {g.x, g.y, g.z, g.w}, be[4] = {b.x, b.y, b.z, b.w};
              float x[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                x[j] = __uint_as_float(r[c & 1][i + j]) + bb[j];
                if (p.norm1 != NORM_NONE) x[j] = fmaf((x[j] - mean1) * rstd1, gg[j], be[j]);
              }
#pragma unroll
              for (int j = 0; j < 4; j += 2) {
                const float2 y = C_ACT ? act_fast2(make_float2(x[j], x[j + 1]), p.act) : make_float2(x[j], x[j + 1]);
                v[i + j] = y.x;
                v[i + j + 1] = y.y;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 x = make_float2(__uint_as_float(r[c & 1][i]), __uint_as_float(r[c & 1][i + 1]));
              const float2 y = C_ACT ? act_fast2(x, p.act) : x;
              v[i] = y.x;
              v[i + 1] = y.y;
            }
          }
          if (C_RESID && p.resid && valid) {
            const float* rp = p.resid + blocked_off(m, n0 + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = *reinterpret_cast<const float4*>(rp + i * 512);
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          if (C_ACC && p.accumulate && valid) {
            const float* ap = p.out1_f32 + blocked_off(m, n0 + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = *reinterpret_cast<const float4*>(ap + i * 512);
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          if (C_F32B && p.f32_mode == 1) {
            if (valid) {
              float* op = p.out1_f32 + blocked_off(m, n0 + c * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(op + i * 512) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          } else if (C_F32R && p.f32_mode == 2) {
            stage_f32(&p.tma_of, v, n0 + c * 32, t0, seq);
          }
          if (C_O1B && p.has_o1_bf16) stage_bf16(&p.tma_o1, v, n0 + c * 32, t0, seq);
          if (C_N2 && p.norm2 != NORM_NONE) {
            uint32_t w[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s2 += v[i];
              ss2 = fmaf(v[i], v[i], ss2);
              w[i] = __float_as_uint(v[i]);
            }
            tmem_st32(taddr + c * 32, w);  // keep v for the LayerNorm2 pass
          }
        }
      }
      if (C_N2 && p.norm2 != NORM_NONE) {
        tmem_st_wait();
        ev.part[half][row_in_tile][0] = s2;
        ev.part[half][row_in_tile][1] = ss2;
        bar_epi();
        s2 += ev.part[half ^ 1][row_in_tile][0];
        ss2 += ev.part[half ^ 1][row_in_tile][1];
        const float mean2 = s2 * (1.0f / kDim);
        const float var2 = fmaxf(ss2 - s2 * mean2, 0.f) * (1.0f / kDim);
        const float rstd2 = rsqrtf(var2 + kEps);
        bar_epi();
        uint32_t r[2][32];
        tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g = *reinterpret_cast<const float4*>(&ev.g2[cbase + c * 32 + i]);
            const float4 b = *reinterpret_cast<const float4*>(&ev.b2[cbase + c * 32 + i]);
            v[i] = fmaf((__uint_as_float(r[c & 1][i]) - mean2) * rstd2, g.x, b.x);
            v[i + 1] = fmaf((__uint_as_float(r[c & 1][i + 1]) - mean2) * rstd2, g.y, b.y);
            v[i + 2] = fmaf((__uint_as_float(r[c & 1][i + 2]) - mean2) * rstd2, g.z, b.z);
            v[i + 3] = fmaf((__uint_as_float(r[c & 1][i + 3]) - mean2) * rstd2, g.w, b.w);
          }
          stage_bf16(&p.tma_o2, v, n0 + c * 32, t0, seq);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (leader) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Arguments as TcGemmArgs (common.cuh), plus the fp32 layout choices:
//   f32_mode 1: out1_f32 / resid / accumulate use the blocked layout (dense row index seq*rows_per_seq + t)
//   f32_mode 2: out1_f32 is row-major (e.out1_map) and written through TMA staging; resid stays blocked.
int launch_gemm_lin(cudaStream_t st, const TcGemmArgs& a, int f32_mode, int n_sm, std::string* err) {
  if (a.N % LBN || a.K % LBK) {
    if (err) *err = "gemm_lin: N must be a multiple of 256 and K of 64";
    return -1;
  }
  const Epilogue& e = a.e;
  if ((e.norm1 != NORM_NONE || e.norm2 != NORM_NONE || e.bias) && a.N != LBN) {
    if (err) *err = "gemm_lin: bias / row norms need N == 256";
    return -1;
  }
  if ((e.resid || e.accumulate || (a.out1_f32 && f32_mode == 1)) && a.N != LBN) {
    if (err) *err = "gemm_lin: blocked fp32 rows need N == 256";
    return -1;
  }
  LinParams p{};
  {
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.rows_per_seq, (uint64_t)a.nseq};
    const uint64_t strides[2] = {(uint64_t)a.a_map.row_stride,
                                 (uint64_t)(a.nseq > 1 ? a.a_map.seq_stride : a.a_map.row_stride * a.rows_per_seq)};
    const uint32_t box[3] = {LBK, LBM, 1};
    if (!make_tmap_bf16(&p.tma_a, a.A, 3, dims, strides, box, err)) return -1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    const uint64_t strides[1] = {(uint64_t)a.K};
    const uint32_t box[2] = {LBK, LBN};
    if (!make_tmap_bf16(&p.tma_b, a.W, 2, dims, strides, box, err)) return -1;
  }
  auto out_map = [&](CUtensorMap* m, void* base, const RowMap& rm, int elem_bytes) {
    const uint64_t dims[3] = {(uint64_t)a.N, (uint64_t)a.rows_per_seq, (uint64_t)a.nseq};
    const uint64_t strides[2] = {(uint64_t)rm.row_stride,
                                 (uint64_t)(a.nseq > 1 ? rm.seq_stride : rm.row_stride * a.rows_per_seq)};
    const uint32_t box[3] = {32, 128, 1};
    return make_tmap(m, base, elem_bytes, 3, dims, strides, box, elem_bytes == 2 ? 64 : 128, err);
  };
  if (a.out1_f32 && f32_mode == 2 && (a.out1_bf16 || e.norm2 != NORM_NONE)) {
    if (err) *err = "gemm_lin: a row-major fp32 output cannot be combined with bf16 outputs";
    return -1;
  }
  if (a.out1_bf16 && !out_map(&p.tma_o1, a.out1_bf16, e.out1_map, 2)) return -1;
  if (e.norm2 != NORM_NONE && !out_map(&p.tma_o2, e.out2, e.out2_map, 2)) return -1;
  if (a.out1_f32 && f32_mode == 2 && !out_map(&p.tma_of, a.out1_f32, e.out1_map, 4)) return -1;
  p.nseq = a.nseq;
  p.rows_per_seq = a.rows_per_seq;
  p.tiles_per_seq = (a.rows_per_seq + LBM - 1) / LBM;
  p.n_tiles_n = a.N / LBN;
  p.num_k_blocks = a.K / LBK;
  p.N = a.N;
  p.bias = e.bias;
  p.norm1 = e.norm1; p.g1 = e.g1; p.b1 = e.b1;
  p.act = e.act;
  p.resid = e.resid;
  p.accumulate = e.accumulate;
  p.out1_f32 = a.out1_f32;
  p.f32_mode = a.out1_f32 ? f32_mode : 0;
  p.has_o1_bf16 = a.out1_bf16 != nullptr;
  p.norm2 = e.norm2; p.g2 = e.g2; p.b2 = e.b2;
  p.fp16 = g_fp16;
  {
    // measured on 512 000 x 256 inputs (tools/lin_probe.py): N = 768 243 -> 191 us and N = 512 169 -> 133 us two tiles
    // ahead, N = 256 97 -> 90 us one tile ahead (two: 95, three or more: slower than none)
    static const int env_pf = [] { const char* v = getenv("VAPB_LIN_PREFETCH"); return v ? atoi(v) : -1; }();
    // with a residual the epilogue already prefetches 128 KB of fp32 rows per tile and A prefetch costs 2-9 %
    const int auto_pf = a.e.resid ? 0 : (a.N > LBN ? 2 : 1);
    p.prefetch = a.K <= 512 ? (env_pf >= 0 ? env_pf : auto_pf) : 0;  // long-K tiles would park hundreds of KB per SM in L2
  }
  const int need = ((e.bias || e.norm1 != NORM_NONE) ? F_PRE : 0) | (e.act != ACT_NONE ? F_ACT : 0) |
                   (e.resid ? F_RESID : 0) | (e.accumulate ? F_ACC : 0) | (p.f32_mode == 1 ? F_F32B : 0) |
                   (p.f32_mode == 2 ? F_F32R : 0) | (p.has_o1_bf16 ? F_O1B : 0) | (e.norm2 != NORM_NONE ? F_N2 : 0);
  typedef void (*KFn)(const LinParams);
  struct Variant { int mask; KFn fn[2]; bool configured[2][64]; int nst; };  // configured: per device
#define VAPB_LIN_VARIANT(M) {M, {gemm_lin_kernel<false, M, 3>, gemm_lin_kernel<true, M, 3>}, {}, 3}
#define VAPB_LIN_VARIANT4(M) {M, {gemm_lin_kernel<false, M, 4>, gemm_lin_kernel<true, M, 3>}, {}, 4}
  static Variant variants[] = {
      VAPB_LIN_VARIANT(F_O1B),                                        // q/k/v, cross k/v, cross q
      VAPB_LIN_VARIANT(F_O1B | F_ACT),                                // FFN in + GELU
      VAPB_LIN_VARIANT(F_RESID | F_F32B | F_N2),                      // attention out-projection
      VAPB_LIN_VARIANT(F_RESID | F_F32B | F_O1B | F_N2),              // FFN out
      VAPB_LIN_VARIANT4(F_PRE | F_ACT | F_O1B),                       // gEncoder convs (bias + ChannelNorm + ReLU), 4 stages
      VAPB_LIN_VARIANT(F_PRE | F_F32R),                               // vap_head
      VAPB_LIN_VARIANT(F_PRE | F_ACT | F_ACC | F_F32B | F_O1B),       // combinator
      VAPB_LIN_VARIANT(F_PRE | F_ACT | F_F32B | F_O1B | F_N2),        // downsample conv
      VAPB_LIN_VARIANT(F_ALL),
  };
#undef VAPB_LIN_VARIANT
#undef VAPB_LIN_VARIANT4
  Variant* v = nullptr;
  for (auto& cand : variants)
    if ((need & ~cand.mask) == 0) { v = &cand; break; }
  const int tiles = p.nseq * p.tiles_per_seq * p.n_tiles_n;
  const int grid = tiles < n_sm ? tiles : n_sm;
  const int wres = (a.K == 4 * LBK && grid >= p.n_tiles_n) ? 1 : 0;
  const int smem = wres ? lin_smem_bytes<true, 3>() : (v->nst == 4 ? lin_smem_bytes<false, 4>() : lin_smem_bytes<false, 3>());
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  cur_dev &= 63;
  if (!v->configured[wres][cur_dev]) {
    if (cudaFuncSetAttribute(v->fn[wres], cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      if (err) *err = "gemm_lin: cannot reserve shared memory";
      return -1;
    }
    v->configured[wres][cur_dev] = true;
  }
  launch_pdl(v->fn[wres], grid, L_THREADS, smem, st, p);
  return 1;
}

}  // namespace vapb
