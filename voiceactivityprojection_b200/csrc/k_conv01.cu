// conv0 + ChannelNorm + ReLU fused into the A-operand producer of conv1's CTA-pair GEMM: the first
// layer's activation (512 B per 6.25 ms frame, 65.5 MB per 20 s chunk) never reaches HBM.
// Reference: vap/encoder_components.py:83-86,99-100 (Conv1d(1,256,k10,s5,p3) + ChannelNorm + ReLU, then
// Conv1d(256,256,k8,s4,p2) + ChannelNorm + ReLU), ChannelNorm :62-70.
//
//   D[256 rows][256 ch] = A[256][2048] W1[256][2048]^T      tcgen05.mma.cta_group::2 M256 N256 K16 (as k_gemm_2sm.cu)
// Row t of conv1 reads conv0 frames 4t-2 .. 4t+5 (tap j = 0..7, 256 channels each), so k-block (j, cb) of A
// (64 channels of tap j) is  A_j[t][c] = y0[4t-2+j][64 cb + c]  and  A_{j+4}[t] = A_j[t+1]: taps j and j+4 hold the
// same frames one output row apart. The M rows of a CTA tile are therefore PERMUTED: MMA row r = 8g + i holds
// t = t0 + 16 i + g. In the 128-byte-swizzled K-major layout (8-row groups of 1024 B) "one output row later" is then
// "one row group later": the operand of tap j+4 is the operand of tap j with its start address advanced by 1024 B.
// The producer builds ONE 17-group block per (j, cb) (groups 0..16; group 16 is group 0 shifted by one row plus the
// frame of row t0+128) and the MMA thread issues both k-blocks from it: every conv0 output is computed 1.06 times
// instead of twice.
//
// conv0 on CUDA cores, one lane per output row: the lane scales its frame's 10 samples by the frame's
// 1/sqrt(var+eps) (closed form x'Gx + 2h.x + s, as k_conv0_tc.cu) and runs 12 packed FMAs per channel pair
// (10 taps, the folded bias d_c * rstd, the norm bias beta_c) whose weight operands come straight from the kernel
// parameter bank as uniform registers (FFMA2 R, R.F32, UR.F32x2, R): no shared-memory operand traffic at all;
// cvt.rn.relu packs two channels. Zero padding frames (f < 0, f >= L0) come out as exact zeros (rstd = 0, flag = 0).
//
// Roles per CTA (512 threads): warp 0 = TMA producer of W1 (both k-blocks of the stage, own 128 output channels),
// warp 1 = MMA issuer (leader CTA), warp 2 = TMEM allocation, warp 3 = the extra row t0+128 (lane = 8 channels),
// warps 4-11 = epilogue (bias -> ChannelNorm -> ReLU -> 16-bit -> swizzled staging -> TMA store that undoes the
// row permutation), warps 12-15 = conv0 producers. The stage's full barrier (leader CTA) counts the W bytes plus one
// arrival per producing warp of both CTAs.
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int F_STAGES = 3;
constexpr int F_A_BYTES = 17 * 1024;              // 17 row groups x (8 rows x 128 B)
constexpr int F_WK_BYTES = 128 * 64 * 2;          // one k-block of this CTA's half of W1
constexpr int F_STAGE_BYTES = F_A_BYTES + 2 * F_WK_BYTES;  // 49 KB
constexpr int F_OFF_STG = F_STAGES * F_STAGE_BYTES;        // [half][2] x 8 KB (128 rows x 64 B, SW64)
constexpr int F_OFF_BAR = F_OFF_STG + 4 * 8192;
constexpr int F_OFF_VEC = F_OFF_BAR + 256;
constexpr int F_WIN = 2585;                        // samples of one tile's window (rows t0 .. t0+128, taps 0..3)
constexpr int F_XS = 2624;                         // floats per window buffer (index skew: i + i / 320)
constexpr int F_THREADS = 512;
constexpr int F_PROD_WARP0 = 12, F_PROD_WARPS = 4;
constexpr int F_FULL_ARRIVALS = 1 + 2 * (F_PROD_WARPS + 1);  // leader's expect_tx + producing warps of both CTAs
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the even CTA

struct F01Vecs {
  float bias[256], g1[256], b1[256];
  float part[2][128][2];
  float G[10][12];  // Conv0Stats, rows padded to three float4 (read from shared memory: as kernel parameters the 111
  float h2[12];     // constants would be hoisted into uniform registers and starve the producers' weight loads)
  float s, pad_[3];
};
constexpr int F_OFF_XSB = F_OFF_VEC + (int)sizeof(F01Vecs);
constexpr int F_OFF_XP = F_OFF_XSB + 2 * F_XS * 4;       // float [2][128][12]: scaled frames of the current tap
constexpr int F_SMEM = F_OFF_XP + 2 * 128 * 12 * 4 + 1024;
static_assert(F_STAGE_BYTES % 1024 == 0 && F_OFF_STG % 1024 == 0, "SW128 operands need 1024-byte aligned bases");
static_assert(F_SMEM <= 232448, "shared memory budget");

struct alignas(64) F01Params {
  CUtensorMap tma_w;   // (2048, 256) 16-bit, box (64, 128), SW128
  CUtensorMap tma_o;   // store_mode 0: (256, 8, 16, tiles, nseq) box (32, 8, 16, 1, 1); 1: (256, 16, rows/16, nseq) box (32, 1, 8, 1); SW64
  float4 wq[12][64];   // [k][c/4]: k < 10 folded taps u_k, k = 10 folded bias d, k = 11 ChannelNorm bias beta
  Conv0Stats cs;
  const float* wav;
  const float* wg;     // the same 12 x 256 table in global memory (extra-row warp)
  const float *bias, *g1, *b1;
  long long n_samples;
  int batch, seq0, nseq, pair_tiles_per_seq;
  int L0, L1;
  int store_mode;
};

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_leader, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquires writes of the peer CTA too
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src, uint32_t src_bytes) {  // src_bytes 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int FP16>
__device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {  // max(., 0) and round to two 16-bit values
  uint32_t r;
  if (FP16) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 1 / sqrt(var + eps) over the 256 conv0 channels of a frame from its 10 samples (vap/encoder_components.py:62-70)
template <typename V>
__device__ __forceinline__ float frame_rstd(const V& cs, const float (&xv)[10]) {
  float ss = cs.s;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float4 g0 = *reinterpret_cast<const float4*>(&cs.G[k][0]), g1 = *reinterpret_cast<const float4*>(&cs.G[k][4]),
                 g2 = *reinterpret_cast<const float4*>(&cs.G[k][8]);
    float y = cs.h2[k];
    y = fmaf(g0.x, xv[0], y); y = fmaf(g0.y, xv[1], y); y = fmaf(g0.z, xv[2], y); y = fmaf(g0.w, xv[3], y);
    y = fmaf(g1.x, xv[4], y); y = fmaf(g1.y, xv[5], y); y = fmaf(g1.z, xv[6], y); y = fmaf(g1.w, xv[7], y);
    y = fmaf(g2.x, xv[8], y); y = fmaf(g2.y, xv[9], y);
    ss = fmaf(xv[k], y, ss);
  }
  return rsqrtf(fmaxf(ss, 0.f) * (1.0f / (kDim - 1)) + kEps);
}

// 16 channels (quarter V of channel block cb; wq0 = 16 cb + 4 V indexes the float4 weight table) of FOUR conv0 output
// rows (MMA rows lane + 32 m) -> the SW128 A block. One uniform weight load feeds four rows: per channel pair and tap
// one FFMA2 per row and a quarter of an LDCU. xp[m] = the row's 10 samples * rstd, rstd, valid flag.
template <int FP16, int V>
__device__ __forceinline__ void produce_q(const F01Params& p, const float (&xp)[4][12], int cb, int lane, uint8_t* ablk) {
  constexpr int v = V;
  const int wq0 = cb * 16 + V * 4;  // cb is a loop counter and V a constant: the weight loads stay on the uniform datapath
  const uint32_t sw = (uint32_t)(lane & 7);
  uint8_t* arow = ablk + (lane >> 3) * 1024 + (lane & 7) * 128;  // row lane + 32 m is 4 m row groups further
  const bool dup = lane >= 1 && lane < 8;                         // rows 1..7 of group 0 are rows 0..6 of group 16
  uint8_t* drow = ablk + 16 * 1024 + ((lane - 1) & 7) * 128;
  const uint32_t dsw = (uint32_t)((lane - 1) & 7);
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // 16-byte chunk = 8 channels
    uint32_t pk[4][4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c4 = wq0 + h * 2 + q;
      float2 a[4], b[4];
      {
        const float4 w = p.wq[11][c4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float2 x = make_float2(xp[m][11], xp[m][11]);
          a[m] = __fmul2_rn(x, make_float2(w.x, w.y));
          b[m] = __fmul2_rn(x, make_float2(w.z, w.w));
        }
      }
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float4 w = p.wq[k][c4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float2 x = make_float2(xp[m][k], xp[m][k]);
          a[m] = __ffma2_rn(x, make_float2(w.x, w.y), a[m]);
          b[m] = __ffma2_rn(x, make_float2(w.z, w.w), b[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        pk[m][2 * q] = pack_relu<FP16>(a[m].x, a[m].y);
        pk[m][2 * q + 1] = pack_relu<FP16>(b[m].x, b[m].y);
      }
    }
    const uint32_t chunk = (uint32_t)(2 * v + h);
#pragma unroll
    for (int m = 0; m < 4; ++m)
      *reinterpret_cast<uint4*>(arow + m * 4096 + ((chunk ^ sw) << 4)) = make_uint4(pk[m][0], pk[m][1], pk[m][2], pk[m][3]);
    if (dup) *reinterpret_cast<uint4*>(drow + ((chunk ^ dsw) << 4)) = make_uint4(pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
  }
}

// Epilogue of one CTA (warps 4-11) as its own function: the kernel body then holds only the producers' code, which
// lets ptxas keep a rotating set of uniform registers for their parameter-bank weight loads (inlined, the epilogue's
// register pressure collapsed that schedule to one load in flight).
struct F01EpiArgs {
  uint32_t smem_base, tmem_base, rank;
  uint8_t* smem_gen;
  int warp, lane, num_pair_tiles, n_clusters, cluster_id;
};
template <int FP16>
__device__ __noinline__ void conv01_epilogue(const F01Params& p, const F01EpiArgs a) {
  const uint32_t smem_base = a.smem_base, tmem_base = a.tmem_base, rank = a.rank;
  uint8_t* smem_gen = a.smem_gen;
  const int warp = a.warp, lane = a.lane, num_pair_tiles = a.num_pair_tiles, n_clusters = a.n_clusters,
            cluster_id = a.cluster_id;
  const uint32_t bar_base = smem_base + F_OFF_BAR;
  auto tfull_bar = [&](int k) { return bar_base + 8u * (8 + k); };
  auto tempty_bar = [&](int k) { return bar_base + 8u * (10 + k); };
  F01Vecs& ev = *reinterpret_cast<F01Vecs*>(smem_gen + F_OFF_VEC);
  auto tile_of = [&](int pt, int* lseq, int* t0) {
    *lseq = pt / p.pair_tiles_per_seq;
    *t0 = (pt % p.pair_tiles_per_seq) * 256 + (int)rank * 128;
  };
  {
  // ===== epilogue (both CTAs, own 128 rows): thread = (accumulator row, column half)
  const int quad = warp & 3, half = (warp - 4) >> 2;
  const int row_in_tile = quad * 32 + lane;                     // MMA row r = 8g + i
  const int dt = 16 * (lane & 7) + quad * 4 + (lane >> 3);      // t - t0 of that row
  const bool leader = (threadIdx.x - 128 - half * 128) == 0;
  const bool lead_warp = ((warp - 4) & 3) == 0;
  const int cbase = half * 128;
  const uint32_t stg_addr = smem_base + F_OFF_STG + half * 16384;
  uint8_t* stg_gen = smem_gen + F_OFF_STG + half * 16384;
  const uint32_t sw64 = (uint32_t)((row_in_tile >> 1) & 3);
  uint32_t stg_cnt = 0;
  auto bar_half = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory"); };
  auto bar_epi = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  const uint32_t tempty_leader0 = mapa(tempty_bar(0), 0), tempty_leader1 = mapa(tempty_bar(1), 0);
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
    int lseq, t0;
    tile_of(pt, &lseq, &t0);
    const bool row_ok = t0 + dt < p.L1, tile_ok = t0 < p.L1;
    mbar_wait(tfull_bar(acc), acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256 + cbase;
    float mean1, rstd1;
    {
      float s = 0.f, ss = 0.f;
      uint32_t rr[2][32];
      tmem_ld32(taddr, rr[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c < 3) tmem_ld32(taddr + (c + 1) * 32, rr[(c + 1) & 1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(rr[c & 1][i]) + ev.bias[cbase + c * 32 + i];
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
      const float var = fmaxf(ss - s * mean1, 0.f) * (1.0f / (kDim - 1));
      rstd1 = rsqrtf(var + kEps);
      bar_epi();
    }
    {
      uint32_t rr[2][32];
      tmem_ld32(taddr, rr[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c < 3) tmem_ld32(taddr + (c + 1) * 32, rr[(c + 1) & 1]);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bi = *reinterpret_cast<const float4*>(&ev.bias[cbase + c * 32 + i]);
          const float4 g = *reinterpret_cast<const float4*>(&ev.g1[cbase + c * 32 + i]);
          const float4 b = *reinterpret_cast<const float4*>(&ev.b1[cbase + c * 32 + i]);
          const float bb[4] = {bi.x, bi.y, bi.z, bi.w}, gg[4] = {g.x, g.y, g.z, g.w}, be[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = fmaf((__uint_as_float(rr[c & 1][i + j]) + bb[j] - mean1) * rstd1, gg[j], be[j]);
            v[i + j] = row_ok ? fmaxf(x, 0.f) : 0.f;  // rows past the sequence end are the next layer's zero padding
          }
        }
        if (p.store_mode == 0) {
          if (leader) bulk_wait_read<1>();
        } else if (lead_warp && lane < 16) {
          bulk_wait_read<1>();
        }
        bar_half();
        const uint32_t boff = (stg_cnt & 1u) * 8192u;
        uint8_t* rowp = stg_gen + boff + (uint32_t)row_in_tile * 64u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack16(v[8 * j], v[8 * j + 1], FP16);
          u.y = pack16(v[8 * j + 2], v[8 * j + 3], FP16);
          u.z = pack16(v[8 * j + 4], v[8 * j + 5], FP16);
          u.w = pack16(v[8 * j + 6], v[8 * j + 7], FP16);
          *reinterpret_cast<uint4*>(rowp + (((uint32_t)j ^ sw64) << 4)) = u;
        }
        fence_proxy_async();
        bar_half();
        // the staging tile is in MMA row order (8g + i); the store's box walks (channel, i, g) so row 8g + i lands on
        // output row t0 + 16 i + g
        if (p.store_mode == 0) {
          if (leader) {
            if (tile_ok) tma_store_5d(&p.tma_o, stg_addr + boff, cbase + c * 32, 0, 0, t0 >> 7, lseq);
            bulk_commit();
          }
        } else if (lead_warp && lane < 16) {
          if (tile_ok) tma_store_4d(&p.tma_o, stg_addr + boff + (uint32_t)lane * 512u, cbase + c * 32, lane, t0 >> 4, lseq);
          bulk_commit();
        }
        ++stg_cnt;
      }
    }
    tc_fence_before();
    mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
  if (p.store_mode == 0) {
    if (leader) bulk_wait<0>();
  } else if (lead_warp && lane < 16) {
    bulk_wait<0>();
  }
  }
}

template <int FP16>
__global__ void __launch_bounds__(F_THREADS, 1) conv01_kernel(const __grid_constant__ F01Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + F_OFF_BAR;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 12;
  F01Vecs& ev = *reinterpret_cast<F01Vecs*>(smem_gen + F_OFF_VEC);
  float* xs = reinterpret_cast<float*>(smem_gen + F_OFF_XSB);
  float* xps = reinterpret_cast<float*>(smem_gen + F_OFF_XP);
  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform (uniform registers for
  // the producers' parameter-bank weight loads)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs, owns the full barriers)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_w);
    prefetch_tmap(&p.tma_o);
    for (int s = 0; s < F_STAGES; ++s) {
      mbar_init(full_bar(s), F_FULL_ARRIVALS);
      mbar_init(empty_bar(s), 1);  // multicast MMA commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);     // multicast MMA commit
      mbar_init(tempty_bar(a), 512);  // the epilogue threads of both CTAs (leader's copy is the one used)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  if (warp >= 4 && warp < 12) {
    const int e = threadIdx.x - 128;
    ev.bias[e] = p.bias ? p.bias[e] : 0.f;
    ev.g1[e] = p.g1[e];
    ev.b1[e] = p.b1[e];
  }
  if (threadIdx.x == 96) {  // static indices only: a dynamically indexed parameter would be copied to local memory
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
      for (int l = 0; l < 12; ++l) ev.G[k][l] = l < 10 ? p.cs.G[k][l < 10 ? l : 0] : 0.f;
      ev.h2[k] = p.cs.h2[k];
    }
    ev.h2[10] = ev.h2[11] = 0.f;
    ev.s = p.cs.s;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM allocations exist before anyone signals across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int num_pair_tiles = p.nseq * p.pair_tiles_per_seq;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  auto tile_of = [&](int pt, int* lseq, int* t0) {
    *lseq = pt / p.pair_tiles_per_seq;
    *t0 = (pt % p.pair_tiles_per_seq) * 256 + (int)rank * 128;
  };

  if (warp == 0) {
    // ===== TMA producer of W1 (both CTAs): k-blocks (j, cb) and (j+4, cb), own 128 output channels
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
        for (int ss = 0; ss < 16; ++ss) {
          const int j = ss >> 2, cb = ss & 3;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 4 * F_WK_BYTES);
          const uint32_t w_dst = smem_base + stage * F_STAGE_BYTES + F_A_BYTES;
          const uint32_t bar_leader = full_bar(stage) & kPeerBitMask;
          tma_load_2d_2sm(w_dst, &p.tma_w, bar_leader, (j * 4 + cb) * 64, (int)rank * 128);
          tma_load_2d_2sm(w_dst + F_WK_BYTES, &p.tma_w, bar_leader, ((j + 4) * 4 + cb) * 64, (int)rank * 128);
          if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only): per stage, taps j and j+4 from the same A block one row group apart
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc_16(256, 256, 0, 0, FP16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int ss = 0; ss < 16; ++ss) {
          mbar_wait_cluster(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * F_STAGE_BYTES, w_addr = a_addr + F_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_16_2sm(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(w_addr + k * 32, 0, 1024),
                        idesc, (ss | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_16_2sm(d_tmem, make_smem_desc_sw128(a_addr + 1024 + k * 32, 0, 1024),
                        make_smem_desc_sw128(w_addr + F_WK_BYTES + k * 32, 0, 1024), idesc, 1);
          umma_commit_2sm(empty_bar(stage));
          if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3 || warp >= F_PROD_WARP0) {
    // ===== conv0 producers (warps 12-15) and the extra row t0 + 128 (warp 3)
    const bool extra = warp == 3;
    const int pw = warp - F_PROD_WARP0;
    const int ptid = pw * 32 + lane;                                    // producers: MMA row r = 8g + i of the x' pass
    const int dt = extra ? 128 : 16 * (ptid & 7) + (ptid >> 3);         // t - t0 of that row
    auto bar_all = [&]() { asm volatile("bar.sync 4, %0;" ::"n"((F_PROD_WARPS + 1) * 32) : "memory"); };
    auto bar_prod = [&]() { asm volatile("bar.sync 5, %0;" ::"n"(F_PROD_WARPS * 32) : "memory"); };
    auto prefetch = [&](int pt, int buf) {  // the tile's waveform window -> xs[buf] (zero outside the signal)
      int lseq, t0;
      tile_of(pt, &lseq, &t0);
      const int seq = p.seq0 + lseq;  // channel-major sequence id: c * batch + item
      const float* x = p.wav + ((long long)(seq % p.batch) * 2 + seq / p.batch) * p.n_samples;
      const long long sbase = 20LL * t0 - 13;
      const uint32_t dst = smem_base + F_OFF_XSB + (uint32_t)buf * (F_XS * 4);
      for (int i = ptid; i < F_WIN; i += F_PROD_WARPS * 32) {
        const long long s = sbase + i;
        const bool ok = s >= 0 && s < p.n_samples;
        cp_async4(dst + 4u * (uint32_t)(i + i / 320), ok ? x + s : x, ok ? 4u : 0u);
      }
      cp_async_commit();
    };
    int stage = 0, it = 0;
    uint32_t phase = 0;
    if (!extra && cluster_id < num_pair_tiles) prefetch(cluster_id, 0);
    for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++it) {
      const int buf = it & 1;
      int lseq, t0;
      tile_of(pt, &lseq, &t0);
      if (!extra) cp_async_wait_all();
      bar_all();  // xs[buf] is complete; everyone is done with xs[buf ^ 1] (read during the previous tile)
      if (!extra && pt + n_clusters < num_pair_tiles) prefetch(pt + n_clusters, buf ^ 1);
      const float* xw = xs + buf * F_XS;
      for (int j = 0; j < 4; ++j) {
        // this thread's frame: samples, 1 / sqrt(var + eps), validity (zero padding frames of conv1 give zeros)
        const int f = 4 * (t0 + dt) - 2 + j;
        const bool valid = f >= 0 && f < p.L0;
        const int i0 = 20 * dt + 5 * j;
        float xv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) xv[k] = xw[(i0 + k) + (i0 + k) / 320];
        const float rstd = valid ? frame_rstd(ev, xv) : 0.f;
        if (!extra) {
          // share the scaled frames: every producer warp needs all 128 rows (it owns 16 channels of each)
          float* xrow = xps + ((j & 1) * 128 + ptid) * 12;
          *reinterpret_cast<float4*>(xrow) = make_float4(xv[0] * rstd, xv[1] * rstd, xv[2] * rstd, xv[3] * rstd);
          *reinterpret_cast<float4*>(xrow + 4) = make_float4(xv[4] * rstd, xv[5] * rstd, xv[6] * rstd, xv[7] * rstd);
          *reinterpret_cast<float4*>(xrow + 8) = make_float4(xv[8] * rstd, xv[9] * rstd, rstd, valid ? 1.f : 0.f);
          bar_prod();  // (double-buffered on j: the reads of j - 1 finished before their owners arrived here)
          float xp[4][12];
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const float* src = xps + ((j & 1) * 128 + lane + 32 * m) * 12;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(src + 4 * q);
              xp[m][4 * q] = t4.x; xp[m][4 * q + 1] = t4.y; xp[m][4 * q + 2] = t4.z; xp[m][4 * q + 3] = t4.w;
            }
          }
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            uint8_t* ablk = smem_gen + stage * F_STAGE_BYTES;
            switch (pw) {  // one instantiation per channel quarter (a warp-dependent index would leave the uniform path)
              case 0: produce_q<FP16, 0>(p, xp, cb, lane, ablk); break;
              case 1: produce_q<FP16, 1>(p, xp, cb, lane, ablk); break;
              case 2: produce_q<FP16, 2>(p, xp, cb, lane, ablk); break;
              default: produce_q<FP16, 3>(p, xp, cb, lane, ablk); break;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(full_bar(stage), 0));
            if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          // row t0 + 128 (group 16, row 7): lane = channels 8 lane .. 8 lane + 7, weights from global memory (L1)
          float acc[8];
          {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.wg + 11 * kDim + lane * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.wg + 11 * kDim + lane * 8 + 4));
            const float fl = valid ? 1.f : 0.f;
            acc[0] = fl * b0.x; acc[1] = fl * b0.y; acc[2] = fl * b0.z; acc[3] = fl * b0.w;
            acc[4] = fl * b1.x; acc[5] = fl * b1.y; acc[6] = fl * b1.z; acc[7] = fl * b1.w;
          }
#pragma unroll
          for (int k = 0; k < 11; ++k) {
            const float xk = (k < 10 ? xv[k % 10] : 1.f) * rstd;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.wg + k * kDim + lane * 8));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.wg + k * kDim + lane * 8 + 4));
            acc[0] = fmaf(xk, w0.x, acc[0]); acc[1] = fmaf(xk, w0.y, acc[1]);
            acc[2] = fmaf(xk, w0.z, acc[2]); acc[3] = fmaf(xk, w0.w, acc[3]);
            acc[4] = fmaf(xk, w1.x, acc[4]); acc[5] = fmaf(xk, w1.y, acc[5]);
            acc[6] = fmaf(xk, w1.z, acc[6]); acc[7] = fmaf(xk, w1.w, acc[7]);
          }
          const uint4 u = make_uint4(pack_relu<FP16>(acc[0], acc[1]), pack_relu<FP16>(acc[2], acc[3]),
                                     pack_relu<FP16>(acc[4], acc[5]), pack_relu<FP16>(acc[6], acc[7]));
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if ((lane >> 3) == cb)
              *reinterpret_cast<uint4*>(smem_gen + stage * F_STAGE_BYTES + 16 * 1024 + 7 * 128 +
                                        ((((uint32_t)lane & 7u) ^ 7u) << 4)) = u;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(full_bar(stage), 0));
            if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    const F01EpiArgs ea{smem_base, tmem_base, rank, smem_gen, warp, lane, num_pair_tiles, n_clusters, cluster_id};
    conv01_epilogue<FP16>(p, ea);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace

// wav: (batch, 2, n_samples) fp32; sequences [seq0, seq0 + nseq) of the channel-major order c * batch + item.
// u / d / beta: folded conv0 parameters (conv0_v2_fold) on the HOST (they travel as kernel parameters) and the same
// table [12][256] = u | d | beta in device memory (wg). w1: conv1 weight [256][8 * 256] (16-bit, tap-major K), bias1 /
// g1 / b1: conv1 bias and ChannelNorm affine (device). out: row t of sequence s at out + s * out_seq_stride +
// (out_pad_rows + t) * 256; the kernel writes whole 128-row tiles (zeros past L1), so every sequence needs
// out_pad_rows + roundup(L1, 128) rows.
int launch_conv01(cudaStream_t st, const float* wav, int batch, long long n_samples, int seq0, int nseq, long long L0,
                  long long L1, const float* host_tab /*[12][256]*/, const float* dev_tab, const Conv0Stats& cs,
                  const void* w1, const float* bias1, const float* g1, const float* b1, void* out,
                  long long out_seq_stride, int out_pad_rows, int n_sm, std::string* err) {
  F01Params p{};
  {
    const uint64_t dims[2] = {(uint64_t)(8 * kDim), (uint64_t)kDim};
    const uint64_t strides[1] = {(uint64_t)(8 * kDim)};
    const uint32_t box[2] = {64, 128};
    if (!make_tmap_bf16(&p.tma_w, w1, 2, dims, strides, box, err)) return -1;
  }
  const long long tiles128 = (L1 + 127) / 128;
  char* obase = static_cast<char*>(out) + (long long)out_pad_rows * kDim * 2;
  static int force_mode = [] { const char* e = getenv("VAPB_CONV01_STORE"); return e ? atoi(e) : -1; }();
  p.store_mode = force_mode == 1 ? 1 : 0;
  if (p.store_mode == 0) {
    const uint64_t dims[5] = {(uint64_t)kDim, 8, 16, (uint64_t)tiles128, (uint64_t)nseq};
    const uint64_t strides[4] = {16ull * kDim, (uint64_t)kDim, 128ull * kDim, (uint64_t)out_seq_stride};
    const uint32_t box[5] = {32, 8, 16, 1, 1};
    std::string e0;
    if (!make_tmap(&p.tma_o, obase, 2, 5, dims, strides, box, 64, &e0)) {
      if (force_mode == 0) { if (err) *err = e0; return -1; }
      p.store_mode = 1;  // the driver refused strides that do not nest: one store per row group instead
    }
  }
  if (p.store_mode == 1) {
    const uint64_t dims[4] = {(uint64_t)kDim, 16, (uint64_t)(tiles128 * 8), (uint64_t)nseq};
    const uint64_t strides[3] = {(uint64_t)kDim, 16ull * kDim, (uint64_t)out_seq_stride};
    const uint32_t box[4] = {32, 1, 8, 1};
    if (!make_tmap(&p.tma_o, obase, 2, 4, dims, strides, box, 64, err)) return -1;
  }
  for (int k = 0; k < 12; ++k)
    for (int c4 = 0; c4 < 64; ++c4)
      p.wq[k][c4] = make_float4(host_tab[k * kDim + 4 * c4], host_tab[k * kDim + 4 * c4 + 1], host_tab[k * kDim + 4 * c4 + 2],
                                host_tab[k * kDim + 4 * c4 + 3]);
  p.cs = cs;
  p.wav = wav;
  p.wg = dev_tab;
  p.bias = bias1; p.g1 = g1; p.b1 = b1;
  p.n_samples = n_samples;
  p.batch = batch; p.seq0 = seq0; p.nseq = nseq;
  p.pair_tiles_per_seq = (int)((L1 + 255) / 256);
  p.L0 = (int)L0; p.L1 = (int)L1;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(conv01_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(conv01_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess) {
      if (err) *err = "conv01: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const int pair_tiles = p.nseq * p.pair_tiles_per_seq;
  int clusters = n_sm / 2;
  if (clusters > pair_tiles) clusters = pair_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(F_THREADS);
  cfg.dynamicSmemBytes = F_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t ce = g_fp16 ? cudaLaunchKernelEx(&cfg, conv01_kernel<1>, p) : cudaLaunchKernelEx(&cfg, conv01_kernel<0>, p);
  if (ce != cudaSuccess) {
    if (err) *err = std::string("conv01 launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 1;
}

}  // namespace vapb
