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
// conv0 itself is a second, tiny tensor-core GEMM per tap j: D0[256 rows][256 ch] = X'_j[256][32] U[256][32]^T, fp16
// operands whatever the mode (the products are internal), K = 32 = (x' hi | x' lo): a producer lane scales its frame's
// 10 samples by the frame's 1/sqrt(var+eps) (closed form x'Gx + 2h.x + s, as k_conv0_tc.cu), splits them into fp16
// hi + lo parts and writes one 64-byte operand row (10 taps, rstd for the folded bias d_c, a validity flag for the norm
// bias beta_c; lo parts in the second half); U holds fp16(u | d | beta) twice. D0 lives in the 256 TMEM columns of the
// conv1 accumulator that is idle at that moment (tap 0: the tile's own accumulator before its main loop starts, taps
// 1-3: the other one after its epilogue has drained it), so conv1 keeps both accumulators. The producers then only move
// D0: tcgen05.ld -> cvt.rn.relu pack -> swizzled A block (the fourth channel block is read early into registers so the
// TMEM columns are released before a ring stage is free). Zero padding frames (f < 0, f >= L0) have x' = 0, flag = 0:
// exact zeros. A first version computed conv0 on CUDA cores with weights streamed from the parameter bank (LDCU +
// FFMA2): correct, but 3-5x too slow (the 12 KB table thrashes the constant cache; profiles/r2_conv01_history.md).
//
// Roles per CTA (512 threads): warp 0 = TMA producer of W1 (both k-blocks of the stage, own 128 output channels),
// warp 1 = MMA issuer (leader CTA: conv1 stages and the conv0 GEMMs, interleaved C C Z C C per tap so neither waits on
// the other), warp 2 = TMEM allocation, warp 3 = the extra row t0+128 (CUDA cores, lane = 8 channels), warps 4-11 =
// epilogue (bias -> ChannelNorm -> ReLU -> 16-bit -> swizzled staging -> TMA store that undoes the row permutation),
// warps 12-15 = conv0 producers (lane = MMA row). The stage's full barrier (leader CTA) counts the W bytes plus one
// arrival per producing warp of both CTAs.
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

// diagnostics (tools/conv01_probe.py): SM-clock timers, compiled only into the probe build (-DF01_TIMERS)
#ifdef F01_TIMERS
#define F01_T(...) __VA_ARGS__
#else
#define F01_T(...)
#endif

constexpr int F_STAGES = 3;
constexpr int F_A_BYTES = 17 * 1024;              // 17 row groups x (8 rows x 128 B)
constexpr int F_WK_BYTES = 128 * 64 * 2;          // one k-block of this CTA's half of W1
constexpr int F_STAGE_BYTES = F_A_BYTES + 2 * F_WK_BYTES;  // 49 KB
constexpr int F_OFF_STG = F_STAGES * F_STAGE_BYTES;        // [epilogue warp][2] x 2 KB (32 rows x 64 B, SW64)
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
constexpr int F_OFF_XA = F_OFF_XSB + 2 * F_XS * 4;       // conv0 A operand: fp16 [4 k-chunks][128 rows][8] (no swizzle)
constexpr int F_OFF_U = F_OFF_XA + 4 * 128 * 16;         // conv0 B operand: fp16 [4 k-chunks][128 own channels][8]
constexpr int F_SMEM = F_OFF_U + 4 * 128 * 16 + 1024;
static_assert(F_STAGE_BYTES % 1024 == 0 && F_OFF_STG % 1024 == 0, "SW128 operands need 1024-byte aligned bases");
static_assert(F_SMEM <= 232448, "shared memory budget");

struct alignas(64) F01Params {
  CUtensorMap tma_w;   // (2048, 256) 16-bit, box (64, 128), SW128
  CUtensorMap tma_o;   // store_mode 0: (256, 8, 16, tiles, nseq) box (32, 8, 4, 1, 1); 1: (256, 16, rows/16, nseq) box (32, 1, 8, 1); SW64
  Conv0Stats cs;
  const void* wav;     // (batch, 2, n_samples) fp32, or int16 PCM (scaled by 1/32768 when read; needs n_samples even)
  const float* wg;     // folded conv0 table [12][256] in global memory: k < 10 taps u_k, k = 10 bias d, k = 11 norm bias beta
  const float *bias, *g1, *b1;
  long long n_samples;
  int batch, seq0, nseq, pair_tiles_per_seq;
  int L0, L1;
  int store_mode;
  long long* dbg;  // diagnostics (tools/conv01_probe.py): SM-clock stamps of cluster 0 / CTA 0, [tile iteration < 4][tap][16 events], or null
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
// Arrival with the default (.release.cta) semantics on a barrier of this or the peer CTA, as CUTLASS's ClusterBarrier::
// arrive: the data it publishes was written by this warp to its OWN shared memory and made visible to the async proxy
// with fence.proxy.async; the .release.cluster form above costs ~1000 cycles per arrival (measured: 38 % of the
// producers' time) and sits on their critical path four times per tap.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_test_cluster(uint32_t bar, uint32_t parity) {  // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
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
// cs.G holds the symmetric matrix folded to its upper triangle (off-diagonal entries doubled, lower triangle zero):
// 55 multiply-adds instead of 100, and only the float4 groups that hold non-zeros are read.
template <typename V>
__device__ __forceinline__ float frame_rstd(const V& cs, const float (&xv)[10]) {
  float ss = cs.s;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    float y = cs.h2[k];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (4 * q + 3 < k) continue;  // this group lies below the diagonal
      const float4 g = *reinterpret_cast<const float4*>(&cs.G[k][4 * q]);
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (4 * q + e >= k && 4 * q + e < 10) y = fmaf(gv[e], xv[4 * q + e], y);
    }
    ss = fmaf(xv[k], y, ss);
  }
  return rsqrtf(fmaxf(ss, 0.f) * (1.0f / (kDim - 1)) + kEps);
}

// Epilogue of one CTA (warps 4-11) as its own function (its register allocation stays separate from the producers').
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
  // ===== epilogue (both CTAs, own 128 rows): thread = (accumulator row, column half). Every warp stages and stores its
  // own 32 rows (2 KB per 32-channel chunk, two buffers): no block-level barrier on the path that releases the
  // accumulator - the next tile's conv0 GEMMs for taps 1-3 are waiting for it.
  const int quad = warp & 3, half = (warp - 4) >> 2;
  const int row_in_tile = quad * 32 + lane;                     // MMA row r = 8g + i
  const int dt = 16 * (lane & 7) + quad * 4 + (lane >> 3);      // t - t0 of that row
  const int cbase = half * 128;
  const uint32_t stg_addr = smem_base + F_OFF_STG + (uint32_t)(warp - 4) * 4096u;
  uint8_t* stg_gen = smem_gen + F_OFF_STG + (warp - 4) * 4096;
  const uint32_t sw64 = (uint32_t)((lane >> 1) & 3);
  uint32_t stg_cnt = 0;
  auto bar_epi = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  const uint32_t tempty_leader0 = mapa(tempty_bar(0), 0), tempty_leader1 = mapa(tempty_bar(1), 0);
  const bool storer = p.store_mode == 0 ? lane == 0 : lane < 4;  // lanes that issue this warp's TMA stores
  // diagnostics: cycles from accumulator-ready to statistics, to release, to tile end; tiles
  F01_T(long long etm[4] = {0, 0, 0, 0}; const bool etm_on = p.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;)
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
    int lseq, t0;
    tile_of(pt, &lseq, &t0);
    const bool row_ok = t0 + dt < p.L1, tile_ok = t0 < p.L1;
    mbar_wait(tfull_bar(acc), acc_phase);
    F01_T(long long te0 = 0, te1 = 0; if (etm_on) te0 = clock64();)
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256 + cbase;
    // Pass 1: mean / variance of (acc + bias) over the row's 256 channels (this thread's 128, then the other half's
    // partial sums through shared memory). Packed fp32 arithmetic throughout: the epilogue sits on the path that
    // releases the accumulator for the next tile's conv0 GEMMs.
    float mean1, rstd1;
    {
      float2 s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, q2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t rr[2][32];
      tmem_ld32(taddr, rr[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c < 3) tmem_ld32(taddr + (c + 1) * 32, rr[(c + 1) & 1]);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bi = *reinterpret_cast<const float4*>(&ev.bias[cbase + c * 32 + i]);
          const float2 v0 = __fadd2_rn(make_float2(__uint_as_float(rr[c & 1][i]), __uint_as_float(rr[c & 1][i + 1])),
                                       make_float2(bi.x, bi.y));
          const float2 v1 = __fadd2_rn(make_float2(__uint_as_float(rr[c & 1][i + 2]), __uint_as_float(rr[c & 1][i + 3])),
                                       make_float2(bi.z, bi.w));
          s2[0] = __fadd2_rn(s2[0], v0);
          s2[1] = __fadd2_rn(s2[1], v1);
          q2[0] = __ffma2_rn(v0, v0, q2[0]);
          q2[1] = __ffma2_rn(v1, v1, q2[1]);
        }
      }
      float s = (s2[0].x + s2[0].y) + (s2[1].x + s2[1].y), ss = (q2[0].x + q2[0].y) + (q2[1].x + q2[1].y);
      ev.part[half][row_in_tile][0] = s;
      ev.part[half][row_in_tile][1] = ss;
      bar_epi();
      s += ev.part[half ^ 1][row_in_tile][0];
      ss += ev.part[half ^ 1][row_in_tile][1];
      mean1 = s * (1.0f / kDim);
      const float var = fmaxf(ss - s * mean1, 0.f) * (1.0f / (kDim - 1));
      rstd1 = rsqrtf(var + kEps);
      bar_epi();
      F01_T(if (etm_on) { te1 = clock64(); etm[0] += te1 - te0; })
    }
    // Pass 2: out = relu(((acc + bias) - mean) * rstd * g + b) as three packed FMAs and a cvt.relu pack per channel
    // pair. All four 32-channel chunks are computed into registers first, so the accumulator is released after the
    // last TMEM load and the staging / TMA stores happen off that path.
    uint32_t pk[4][16];
    {
      const float2 r2 = make_float2(rstd1, rstd1), m2 = make_float2(-mean1 * rstd1, -mean1 * rstd1);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {  // 16 columns at a time (register budget: the 64 packed outputs stay live)
        const int c = c8 >> 1, i0 = (c8 & 1) * 16;
        uint32_t rr[16];
        tmem_ld16(taddr + c8 * 16, rr);
        tmem_ld_wait();
        if (c8 == 7) {
          tc_fence_before();
          mbar_arrive_remote(acc == 0 ? tempty_leader0 : tempty_leader1);
          F01_T(if (etm_on) { const long long t2 = clock64(); etm[1] += t2 - te1; te1 = t2; })
        }
#pragma unroll
        for (int ii = 0; ii < 16; ii += 4) {
          const int i = i0 + ii;
          const float4 bi = *reinterpret_cast<const float4*>(&ev.bias[cbase + c * 32 + i]);
          const float4 g = *reinterpret_cast<const float4*>(&ev.g1[cbase + c * 32 + i]);
          const float4 b = *reinterpret_cast<const float4*>(&ev.b1[cbase + c * 32 + i]);
          float2 t0v = __ffma2_rn(make_float2(__uint_as_float(rr[ii]), __uint_as_float(rr[ii + 1])), r2, m2);
          float2 t1v = __ffma2_rn(make_float2(__uint_as_float(rr[ii + 2]), __uint_as_float(rr[ii + 3])), r2, m2);
          t0v = __ffma2_rn(make_float2(bi.x, bi.y), r2, t0v);
          t1v = __ffma2_rn(make_float2(bi.z, bi.w), r2, t1v);
          t0v = __ffma2_rn(t0v, make_float2(g.x, g.y), make_float2(b.x, b.y));
          t1v = __ffma2_rn(t1v, make_float2(g.z, g.w), make_float2(b.z, b.w));
          pk[c][i >> 1] = pack_relu<FP16>(t0v.x, t0v.y);
          pk[c][(i >> 1) + 1] = pack_relu<FP16>(t1v.x, t1v.y);
        }
      }
    }
    if (!row_ok) {  // rows past the sequence end are the next layer's zero padding (last tile of a sequence only)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[c][i] = 0u;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (storer) bulk_wait_read<1>();  // the store that read this buffer two chunks ago has drained it
      __syncwarp();
      const uint32_t boff = (stg_cnt & 1u) * 2048u;
      uint8_t* rowp = stg_gen + boff + (uint32_t)lane * 64u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(rowp + (((uint32_t)j ^ sw64) << 4)) =
            make_uint4(pk[c][4 * j], pk[c][4 * j + 1], pk[c][4 * j + 2], pk[c][4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      // the warp's staging tile is in MMA row order (lane = 8 (g - 4 quad) + i); the store's box walks (channel, i, g)
      // so that row lands on output row t0 + 16 i + g
      if (storer) {
        if (tile_ok) {
          if (p.store_mode == 0) tma_store_5d(&p.tma_o, stg_addr + boff, cbase + c * 32, 0, 4 * quad, t0 >> 7, lseq);
          else tma_store_4d(&p.tma_o, stg_addr + boff + (uint32_t)lane * 512u, cbase + c * 32, 4 * quad + lane, t0 >> 4, lseq);
        }
        bulk_commit();
      }
      ++stg_cnt;
    }
    F01_T(if (etm_on) { etm[2] += clock64() - te1; etm[3] += 1; })
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
  F01_T(if (etm_on) for (int k = 0; k < 4; ++k) p.dbg[264 + k] = etm[k];)
  if (storer) bulk_wait<0>();
  }
}

template <int FP16, int PCM16>
__global__ void __launch_bounds__(F_THREADS, 1) conv01_kernel(const __grid_constant__ F01Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + F_OFF_BAR;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + a); };
  const uint32_t xa_full_bar = bar_base + 8u * 12;  // leader: x' operand rows of both CTAs written (8 warps)
  const uint32_t d0_full_bar = bar_base + 8u * 13;  // both: the conv0 GEMM of the current tap has completed
  const uint32_t d0_free_bar = bar_base + 8u * 14;  // leader: both CTAs' producers have read D0 out of TMEM (8 warps)
  const uint32_t win_full_bar = bar_base + 8u * 15;  // local: the waveform window of tile n is complete (producers -> extra-row warp)
  const uint32_t win_free_bar = bar_base + 8u * 16;  // local: the extra-row warp has read the window of tile n
  const uint32_t tmem_slot = bar_base + 8u * 20;
  F01Vecs& ev = *reinterpret_cast<F01Vecs*>(smem_gen + F_OFF_VEC);
  float* xs = reinterpret_cast<float*>(smem_gen + F_OFF_XSB);
  // sample i of the window in buffer buf. fp32: word i + i / 320 (the skew spreads the lanes' 20-sample stride over the
  // banks). int16 PCM: the window is loaded as aligned sample pairs starting one sample early, element i + 1 at
  // half-word (i + 1) + 2 ((i + 1) / 320); the scaling by 2^-15 is exact.
  auto win = [&](int buf, int i) -> float {
    if (PCM16) {
      const int e = i + 1;
      return (float)reinterpret_cast<const int16_t*>(xs + buf * F_XS)[e + 2 * (e / 320)] * (1.0f / 32768.0f);
    }
    return xs[buf * F_XS + i + i / 320];
  };
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs, owns the cross-CTA barriers)

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
    mbar_init(xa_full_bar, 2 * F_PROD_WARPS);
    mbar_init(d0_full_bar, 1);  // multicast MMA commit
    mbar_init(d0_free_bar, 2 * F_PROD_WARPS);
    mbar_init(win_full_bar, 1);
    mbar_init(win_free_bar, 1);
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
      for (int l = 0; l < 12; ++l)  // upper triangle, off-diagonal entries doubled (G is symmetric)
        ev.G[k][l] = (l < 10 && l >= k) ? (l == k ? 1.f : 2.f) * p.cs.G[k][l < 10 ? l : 0] : 0.f;
      ev.h2[k] = p.cs.h2[k];
    }
    ev.h2[10] = ev.h2[11] = 0.f;
    ev.s = p.cs.s;
  }
  // conv0 B operand of this CTA's 128 channels: fp16 [k chunk of 8][channel][8], K = (u0..9, d, beta, 0 x4 | u0..9, d, 0 x5)
  for (int i = threadIdx.x; i < 128 * 32; i += F_THREADS) {
    const int n = i >> 5, k = i & 31, kk = k & 15;
    const float v = (kk < 11 || k == 11) ? __ldg(p.wg + kk * kDim + (int)rank * 128 + n) : 0.f;
    *reinterpret_cast<__half*>(smem_gen + F_OFF_U + (k >> 3) * 2048 + n * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers, operands and TMEM allocations exist before anyone signals across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int num_pair_tiles = p.nseq * p.pair_tiles_per_seq;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  auto tile_of = [&](int pt, int* lseq, int* t0) {
    *lseq = pt / p.pair_tiles_per_seq;
    *t0 = (pt % p.pair_tiles_per_seq) * 256 + (int)rank * 128;
  };
  // TMEM columns of the conv0 result of (tile iteration it, tap j): tap 0 sits in the tile's own conv1 accumulator
  // (it & 1) before that tile's main loop starts, taps 1-3 in the other one
  auto d0_acc = [&](int it, int j) { return j == 0 ? (it & 1) : ((it + 1) & 1); };
#ifdef F01_TIMERS
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
  auto stamp = [&](int it, int j, int e) {
    if (dbg_on && it < 4) p.dbg[(it * 4 + j) * 16 + e] = clock64();
  };
#else
  auto stamp = [&](int, int, int) {};
#endif

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
    // ===== MMA issuer (leader CTA only). Per tap: C C Z C C, where C = one ring stage of conv1 (taps j and j+4 from the
    // same A block one row group apart) and Z = the conv0 GEMM of the NEXT tap, issued while two stages are still queued
    if (lane == 0 && rank == 0) {
      const uint32_t idesc1 = make_idesc_16(256, 256, 0, 0, FP16);
      const uint32_t idesc0 = make_idesc_16(256, 256, 0, 0, 1);  // conv0: fp16 operands in either mode
      int stage = 0;
      uint32_t phase = 0, z = 0;  // z = conv0 GEMMs issued so far
      // true when the next conv0 GEMM could be issued without blocking (same conditions as issue_z waits for)
      auto z_ready = [&](int it, int j) {
        if (j == 1 && !mbar_test_cluster(tempty_bar(d0_acc(it, j)), (uint32_t)((((it + 1) >> 1) - 1) & 1))) return false;
        return mbar_test_cluster(xa_full_bar, z & 1u) && mbar_test_cluster(d0_free_bar, (z - 1u) & 1u);
      };
      auto issue_z = [&](int it, int j) {
        const int a = d0_acc(it, j);
        stamp(it, j, 9);
        // taps 1-3 go to the other accumulator: the epilogues of tile it-1 (both CTAs) must have drained it. It is the
        // k-th release of that barrier, k = (it+1)/2 (k = 0: nothing to wait for, the parity trick passes)
        if (j == 1) mbar_wait(tempty_bar(a), (uint32_t)((((it + 1) >> 1) - 1) & 1));
        mbar_wait_cluster(xa_full_bar, z & 1u);
        stamp(it, j, 10);
        mbar_wait_cluster(d0_free_bar, (z - 1u) & 1u);  // the previous conv0 result has left TMEM (z = 0 passes)
        stamp(it, j, 11);
        tc_fence_after();
        const uint32_t xa = smem_base + F_OFF_XA, ub = smem_base + F_OFF_U;
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
          umma_16_2sm(tmem_base + a * 256, make_smem_desc_nosw(xa + s2 * 4096, 2048, 128),
                      make_smem_desc_nosw(ub + s2 * 4096, 2048, 128), idesc0, s2 != 0);
        umma_commit_2sm(d0_full_bar);
        ++z;
      };
      auto issue_c = [&](uint32_t d_tmem, int ss, int it) {
        mbar_wait_cluster(full_bar(stage), phase);
        stamp(it, ss >> 2, 12 + (ss & 3));
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * F_STAGE_BYTES, w_addr = a_addr + F_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_16_2sm(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(w_addr + k * 32, 0, 1024),
                      idesc1, (ss | k) != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_16_2sm(d_tmem, make_smem_desc_sw128(a_addr + 1024 + k * 32, 0, 1024),
                      make_smem_desc_sw128(w_addr + F_WK_BYTES + k * 32, 0, 1024), idesc1, 1);
        umma_commit_2sm(empty_bar(stage));
        if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
      };
      int it = 0;
      if (cluster_id < num_pair_tiles) issue_z(0, 0);
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));  // epilogues of tile it-2 have drained it
        mbar_wait_cluster(d0_free_bar, 0u);  // conv0 result 4 it (tap 0 of this tile, in this accumulator) has been read out
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const bool more = pt + n_clusters < num_pair_tiles;
        for (int j = 0; j < 4; ++j) {
          // C C Z C C, but a conv0 GEMM that is not ready yet (tap 1 waits for the previous tile's epilogue) lets the
          // remaining stages of this tap go first instead of idling the tensor pipe behind a blocked issuer
          const bool has_z = j < 3 || more;
          const int zi = j < 3 ? it : it + 1, zj = j < 3 ? j + 1 : 0;
          issue_c(d_tmem, 4 * j, it);
          issue_c(d_tmem, 4 * j + 1, it);
          int c_next = 2;
          while (has_z && c_next < 4 && !z_ready(zi, zj)) issue_c(d_tmem, 4 * j + c_next++, it);
          if (has_z) issue_z(zi, zj);
          while (c_next < 4) issue_c(d_tmem, 4 * j + c_next++, it);
        }
        umma_commit_2sm(tfull_bar(acc));
      }
    }
  } else if (warp == 3) {
    // ===== the extra row t0 + 128 (group 16, row 7 of every A block): CUDA cores, lane = channels 8 lane .. 8 lane + 7,
    // fp32 weights from global memory (L1-resident)
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++it) {
      int lseq, t0;
      tile_of(pt, &lseq, &t0);
      // The tile's window xs[it & 1]: wait for the producers, take the 25 samples of the four taps at once and hand the
      // buffer back (mbarriers, not a block barrier: this warp runs up to three ring stages behind the producers, which
      // must not wait for it - they would be waiting for stages that only they can fill).
      mbar_wait(win_full_bar, (uint32_t)(it & 1));
      float xall[25];
#pragma unroll
      for (int k = 0; k < 25; ++k) xall[k] = win(it & 1, 2560 + k);
      __syncwarp();
      if (lane == 0) mbar_arrive(win_free_bar);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = 4 * (t0 + 128) - 2 + j;
        const bool valid = f >= 0 && f < p.L0;
        float xv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) xv[k] = xall[5 * j + k];
        const float rstd = valid ? frame_rstd(ev, xv) : 0.f;
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
          if (lane == 0) mbar_arrive_remote(mapa(full_bar(stage), 0));
          if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= F_PROD_WARP0) {
    // ===== conv0 producers (warps 12-15): lane = MMA row r = 8g + i  <->  t = t0 + 16 i + g
    const int pw = warp - F_PROD_WARP0;  // = TMEM lane quadrant of the warp (12 % 4 == 0)
    const int r = pw * 32 + lane;
    const int dt = 16 * (r & 7) + (r >> 3);  // t - t0 of the row
    auto bar_prod = [&]() { asm volatile("bar.sync 5, %0;" ::"n"(F_PROD_WARPS * 32) : "memory"); };
    auto prefetch = [&](int pt, int buf) {  // the tile's waveform window -> xs[buf] (zero outside the signal)
      int lseq, t0;
      tile_of(pt, &lseq, &t0);
      const int seq = p.seq0 + lseq;  // channel-major sequence id: c * batch + item
      const long long row = ((long long)(seq % p.batch) * 2 + seq / p.batch) * p.n_samples;
      const long long sbase = 20LL * t0 - 13;
      const uint32_t dst = smem_base + F_OFF_XSB + (uint32_t)buf * (F_XS * 4);
      if (PCM16) {
        // aligned pairs (s, s + 1), s even, from sbase - 1; signal start and end are even, so a pair is in or out as a whole
        const int16_t* x = static_cast<const int16_t*>(p.wav) + row;
        for (int q = r; q < (F_WIN + 2) / 2; q += F_PROD_WARPS * 32) {
          const long long s = sbase - 1 + 2 * q;
          const bool ok = s >= 0 && s < p.n_samples;
          cp_async4(dst + 2u * (uint32_t)(2 * q + 2 * ((2 * q) / 320)), reinterpret_cast<const float*>(ok ? x + s : x),
                    ok ? 4u : 0u);
        }
      } else {
        const float* x = static_cast<const float*>(p.wav) + row;
        for (int i = r; i < F_WIN; i += F_PROD_WARPS * 32) {
          const long long s = sbase + i;
          const bool ok = s >= 0 && s < p.n_samples;
          cp_async4(dst + 4u * (uint32_t)(i + i / 320), ok ? x + s : x, ok ? 4u : 0u);
        }
      }
      cp_async_commit();
    };
    const uint32_t xa_full_leader = mapa(xa_full_bar, 0), d0_free_leader = mapa(d0_free_bar, 0);
    // diagnostics: cycles this warp spent per phase, summed over the kernel (registers; written once at the end)
#ifdef F01_TIMERS
    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc0 = 0;
    const bool tm_on = p.dbg != nullptr && blockIdx.x == 0 && pw == 0;
    auto tick = [&]() { if (tm_on) tc0 = clock64(); };
    auto tock = [&](int k) { if (tm_on) { const long long t1 = clock64(); tm[k] += t1 - tc0; tc0 = t1; } };
#else
    auto tick = [&]() {};
    auto tock = [&](int) {};
#endif
    // X: operand row of the conv0 GEMM of tap j of the tile starting at t0 (window buffer buf): this thread's frame
    // scaled by its 1 / sqrt(var + eps), as fp16 hi parts (10 taps, rstd, validity flag) in k 0..15 and lo parts in
    // k 16..31. Zero padding frames of conv1 (f < 0, f >= L0) give an all-zero row. The caller has seen the previous
    // conv0 GEMM complete (d0_full), so the single operand buffer is free.
    auto make_x = [&](int t0, int buf, int j) {
      const int f = 4 * (t0 + dt) - 2 + j;
      const bool valid = f >= 0 && f < p.L0;
      const int i0 = 20 * dt + 5 * j;
      float xv[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) xv[k] = win(buf, i0 + k);
      const float rstd = valid ? frame_rstd(ev, xv) : 0.f;
      uint32_t hi[6], lo[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float v0 = q < 5 ? xv[2 * q < 10 ? 2 * q : 0] * rstd : rstd;
        const float v1 = q < 5 ? xv[2 * q + 1 < 10 ? 2 * q + 1 : 0] * rstd : (valid ? 1.f : 0.f);
        const __half2 h = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v0 - hf.x, q < 5 ? v1 - hf.y : 0.f);
        hi[q] = *reinterpret_cast<const uint32_t*>(&h);
        lo[q] = *reinterpret_cast<const uint32_t*>(&l);
      }
      uint8_t* xrow = smem_gen + F_OFF_XA + r * 16;
      *reinterpret_cast<uint4*>(xrow) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(xrow + 2048) = make_uint4(hi[4], hi[5], 0u, 0u);
      *reinterpret_cast<uint4*>(xrow + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(xrow + 6144) = make_uint4(lo[4], lo[5], 0u, 0u);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(xa_full_leader);
    };
    int stage = 0, it = 0;
    uint32_t phase = 0, z = 0;
    if (cluster_id < num_pair_tiles) {
      int lseq, t0;
      tile_of(cluster_id, &lseq, &t0);
      prefetch(cluster_id, 0);
      cp_async_wait_all();
      bar_prod();
      if (r == 0) mbar_arrive(win_full_bar);
      if (cluster_id + n_clusters < num_pair_tiles) prefetch(cluster_id + n_clusters, 1);
      make_x(t0, 0, 0);
    }
    for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++it) {
      int lseq, t0;
      tile_of(pt, &lseq, &t0);
      for (int j = 0; j < 4; ++j, ++z) {
        tick();
        mbar_wait(d0_full_bar, z & 1u);  // the conv0 GEMM of this tap is in TMEM (and has released the operand buffer)
        tock(2);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(pw * 32) << 16) + d0_acc(it, j) * 256;
        const bool dup = r >= 1 && r < 8;  // rows 1..7 of group 0 are rows 0..6 of group 16
        const uint32_t sw = (uint32_t)(r & 7), dsw = (uint32_t)((r - 1) & 7);
        // 64 channels of this row: TMEM -> ReLU -> 16-bit, four 16-column loads (register budget)
        auto load_pack = [&](int cb, uint32_t (&pk)[32]) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t raw[16];
            tmem_ld16(taddr + cb * 64 + h * 16, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pk[h * 8 + i] = pack_relu<FP16>(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1]));
          }
        };
        auto store_blk = [&](const uint32_t (&pk)[32]) {
          tock(3);
          mbar_wait(empty_bar(stage), phase ^ 1);
          tock(4);
          uint8_t* ablk = smem_gen + stage * F_STAGE_BYTES;
          uint8_t* arow = ablk + (r >> 3) * 1024 + (r & 7) * 128;
          uint8_t* drow = ablk + 16 * 1024 + ((r - 1) & 7) * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            *reinterpret_cast<uint4*>(arow + (((uint32_t)c ^ sw) << 4)) = u;
            if (dup) *reinterpret_cast<uint4*>(drow + (((uint32_t)c ^ dsw) << 4)) = u;
          }
          tock(5);
          fence_proxy_async();
          tock(6);
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa(full_bar(stage), 0));
          if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
          tock(7);
        };
        // The last channel block is read FIRST and held in registers: D0 has then left TMEM as soon as block 2 is
        // loaded (after block 1's stage came free), one stage time before the ring could take block 3 - the next
        // conv0 GEMM and, for tap 0, the tile's conv1 main loop are waiting for exactly that.
        uint32_t held[32];
        load_pack(3, held);
#pragma unroll 1
        for (int cb = 0; cb < 3; ++cb) {
          uint32_t pk[32];
          load_pack(cb, pk);
          if (cb == 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(d0_free_leader);
          }
          store_blk(pk);
          if (cb == 0) {
            // the NEXT tap's operand, after this tap's first block is on its way (the stages this tap still owes are
            // not due before one, two and three stage times from now; its own GEMM is issued two stages from now)
            tock(0);
            if (j < 3) {
              make_x(t0, it & 1, j + 1);
            } else if (pt + n_clusters < num_pair_tiles) {
              int lseq1, t1;
              tile_of(pt + n_clusters, &lseq1, &t1);
              // Window hand-over. One thread waits for the extra-row warp to have left this tile's window (it read it
              // at its tile start) BEFORE the block barrier, and only after the barrier is the next window published:
              // the extra-row warp can then never complete a second phase of win_free while a producer still waits
              // for the first (a parity wait cannot tell phase n from phase n + 2; with every producer waiting on its
              // own after the publication, a thread delayed by ~1000 cycles - an instruction-cache miss under memory
              // pressure from concurrent kernels - deadlocked the kernel: VAPB_PIPE=4 hung within 40 calls).
              const bool more2 = pt + 2 * n_clusters < num_pair_tiles;
              cp_async_wait_all();
              if (r == 0 && more2) mbar_wait(win_free_bar, (uint32_t)(it & 1));
              bar_prod();  // the next window is complete; every producer and the extra-row warp have left this tile's
              if (r == 0) mbar_arrive(win_full_bar);
              if (more2) prefetch(pt + 2 * n_clusters, it & 1);
              make_x(t1, (it + 1) & 1, 0);
            }
            tock(1);
          }
        }
        store_blk(held);
      }
    }
    F01_T(if (tm_on && lane == 0) for (int k = 0; k < 8; ++k) p.dbg[256 + k] = tm[k];)
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
// dev_tab: folded conv0 parameters (conv0_v2_fold) [12][256] = u (10 taps) | d | beta in device memory (host_tab, the
// same table on the host, is not used any more: the first version passed it as kernel parameters). w1: conv1 weight [256][8 * 256] (16-bit, tap-major K), bias1 /
// g1 / b1: conv1 bias and ChannelNorm affine (device). out: row t of sequence s at out + s * out_seq_stride +
// (out_pad_rows + t) * 256; the kernel writes whole 128-row tiles (zeros past L1), so every sequence needs
// out_pad_rows + roundup(L1, 128) rows.
int launch_conv01(cudaStream_t st, const void* wav, int wav_pcm16, int batch, long long n_samples, int seq0, int nseq, long long L0,
                  long long L1, const float* host_tab /*[12][256]*/, const float* dev_tab, const Conv0Stats& cs,
                  const void* w1, const float* bias1, const float* g1, const float* b1, void* out,
                  long long out_seq_stride, int out_pad_rows, int n_sm, std::string* err, long long* dbg) {
  F01Params p{};
  p.dbg = dbg;
  if (wav_pcm16 && ((n_samples & 1) || (reinterpret_cast<uintptr_t>(wav) & 3))) {
    if (err) *err = "conv01: int16 PCM input needs an even n_samples and a 4-byte aligned buffer";
    return -1;
  }
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
    const uint32_t box[5] = {32, 8, 4, 1, 1};  // one epilogue warp's 32 rows: 4 row groups g, 8 rows i each
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
    if (cudaFuncSetAttribute(conv01_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(conv01_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(conv01_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(conv01_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess) {
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
  void (*kern)(F01Params) = wav_pcm16 ? (g_fp16 ? conv01_kernel<1, 1> : conv01_kernel<0, 1>)
                                       : (g_fp16 ? conv01_kernel<1, 0> : conv01_kernel<0, 0>);
  const cudaError_t ce = cudaLaunchKernelEx(&cfg, kern, p);
  if (ce != cudaSuccess) {
    if (err) *err = std::string("conv01 launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 1;
}

}  // namespace vapb
