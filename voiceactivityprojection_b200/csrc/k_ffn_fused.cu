// Fused feed-forward block of a transformer layer on tcgen05:
//   x_out = x + W2 GELU(W1 z),  xs = bf16(x_out),  z_next = LayerNorm_next(x_out)
// (vap/modules.py:9-21 ffn_block without biases, :274 residual; z = ln_ffnetwork(x) comes from the
// attention out-projection's epilogue.) The 768-wide hidden activation never leaves the SM: the
// unfused pair of GEMMs wrote and re-read it through HBM (1.57 GB per layer at B = 256), which was
// half of their traffic.
//
// One persistent CTA per SM, 128 rows per tile:
//   Z   (A operand of GEMM 1, 4 k-blocks x 16 KB, TMA)            resident for the tile
//   for j in 0..5 (hidden columns 128 j .. 128 j + 127), software-pipelined two chunks deep:
//     ACC_H[j&1] = Z W1[j]^T          16 x tcgen05.mma M128 N128 K16, W1 k-blocks streamed through a ring
//     H[j&1]     = bf16(GELU(ACC_H))  epilogue warps: TMEM -> registers -> shared memory in the SW128 K-major
//                                     operand layout (what TMA would have written); runs while the tensor pipe
//                                     computes ACC_H of the next chunk
//     ACC_O     += H W2[:, j]^T       8 x tcgen05.mma M128 N256 K16, W2 k-blocks through the same ring
//   final epilogue on ACC_O: + residual (row-blocked fp32) -> x_out, bf16 shadow, LayerNorm -> z_next
// TMEM: ACC_H 2 x 128 columns, ACC_O 256 columns. The staging tiles of the final epilogue alias the H buffers.
//
// Two CTAs of a cluster work as a pair on 256 rows (tcgen05 cta_group::2, issued by the even CTA): each CTA holds ITS
// 128 rows of Z / H and its accumulators, and HALF of every weight tile (the B operand of a pair MMA is split by output
// channel over the two CTAs). Why: a tile needs all of W1 and W2 (786 KB from L2) and the ring that carries them is
// bounded by shared memory to 80 KB; with every CTA pulling whole weight tiles the ring covered ~0.6 of a chunk and the
// four later chunks of a tile ran at the ring's latency-bound rate (tools/ffn_probe.py, round 2: 4500 clk per chunk
// against 2400 clk of MMAs; neither halving the bytes per load nor multicasting the tiles changed that - only bytes in
// flight per CTA count). With half tiles the same ring holds 1.25 chunks.
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int FK_BYTES = 128 * 64 * 2;   // one A k-block: 128 rows x 64 k, 16 KB
constexpr int FW_BYTES = 128 * 64 * 2;   // one ring stage, 16 KB: this CTA's half of two W1 k-blocks (2 x 64 n x 64 k) or of one W2 k-block (128 n x 64 k)
constexpr int FW_STAGES = 5;
constexpr int F_OFF_Z = 0;
constexpr int F_OFF_H = 4 * FK_BYTES;                 // also: staging [half][2] x 8 KB, then LN vectors / partials
constexpr int F_OFF_W = F_OFF_H + 4 * FK_BYTES;
constexpr int F_OFF_VEC = F_OFF_W + FW_STAGES * FW_BYTES;   // FfnVecs (resident)
constexpr int F_OFF_BAR = F_OFF_VEC + 6144;
constexpr int F_SMEM = F_OFF_BAR + 256 + 1024 /*alignment slack*/;
constexpr int F_H_STG = 0;                            // the final epilogue's staging tiles alias the H region: [quarter][2] x 8 KB
constexpr int F_THREADS = 640;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-19 epilogue (4 warps per TMEM lane quadrant)
constexpr int F_EPI = 512;

struct alignas(64) FfnParams {
  CUtensorMap tma_z;    // (256, M) bf16, box (64, 128)
  CUtensorMap tma_w1;   // (256 k, 768 n) bf16, box (64, 64): this CTA's half of a 128-channel chunk
  CUtensorMap tma_w2;   // (768 k, 256 n) bf16, box (64, 128)
  CUtensorMap tma_xs;   // (256, M) bf16 out, box (32, 32), SW64: one store per epilogue warp
  CUtensorMap tma_zn;   // (256, M) bf16 out (LayerNorm of x_out), box (32, 32), SW64
  int M, num_tiles;
  const float* resid;   // row-blocked fp32
  float* x_out;         // row-blocked fp32
  int has_ln;
  const float *g2, *b2;
  // VAD head of the last layer (vap/model.py:258-259: Linear(256, 1) on each channel's output), taken from the rows of
  // x_out while they are in registers: saves a kernel that re-read the whole residual stream. Only without has_ln (the
  // weight vector lives where the LayerNorm vectors would).
  const float *vad_w, *vad_b;
  float *vad_logits, *vad_sig;  // (batch, T, 2), either may be null
  int vad_batch, vad_T;
  int fp16;
  long long* dbg;  // optional [32 tiles][16] SM-clock samples of CTA 0's first epilogue thread (tools/ffn_probe.py)
};

struct FfnVecs {
  float g2[256], b2[256];
  float part[4][128][2];
};

__device__ __forceinline__ float2 gelu_poly2_f(float2 x) {  // same polynomial as k_gemm_lin.cu
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

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_leader, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs when all prior MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {  // one whole warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ long long blocked_off_f(long long m, int c) {
  return (((m >> 7) * 64 + (c >> 2)) * 128 + (m & 127)) * 4 + (c & 3);
}

__global__ void __launch_bounds__(F_THREADS, 1) ffn_fused_kernel(const __grid_constant__ FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + F_OFF_BAR;
  const uint32_t z_full = bar_base, z_empty = bar_base + 8;
  auto w_full = [&](int s) { return bar_base + 8u * (2 + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (8 + s); };
  auto acch_full = [&](int b) { return bar_base + 8u * (14 + b); };  // GEMM 1 of a chunk has completed
  auto h_full = [&](int b) { return bar_base + 8u * (16 + b); };     // epilogue wrote H[b] (and no longer reads ACC_H[b]); 512
  auto h_empty = [&](int b) { return bar_base + 8u * (18 + b); };    // GEMM 2 of a chunk has completed (H[b] may be overwritten)
  const uint32_t acco_full = bar_base + 8 * 20;  // GEMM 2 of the last chunk has completed
  const uint32_t acco_empty = bar_base + 8 * 21; // final epilogue no longer reads ACC_O; 512 arrivals
  const uint32_t tmem_slot = bar_base + 8 * 22;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_z);
    prefetch_tmap(&p.tma_w1);
    prefetch_tmap(&p.tma_w2);
    // *_full of operands: the leader's copy is the one used (its expect_tx arrival, bytes from both CTAs);
    // *_empty / acc*_full: multicast MMA commits, every CTA waits on its own copy;
    // h_full / acco_empty: the epilogue threads of BOTH CTAs arrive on the leader's copy
    mbar_init(z_full, 1);
    mbar_init(z_empty, 1);
    for (int s = 0; s < FW_STAGES; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acch_full(b), 1);
      mbar_init(h_full(b), 2 * F_EPI);
      mbar_init(h_empty(b), 1);
    }
    mbar_init(acco_full, 1);
    mbar_init(acco_empty, 2 * F_EPI);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM allocations exist before anyone signals across the pair
  tc_fence_after();
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int num_pair_tiles = (p.num_tiles + 1) / 2;
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t t_acch = tmem_base /* 2 x 128 columns */, t_acco = tmem_base + 256;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own Z rows, then this CTA's half of the W stages in the order the MMA thread
    // consumes them: W1[0], W1[1], then for j = 0..5: W2[j], W1[j+2]. A chunk of W1 is two stages (two k-blocks of
    // 64 n x 64 k each), a chunk of W2 two stages (one k-block of 128 n x 64 k each).
    if (lane == 0) {
      uint32_t wc = 0, it = 0;
      auto load_w1 = [&](int j) {
        for (int i = 0; i < 2; ++i, ++wc) {
          const int s = wc % FW_STAGES;
          mbar_wait(w_empty(s), ((wc / FW_STAGES) & 1u) ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(w_full(s), 2 * FW_BYTES);
          const uint32_t bar_leader = w_full(s) & kPeerBitMask;
          for (int h = 0; h < 2; ++h)
            tma_load_2d_2sm(smem_base + F_OFF_W + s * FW_BYTES + h * (FW_BYTES / 2), &p.tma_w1, bar_leader,
                            (i * 2 + h) * 64, j * 128 + (int)rank * 64);
        }
      };
      auto load_w2 = [&](int j) {
        for (int kb = 0; kb < 2; ++kb, ++wc) {
          const int s = wc % FW_STAGES;
          mbar_wait(w_empty(s), ((wc / FW_STAGES) & 1u) ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(w_full(s), 2 * FW_BYTES);
          tma_load_2d_2sm(smem_base + F_OFF_W + s * FW_BYTES, &p.tma_w2, w_full(s) & kPeerBitMask, j * 128 + kb * 64,
                          (int)rank * 128);
        }
      };
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++it) {
        const int tile = pt * 2 + (int)rank;  // a pair's odd tile may not exist: TMA fills its rows with zeros
        mbar_wait(z_empty, (it & 1u) ^ 1u);
        if (rank == 0) mbar_arrive_expect_tx(z_full, 2 * 4 * FK_BYTES);
        for (int kb = 0; kb < 4; ++kb)
          tma_load_2d_2sm(smem_base + F_OFF_Z + kb * FK_BYTES, &p.tma_z, z_full & kPeerBitMask, kb * 64, tile * 128);
        load_w1(0);
        load_w1(1);
        for (int j = 0; j < 6; ++j) {
          load_w2(j);
          if (j + 2 < 6) load_w1(j + 2);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only): pair MMAs, M = 256 (128 rows in each CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc1 = make_idesc_16(256, 128, 0, 0, p.fp16), idesc2 = make_idesc_16(256, 256, 0, 0, p.fp16);
      uint32_t wc = 0, it = 0, hcnt[2] = {0, 0} /*h_full phases seen per buffer*/, n_o = 0;
      long long t_w = 0, t_h = 0;  // probe: cycles spent waiting for weights / for the epilogue
      const bool mdbg = p.dbg && blockIdx.x == 0;
      auto gemm1 = [&](int b) {  // ACC_H[b] = Z W1[chunk]^T
        for (int i = 0; i < 2; ++i, ++wc) {
          const int s = wc % FW_STAGES;
          const long long tw0 = mdbg ? clock64() : 0;
          mbar_wait(w_full(s), (wc / FW_STAGES) & 1u);
          if (mdbg) t_w += clock64() - tw0;
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t a_addr = smem_base + F_OFF_Z + (i * 2 + h) * FK_BYTES;
            const uint32_t b_addr = smem_base + F_OFF_W + s * FW_BYTES + h * (FW_BYTES / 2);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_2sm(t_acch + b * 128, make_smem_desc_sw128(a_addr + k * 32, 0, 1024),
                            make_smem_desc_sw128(b_addr + k * 32, 0, 1024), idesc1, (i | h | k) != 0);
          }
          umma_commit_2sm(w_empty(s));
        }
        umma_commit_2sm(acch_full(b));
      };
      auto gemm2 = [&](int b, bool first) {  // ACC_O (+)= H[b] W2[:, chunk]^T, one N = 256 MMA per K step
        for (int kb = 0; kb < 2; ++kb, ++wc) {
          const int s = wc % FW_STAGES;
          const long long tw0 = mdbg ? clock64() : 0;
          mbar_wait(w_full(s), (wc / FW_STAGES) & 1u);
          if (mdbg) t_w += clock64() - tw0;
          tc_fence_after();
          const uint32_t a_addr = smem_base + F_OFF_H + (b * 2 + kb) * FK_BYTES, b_addr = smem_base + F_OFF_W + s * FW_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2sm(t_acco, make_smem_desc_sw128(a_addr + k * 32, 0, 1024),
                          make_smem_desc_sw128(b_addr + k * 32, 0, 1024), idesc2, !(first && kb == 0 && k == 0));
          umma_commit_2sm(w_empty(s));
        }
        umma_commit_2sm(h_empty(b));
      };
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++it) {
        mbar_wait(z_full, it & 1u);
        tc_fence_after();
        // both ACC_H buffers are free: the epilogue of the previous tile's last two chunks arrived on h_full (waited)
        gemm1(0);
        gemm1(1);
        for (int j = 0; j < 6; ++j) {
          const int b = j & 1;
          const long long th0 = mdbg ? clock64() : 0;
          mbar_wait(h_full(b), hcnt[b] & 1u);  // H[b] is in shared memory and ACC_H[b] has been read, in both CTAs
          if (mdbg) t_h += clock64() - th0;
          ++hcnt[b];
          tc_fence_after();
          if (j == 0) {
            mbar_wait(acco_empty, (n_o & 1u) ^ 1u);  // previous tile's final epilogue has read ACC_O
            tc_fence_after();
          }
          gemm2(b, j == 0);
          if (j + 2 < 6) {
            gemm1(b);
            if (j + 2 == 5) umma_commit_2sm(z_empty);  // last read of Z for this tile
          }
          if (j == 5) {
            umma_commit_2sm(acco_full);
            ++n_o;
          }
        }
        if (mdbg && it < 32) {
          p.dbg[it * 16 + 11] = t_w;
          p.dbg[it * 16 + 12] = t_h;
          p.dbg[it * 16 + 13] = clock64();
          t_w = t_h = 0;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps, thread = (accumulator row, column quarter). Two warps per scheduler could not hide
    // the latencies of this epilogue (ncu: 0.27 IPC per scheduler, 24 us per tile against 7.6 us of MMAs).
    const int quad = warp & 3, qtr = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    uint8_t* h_gen = smem_gen + F_OFF_H;
    const uint32_t sw128 = (uint32_t)(row & 7), sw64 = (uint32_t)((row >> 1) & 3);
    FfnVecs& ev = *reinterpret_cast<FfnVecs*>(smem_gen + F_OFF_VEC);
    const uint32_t stg_addr = smem_base + F_OFF_H + F_H_STG + qtr * 16384;  // 2 x (128 rows x 64 B) per quarter
    uint8_t* stg_gen = h_gen + F_H_STG + qtr * 16384;
    uint32_t stg_cnt = 0, hcnt[2] = {0, 0} /*chunks done per ACC_H / H buffer*/, n_o = 0;
    if (p.has_ln && threadIdx.x - 128 < 256) {  // LayerNorm affine vectors, resident for the kernel
      ev.g2[threadIdx.x - 128] = p.g2[threadIdx.x - 128];
      ev.b2[threadIdx.x - 128] = p.b2[threadIdx.x - 128];
    }
    if (p.vad_w && threadIdx.x - 128 < 256) ev.g2[threadIdx.x - 128] = p.vad_w[threadIdx.x - 128];
    auto bar_epi = [&]() { asm volatile("bar.sync 1, 512;" ::: "memory"); };
    // every warp stages and stores its own 32 rows x 32 columns (2 KB, two buffers): no barrier couples the warps
    const uint32_t wstg = (uint32_t)quad * 4096u;
    auto stage_bf16 = [&](const CUtensorMap* map, const float (&v)[32], int c0, int c1) {
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      const uint32_t boff = wstg + (stg_cnt & 1u) * 2048u;
      uint8_t* rowp = stg_gen + boff + (uint32_t)lane * 64u;
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
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(map, stg_addr + boff, c0, c1 + quad * 32);
        bulk_commit();
      }
      ++stg_cnt;
    };

    const uint32_t h_full_leader[2] = {mapa(h_full(0), 0), mapa(h_full(1), 0)};
    const uint32_t acco_empty_leader = mapa(acco_empty, 0);
    int dbg_t = 0;
    for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters, ++dbg_t) {
      const int tile = pt * 2 + (int)rank;
      const bool dbg = p.dbg && blockIdx.x == 0 && threadIdx.x == 128 && dbg_t < 32;
      long long* dq = p.dbg + dbg_t * 16;
      if (dbg) dq[0] = clock64();
      const long long m = (long long)tile * 128 + row;
      const bool valid = m < p.M;
      const int cbase = qtr * 64;  // this thread's 64 output columns in the final epilogue
      // pull this tile's residual rows into L2 now: the final epilogue reads them ~10 us later and otherwise stalls on
      // DRAM latency (ncu: the residual adds were the top long-scoreboard stall)
      if (valid && (lane & 7) == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + blocked_off_f(m, cbase + c * 4)));
      }
      // ---- six hidden chunks of 128: ACC_H[b] -> GELU -> H[b] (bf16, SW128 K-major); this thread: 32 columns
      for (int j = 0; j < 6; ++j) {
        const int b = j & 1;
        mbar_wait(acch_full(b), hcnt[b] & 1u);
        if (dbg && j == 0) dq[1] = clock64();
        // H[b] is free: GEMM 2 of the chunk that used it two chunks ago has completed (first uses: nothing to wait for) ...
        mbar_wait(h_empty(b), (hcnt[b] & 1u) ^ 1u);
        ++hcnt[b];
        // ... and, for the first chunk of a tile, the TMA stores of the previous final epilogue have drained the region
        if (j == 0) {
          if (lane == 0) bulk_wait_read<0>();
          bar_epi();
        }
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(t_acch + lane_off + b * 128 + qtr * 32, r);
        tmem_ld_wait();
        // columns 32 qtr .. +31 of the chunk: k-block qtr / 2, 16-byte chunks 4 (qtr & 1) .. +3 of the row
        uint8_t* rowp = h_gen + (b * 2 + (qtr >> 1)) * FK_BYTES + (uint32_t)row * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            y[e] = gelu_poly2_f(make_float2(__uint_as_float(r[8 * q + 2 * e]), __uint_as_float(r[8 * q + 2 * e + 1])));
          uint4 u;
          u.x = pack16(y[0].x, y[0].y, p.fp16);
          u.y = pack16(y[1].x, y[1].y, p.fp16);
          u.z = pack16(y[2].x, y[2].y, p.fp16);
          u.w = pack16(y[3].x, y[3].y, p.fp16);
          *reinterpret_cast<uint4*>(rowp + ((((uint32_t)((qtr & 1) * 4 + q)) ^ sw128) << 4)) = u;
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive_remote(h_full_leader[b]);
        if (dbg) dq[2 + j] = clock64();
      }
      // ---- final epilogue on ACC_O: this thread's 64 columns in two chunks of 32
      mbar_wait(acco_full, n_o & 1u);
      if (dbg) dq[8] = clock64();
      ++n_o;
      tc_fence_after();
      // GEMM 2 of the last chunk is complete, so the H region is free: the staging tiles live there
      const uint32_t taddr = t_acco + lane_off + cbase;
      float s2 = 0.f, ss2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (valid) {
          const float* rp = p.resid + blocked_off_f(m, cbase + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(rp + i * 512);
            v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
          }
          float* op = p.x_out + blocked_off_f(m, cbase + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(op + i * 512) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        stage_bf16(&p.tma_xs, v, cbase + c * 32, tile * 128);
        if (p.has_ln) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s2 += v[i];
            ss2 = fmaf(v[i], v[i], ss2);
            r[i] = __float_as_uint(v[i]);
          }
          tmem_st32(taddr + c * 32, r);  // keep v for the LayerNorm pass
        } else if (p.vad_w) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(&ev.g2[cbase + c * 32 + i]);
            s2 = fmaf(v[i], wv.x, s2);
            s2 = fmaf(v[i + 1], wv.y, s2);
            s2 = fmaf(v[i + 2], wv.z, s2);
            s2 = fmaf(v[i + 3], wv.w, s2);
          }
        }
      }
      if (!p.has_ln && p.vad_w) {  // this row's VAD logit: the four column quarters meet in shared memory
        ev.part[qtr][row][0] = s2;
        bar_epi();
        if (qtr == 0 && valid) {
          const float sv = ((ev.part[0][row][0] + ev.part[1][row][0]) + (ev.part[2][row][0] + ev.part[3][row][0])) + p.vad_b[0];
          const long long seq = m / p.vad_T, t = m % p.vad_T;
          const long long ch = seq / p.vad_batch, b = seq % p.vad_batch;
          const long long o = (b * p.vad_T + t) * 2 + ch;
          if (p.vad_logits) p.vad_logits[o] = sv;
          if (p.vad_sig) p.vad_sig[o] = 1.0f / (1.0f + expf(-sv));
        }
      }
      if (dbg) dq[9] = clock64();
      if (p.has_ln) {
        tmem_st_wait();
        ev.part[qtr][row][0] = s2;
        ev.part[qtr][row][1] = ss2;
        bar_epi();
        s2 = (ev.part[0][row][0] + ev.part[1][row][0]) + (ev.part[2][row][0] + ev.part[3][row][0]);
        ss2 = (ev.part[0][row][1] + ev.part[1][row][1]) + (ev.part[2][row][1] + ev.part[3][row][1]);
        const float mean2 = s2 * (1.0f / kDim);
        const float var2 = fmaxf(ss2 - s2 * mean2, 0.f) * (1.0f / kDim);
        const float rstd2 = rsqrtf(var2 + kEps);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 g = *reinterpret_cast<const float4*>(&ev.g2[cbase + c * 32 + i]);
            const float4 b = *reinterpret_cast<const float4*>(&ev.b2[cbase + c * 32 + i]);
            v[i] = fmaf((__uint_as_float(r[i]) - mean2) * rstd2, g.x, b.x);
            v[i + 1] = fmaf((__uint_as_float(r[i + 1]) - mean2) * rstd2, g.y, b.y);
            v[i + 2] = fmaf((__uint_as_float(r[i + 2]) - mean2) * rstd2, g.z, b.z);
            v[i + 3] = fmaf((__uint_as_float(r[i + 3]) - mean2) * rstd2, g.w, b.w);
          }
          stage_bf16(&p.tma_zn, v, cbase + c * 32, tile * 128);
        }
      }
      tc_fence_before();
      mbar_arrive_remote(acco_empty_leader);
      if (dbg) dq[10] = clock64();
    }
    if (lane == 0) bulk_wait<0>();
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

// z: bf16 (M, 256) dense; w1: bf16 [768][256]; w2: bf16 [256][768]; resid / x_out: row-blocked fp32 (M padded to 128);
// xs: bf16 (M, 256) = x_out; zn: bf16 (M, 256) = LayerNorm(x_out; g2, b2) or null.
int launch_ffn_fused(cudaStream_t st, const __nv_bfloat16* z, const __nv_bfloat16* w1, const __nv_bfloat16* w2,
                     const float* resid, float* x_out, __nv_bfloat16* xs, __nv_bfloat16* zn, const float* g2,
                     const float* b2, int M, int n_sm, std::string* err, long long* dbg, const FfnVad* vad) {
  FfnParams p{};
  if (vad && !zn) {
    p.vad_w = vad->w;
    p.vad_b = vad->b;
    p.vad_logits = vad->logits;
    p.vad_sig = vad->sig;
    p.vad_batch = vad->batch;
    p.vad_T = vad->T;
  }
  auto map2 = [&](CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t b0, uint32_t b1, int sw) {
    const uint64_t dims[2] = {inner, outer};
    const uint64_t strides[1] = {inner};
    const uint32_t box[2] = {b0, b1};
    return make_tmap(m, base, 2, 2, dims, strides, box, sw, err);
  };
  if (!map2(&p.tma_z, z, 256, (uint64_t)M, 64, 128, 128)) return -1;
  if (!map2(&p.tma_w1, w1, 256, 768, 64, 64, 128)) return -1;
  if (!map2(&p.tma_w2, w2, 768, 256, 64, 128, 128)) return -1;
  if (!map2(&p.tma_xs, xs, 256, (uint64_t)M, 32, 32, 64)) return -1;
  if (zn && !map2(&p.tma_zn, zn, 256, (uint64_t)M, 32, 32, 64)) return -1;
  p.M = M;
  p.num_tiles = (M + 127) / 128;
  p.resid = resid;
  p.x_out = x_out;
  p.has_ln = zn != nullptr;
  p.g2 = g2;
  p.b2 = b2;
  p.dbg = dbg;
  p.fp16 = g_fp16;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM) != cudaSuccess) {
      if (err) *err = "ffn_fused: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  int pairs = (p.num_tiles + 1) / 2;
  if (pairs > n_sm / 2) pairs = n_sm / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(F_THREADS);
  cfg.dynamicSmemBytes = F_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // the kernel has griddepcontrol.wait before its first load
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const cudaError_t ce = cudaLaunchKernelEx(&cfg, ffn_fused_kernel, p);
  if (ce != cudaSuccess) {
    if (err) *err = std::string("ffn_fused launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 1;
}

}  // namespace vapb
