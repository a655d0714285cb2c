// CTA-pair (cta_group::2) BF16 implicit GEMM for the K >= 1024 convolutions of the CPC gEncoder
// (vap/encoder_components.py:85-92,100-103: Conv1d + ChannelNorm + ReLU).
//
//   D[256 rows][256 ch] (+)= A[256][K] W[256][K]^T      one tcgen05.mma.cta_group::2 M256 N256 K16
// Two CTAs of a cluster (the two SMs of a TPC) share one MMA: each CTA holds ITS 128 rows of A and
// HALF of the W tile (128 of the 256 output channels) in shared memory and gets its 128 accumulator
// rows in its own TMEM. Per CTA a k-block is 16 KB of A + 16 KB of W instead of 16 + 32 KB, so the
// same shared memory holds a 5-deep operand ring plus double-buffered output staging (the 1-CTA kernel idled the tensor pipe 29 % of the
// time with 3 stages, 22 % with 4) and W is fetched from L2 once per CTA pair.
//
// Roles per CTA: warp 0 = TMA producer (both CTAs load; every transaction completes on the LEADER's
// full barrier), warp 1 = MMA issuer (leader CTA only), warp 2 = TMEM allocation (cta_group::2, both),
// warps 4-11 = epilogue on the CTA's own 128 rows (bias -> Channel/LayerNorm -> ReLU -> bf16 ->
// swizzled staging -> TMA store). MMA completion is multicast to both CTAs' empty / accumulator-full
// barriers; both epilogues release the accumulator on the leader's barrier.
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int G2_STAGES = 5;
constexpr int G2_A_BYTES = 128 * 64 * 2, G2_B_BYTES = 128 * 64 * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;  // 32 KB
constexpr int G2_OFF_STG = G2_STAGES * G2_STAGE_BYTES;   // [half][2] x 8 KB (128 rows x 64 B, SW64)
constexpr int G2_OFF_BAR = G2_OFF_STG + 4 * 8192;
constexpr int G2_OFF_VEC = G2_OFF_BAR + 256;
constexpr int G2_THREADS = 384;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even CTA of the pair

struct G2Vecs {
  float bias[256], g1[256], b1[256];
  float part[2][128][2];
};
constexpr int G2_SMEM = G2_OFF_VEC + (int)sizeof(G2Vecs) + 1024;

struct alignas(64) G2Params {
  CUtensorMap tma_a;   // (K, rows, nseq) bf16, box (64, 128, 1), SW128 (rows may overlap: implicit conv)
  CUtensorMap tma_b;   // (K, 256) bf16, box (64, 128), SW128
  CUtensorMap tma_o;   // (256, rows, nseq) bf16, box (32, 128, 1), SW64
  int nseq, rows_per_seq, pair_tiles_per_seq, num_k_blocks;
  const float* bias;
  int norm1;
  const float *g1, *b1;
  int act;
  int fp16;
};

__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_leader, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {  // one whole warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __launch_bounds__(G2_THREADS, 1) gemm_2sm_kernel(const __grid_constant__ G2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + G2_OFF_BAR;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (16 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (18 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 20;
  G2Vecs& ev = *reinterpret_cast<G2Vecs*>(smem_gen + G2_OFF_VEC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_b);
    prefetch_tmap(&p.tma_o);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // the leader's expect_tx arrival; bytes come from both CTAs
      mbar_init(empty_bar(s), 1);  // multicast MMA commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);     // multicast MMA commit
      mbar_init(tempty_bar(a), 512);  // the epilogue threads of both CTAs (leader's copy is the one used)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  if (warp >= 4) {
    const int e = threadIdx.x - 128;
    ev.bias[e] = p.bias ? p.bias[e] : 0.f;
    ev.g1[e] = p.norm1 != NORM_NONE ? p.g1[e] : 1.f;
    ev.b1[e] = p.norm1 != NORM_NONE ? p.b1[e] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM allocations exist before anyone signals across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int num_pair_tiles = p.nseq * p.pair_tiles_per_seq;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own 128 output channels of W
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
        const int seq = pt / p.pair_tiles_per_seq, t0 = (pt % p.pair_tiles_per_seq) * 256 + (int)rank * 128;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * G2_STAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * G2_STAGE_BYTES;
          const uint32_t bar_leader = full_bar(stage) & kPeerBitMask;
          tma_load_3d_2sm(a_dst, &p.tma_a, bar_leader, kb * 64, t0, seq);
          tma_load_2d_2sm(a_dst + G2_A_BYTES, &p.tma_b, bar_leader, kb * 64, (int)rank * 128);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc_16(256, 256, 0, 0, p.fp16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * G2_STAGE_BYTES, b_addr = a_addr + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2sm(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(b_addr + k * 32, 0, 1024),
                          idesc, (kb | k) != 0);
          umma_commit_2sm(empty_bar(stage));
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, own 128 rows): thread = (accumulator row, column half)
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const bool leader = (threadIdx.x - 128 - half * 128) == 0;
    const int cbase = half * 128;
    const uint32_t stg_addr = smem_base + G2_OFF_STG + half * 16384;
    uint8_t* stg_gen = smem_gen + G2_OFF_STG + half * 16384;
    const uint32_t sw64 = (uint32_t)((row_in_tile >> 1) & 3);
    uint32_t stg_cnt = 0;
    auto bar_half = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory"); };
    auto bar_epi = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const uint32_t tempty_leader0 = mapa(tempty_bar(0), 0), tempty_leader1 = mapa(tempty_bar(1), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pt = cluster_id; pt < num_pair_tiles; pt += n_clusters) {
      const int seq = pt / p.pair_tiles_per_seq, t0 = (pt % p.pair_tiles_per_seq) * 256 + (int)rank * 128;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256 + cbase;
      float mean1 = 0.f, rstd1 = 1.f;
      if (p.norm1 != NORM_NONE) {
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
      {
        uint32_t r[2][32];
        tmem_ld32(taddr, r[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
          if (c < 3) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bi = *reinterpret_cast<const float4*>(&ev.bias[cbase + c * 32 + i]);
            const float4 g = *reinterpret_cast<const float4*>(&ev.g1[cbase + c * 32 + i]);
            const float4 b = *reinterpret_cast<const float4*>(&ev.b1[cbase + c * 32 + i]);
            const float bb[4] = {bi.x, bi.y, bi.z, bi.w}, gg[4] = {g.x, g.y, g.z, g.w}, be[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x = __uint_as_float(r[c & 1][i + j]) + bb[j];
              if (p.norm1 != NORM_NONE) x = fmaf((x - mean1) * rstd1, gg[j], be[j]);
              v[i + j] = p.act == ACT_RELU ? fmaxf(x, 0.f) : x;
            }
          }
          if (leader) bulk_wait_read<1>();
          bar_half();
          const uint32_t boff = (stg_cnt & 1u) * 8192u;
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
          bar_half();
          if (leader) {
            tma_store_3d(&p.tma_o, stg_addr + boff, cbase + c * 32, t0, seq);
            bulk_commit();
          }
          ++stg_cnt;
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (leader) bulk_wait<0>();
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

// Operands as TcGemmArgs (common.cuh); supports the conv epilogue only: bias -> norm1 -> ReLU -> bf16 (out1_bf16).
int launch_gemm_2sm(cudaStream_t st, const TcGemmArgs& a, int n_sm, std::string* err) {
  const Epilogue& e = a.e;
  if (a.N != 256 || a.K % 64 || !a.out1_bf16 || a.out1_f32 || e.resid || e.accumulate || e.norm2 != NORM_NONE ||
      (e.act != ACT_RELU && e.act != ACT_NONE)) {
    if (err) *err = "gemm_2sm: unsupported problem (N = 256, K % 64 == 0, bias/norm1/ReLU -> bf16 only)";
    return -1;
  }
  G2Params p{};
  {
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.rows_per_seq, (uint64_t)a.nseq};
    const uint64_t strides[2] = {(uint64_t)a.a_map.row_stride,
                                 (uint64_t)(a.nseq > 1 ? a.a_map.seq_stride : a.a_map.row_stride * a.rows_per_seq)};
    const uint32_t box[3] = {64, 128, 1};
    if (!make_tmap_bf16(&p.tma_a, a.A, 3, dims, strides, box, err)) return -1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    const uint64_t strides[1] = {(uint64_t)a.K};
    const uint32_t box[2] = {64, 128};
    if (!make_tmap_bf16(&p.tma_b, a.W, 2, dims, strides, box, err)) return -1;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.N, (uint64_t)a.rows_per_seq, (uint64_t)a.nseq};
    const uint64_t strides[2] = {(uint64_t)e.out1_map.row_stride,
                                 (uint64_t)(a.nseq > 1 ? e.out1_map.seq_stride : e.out1_map.row_stride * a.rows_per_seq)};
    const uint32_t box[3] = {32, 128, 1};
    if (!make_tmap(&p.tma_o, a.out1_bf16, 2, 3, dims, strides, box, 64, err)) return -1;
  }
  p.nseq = a.nseq;
  p.rows_per_seq = a.rows_per_seq;
  p.pair_tiles_per_seq = (a.rows_per_seq + 255) / 256;
  p.num_k_blocks = a.K / 64;
  p.bias = e.bias;
  p.norm1 = e.norm1; p.g1 = e.g1; p.b1 = e.b1;
  p.act = e.act;
  p.fp16 = g_fp16;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM) != cudaSuccess) {
      if (err) *err = "gemm_2sm: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const int pair_tiles = p.nseq * p.pair_tiles_per_seq;
  int clusters = n_sm / 2;
  if (clusters > pair_tiles) clusters = pair_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = G2_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t ce = cudaLaunchKernelEx(&cfg, gemm_2sm_kernel, p);
  if (ce != cudaSuccess) {
    if (err) *err = std::string("gemm_2sm launch: ") + cudaGetErrorString(ce);
    return -1;
  }
  return 1;
}

}  // namespace vapb
