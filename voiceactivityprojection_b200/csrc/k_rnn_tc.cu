// BF16 tensor-core recurrence of CPC's autoregressive net gAR (LSTM or GRU,
// 256 -> 256, zero initial state; vap/encoder_components.py:140-159 -> nn.LSTM /
// nn.GRU with PyTorch gate order i,f,g,o / r,z,n).
//
// The recurrence is a chain of T dependent steps, so the design minimises the
// latency of one step and keeps every weight on chip for the whole sequence:
//
// * A thread-block cluster of 8 CTAs owns 16*NG sequences. CTA r owns hidden units
//   [32r, 32r+32) of all four gate blocks: its 128 x 512 bf16 slice of
//   [W_ih | W_hh] (128 KB) is written ONCE into tensor memory (256 TMEM columns)
//   and is the A operand (M = 128 gate rows, read from TMEM) of every
//   tcgen05.mma of the sequence; the B operand (N = 16 sequences, K-major,
//   unswizzled core matrices in shared memory) is [x_t ; h_{t-1}], so the input
//   projection is fused into the step and no (T, 4*256) projection buffer exists.
//   (With A in shared memory every step re-reads the 128 KB slice and the step
//   is shared-memory-bandwidth bound; measured 2x slower.)
// * The sequences of a cluster form NG (2..4) independent groups of 16 that
//   interleave: while one group's gate math runs on the CUDA cores / MUFU, the
//   other groups' MMAs run on the tensor pipe. Accumulators (128 lanes x 16 fp32
//   columns, two per group) live in TMEM next to the weights. At most 15 clusters
//   of 8 are co-resident on a B200 (cudaOccupancyMaxActiveClusters), so NG is
//   chosen to cover the batch in one wave.
// * Gate math: TMEM lane = gate row, so warp q of a group holds gate q of 32
//   units x 16 sequences and applies its non-linearity without divergence; the
//   four gates of a unit meet through a small shared-memory exchange; c_t (h_t
//   for the GRU) stays in fp32 registers for all T steps.
// * h_t (bf16) is written once to a 1 KB staging tile that is already in the
//   operand layout, then pushed by ONE thread as 8 bulk async copies
//   (cp.async.bulk shared::cta -> shared::cluster) into every CTA's h buffer of
//   the cluster, each completing transaction bytes on the destination's mbarrier,
//   and (from a row-major copy) as one TMA store to the (seq, t, 256) output in HBM.
//   x_t tiles arrive by TMA as 128-byte-swizzled K-major rows (a 16-byte-granular
//   box that wrote the unswizzled layout directly kept the TMA unit busy for most
//   of a step and delayed the latency-critical h copies behind it).
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int RC = 8;            // CTAs per cluster
constexpr int RXST = 2;          // x tiles in flight: the two time steps of a pair (one MMA series covers both)
constexpr int R_ACC_COL = 256;   // TMEM: weights in columns [0,256), accumulators behind them

// NG groups of NB sequences per cluster.
template <int NG, int NB>
struct RnnCfg {
  static constexpr int THREADS = NG * 128 + 64;    // 4 gate warps per group + TMA warp + MMA warp
  static constexpr int TILE = NB * 256 * 2;        // [32 k-chunks][NB seq][16 B]
  static constexpr int SLICE = NB * 32 * 2;        // this CTA's 32 units of one h tile (4 k-chunks)
  static constexpr int EXS = NB + 4;               // padded sequence stride (floats) of the gate exchange
  static constexpr int EX_BYTES = 4 * 32 * EXS * 4;
  static constexpr int OFF_X = 0;
  static constexpr int XTILE = NG * TILE;          // x_t of all groups: 4 k-blocks of [NG*NB seq][128 B], SW128
  static constexpr int OSTG = NB * 64;             // row-major copy of the slice for the HBM store: [NB seq][64 B]
  static constexpr int OFF_H = OFF_X + RXST * XTILE;
  static constexpr int OFF_EX = OFF_H + NG * 2 * TILE;
  static constexpr int OFF_STG = OFF_EX + NG * EX_BYTES;
  static constexpr int OFF_OSTG = OFF_STG + NG * 2 * SLICE;
  static constexpr int OFF_BAR = OFF_OSTG + NG * 2 * OSTG;
  static constexpr int SMEM = OFF_BAR + 512 + 1024 /*alignment slack*/;
};

struct alignas(64) RnnParams {
  const __nv_bfloat16* w_cat;  // [1024 rows][512] bf16
  CUtensorMap tma_x;    // (256 k, T, nseq) bf16, SW128, box (64, 1, NG*NB)
  CUtensorMap tma_out;  // (256 k, T, nseq) bf16, no swizzle, box (32, 1, NB)
  const float* bias;    // [8][128]
  int nseq, T;
  int fp16;
  long long* dbg;  // optional [steps][8] clock samples of cluster 0 / rank 0 / group 0 (diagnostics)
};

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int KIND, int NG, int NB>
__global__ void __launch_bounds__(RnnCfg<NG, NB>::THREADS, 1) rnn_tc_kernel(const __grid_constant__ RnnParams p) {
  using Cfg = RnnCfg<NG, NB>;
  constexpr int OFF_X = Cfg::OFF_X, OFF_H = Cfg::OFF_H, OFF_EX = Cfg::OFF_EX, OFF_STG = Cfg::OFF_STG;
  constexpr int R_TILE = Cfg::TILE, R_SLICE = Cfg::SLICE, R_EXS = Cfg::EXS, R_EX_BYTES = Cfg::EX_BYTES, RNB = NB;
  constexpr int TMA_WARP = NG * 4, MMA_WARP = NG * 4 + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::OFF_BAR;
  // barriers
  auto xfull = [&](int s) { return bar_base + 8u * s; };
  auto xempty = [&](int s) { return bar_base + 8u * (8 + s); };
  auto hfull = [&](int g, int b) { return bar_base + 8u * (16 + g * 2 + b); };
  auto accfull = [&](int g, int b) { return bar_base + 8u * (24 + g * 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int seq0 = (int)cluster_id_x() * (NG * RNB);
  const int T = p.T;

  if (warp == MMA_WARP && lane == 0) {
    prefetch_tmap(&p.tma_x);
    prefetch_tmap(&p.tma_out);
    for (int s = 0; s < RXST; ++s) {
      mbar_init(xfull(s), 1);
      mbar_init(xempty(s), 1);
    }
    for (int g = 0; g < NG; ++g) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(hfull(g, b), 1);
        mbar_init(accfull(g, b), 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == TMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp < 4) {
    // weights -> TMEM: thread (quadrant warp, lane) owns gate row m = 32*warp + lane; a 32-bit TMEM
    // column holds two consecutive k (low half = even k), which is the row's memory order
    const uint4* wrow = reinterpret_cast<const uint4*>(p.w_cat + ((size_t)rank * 128 + warp * 32 + lane) * 512);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 v = __ldg(wrow + c * 8 + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anyone sends to them
  tc_fence_after();

  if (warp == TMA_WARP) {
    // ===== TMA producer: x_t tiles (all groups of the cluster in one box)
    if (lane == 0) {
      // x of TWO time steps per stage: rows [0, NG*NB) of every k-block are step 2 pr, rows [NG*NB, 2 NG*NB) step 2 pr + 1
      for (int pr = 0; 2 * pr < T; ++pr) {
        const int nst = 2 * pr + 1 < T ? 2 : 1;
        mbar_wait(xempty(0), ((uint32_t)pr & 1u) ^ 1u);
        mbar_arrive_expect_tx(xfull(0), nst * Cfg::XTILE);
        for (int tl = 0; tl < nst; ++tl)
          for (int kb = 0; kb < 4; ++kb)
            tma_load_3d(smem_base + OFF_X + kb * (2 * NG * NB * 128) + tl * (NG * NB * 128), &p.tma_x, xfull(0), kb * 64,
                        2 * pr + tl, seq0);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer. One tcgen05.mma costs ~62 cycles for any N <= 128 (the 128 x 16 weight operand
    // streams at 64 B/clk) and that stream is what a step's time is made of (16 MMAs per series: input half +
    // one recurrent half per group = 3 x 991 of ~3600 clk at NG = 2). So the input half W_ih x_t is issued once
    // for all groups AND for two time steps (N = 2*NG*NB <= 128: x does not depend on the recurrence), into a
    // ring of four accumulator sets; only the recurrent half W_hh h_{t-1} is per group and per step.
    if (lane == 0) {
      const uint32_t idesc_h = make_idesc_16(128, NB, 0, 0, p.fp16);
      const uint32_t idesc_x = make_idesc_16(128, 2 * NG * NB, 0, 0, p.fp16);
      auto acc_col = [&](int t, int g) { return tmem_base + R_ACC_COL + (t & 3) * (NG * NB) + g * NB; };
      // input half of steps 2 pr and 2 pr + 1 into accumulator sets (2 pr) & 3 and the one after it. Both sets are free:
      // their last readers were the gates of steps 2 pr - 4 and 2 pr - 3. With an odd T the last pair's second half
      // multiplies stale x rows into a set nobody reads.
      auto x_pair = [&](int pr) {
        mbar_wait(xfull(0), (uint32_t)pr & 1u);
        tc_fence_after();
        const uint32_t b_tile = smem_base + OFF_X;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint64_t bd = make_smem_desc_sw128(b_tile + (j >> 2) * (2 * NG * NB * 128) + (j & 3) * 32, 0, 1024);
          umma_bf16_ts(acc_col(2 * pr, 0), tmem_base + j * 8, bd, idesc_x, j == 0 ? 0u : 1u);
        }
        umma_commit(xempty(0));
      };
      for (int g = 0; g < NG; ++g)
        for (int b = 0; b < 2; ++b)
          if (b + 1 < T) mbar_arrive_expect_tx(hfull(g, b), R_TILE);  // h_b will arrive
      x_pair(0);
      for (int g = 0; g < NG; ++g) umma_commit(accfull(g, 0));
      for (int t = 1; t < T; ++t) {
        const int hb = (t - 1) & 1;
        const uint32_t hph = (uint32_t)((t - 1) >> 1) & 1u;
        for (int g = 0; g < NG; ++g) {
          mbar_wait(hfull(g, hb), hph);
          tc_fence_after();
          if (p.dbg && g == 0 && blockIdx.x == 0 && t >= 64 && t < 96) p.dbg[(t - 64) * 8 + 0] = clock64();
          if (t + 1 < T - 1) mbar_arrive_expect_tx(hfull(g, hb), R_TILE);  // re-arm for h_{t+1}
          const uint32_t b_tile = smem_base + OFF_H + (g * 2 + hb) * R_TILE;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint64_t bd = make_smem_desc_nosw(b_tile + j * (2 * NB * 16), NB * 16, 128);
            umma_bf16_ts(acc_col(t, g), tmem_base + 128 + j * 8, bd, idesc_h, 1u);
          }
          umma_commit(accfull(g, t & 1));
          if (p.dbg && g == 0 && blockIdx.x == 0 && t >= 64 && t < 96) p.dbg[(t - 64) * 8 + 1] = clock64();
        }
        if (((t + 1) & 1) == 0 && t + 1 < T) x_pair((t + 1) >> 1);
      }
    }
  } else {
    // ===== gate warps: group g = warp / 4, gate row block q = warp % 4 (== TMEM lane quadrant)
    const int g = warp >> 2, q = warp & 3;
    const int tid = threadIdx.x & 127;
    const float bias = p.bias[rank * 128 + q * 32 + lane];
    float* ex = reinterpret_cast<float*>(smem_gen + OFF_EX + g * R_EX_BYTES);
    const uint32_t stg_base = smem_base + OFF_STG + g * 2 * R_SLICE;
    uint8_t* stg_gen = smem_gen + OFF_STG + g * 2 * R_SLICE;
    const uint32_t ostg_base = smem_base + Cfg::OFF_OSTG + g * 2 * Cfg::OSTG;
    uint8_t* ostg_gen = smem_gen + Cfg::OFF_OSTG + g * 2 * Cfg::OSTG;
    // cell phase: sequence cb, units UPT*cq .. +UPT-1 (a UPT*2-byte piece of one 16-byte operand chunk)
    constexpr int UPT = NB / 4;
    const int cb = tid % NB, cq = tid / NB;
    const uint32_t stg_off = ((cq * UPT) >> 3) * (NB * 16) + cb * 16 + ((cq * UPT) & 7) * 2;
    const uint32_t ostg_off = cb * 64 + cq * UPT * 2;
    float state[UPT];  // LSTM: c ; GRU: h
#pragma unroll
    for (int j = 0; j < UPT; ++j) state[j] = 0.f;
    // threads 0..7 of the group each push the slice to one CTA of the cluster; thread 8 stores it to HBM
    uint32_t dst_h = 0, dst_bar = 0;
    if (tid < RC) {
      dst_h = mapa(smem_base + OFF_H + g * 2 * R_TILE + rank * R_SLICE, tid);
      dst_bar = mapa(hfull(g, 0), tid);
    }
    for (int t = 0; t < T; ++t) {
      const int par = t & 1;
      mbar_wait(accfull(g, par), (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const bool dbg = p.dbg && tid == 0 && g == 0 && blockIdx.x == 0 && t >= 64 && t < 96;
      if (dbg) p.dbg[(t - 64) * 8 + 2] = clock64();
      const uint32_t acc_addr = tmem_base + ((uint32_t)(q * 32) << 16) + R_ACC_COL + (t & 3) * (NG * NB) + g * NB;
      float* row = ex + (q * 32 + lane) * R_EXS;
#pragma unroll
      for (int c = 0; c < NB / 16; ++c) {
        uint32_t r[16];
        tmem_ld16(acc_addr + c * 16, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = __uint_as_float(r[i]) + bias;
          if (KIND == 0)
            v[i] = (q == 2) ? tanh_fast(a) : sigmoid_fast(a);
          else
            v[i] = (q < 2) ? sigmoid_fast(a) : a;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(row + c * 16 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      tc_fence_before();
      // the TMA store that read staging[par] two steps ago must have drained it
      if (dbg) p.dbg[(t - 64) * 8 + 3] = clock64();
      if (tid == 8) bulk_wait_read<1>();
      named_bar_sync<128>(1 + g);
      if (dbg) p.dbg[(t - 64) * 8 + 4] = clock64();
      float h[UPT];
#pragma unroll
      for (int j = 0; j < UPT; ++j) {
        const int u = UPT * cq + j;
        const float g0 = ex[(0 * 32 + u) * R_EXS + cb], g1 = ex[(1 * 32 + u) * R_EXS + cb];
        const float g2 = ex[(2 * 32 + u) * R_EXS + cb], g3 = ex[(3 * 32 + u) * R_EXS + cb];
        if (KIND == 0) {
          state[j] = fmaf(g1, state[j], g0 * g2);
          h[j] = g3 * tanh_fast(state[j]);
        } else {
          const float n = tanh_fast(fmaf(g0, g3, g2));
          h[j] = fmaf(g1, state[j] - n, n);  // (1-z)*n + z*h
          state[j] = h[j];
        }
      }
      if (UPT == 4) {
        uint2 pk;
        pk.x = pack16(h[0], h[1], p.fp16);
        pk.y = pack16(h[2], h[3], p.fp16);
        *reinterpret_cast<uint2*>(stg_gen + par * R_SLICE + stg_off) = pk;
        *reinterpret_cast<uint2*>(ostg_gen + par * Cfg::OSTG + ostg_off) = pk;
      } else {
        uint4 pk;
        pk.x = pack16(h[0], h[1], p.fp16);
        pk.y = pack16(h[2], h[3], p.fp16);
        pk.z = pack16(h[4 % UPT], h[5 % UPT], p.fp16);
        pk.w = pack16(h[6 % UPT], h[7 % UPT], p.fp16);
        *reinterpret_cast<uint4*>(stg_gen + par * R_SLICE + stg_off) = pk;
        *reinterpret_cast<uint4*>(ostg_gen + par * Cfg::OSTG + ostg_off) = pk;
      }
      fence_proxy_async();
      named_bar_sync<128>(1 + g);
      if (dbg) p.dbg[(t - 64) * 8 + 5] = clock64();
      const uint32_t src = stg_base + par * R_SLICE;
      if (tid < RC) {
        if (t + 1 < T) bulk_copy_to_cluster(dst_h + par * R_TILE, src, R_SLICE, dst_bar + par * 8);
        if (dbg) p.dbg[(t - 64) * 8 + 6] = clock64();
      } else if (tid == 8) {
        tma_store_3d(&p.tma_out, ostg_base + par * Cfg::OSTG, (int)rank * 32, t, seq0 + g * RNB);
        bulk_commit();
      }
    }
    if (tid == 8) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();  // no CTA exits while a peer may still address its shared memory
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Packs one layer's weights for the kernel above. Output row R = r*128 + q*32 + u
// (cluster rank r, gate block q, unit u of the rank): LSTM q = i,f,g,o with
// [W_ih | W_hh]; GRU q = r, z, n_x ([W_in | 0]), n_h ([0 | W_hn]).
void rnn_tc_pack(int kind, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                 float* w_cat /*[1024][512]*/, float* bias /*[1024]*/) {
  for (int r = 0; r < 8; ++r)
    for (int q = 0; q < 4; ++q)
      for (int u = 0; u < 32; ++u) {
        const int R = r * 128 + q * 32 + u;
        const int unit = 32 * r + u;
        float* o = w_cat + (size_t)R * 512;
        int src;
        bool use_x = true, use_h = true;
        if (kind == 0) {
          src = q * kDim + unit;
          bias[R] = b_ih[src] + b_hh[src];
        } else {
          src = (q < 2 ? q : 2) * kDim + unit;
          if (q < 2) bias[R] = b_ih[src] + b_hh[src];
          else if (q == 2) { bias[R] = b_ih[src]; use_h = false; }
          else { bias[R] = b_hh[src]; use_x = false; }
        }
        for (int k = 0; k < kDim; ++k) {
          o[k] = use_x ? w_ih[(size_t)src * kDim + k] : 0.f;
          o[kDim + k] = use_h ? w_hh[(size_t)src * kDim + k] : 0.f;
        }
      }
}

template <int KIND, int NG, int NB>
static int launch_rnn_cfg(cudaStream_t st, RnnParams& p, const void* x, long long x_seq_stride, long long x_row_stride,
                          void* out, long long out_seq_stride, std::string* err) {
  using Cfg = RnnCfg<NG, NB>;
  {
    const uint64_t dims[3] = {256, (uint64_t)p.T, (uint64_t)p.nseq};
    const uint64_t strides[2] = {(uint64_t)x_row_stride, (uint64_t)x_seq_stride};
    const uint32_t box[3] = {64, 1, NG * NB};
    if (!make_tmap(&p.tma_x, x, 2, 3, dims, strides, box, 128, err)) return -1;
  }
  {
    const uint64_t dims[3] = {256, (uint64_t)p.T, (uint64_t)p.nseq};
    const uint64_t strides[2] = {(uint64_t)kDim, (uint64_t)out_seq_stride};
    const uint32_t box[3] = {32, 1, NB};
    if (!make_tmap(&p.tma_out, out, 2, 3, dims, strides, box, 0, err)) return -1;
  }
  auto kern = rnn_tc_kernel<KIND, NG, NB>;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
      if (err) *err = "rnn_tc: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const int n_clusters = (p.nseq + NG * NB - 1) / (NG * NB);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_clusters * RC));
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    if (err) *err = std::string("rnn_tc launch: ") + cudaGetErrorString(e);
    return -1;
  }
  return 1;
}

// `shape` (diagnostics): 0 = auto, else 100*NB + NG with (NB, NG) in {(16,2), (16,4), (32,2)}.
int launch_rnn_tc(cudaStream_t st, int kind, const __nv_bfloat16* x, long long x_seq_stride, long long x_row_stride,
                  const __nv_bfloat16* w_cat, const float* bias, __nv_bfloat16* out, long long out_seq_stride,
                  int nseq, int T, std::string* err, long long* dbg, int shape) {
  RnnParams p{};
  p.dbg = dbg;
  p.fp16 = g_fp16;
  p.w_cat = w_cat;
  p.bias = bias;
  p.nseq = nseq;
  p.T = T;
  // One MMA costs ~62 cycles whatever N <= 128 is (its 128 x 16 weight operand streams at 64 B/clk), so a
  // step's tensor time is per GROUP: two groups (enough to ping-pong) of as many sequences as the batch
  // needs to fit one wave of <= 15 co-resident clusters.
  if (shape == 0) shape = nseq <= 15 * 32 ? 1602 : 3202;
#define VAPB_RNN_CASE(K, NG_, NB_) \
  if (kind == K && shape == 100 * NB_ + NG_) \
    return launch_rnn_cfg<K, NG_, NB_>(st, p, x, x_seq_stride, x_row_stride, out, out_seq_stride, err);
  VAPB_RNN_CASE(0, 2, 16) VAPB_RNN_CASE(0, 4, 16) VAPB_RNN_CASE(0, 2, 32)
  VAPB_RNN_CASE(1, 2, 16) VAPB_RNN_CASE(1, 4, 16) VAPB_RNN_CASE(1, 2, 32)
#undef VAPB_RNN_CASE
  if (err) *err = "rnn_tc: unsupported kind/shape";
  return -1;
}

}  // namespace vapb
