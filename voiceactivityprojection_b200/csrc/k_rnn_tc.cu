// BF16 tensor-core recurrence of CPC's autoregressive net gAR (LSTM or GRU,
// 256 -> 256, zero initial state; vap/encoder_components.py:140-159 -> nn.LSTM /
// nn.GRU with PyTorch gate order i,f,g,o / r,z,n).
//
// The recurrence is a chain of T dependent steps, so the design minimises the
// latency of one step and keeps every weight on chip for the whole sequence:
//
// * A thread-block cluster of 8 CTAs owns 32 sequences. CTA r owns hidden units
//   [32r, 32r+32) of all four gate blocks: its 128 x 512 bf16 slice of
//   [W_ih | W_hh] (128 KB) sits in shared memory for all T steps and is the A
//   operand (M = 128 gate rows) of tcgen05.mma; the B operand (N = 16 sequences,
//   K-major, unswizzled core matrices) is [x_t ; h_{t-1}], so the input
//   projection is fused into the step and no (T, 4*256) projection buffer exists.
// * The 32 sequences of a cluster form two independent groups of 16 that
//   ping-pong: while group 0's gate math runs on the CUDA cores / MUFU, group 1's
//   MMAs run on the tensor pipe. Accumulators (128 lanes x 16 fp32 columns, two
//   per group) live in TMEM.
// * Gate math: TMEM lane = gate row, so warp q of a group holds gate q of 32
//   units x 16 sequences and applies its non-linearity without divergence; the
//   four gates of a unit meet through a small shared-memory exchange; c_t (h_t
//   for the GRU) stays in fp32 registers for all T steps.
// * h_t (bf16) is written once to a 1 KB staging tile that is already in the
//   operand layout, then pushed by ONE thread as 8 bulk async copies
//   (cp.async.bulk shared::cta -> shared::cluster) into every CTA's h buffer of
//   the cluster, each completing transaction bytes on the destination's mbarrier,
//   and as one TMA store to the (seq, t, 256) output in HBM. x_t tiles arrive by
//   TMA (4-D map that writes the unswizzled core-matrix layout directly).
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int RC = 8;            // CTAs per cluster
constexpr int RNB = 16;          // sequences per group
constexpr int RGROUPS = 2;       // groups per CTA
constexpr int RXST = 2;          // x tile stages per group
constexpr int R_A_BYTES = 128 * 512 * 2;
constexpr int R_TILE = RNB * 256 * 2;  // 8 KB: [32 k-chunks][16 seq][16 B]
constexpr int R_SLICE = RNB * 32 * 2;  // 1 KB: this CTA's 32 units of one h tile
constexpr int R_EXS = 20;        // padded sequence stride (floats) of the gate exchange
constexpr int R_EX_BYTES = 4 * 32 * R_EXS * 4;
constexpr int R_THREADS = 320;   // 8 gate warps + TMA warp + MMA warp

constexpr int OFF_A = 0;
constexpr int OFF_X = OFF_A + R_A_BYTES;
constexpr int OFF_H = OFF_X + RGROUPS * RXST * R_TILE;
constexpr int OFF_EX = OFF_H + RGROUPS * 2 * R_TILE;
constexpr int OFF_STG = OFF_EX + RGROUPS * R_EX_BYTES;
constexpr int OFF_BAR = OFF_STG + RGROUPS * 2 * R_SLICE;
constexpr int R_SMEM = OFF_BAR + 256 + 1024 /*alignment slack*/;

struct alignas(64) RnnParams {
  CUtensorMap tma_w;    // [1024 rows][512] bf16, SW128, box (64, 128)
  CUtensorMap tma_x;    // (8, nseq, 32, T) bf16, no swizzle, box (8, 16, 32, 1)
  CUtensorMap tma_out;  // (8, nseq, 32, T) bf16, no swizzle, box (8, 16, 4, 1)
  const float* bias;    // [8][128]
  int nseq, T;
};

template <int KIND>
__global__ void __launch_bounds__(R_THREADS, 1) rnn_tc_kernel(const __grid_constant__ RnnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + OFF_BAR;
  // barriers
  auto xfull = [&](int g, int s) { return bar_base + 8u * (g * RXST + s); };
  auto xempty = [&](int g, int s) { return bar_base + 8u * (4 + g * RXST + s); };
  auto hfull = [&](int g, int b) { return bar_base + 8u * (8 + g * 2 + b); };
  auto accfull = [&](int g, int b) { return bar_base + 8u * (12 + g * 2 + b); };
  const uint32_t wfull = bar_base + 8u * 16;
  const uint32_t tmem_slot = bar_base + 8u * 17;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int seq0 = (int)cluster_id_x() * (RGROUPS * RNB);
  const int T = p.T;

  if (warp == 9 && lane == 0) {
    prefetch_tmap(&p.tma_w);
    prefetch_tmap(&p.tma_x);
    prefetch_tmap(&p.tma_out);
    for (int g = 0; g < RGROUPS; ++g) {
      for (int s = 0; s < RXST; ++s) {
        mbar_init(xfull(g, s), 1);
        mbar_init(xempty(g, s), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(hfull(g, b), 1);
        mbar_init(accfull(g, b), 1);
      }
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anyone sends to them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 8) {
    // ===== TMA producer: weights once, then x_t tiles for both groups
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, R_A_BYTES);
      for (int kb = 0; kb < 8; ++kb)
        tma_load_2d(smem_base + OFF_A + kb * 16384, &p.tma_w, wfull, kb * 64, (int)rank * 128);
      for (int t = 0; t < T; ++t) {
        const int s = t % RXST;
        const uint32_t ph = (uint32_t)(t / RXST) & 1u;
        for (int g = 0; g < RGROUPS; ++g) {
          mbar_wait(xempty(g, s), ph ^ 1u);
          mbar_arrive_expect_tx(xfull(g, s), R_TILE);
          tma_load_4d(smem_base + OFF_X + (g * RXST + s) * R_TILE, &p.tma_x, xfull(g, s), 0, seq0 + g * RNB, 0, t);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, RNB, 0, 0);
      const uint32_t a_base = smem_base + OFF_A;
      auto issue_half = [&](uint32_t acc, uint32_t b_tile, int kb0, bool fresh) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint64_t ad = make_smem_desc_sw128(a_base + (kb0 + (j >> 2)) * 16384 + (j & 3) * 32, 0, 1024);
          const uint64_t bd = make_smem_desc_nosw(b_tile + j * 512, 256, 128);
          umma_bf16(acc, ad, bd, idesc, (fresh && j == 0) ? 0u : 1u);
        }
      };
      auto x_part = [&](int g, int t) {
        const int s = t % RXST;
        mbar_wait(xfull(g, s), (uint32_t)(t / RXST) & 1u);
        tc_fence_after();
        issue_half(tmem_base + (g * 2 + (t & 1)) * RNB, smem_base + OFF_X + (g * RXST + s) * R_TILE, 0, true);
        umma_commit(xempty(g, s));
      };
      for (int g = 0; g < RGROUPS; ++g)
        for (int b = 0; b < 2; ++b)
          if (b + 1 < T) mbar_arrive_expect_tx(hfull(g, b), R_TILE);  // h_b will arrive
      mbar_wait(wfull, 0);
      tc_fence_after();
      for (int g = 0; g < RGROUPS; ++g) {
        x_part(g, 0);
        umma_commit(accfull(g, 0));
        if (T > 1) x_part(g, 1);
      }
      for (int t = 1; t < T; ++t) {
        const int hb = (t - 1) & 1;
        const uint32_t hph = (uint32_t)((t - 1) >> 1) & 1u;
        for (int g = 0; g < RGROUPS; ++g) {
          mbar_wait(hfull(g, hb), hph);
          tc_fence_after();
          if (t + 1 < T - 1) mbar_arrive_expect_tx(hfull(g, hb), R_TILE);  // re-arm for h_{t+1}
          issue_half(tmem_base + (g * 2 + (t & 1)) * RNB, smem_base + OFF_H + (g * 2 + hb) * R_TILE, 4, false);
          umma_commit(accfull(g, t & 1));
          if (t + 1 < T) x_part(g, t + 1);
        }
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
    const int cb = tid & 15, cq = tid >> 4;  // cell phase: sequence cb, units 4cq..4cq+3
    const uint32_t stg_off = (cq >> 1) * 256 + (cb >> 3) * 128 + (cb & 7) * 16 + (cq & 1) * 8;
    float state[4] = {0.f, 0.f, 0.f, 0.f};  // LSTM: c ; GRU: h
    uint32_t dst_h[RC], dst_bar[RC];
    if (tid == 0) {
#pragma unroll
      for (int d = 0; d < RC; ++d) {
        dst_h[d] = mapa(smem_base + OFF_H + g * 2 * R_TILE + rank * R_SLICE, d);
        dst_bar[d] = mapa(hfull(g, 0), d);
      }
    }
    for (int t = 0; t < T; ++t) {
      const int par = t & 1;
      mbar_wait(accfull(g, par), (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (g * 2 + par) * RNB, r);
      tmem_ld_wait();
      tc_fence_before();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = __uint_as_float(r[i]) + bias;
        if (KIND == 0)
          v[i] = (q == 2) ? tanh_fast(a) : sigmoid_fast(a);
        else
          v[i] = (q < 2) ? sigmoid_fast(a) : a;
      }
      float* row = ex + (q * 32 + lane) * R_EXS;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(row + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      // the TMA store that read staging[par] two steps ago must have drained it
      if (tid == 0) bulk_wait_read<1>();
      named_bar_sync<128>(1 + g);
      float h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = 4 * cq + j;
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
      uint2 pk;
      pk.x = pack_bf16(h[0], h[1]);
      pk.y = pack_bf16(h[2], h[3]);
      *reinterpret_cast<uint2*>(stg_gen + par * R_SLICE + stg_off) = pk;
      fence_proxy_async();
      named_bar_sync<128>(1 + g);
      if (tid == 0) {
        const uint32_t src = stg_base + par * R_SLICE;
        if (t + 1 < T) {
#pragma unroll
          for (int d = 0; d < RC; ++d)
            bulk_copy_to_cluster(dst_h[d] + par * R_TILE, src, R_SLICE, dst_bar[d] + par * 8);
        }
        tma_store_4d(&p.tma_out, src, 0, seq0 + g * RNB, (int)rank * 4, t);
        bulk_commit();
      }
    }
    if (tid == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();  // no CTA exits while a peer may still address its shared memory
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
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

int launch_rnn_tc(cudaStream_t st, int kind, const __nv_bfloat16* x, long long x_seq_stride, long long x_row_stride,
                  const __nv_bfloat16* w_cat, const float* bias, __nv_bfloat16* out, long long out_seq_stride,
                  int nseq, int T, std::string* err) {
  RnnParams p{};
  {
    const uint64_t dims[2] = {512, 1024};
    const uint64_t strides[1] = {512};
    const uint32_t box[2] = {64, 128};
    if (!make_tmap(&p.tma_w, w_cat, 2, 2, dims, strides, box, 128, err)) return -1;
  }
  {
    const uint64_t dims[4] = {8, (uint64_t)nseq, 32, (uint64_t)T};
    const uint64_t strides[3] = {(uint64_t)x_seq_stride, 8, (uint64_t)x_row_stride};
    const uint32_t box[4] = {8, RNB, 32, 1};
    if (!make_tmap(&p.tma_x, x, 2, 4, dims, strides, box, 0, err)) return -1;
  }
  {
    const uint64_t dims[4] = {8, (uint64_t)nseq, 32, (uint64_t)T};
    const uint64_t strides[3] = {(uint64_t)out_seq_stride, 8, (uint64_t)kDim};
    const uint32_t box[4] = {8, RNB, 4, 1};
    if (!make_tmap(&p.tma_out, out, 2, 4, dims, strides, box, 0, err)) return -1;
  }
  p.bias = bias;
  p.nseq = nseq;
  p.T = T;
  auto kern = kind == 0 ? rnn_tc_kernel<0> : rnn_tc_kernel<1>;
  static bool configured[2] = {false, false};
  if (!configured[kind]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM) != cudaSuccess) {
      if (err) *err = "rnn_tc: cannot reserve shared memory";
      return -1;
    }
    configured[kind] = true;
  }
  const int n_clusters = (nseq + RGROUPS * RNB - 1) / (RGROUPS * RNB);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_clusters * RC));
  cfg.blockDim = dim3(R_THREADS);
  cfg.dynamicSmemBytes = R_SMEM;
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

}  // namespace vapb
