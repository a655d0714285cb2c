// vap_head Linear(256, 256) fused with everything VapGPT.probs derives from its logits (north_star item (3)):
//   logits = x W^T + b                                   vap/model.py:261
//   probs = softmax(logits), H = -sum p log2 p            vap/model.py:189,200-202
//   p_now / p_future = codebook marginals                 vap/objective.py:184-204 (class index = 8 bits, LSB first;
//                                                         bits 0-3 speaker 0 bins 0-3, bits 4-7 speaker 1), model.py:205-210
//   arg-max class, logsumexp (for the loss kernel), the bulk driver's counters
// One persistent CTA per SM: the 128 KB weight stays in shared memory, A tiles (128 rows x 256, 16-bit) stream through a
// TMA ring, tcgen05.mma M128 N256 K16 into two TMEM accumulators, and eight epilogue warps (thread = accumulator row x
// column half) turn the 256 fp32 logits of a row into its outputs without the logits ever visiting HBM - they are
// written only when the caller asks for them (forward(), or probs() with the loss). The marginals use the bit
// structure of the class index instead of a 256 x 4 weight table: a 57-add butterfly per 32 exponentials yields the sum
// of the terms whose index has bit b set (b = 0..4), the chunk / half totals give bits 5..7; p_now / p_future are sums of
// those eight "bit sums" over the requested bins. Entropy is closed form: H = log2 S - (log2 e / S) sum e_i (x_i - m).
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int H_BM = 128, H_N = 256, H_K = 256, H_BK = 64, H_STAGES = 4;
constexpr int H_W_BYTES = H_N * H_K * 2;          // 128 KB: 4 k-blocks of (256 rows x 64 k), SW128
constexpr int H_A_BYTES = H_BM * H_BK * 2;        // 16 KB
constexpr int H_OFF_A = H_W_BYTES;
constexpr int H_OFF_BAR = H_OFF_A + H_STAGES * H_A_BYTES;
constexpr int H_OFF_VEC = H_OFF_BAR + 256;
constexpr int H_THREADS = 384;
constexpr float kL2E = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct HeadVecs {
  float bias[256];
  float part[2][128][12];   // per-row exchange between the two column halves
  unsigned int cnt[258];    // arg-max class histogram + active frames per channel (flushed once per CTA)
};
constexpr int H_SMEM = H_OFF_VEC + (int)sizeof(HeadVecs) + 1024;
static_assert(H_SMEM <= 232448, "shared memory budget");

struct alignas(64) HeadParams {
  CUtensorMap tma_a;   // (256, rows) 16-bit, box (64, 128), SW128
  CUtensorMap tma_w;   // (256 k, 256 n) 16-bit, box (64, 256), SW128
  const float* bias;
  long long rows;
  int now_lo, now_hi, fut_lo, fut_hi;
  float *logits, *probs, *p_now, *p_future, *H, *lse;
  uint8_t* argmax;
  unsigned long long* counters;
  const float* vad_sig;  // (rows, 2), with counters
  int fp16;
};

__global__ void __launch_bounds__(H_THREADS, 1) head_probs_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + H_OFF_BAR;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + a); };
  const uint32_t w_bar = bar_base + 8u * 12;
  const uint32_t tmem_slot = bar_base + 8u * 14;
  HeadVecs& ev = *reinterpret_cast<HeadVecs*>(smem_gen + H_OFF_VEC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma_a);
    prefetch_tmap(&p.tma_w);
    for (int s = 0; s < H_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp >= 4) {
    const int e = threadIdx.x - 128;
    ev.bias[e] = p.bias ? p.bias[e] : 0.f;
    ev.cnt[e] = 0u;
    if (e < 2) ev.cnt[256 + e] = 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int num_tiles = (int)((p.rows + H_BM - 1) / H_BM);

  if (warp == 0) {
    if (lane == 0) {
      // the weight once (parameters: not written by the previous kernel, so ahead of the dependency wait), then A tiles
      mbar_arrive_expect_tx(w_bar, H_W_BYTES);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem_base + kb * (H_N * H_BK * 2), &p.tma_w, w_bar, kb * H_BK, 0);
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), H_A_BYTES);
          tma_load_2d(smem_base + H_OFF_A + stage * H_A_BYTES, &p.tma_a, full_bar(stage), kb * H_BK, tile * H_BM);
          if (++stage == H_STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_16(H_BM, H_N, 0, 0, p.fp16);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      mbar_wait(w_bar, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * H_N;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + H_OFF_A + stage * H_A_BYTES, w_addr = smem_base + kb * (H_N * H_BK * 2);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(w_addr + k * 32, 0, 1024),
                      idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));
          if (++stage == H_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = (row of the tile, column half)
    pdl_wait();  // vad_sig (counters) was written by an earlier kernel of the stream
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int cbase = half * 128;
    auto epi_bar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // bins of the two marginals as masks over the 4 bits of a speaker
    uint32_t mask_now = 0, mask_fut = 0;
    for (int b = p.now_lo; b <= p.now_hi; ++b) mask_now |= 1u << b;
    for (int b = p.fut_lo; b <= p.fut_hi; ++b) mask_fut |= 1u << b;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * H_BM + row_in_tile;
      const bool valid = row < p.rows;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * H_N + cbase;
      // ---- pass 1: row maximum, arg-max (first index wins), minimum (for the reference's 0 * -inf = NaN entropy)
      float mx = -INFINITY, mn = INFINITY;
      int best = 0;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = __uint_as_float(r[i]) + ev.bias[cbase + c * 32 + i];
          if (v[i] > mx) { mx = v[i]; best = c * 32 + i; }
          mn = fminf(mn, v[i]);
        }
        if (p.logits && valid) {
          float* lg = p.logits + row * kClasses + cbase + c * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(lg + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
      float* mine = ev.part[half][row_in_tile];
      const float* other = ev.part[half ^ 1][row_in_tile];
      mine[0] = mx;
      mine[1] = __int_as_float(cbase + best);
      mine[2] = mn;
      epi_bar();
      {
        const float omx = other[0];
        const int oidx = __float_as_int(other[1]);
        int idx = cbase + best;
        // ties go to the lower class index (torch.argmax): the other half wins a tie only if it is the lower half
        if (omx > mx || (omx == mx && oidx < idx)) { mx = omx; idx = oidx; }
        best = idx;
        mn = fminf(mn, other[2]);
      }
      epi_bar();
      // ---- pass 2: exponentials, their sum, sum e (x - m), and the eight bit sums
      const float m2 = mx * kL2E;
      float S = 0.f, E1 = 0.f, B[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = __uint_as_float(r[i]) + ev.bias[cbase + c * 32 + i];
          e[i] = ex2_fast(fmaf(x, kL2E, -m2));
          E1 = fmaf(e[i], x - mx, E1);
        }
        // butterfly: after level b, e[0 .. 32 >> (b+1)) hold pair sums and the odd partners have been added to B[b]
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          const int n = 32 >> (b + 1);
          float odd = 0.f;
#pragma unroll
          for (int j = 0; j < n; ++j) {
            odd += e[2 * j + 1];
            e[j] = e[2 * j] + e[2 * j + 1];
          }
          B[b] += odd;
        }
        S += e[0];
        if (c & 1) B[5] += e[0];
        if (c & 2) B[6] += e[0];
      }
      if (half) B[7] = S;
      mine[0] = S;
      mine[1] = E1;
#pragma unroll
      for (int b = 0; b < 8; ++b) mine[2 + b] = B[b];
      epi_bar();
      S += other[0];
      E1 += other[1];
#pragma unroll
      for (int b = 0; b < 8; ++b) B[b] += other[2 + b];
      epi_bar();
      const float inv = 1.0f / S;
      // ---- pass 3 (only when the probabilities themselves are wanted)
      if (p.probs) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (valid) {
            float* pr = p.probs + row * kClasses + cbase + c * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o;
              o.x = ex2_fast(fmaf(__uint_as_float(r[i]) + ev.bias[cbase + c * 32 + i], kL2E, -m2)) * inv;
              o.y = ex2_fast(fmaf(__uint_as_float(r[i + 1]) + ev.bias[cbase + c * 32 + i + 1], kL2E, -m2)) * inv;
              o.z = ex2_fast(fmaf(__uint_as_float(r[i + 2]) + ev.bias[cbase + c * 32 + i + 2], kL2E, -m2)) * inv;
              o.w = ex2_fast(fmaf(__uint_as_float(r[i + 3]) + ev.bias[cbase + c * 32 + i + 3], kL2E, -m2)) * inv;
              *reinterpret_cast<float4*>(pr + i) = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // ---- per-row outputs (one thread of the pair)
      if (half == 0 && valid) {
        if (p.H) {
          // the reference computes -sum p log2 p term by term: a probability that underflows to 0 makes it NaN
          const bool zero_p = ex2_fast(fmaf(mn, kL2E, -m2)) * inv == 0.f;
          p.H[row] = zero_p ? __int_as_float(0x7fc00000) : __log2f(S) - kL2E * inv * E1;
        }
        float n0 = 0.f, n1 = 0.f, f0 = 0.f, f1 = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (mask_now >> b & 1u) { n0 += B[b]; n1 += B[4 + b]; }
          if (mask_fut >> b & 1u) { f0 += B[b]; f1 += B[4 + b]; }
        }
        if (p.p_now) {
          const float a = n0 * inv, b = n1 * inv, d = (a + b) + 1e-5f;
          *reinterpret_cast<float2*>(p.p_now + row * 2) = make_float2(a / d, b / d);
        }
        if (p.p_future) {
          const float a = f0 * inv, b = f1 * inv, d = (a + b) + 1e-5f;
          *reinterpret_cast<float2*>(p.p_future + row * 2) = make_float2(a / d, b / d);
        }
        if (p.lse) p.lse[row] = mx + kLn2 * __log2f(S);
        if (p.argmax) p.argmax[row] = (uint8_t)best;
        if (p.counters) {
          atomicAdd(&ev.cnt[best], 1u);
          const float2 v = *reinterpret_cast<const float2*>(p.vad_sig + row * 2);
          if (v.x >= 0.5f) atomicAdd(&ev.cnt[256], 1u);
          if (v.y >= 0.5f) atomicAdd(&ev.cnt[257], 1u);
        }
      }
    }
    if (p.counters) {
      epi_bar();
      const int e = threadIdx.x - 128;
      if (ev.cnt[e]) atomicAdd(&p.counters[e], (unsigned long long)ev.cnt[e]);
      if (e < 2 && ev.cnt[256 + e]) atomicAdd(&p.counters[256 + e], (unsigned long long)ev.cnt[256 + e]);
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

// x: 16-bit (rows, 256) dense; w: 16-bit [256 classes][256]; bias fp32 [256]. Every output is optional (device fp32 unless
// noted): logits (rows,256), probs (rows,256), p_now / p_future (rows,2), H (rows), lse (rows), argmax (rows) uint8,
// counters unsigned long long [258] accumulated (needs vad_sig (rows,2)). Returns launches or -1.
int launch_head_probs(cudaStream_t st, const void* x, const void* w, const float* bias, long long rows, int now_lo,
                      int now_hi, int fut_lo, int fut_hi, float* logits, float* probs, float* p_now, float* p_future,
                      float* H, float* lse, uint8_t* argmax, unsigned long long* counters, const float* vad_sig, int n_sm,
                      std::string* err) {
  HeadParams p{};
  {
    const uint64_t dims[2] = {(uint64_t)H_K, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)H_K};
    const uint32_t box[2] = {H_BK, H_BM};
    if (!make_tmap_bf16(&p.tma_a, x, 2, dims, strides, box, err)) return -1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)H_K, (uint64_t)H_N};
    const uint64_t strides[1] = {(uint64_t)H_K};
    const uint32_t box[2] = {H_BK, H_N};
    if (!make_tmap_bf16(&p.tma_w, w, 2, dims, strides, box, err)) return -1;
  }
  p.bias = bias;
  p.rows = rows;
  p.now_lo = now_lo; p.now_hi = now_hi; p.fut_lo = fut_lo; p.fut_hi = fut_hi;
  p.logits = logits; p.probs = probs; p.p_now = p_now; p.p_future = p_future; p.H = H; p.lse = lse;
  p.argmax = argmax;
  p.counters = counters;
  p.vad_sig = vad_sig;
  p.fp16 = g_fp16;
  static bool configured_on[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(head_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM) != cudaSuccess) {
      if (err) *err = "head_probs: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const long long tiles = (rows + H_BM - 1) / H_BM;
  const int grid = tiles < n_sm ? (int)tiles : n_sm;
  launch_pdl(head_probs_kernel, grid, H_THREADS, H_SMEM, st, p);
  return 1;
}

}  // namespace vapb
