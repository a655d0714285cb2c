// fp32-class fused causal ALiBi attention on the tensor cores, for the `fp32_tc` mode: every operand of the two
// contractions is an fp16 (hi, lo) pair and each contraction is three MMA series (hi*hi + hi*lo + lo*hi, fp32
// accumulation; the dropped lo*lo term is 2^-22 relative), the softmax is fp32.
// Reference: vap/modules.py:82-110 (scores, softmax, PV), :169-202 (bias 1 + m_h*j on allowed positions, -inf above
// the diagonal), :52 (scale 1/16). Same structure as k_attn_tc.cu, one pipeline per CTA:
//
//   split16_kernel          q | k | v (fp32, as the projection GEMM wrote them) -> hi and lo fp16 planes of the same layout
//   S[128 q][128 k]  = Qh Kh^T + Qh Kl^T + Ql Kh^T          12 x tcgen05.mma, operands K-major SW128 in smem (TMA)
//   softmax                  a thread owns half a query row; one pass over S against the running reference m (exact
//                            two-pass route on the first tile of an item and whenever some row's p sum to more than
//                            2^8); p -> (hi, lo) fp16 pairs in TMEM
//   O[128 q][64 d]  += Ph Vh + Ph Vl + Pl Vh                24 x tcgen05.mma, A from TMEM, B = V planes MN-major SW128
//   O / l -> fp32 out
// QK^T of the next tile is issued ahead of PV of the current one; the softmax warps wait for pv_done before they
// overwrite P or rescale O. TMEM: S [0,128), Ph [128,192), Pl [192,256), O [256,320).
#include <cuda_fp16.h>

#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int X_TILE = 128 * 64 * 2;  // one 128-row x 64-column fp16 tile
// warp 0: TMA producer; warp 1: MMA issuer; warp 2: TMEM allocation; warp 3 idle; warps 4-11: softmax
constexpr int X_THREADS = 384;
constexpr int X_OFF_Q = 0;                        // [hi, lo]
constexpr int X_OFF_K = 2 * X_TILE;               // [2 stages][hi, lo]
constexpr int X_OFF_V = X_OFF_K + 4 * X_TILE;     // [2 stages][hi, lo]
constexpr int X_OFF_X = X_OFF_V + 4 * X_TILE;     // exchange between the two half-row threads: float [half][128]
constexpr int X_OFF_BAR = X_OFF_X + 2 * 128 * 4;
constexpr int X_SMEM = X_OFF_BAR + 256 + 1024 /*alignment slack*/;
constexpr int X_COL_PH = 128, X_COL_PL = 192, X_COL_O = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(64) AttnX3Params {
  CUtensorMap tqh, tql, tkh, tkl, tvh, tvl;  // (256 head*d, T, nseq) fp16, SW128, box (64, 128, 1)
  float* out;                                // (nseq*T, 256) fp32
  const float* slopes;                       // [4]
  int nseq, T, nqt, n_items, cross, unit_major;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
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

// p0, p1 -> packed fp16 hi pair and packed fp16 pair of the remainders
__device__ __forceinline__ void split_pair(float p0, float p1, uint32_t* hi, uint32_t* lo) {
  const __half2 h = __floats2half2_rn(p0, p1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
  *hi = *reinterpret_cast<const uint32_t*>(&h);
  *lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct Item {
  int qi, seq, head;
};
// k-th work item of this CTA: all query tiles of one (sequence, head) back to back (longest first, so its K/V re-reads
// hit L2), the head rotating with the round; small batches spread single tiles
__device__ __forceinline__ bool next_item(const AttnX3Params& p, int k, Item* it) {
  const int pipe = blockIdx.x, n_pipes = gridDim.x;
  if (p.unit_major) {
    const int round = k / p.nqt;
    const int u = pipe + round * n_pipes;
    if (u >= p.nseq * 4) return false;
    it->qi = p.nqt - 1 - k % p.nqt;
    it->seq = u >> 2;
    it->head = (u + round) & 3;
    return true;
  }
  const int item = pipe + k * n_pipes;
  if (item >= p.n_items) return false;
  const int per_q = p.nseq * 4;
  it->qi = p.nqt - 1 - item / per_q;
  const int rem = item % per_q;
  it->seq = rem >> 2;
  it->head = rem & 3;
  return true;
}

__global__ void __launch_bounds__(256) split16_kernel(const float4* __restrict__ in, uint2* __restrict__ hi,
                                                      uint2* __restrict__ lo, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    uint2 h, l;
    split_pair(v.x, v.y, &h.x, &l.x);
    split_pair(v.z, v.w, &h.y, &l.y);
    hi[i] = h;
    lo[i] = l;
  }
}

__global__ void __launch_bounds__(X_THREADS, 1) attention_x3_kernel(const __grid_constant__ AttnX3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + X_OFF_BAR;
  const uint32_t q_full = bar_base, q_empty = bar_base + 8;
  auto kv_full = [&](int st) { return bar_base + 8u * (2 + st); };
  auto kv_empty = [&](int st) { return bar_base + 8u * (4 + st); };
  const uint32_t s_full = bar_base + 8 * 6, p_full = bar_base + 8 * 7, o_final = bar_base + 8 * 8, pv_done = bar_base + 8 * 9;
  const uint32_t tmem_slot = bar_base + 8 * 10;
  float* slope_s = reinterpret_cast<float*>(smem_gen + X_OFF_BAR + 8 * 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tqh);
    prefetch_tmap(&p.tql);
    prefetch_tmap(&p.tkh);
    prefetch_tmap(&p.tkl);
    prefetch_tmap(&p.tvh);
    prefetch_tmap(&p.tvl);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int st = 0; st < 2; ++st) {
      mbar_init(kv_full(st), 1);
      mbar_init(kv_empty(st), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(o_final, 1);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 3 && lane < 4) slope_s[lane] = p.slopes[lane];
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      uint32_t kvc = 0;
      auto load_q = [&](const Item& it, uint32_t n) {
        mbar_wait(q_empty, (n & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full, 2 * X_TILE);
        tma_load_3d(smem_base + X_OFF_Q, &p.tqh, q_full, it.head * 64, it.qi * 128, it.seq);
        tma_load_3d(smem_base + X_OFF_Q + X_TILE, &p.tql, q_full, it.head * 64, it.qi * 128, it.seq);
      };
      Item it, nx;
      bool have = next_item(p, 0, &it);
      if (have) load_q(it, 0);
      for (uint32_t k = 0; have; ++k) {
        const bool have_next = next_item(p, (int)k + 1, &nx);
        const int kvseq = p.cross ? (it.seq + p.nseq / 2) % p.nseq : it.seq;
        const int col = it.head * 64;
        for (int kt = it.qi; kt >= 0; --kt) {
          const int st = kvc & 1;
          mbar_wait(kv_empty(st), ((kvc >> 1) & 1u) ^ 1u);
          ++kvc;
          mbar_arrive_expect_tx(kv_full(st), 4 * X_TILE);
          tma_load_3d(smem_base + X_OFF_K + (st * 2 + 0) * X_TILE, &p.tkh, kv_full(st), col, kt * 128, kvseq);
          tma_load_3d(smem_base + X_OFF_K + (st * 2 + 1) * X_TILE, &p.tkl, kv_full(st), col, kt * 128, kvseq);
          tma_load_3d(smem_base + X_OFF_V + (st * 2 + 0) * X_TILE, &p.tvh, kv_full(st), col, kt * 128, kvseq);
          tma_load_3d(smem_base + X_OFF_V + (st * 2 + 1) * X_TILE, &p.tvl, kv_full(st), col, kt * 128, kvseq);
        }
        if (have_next) load_q(nx, k + 1);  // the Q buffer is released by the item's last QK
        it = nx;
        have = have_next;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_16(128, 128, 0, 0, 1);
      const uint32_t idesc_pv = make_idesc_16(128, 64, 0, 1, 1);  // B = V is MN-major (d contiguous)
      const uint32_t qa = smem_base + X_OFF_Q;
      uint32_t n_item = 0, kq = 0, kpv = 0, pc = 0;
      auto issue_qk = [&]() {
        const int st = kq & 1;
        mbar_wait(kv_full(st), (kq >> 1) & 1u);
        ++kq;
        tc_fence_after();
        const uint32_t ka = smem_base + X_OFF_K + st * 2 * X_TILE;
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // Qh Kh, Qh Kl, Ql Kh
          const uint32_t a = qa + (c == 2 ? X_TILE : 0), b = ka + (c == 1 ? X_TILE : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, make_smem_desc_sw128(a + k * 32, 0, 1024), make_smem_desc_sw128(b + k * 32, 0, 1024),
                      idesc_qk, (c | k) != 0);
        }
        umma_commit(s_full);
      };
      Item it;
      for (int k = 0; next_item(p, k, &it); ++k, ++n_item) {
        const int ntiles = it.qi + 1;
        mbar_wait(q_full, n_item & 1u);
        issue_qk();
        if (ntiles == 1) umma_commit(q_empty);
        for (int n = 0; n < ntiles; ++n) {
          mbar_wait(p_full, pc & 1u);
          ++pc;
          tc_fence_after();
          if (n + 1 < ntiles) {
            issue_qk();
            if (n + 2 == ntiles) umma_commit(q_empty);  // that was the item's last QK
          }
          const int st = kpv & 1;
          ++kpv;
          const uint32_t va = smem_base + X_OFF_V + st * 2 * X_TILE;
#pragma unroll
          for (int c = 0; c < 3; ++c) {  // Ph Vh, Ph Vl, Pl Vh
            const uint32_t a = tmem_base + (c == 2 ? X_COL_PL : X_COL_PH), b = va + (c == 1 ? X_TILE : 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16_ts(tmem_base + X_COL_O, a + kk * 8, make_smem_desc_sw128(b + kk * 2048, 1024, 1024), idesc_pv,
                          (n | c | kk) != 0);
          }
          umma_commit(kv_empty(st));
          umma_commit(pv_done);
          if (n + 1 == ntiles) umma_commit(o_final);
        }
      }
    }
  } else if (warp >= 4) {
    // ===== softmax: thread = (query row, half of the tile's keys)
    const int quad = warp & 3, ch = ((warp - 4) >> 2) & 1;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_s = t_row + ch * 64, t_o = t_row + X_COL_O + ch * 32;
    const uint32_t t_ph = t_row + X_COL_PH + ch * 32, t_pl = t_row + X_COL_PL + ch * 32;
    float* xs = reinterpret_cast<float*>(smem_gen + X_OFF_X);  // [half][128]
    float* x_own = xs + ch * 128 + row;
    const float* x_oth = xs + (ch ^ 1) * 128 + row;
    auto slot_bar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    auto slot_any = [&](bool v) {  // OR over the 256 softmax threads (the same named barrier)
      uint32_t r;
      asm volatile(
          "{\n\t.reg .pred pi, po;\n\tsetp.ne.b32 pi, %1, 0;\n\tbar.red.or.pred po, 1, 256, pi;\n\t"
          "selp.u32 %0, 1, 0, po;\n\t}"
          : "=r"(r)
          : "r"((uint32_t)v)
          : "memory");
      return r != 0;
    };
    auto wait_pv = [&](uint32_t tile) {  // PV of the previous tile has completed (P may be overwritten, O rescaled)
      if (tile > 0) mbar_wait(pv_done, (tile - 1) & 1u);
    };
    auto store_p = [&](int ci, const uint32_t (&ph)[16], const uint32_t (&pl)[16]) {
      tmem_st16(t_ph + ci * 16, ph);
      tmem_st16(t_pl + ci * 16, pl);
    };
    constexpr float SC = 0.0625f * kLog2e;
    uint32_t sc_cnt = 0, oc_cnt = 0;
    Item it;
    for (int k = 0; next_item(p, k, &it); ++k) {
      const int head = it.head;
      const float slope2 = slope_s[head] * kLog2e;
      float m = -INFINITY, l = 0.f;  // m: the row's reference (both half-row threads hold the same); l: partial sum
      for (int n = 0; n <= it.qi; ++n) {
        const int k0 = (it.qi - n) * 128;
        const bool diag = n == 0;
        // this thread's two 32-key chunks are global chunks 2*ch and 2*ch+1; on the diagonal tile chunk g is
        // visible to this warp's rows iff g <= quad, and chunk g == quad holds the diagonal itself
        const int g0 = 2 * ch;
        const int nvis = diag ? min(max(quad + 1 - g0, 0), 2) : 2;
        const float base = fmaf(slope2, (float)(k0 + 64 * ch), kLog2e);
        mbar_wait(s_full, sc_cnt & 1u);
        const uint32_t tile = sc_cnt++;
        tc_fence_after();
        float ps = 0.f;
        bool done = false;
        if (!diag) {
          // ---- one pass against the running reference (finite after the diagonal tile: a row always sees itself)
          const float base_m = base - m;
#pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            uint32_t r[32], ph[16], pl[16];
            tmem_ld32(t_s + ci * 32, r);
            tmem_ld_wait();
            const float cb = fmaf(slope2, (float)(ci * 32), base_m);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2_fast(fmaf(__uint_as_float(r[i]), SC, fmaf(slope2, (float)i, cb)));
              const float p1 = ex2_fast(fmaf(__uint_as_float(r[i + 1]), SC, fmaf(slope2, (float)(i + 1), cb)));
              ps += p0 + p1;
              split_pair(p0, p1, &ph[i >> 1], &pl[i >> 1]);
            }
            if (ci == 0) wait_pv(tile);
            store_p(ci, ph, pl);
          }
          // the reference still holds if no p exceeds 2^8: the p are non-negative, so the sum bounds each of them
          if (!slot_any(!(ps <= 256.0f))) {
            done = true;
          } else {
            ps = 0.f;
            tmem_st_wait();  // the exact route rewrites P: its stores must not overtake the ones just issued
          }
        }
        if (!done) {
          // ---- exact route. Pass A: the row maximum of t = SC * s_j + bias_j (log2 domain) over the visible keys (the exact
          // one: with a mere upper bound p can end up many binades below 1, where its fp16 lo part is lost)
          float mx = -INFINITY;
#pragma unroll 1
          for (int ci = 0; ci < nvis; ++ci) {
            uint32_t r[32];
            tmem_ld32(t_s + ci * 32, r);
            tmem_ld_wait();
            const float cb = fmaf(slope2, (float)(ci * 32), base);
            const bool dchunk = diag && g0 + ci == quad;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = fmaf(__uint_as_float(r[i]), SC, fmaf(slope2, (float)i, cb));
              if (!dchunk || i <= lane) mx = fmaxf(mx, t);
            }
          }
          *x_own = mx;
          slot_bar();
          mx = fmaxf(mx, *x_oth);
          const float m_new = fmaxf(m, mx);
          wait_pv(tile);
          tc_fence_after();
          if (n > 0 && __any_sync(0xffffffffu, m_new > m)) {
            const float alpha = ex2_fast(m - m_new);
            l *= alpha;
            uint32_t r[32];
            tmem_ld32(t_o, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(t_o, r);
          }
          m = m_new;
          // pass B: p = 2^(t - m), partial row sum, (hi, lo) -> TMEM
          const float base_m = base - m;
#pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            uint32_t ph[16], pl[16];
            if (ci < nvis) {
              uint32_t r[32];
              tmem_ld32(t_s + ci * 32, r);
              tmem_ld_wait();
              const float cb = fmaf(slope2, (float)(ci * 32), base_m);
              const bool dchunk = diag && g0 + ci == quad;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float p0 = ex2_fast(fmaf(__uint_as_float(r[i]), SC, fmaf(slope2, (float)i, cb)));
                float p1 = ex2_fast(fmaf(__uint_as_float(r[i + 1]), SC, fmaf(slope2, (float)(i + 1), cb)));
                if (dchunk && i > lane) p0 = 0.f;  // keys above the diagonal
                if (dchunk && i + 1 > lane) p1 = 0.f;
                ps += p0 + p1;
                split_pair(p0, p1, &ph[i >> 1], &pl[i >> 1]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) ph[i] = pl[i] = 0u;
            }
            store_p(ci, ph, pl);
          }
          // the exchange slots are rewritten by the next exact tile: every thread has read them (s_full of the next tile
          // follows all p_full arrivals)
        }
        l += ps;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // epilogue: O / l -> out[(seq*T + q), head*64 + 32*ch .. +32) fp32. o_final also says that every thread has arrived
      // on the last p_full, i.e. has read the last tile's exchange slots.
      mbar_wait(o_final, oc_cnt & 1u);
      ++oc_cnt;
      tc_fence_after();
      *x_own = l;
      slot_bar();
      l += *x_oth;
      const float inv = 1.0f / l;
      const int q = it.qi * 128 + row;
      float* dst = p.out + ((long long)it.seq * p.T + q) * kDim + head * 64 + ch * 32;
      {
        uint32_t r[32];
        tmem_ld32(t_o, r);
        tmem_ld_wait();
        tc_fence_before();  // the O reads are ordered before the next item's p_full arrivals
        slot_bar();         // ... and the row sums have been read before the next item's first exchange
        if (q < p.T) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(dst + 4 * i) =
                make_float4(__uint_as_float(r[4 * i]) * inv, __uint_as_float(r[4 * i + 1]) * inv,
                            __uint_as_float(r[4 * i + 2]) * inv, __uint_as_float(r[4 * i + 3]) * inv);
        }
      }
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

// q / k / v: fp32 rows at ptr + (seq*T + t)*row_stride (256 = 4 heads x 64 columns each); `planes`: scratch of
// 2 * 2 bytes per element of q, k and v for the fp16 hi / lo copies; out: dense (nseq*T, 256) fp32.
// q_cols / kv_cols: width of the contiguous buffers q and k|v live in (self: q|k|v in one of 768; cross: 256 and 512).
int launch_attention_x3(cudaStream_t st, const float* qbuf, int q_cols, const float* kvbuf, int kv_cols, int k_off,
                        int v_off, void* planes, float* out, int nseq, int T, int n_heads, const float* slopes, int cross,
                        int n_sm, std::string* err) {
  if (n_heads != 4 || n_heads * 64 != kDim) {
    if (err) *err = "attention_x3: needs 4 heads of 64";
    return -1;
  }
  if (cross && (nseq % 2)) {
    if (err) *err = "attention_x3: cross attention needs both channels";
    return -1;
  }
  const long long rows = (long long)nseq * T;
  // planes: [q hi][q lo][kv hi][kv lo] (self-attention: q buffer == kv buffer, split once)
  __half* qh = reinterpret_cast<__half*>(planes);
  __half* ql = qh + rows * q_cols;
  __half *kvh = qh, *kvl = ql;
  int launches = 0;
  auto split = [&](const float* in, __half* hi, __half* lo, long long n) {
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split16_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(hi),
                                                     reinterpret_cast<uint2*>(lo), n4);
    ++launches;
  };
  split(qbuf, qh, ql, rows * q_cols);
  if (kvbuf != qbuf) {
    kvh = ql + rows * q_cols;
    kvl = kvh + rows * kv_cols;
    split(kvbuf, kvh, kvl, rows * kv_cols);
  }
  AttnX3Params p{};
  auto mk = [&](CUtensorMap* m, const __half* base, long long rs) {
    const uint32_t box[3] = {64, 128, 1};
    const uint64_t dims[3] = {(uint64_t)kDim, (uint64_t)T, (uint64_t)nseq};
    const uint64_t strides[2] = {(uint64_t)rs, (uint64_t)rs * (uint64_t)T};
    return make_tmap(m, base, 2, 3, dims, strides, box, 128, err);
  };
  if (!mk(&p.tqh, qh, q_cols) || !mk(&p.tql, ql, q_cols) || !mk(&p.tkh, kvh + k_off, kv_cols) ||
      !mk(&p.tkl, kvl + k_off, kv_cols) || !mk(&p.tvh, kvh + v_off, kv_cols) || !mk(&p.tvl, kvl + v_off, kv_cols))
    return -1;
  p.out = out;
  p.slopes = slopes;
  p.nseq = nseq;
  p.T = T;
  p.nqt = (T + 127) / 128;
  p.n_items = p.nqt * nseq * 4;
  p.cross = cross;
  p.unit_major = nseq * 4 >= 2 * n_sm;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(attention_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM) != cudaSuccess) {
      if (err) *err = "attention_x3: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const int grid = p.n_items < n_sm ? p.n_items : n_sm;
  attention_x3_kernel<<<grid, X_THREADS, X_SMEM, st>>>(p);
  return launches + 1;
}

}  // namespace vapb
