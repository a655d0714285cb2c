// 16-bit fused causal ALiBi attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
// Reference: vap/modules.py:82-110 (scores, softmax, PV), :169-202 (bias 1 + m_h*j on
// allowed positions, -inf above the diagonal), :52 (scale 1/sqrt(dim) = 1/16).
// Self- and cross-attention share the kernel; cross-attention reads K/V rows of the
// other speaker channel's sequence (vap/modules.py:287-289).
//
// Work item = (sequence, head pair, 128-query tile). One persistent CTA per SM walks the
// items from the longest (last query tile) to the shortest. An item runs as FOUR independent
// chains ("slots"): slot = (head of the pair, key parity). The keys of a head are cut into
// 64-key half tiles; the even ones belong to one slot and the odd ones to the other, each with
// its own running maximum, row sum and O accumulator (split-K as in flash decoding), merged
// once per item. A slot's chain is  QK^T -> softmax -> PV -> QK^T of its next half tile;
// nothing in it waits for another slot (no exchange barrier: a softmax thread owns a whole
// 64-key row), so four chains overlap each other's tensor-pipe, MUFU and barrier latencies.
// Round-2 timeline (tools/attn_probe.py) of the two-slot version this replaces: every softmax
// warp waited 2300 of 3300 clk for its S tile because one issuing thread served both slots
// and a PV with its A operand in TMEM blocks the issuer for ~600 clk.
//
// Per slot and 64-key half tile (visited from the diagonal DOWN to key 0: the ALiBi bias grows
// with the key index, so the running row maximum is almost always set by the first tile and the
// accumulator rescale below is rare):
//   S[128 q][64 k]   = Q K^T      tcgen05.mma, Q and K tiles K-major SW128 in smem (TMA)
//   softmax          : one thread per query row (TMEM lane); two passes over S with
//                      tcgen05.ld (row max, then exp2 / row sum); P is written back over the
//                      first half of S as packed 16-bit pairs
//   O[128 q][64 d]  += P V        tcgen05.mma, A = P from TMEM, B = V tile MN-major SW128
// The T x T score matrix never exists; O stays in TMEM for the whole row of tiles.
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int AT_QTILE = 128 * 64 * 2;  // 128 query rows x 64 d of one head, 16-bit
constexpr int AT_KTILE = 64 * 64 * 2;   // 64 keys x 64 d of one head (K or V half tile)
// warp 0 / 1: TMA producer of head 0 / 1 of the pair; warps 2, 3, 20, 21: MMA issuer of slot 0..3 (warp 2 also owns
// the TMEM allocation); warps 4-19: softmax, four warps (= the four TMEM lane quadrants) per slot.
constexpr int AT_THREADS = 704;
constexpr int AT_OFF_Q = 0;                          // [head][2 buffers]
constexpr int AT_OFF_K = 4 * AT_QTILE;               // [slot][2 stages]
constexpr int AT_OFF_V = AT_OFF_K + 8 * AT_KTILE;    // [slot][2 stages]
constexpr int AT_OFF_X = AT_OFF_V + 8 * AT_KTILE;    // (m, l) exchange of the merge: float2 [slot][128]
constexpr int AT_OFF_BAR = AT_OFF_X + 4 * 128 * 8;
constexpr int AT_SMEM = AT_OFF_BAR + 512 + 1024 /*alignment slack*/;
// TMEM columns of a slot: S fp32 [0,64) with P (16-bit pairs) written over [0,32); O fp32 [64,128)
constexpr int AT_COL_O = 64, AT_SLOT_COLS = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(64) AttnParams {
  CUtensorMap tq, tk, tv;  // (256 head*d, T, nseq) 16-bit, SW128; box (64, 128, 1) for Q, (64, 64, 1) for K and V
  __nv_bfloat16* out;      // (nseq*T, 256)
  const float* slopes;     // [n_heads]
  int nseq, T, nqt, n_items, cross;  // n_items = (sequence, head, query tile) triples
  int fp16;
  long long* dbg;  // optional [64][8] SM-clock samples of CTA 0 (diagnostics, tools/attn_probe.py)
  int pair_major;  // large batches: a CTA walks all query tiles of one (sequence, head pair) back to back, so the
                   // pair's K/V (re-read once per query tile) stay in L2; small batches: spread single tiles
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
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

struct Item {
  int qi, seq, head;
};
// k-th work item of pipeline h (= head slot pair: two of the CTA's four slots) of this CTA; false when it is done.
// The two pipelines of a CTA are independent (own producer, issuers, softmax warps, Q buffers), so the schedule is over
// 2 * gridDim.x pipelines. The head of a unit rotates with the round: the steep ALiBi heads skip most far tiles, and a
// pipeline that always drew the same head would finish long before (or after) the others.
__device__ __forceinline__ bool next_item(const AttnParams& p, int k, int h, Item* it) {
  const int pipe = blockIdx.x * 2 + h, n_pipes = gridDim.x * 2;
  if (p.pair_major) {
    const int round = k / p.nqt;
    const int u = pipe + round * n_pipes;  // unit = (sequence, head): all its query tiles back to back
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

// FP16: 16-bit format of Q / K / V / P / out (compile time: a run-time flag made every pack two predicated instructions)
template <int FP16>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + AT_OFF_BAR;
  auto q_full = [&](int h, int b) { return bar_base + 8u * (h * 2 + b); };
  auto q_empty = [&](int h, int b) { return bar_base + 8u * (4 + h * 2 + b); };
  auto kv_full = [&](int sl, int st) { return bar_base + 8u * (8 + sl * 2 + st); };
  auto kv_empty = [&](int sl, int st) { return bar_base + 8u * (16 + sl * 2 + st); };
  auto s_full = [&](int sl) { return bar_base + 8u * (24 + sl); };
  auto p_full = [&](int sl) { return bar_base + 8u * (28 + sl); };
  auto o_final = [&](int sl) { return bar_base + 8u * (32 + sl); };
  const uint32_t tmem_slot = bar_base + 8u * 36;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tq);
    prefetch_tmap(&p.tk);
    prefetch_tmap(&p.tv);
    for (int h = 0; h < 2; ++h)
      for (int b = 0; b < 2; ++b) {
        mbar_init(q_full(h, b), 1);
        mbar_init(q_empty(h, b), 2);  // released by the last PV of both slots of the head
      }
    for (int sl = 0; sl < 4; ++sl) {
      for (int st = 0; st < 2; ++st) {
        mbar_init(kv_full(sl, st), 1);
        mbar_init(kv_empty(sl, st), 1);
      }
      mbar_init(s_full(sl), 1);
      mbar_init(p_full(sl), 128);
      mbar_init(o_final(sl), 1);
    }
    fence_barrier_init();
  }
  if (warp == 3 && lane < 4) reinterpret_cast<float*>(smem_gen + AT_OFF_BAR + 8 * 38)[lane] = p.slopes[lane];
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp < 2) {
    // ===== TMA producer of head `warp` of the pair: the head's Q tile and the K/V half tiles of its two slots.
    // The Q tile of the NEXT item goes behind the first K/V half tiles of the current one: its buffer is released
    // by the end of the previous item, and waiting for that before any K/V request would drain the K/V ring at
    // every item boundary.
    if (lane == 0) {
      const int h = warp;
      uint32_t kvc = 0;  // half-tile steps issued (the two slots of a head move together)
      auto load_q = [&](const Item& it, uint32_t n) {
        const int b = n & 1;
        mbar_wait(q_empty(h, b), ((n >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full(h, b), AT_QTILE);
        tma_load_3d(smem_base + AT_OFF_Q + (h * 2 + b) * AT_QTILE, &p.tq, q_full(h, b), it.head * 64,
                    it.qi * 128, it.seq);
      };
      Item it, nx;
      bool have = next_item(p, 0, h, &it);
      if (have) load_q(it, 0);
      for (uint32_t k = 0; have; ++k) {
        const bool have_next = next_item(p, (int)k + 1, h, &nx);
        const int kvseq = p.cross ? (it.seq + p.nseq / 2) % p.nseq : it.seq;
        const int col = it.head * 64;
        for (int kt = it.qi; kt >= 0; --kt) {
          const int st = kvc & 1;
          const uint32_t par = ((kvc >> 1) & 1u) ^ 1u;
          ++kvc;
          for (int kp = 0; kp < 2; ++kp) {
            const int sl = h * 2 + kp;
            mbar_wait(kv_empty(sl, st), par);
            mbar_arrive_expect_tx(kv_full(sl, st), 2 * AT_KTILE);
            tma_load_3d(smem_base + AT_OFF_K + (sl * 2 + st) * AT_KTILE, &p.tk, kv_full(sl, st), col,
                        (2 * kt + kp) * 64, kvseq);
            tma_load_3d(smem_base + AT_OFF_V + (sl * 2 + st) * AT_KTILE, &p.tv, kv_full(sl, st), col,
                        (2 * kt + kp) * 64, kvseq);
          }
          if (kt == it.qi && have_next) load_q(nx, k + 1);
        }
        it = nx;
        have = have_next;
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 20) {
    // ===== MMA issuers, one thread per slot. Each walks its slot's chain p_full -> PV -> QK of the next half tile
    // on its own; tcgen05.commit tracks the issuing thread's own MMAs, so the chains only meet in the tensor pipe.
    if (lane == 0) {
      const int sl = warp < 4 ? warp - 2 : warp - 18, h = sl >> 1;
      const uint32_t idesc_qk = make_idesc_16(128, 64, 0, 0, FP16);
      const uint32_t idesc_pv = make_idesc_16(128, 64, 0, 1, FP16);  // B = V is MN-major (d contiguous)
      const uint32_t t_slot = tmem_base + sl * AT_SLOT_COLS;
      uint32_t n_item = 0, kvc = 0, pc = 0;
      auto issue_qk = [&]() {
        const int st = kvc & 1;
        mbar_wait(kv_full(sl, st), (kvc >> 1) & 1u);
        tc_fence_after();
        const uint32_t qa = smem_base + AT_OFF_Q + (h * 2 + (n_item & 1)) * AT_QTILE;
        const uint32_t ka = smem_base + AT_OFF_K + (sl * 2 + st) * AT_KTILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_slot, make_smem_desc_sw128(qa + k * 32, 0, 1024), make_smem_desc_sw128(ka + k * 32, 0, 1024),
                    idesc_qk, k != 0);
        umma_commit(s_full(sl));
      };
      Item it;
      for (int k = 0; next_item(p, k, h, &it); ++k, ++n_item) {
        const int ntiles = it.qi + 1;
        mbar_wait(q_full(h, n_item & 1), (n_item >> 1) & 1u);
        issue_qk();
        for (int n = 0; n < ntiles; ++n) {
          const bool dbg = p.dbg && blockIdx.x == 0 && sl == 0 && pc >= 40 && pc < 104;
          if (dbg) p.dbg[(pc - 40) * 8 + 5] = clock64();
          mbar_wait(p_full(sl), pc & 1u);
          if (dbg) p.dbg[(pc - 40) * 8 + 6] = clock64();
          ++pc;
          tc_fence_after();
          const int st = kvc & 1;
          const uint32_t va = smem_base + AT_OFF_V + (sl * 2 + st) * AT_KTILE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ts(t_slot + AT_COL_O, t_slot + kk * 8, make_smem_desc_sw128(va + kk * 2048, 1024, 1024), idesc_pv,
                         (n | kk) != 0);
          umma_commit(kv_empty(sl, st));
          ++kvc;
          if (dbg) p.dbg[(pc - 41) * 8 + 7] = clock64();
          if (n + 1 < ntiles) {
            issue_qk();
          } else {
            umma_commit(q_empty(h, n_item & 1));
            umma_commit(o_final(sl));
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== softmax: slot = (head of the pair, key parity); thread = one query row of the slot's 64-key half tiles
    const int sl = (warp - 4) >> 2, h = sl >> 1, kp = sl & 1, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_s = tmem_base + ((uint32_t)(quad * 32) << 16) + sl * AT_SLOT_COLS;
    const uint32_t t_o = t_s + AT_COL_O;
    const uint32_t t_o_other = tmem_base + ((uint32_t)(quad * 32) << 16) + (sl ^ 1) * AT_SLOT_COLS + AT_COL_O;
    float2* xch = reinterpret_cast<float2*>(smem_gen + AT_OFF_X);  // [slot][128]
    auto head_bar = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + h) : "memory"); };
    constexpr float SC = 0.0625f * kLog2e;
    uint32_t sc_cnt = 0, oc_cnt = 0;
    const float* slope_s = reinterpret_cast<const float*>(smem_gen + AT_OFF_BAR + 8 * 38);
    Item it;
    for (int k = 0; next_item(p, k, h, &it); ++k) {
      const int head = it.head;
      const float slope2 = slope_s[head] * kLog2e;
      float m = -INFINITY, l = 0.f;  // running maximum (log2 domain) and row sum over this slot's keys
      for (int n = 0; n <= it.qi; ++n) {
        const int k0 = (2 * (it.qi - n) + kp) * 64;  // first key of the half tile
        const bool diag = n == 0;
        // On the diagonal tile the half tile starts kp*64 keys into the query tile: 32-key chunk c is visible to this
        // warp's rows iff its first key kp*64 + 32c <= quad*32, and holds the diagonal itself when equal.
        const int dchunk = quad - 2 * kp;  // chunk index of the diagonal for this warp (may be < 0 or > 1)
        const int nvis = diag ? min(max(dchunk + 1, 0), 2) : 2;
        const float base = fmaf(slope2, (float)k0, kLog2e);
        const bool dbg = p.dbg && blockIdx.x == 0 && sl == 0 && quad == 0 && lane == 0 && sc_cnt >= 40 && sc_cnt < 104;
        const uint32_t di = (sc_cnt - 40) * 8;
        if (dbg) p.dbg[di + 0] = clock64();
        mbar_wait(s_full(sl), sc_cnt & 1u);
        if (dbg) p.dbg[di + 1] = clock64();
        ++sc_cnt;
        tc_fence_after();
        // The kernel is bound by TMEM reads (64 B/clk/SM: one pass over a half tile of S is 32 KB = 512 clk), so S is
        // read ONCE where that is possible: from the second tile on the scores are exponentiated against the running
        // reference m as they are (softmax is shift-invariant; m does not have to be the exact maximum), and the tile
        // maximum is tracked on the side. Only if some row exceeds the reference by more than 8 binades (P is a 16-bit
        // float: fp16 tops out at 2^16) is the tile redone the exact way below, with the P stores held back until then
        // because P overwrites S.
        float2 ps2 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(SC, SC), step2 = make_float2(2.f * slope2, 2.f * slope2);
        bool done = false;
        if (!diag && !__any_sync(0xffffffffu, m == -INFINITY)) {
          const float base_m = base - m;
          uint32_t pk[2][16];
          float tmax = -INFINITY;
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            uint32_t r[32];
            tmem_ld32(t_s + ci * 32, r);
            tmem_ld_wait();
            const float cb = fmaf(slope2, (float)(ci * 32), base_m);
            float2 bias2 = make_float2(cb, cb + slope2);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 t = __ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, bias2);
              bias2 = __fadd2_rn(bias2, step2);
              tmax = fmaxf(tmax, fmaxf(t.x, t.y));
              const float p0 = ex2_fast(t.x), p1 = ex2_fast(t.y);
              ps2 = __fadd2_rn(ps2, make_float2(p0, p1));
              pk[ci][i >> 1] = pack16(p0, p1, FP16);
            }
          }
          if (!__any_sync(0xffffffffu, tmax > 8.0f)) {
            tmem_st16(t_s, pk[0]);
            tmem_st16(t_s + 16, pk[1]);
            done = true;
          } else {
            ps2 = make_float2(0.f, 0.f);
          }
          if (dbg) p.dbg[di + 2] = clock64();
          if (dbg) p.dbg[di + 3] = clock64();
        }
        if (!done) {
          // pass A: upper bound of the row maximum (log2 domain): SC * max_j s_j + bias of the last visible key
          float mx = -INFINITY;
  #pragma unroll 1
          for (int ci = 0; ci < nvis; ++ci) {
            uint32_t r[32];
            tmem_ld32(t_s + ci * 32, r);
            tmem_ld_wait();
            if (diag && ci == dchunk) {
  #pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i <= lane ? __uint_as_float(r[i]) : -INFINITY);
            } else {
  #pragma unroll
              for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
            }
          }
          if (dbg) p.dbg[di + 2] = clock64();
          if (dbg) p.dbg[di + 3] = clock64();
          // last visible key of the row in this half tile (diag: the query itself, at most key 63 of the half tile)
          const int last = diag ? min(row - kp * 64, 63) : 63;
          const float b_last = fmaf(slope2, (float)last, base);
          const float m_new = fmaxf(m, fmaf(mx, SC, b_last));  // nvis == 0: mx = -inf, m stays
          if (n > 0 && __any_sync(0xffffffffu, m_new > m)) {
            // rare: rescale the accumulator row (PV of the previous tile has completed: s_full was committed after it)
            const float alpha = ex2_fast(m - m_new);  // m = -inf (nothing seen so far): 0, and O is 0
            l *= alpha;
  #pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32];
              tmem_ld32(t_o + c * 32, r);
              tmem_ld_wait();
  #pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st32(t_o + c * 32, r);
            }
          }
          // ALiBi makes far tiles of the steep heads irrelevant: if even the tile's upper bound is more than 40 binades
          // below the running maximum for every row of the warp, every p would be < 2^-40 of the row's largest term
          // (fp32 cannot see it in the row sum or in O), so the exponentials are skipped and P is written as zeros.
          const bool negligible = n > 0 && __all_sync(0xffffffffu, fmaf(mx, SC, b_last) < m - 40.0f);
          m = m_new;
          // pass B: p = 2^(t - m), row sum, P -> TMEM as 16-bit pairs over S columns [0, 32) (column c holds keys 2c, 2c+1;
          // chunk 1 of S is in registers before chunk 0's P lands on columns [0, 16), and its own P goes to [16, 32))
          const float base_m = base - m;
          const int nexp = negligible ? 0 : nvis;
  #pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            uint32_t pk[16];
            if (ci < nexp) {
              uint32_t r[32];
              tmem_ld32(t_s + ci * 32, r);
              tmem_ld_wait();
              const float cb = fmaf(slope2, (float)(ci * 32), base_m);
              // bias of the key pair (i, i + 1), stepped by 2 * slope per pair (no per-element constant to materialise)
              float2 bias2 = make_float2(cb, cb + slope2);
              if (diag && ci == dchunk) {
  #pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float2 t = __ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, bias2);
                  bias2 = __fadd2_rn(bias2, step2);
                  float p0 = ex2_fast(t.x), p1 = ex2_fast(t.y);
                  if (i > lane) p0 = 0.f;  // keys above the diagonal
                  if (i + 1 > lane) p1 = 0.f;
                  ps2 = __fadd2_rn(ps2, make_float2(p0, p1));
                  pk[i >> 1] = pack16(p0, p1, FP16);
                }
              } else {
  #pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float2 t = __ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, bias2);
                  bias2 = __fadd2_rn(bias2, step2);
                  const float p0 = ex2_fast(t.x), p1 = ex2_fast(t.y);
                  ps2 = __fadd2_rn(ps2, make_float2(p0, p1));
                  pk[i >> 1] = pack16(p0, p1, FP16);
                }
              }
            } else {
  #pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = 0u;
            }
            tmem_st16(t_s + ci * 16, pk);
          }
        }
        l += ps2.x + ps2.y;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(sl));
        if (dbg) p.dbg[di + 4] = clock64();
      }
      // ===== merge of the head's two slots and epilogue: this thread writes d columns [32 kp, 32 kp + 32) of its row
      const bool dbe = p.dbg && blockIdx.x == 0 && sl == 0 && quad == 0 && lane == 0 && k < 60;
      long long* de = p.dbg + 512 + k * 8;
      if (dbe) de[0] = clock64();
      xch[sl * 128 + row] = make_float2(m, l);
      head_bar();
      if (dbe) de[1] = clock64();
      const float2 o = xch[(sl ^ 1) * 128 + row];
      const float mm = fmaxf(m, o.x);  // finite: the even slot always sees the row's own key
      const float w_own = ex2_fast(m - mm), w_oth = ex2_fast(o.x - mm);  // 2^-inf = 0 for a slot that saw no key
      const float inv = 1.0f / fmaf(w_own, l, w_oth * o.y);
      const float a_own = w_own * inv, a_oth = w_oth * inv;
      mbar_wait(o_final(sl), oc_cnt & 1u);
      mbar_wait(o_final(sl ^ 1), oc_cnt & 1u);
      ++oc_cnt;
      tc_fence_after();
      if (dbe) de[2] = clock64();
      const int q = it.qi * 128 + row;
      __nv_bfloat16* dst = p.out + ((long long)it.seq * p.T + q) * kDim + head * 64 + kp * 32;
      {
        uint32_t r[32], r2[32];
        tmem_ld32(t_o + kp * 32, r);
        tmem_ld32(t_o_other + kp * 32, r2);
        tmem_ld_wait();
        tc_fence_before();
        if (dbe) de[3] = clock64();
        // the partner slot's first PV of the next item overwrites the accumulator this thread has just read: its
        // p_full is arrived by the partner's threads, so both slots meet here before either goes on
        head_bar();
        if (dbe) de[4] = clock64();
        if (q < p.T) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[j] = fmaf(__uint_as_float(r[8 * i + j]), a_own, __uint_as_float(r2[8 * i + j]) * a_oth);
            uint4 u;
            u.x = pack16(f[0], f[1], FP16);
            u.y = pack16(f[2], f[3], FP16);
            u.z = pack16(f[4], f[5], FP16);
            u.w = pack16(f[6], f[7], FP16);
            *reinterpret_cast<uint4*>(dst + 8 * i) = u;
          }
        }
        if (dbe) de[5] = clock64();
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

// q/k/v: bf16 rows of 256 (= n_heads*64) at ptr + (seq*T + t)*row_stride; out: dense (nseq*T, 256) bf16.
int launch_attention_tc(cudaStream_t st, const __nv_bfloat16* q, long long q_row_stride, const __nv_bfloat16* k,
                        const __nv_bfloat16* v, long long kv_row_stride, __nv_bfloat16* out, int nseq, int T,
                        int n_heads, const float* slopes, int cross, int n_sm, std::string* err, long long* dbg) {
  if (n_heads != 4 || n_heads * 64 != kDim) {
    if (err) *err = "attention_tc: needs 4 heads of 64";
    return -1;
  }
  if (cross && (nseq % 2)) {
    if (err) *err = "attention_tc: cross attention needs both channels";
    return -1;
  }
  AttnParams p{};
  auto mk = [&](CUtensorMap* m, const void* base, long long rs, uint32_t rows) {
    const uint32_t box[3] = {64, rows, 1};
    const uint64_t dims[3] = {(uint64_t)kDim, (uint64_t)T, (uint64_t)nseq};
    const uint64_t strides[2] = {(uint64_t)rs, (uint64_t)rs * (uint64_t)T};
    return make_tmap(m, base, 2, 3, dims, strides, box, 128, err);
  };
  if (!mk(&p.tq, q, q_row_stride, 128) || !mk(&p.tk, k, kv_row_stride, 64) || !mk(&p.tv, v, kv_row_stride, 64)) return -1;
  p.out = out;
  p.slopes = slopes;
  p.nseq = nseq;
  p.T = T;
  p.nqt = (T + 127) / 128;
  p.n_items = p.nqt * nseq * 4;
  p.cross = cross;
  p.dbg = dbg;
  p.fp16 = g_fp16;
  p.pair_major = nseq * 4 >= 4 * n_sm;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) {
      if (err) *err = "attention_tc: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const int grid = (p.n_items + 1) / 2 < n_sm ? (p.n_items + 1) / 2 : n_sm;
  if (g_fp16) launch_pdl(attention_tc_kernel<1>, grid, AT_THREADS, AT_SMEM, st, p);
  else launch_pdl(attention_tc_kernel<0>, grid, AT_THREADS, AT_SMEM, st, p);
  return 1;
}

}  // namespace vapb
