// 16-bit fused causal ALiBi attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
// Reference: vap/modules.py:82-110 (scores, softmax, PV), :169-202 (bias 1 + m_h*j on
// allowed positions, -inf above the diagonal), :52 (scale 1/sqrt(dim) = 1/16).
// Self- and cross-attention share the kernel; cross-attention reads K/V rows of the
// other speaker channel's sequence (vap/modules.py:287-289).
//
// Work item = (sequence, head, 128-query tile). A persistent CTA per SM runs TWO independent
// pipelines ("slots"), each with its own TMA producer thread, MMA issuing thread, eight softmax
// warps, buffers and item list (all query tiles of one (sequence, head) back to back, longest
// first, so its K/V re-reads hit L2). While the softmax warps of one slot are on the CUDA cores,
// the other slot's QK^T and PV are on the tensor pipe; nothing in a slot's chain
//   QK^T -> softmax -> PV -> QK^T of the next key tile
// waits for the other slot.
//
// Per slot and 128-key tile (visited from the diagonal DOWN to key 0: the ALiBi bias grows with
// the key index, so the first tile almost always sets the row maximum):
//   S[128 q][128 k]  = Q K^T      tcgen05.mma, Q and K tiles K-major SW128 in smem (TMA)
//   softmax          : a thread owns half a query row (64 keys; warps 4 apart share a TMEM lane
//                      quadrant). S is read with tcgen05.ld ONCE: p = 2^(t - m) against the
//                      running reference m (softmax is shift-invariant, m need not be the exact
//                      maximum of the row so far), and only the first tile of an item - or a tile in which some row's
//                      p sum to more than 2^8, i.e. the reference has fallen far behind - takes the
//                      exact two-pass route. P goes back to TMEM as packed 16-bit pairs.
//   O[128 q][64 d]  += P V        tcgen05.mma, A = P from TMEM, B = V tile MN-major SW128
// QK^T of the NEXT tile is issued before PV of the current one (it only needs S), so the softmax
// warps are back at work while PV is still on the tensor pipe; they wait for it (pv_done) before
// they overwrite P or rescale O. The T x T score matrix never exists; O stays in TMEM for the
// whole row of tiles.
//
// What bounds it (tools/attn_probe.py, tools/attn_time.py, round 2): TMEM READ bandwidth,
// 64 B/clk/SM. The first version read S twice (2 x 64 KB per tile) and fed P to the PV MMAs from
// TMEM (32 KB per tile): 160 KB per tile = 2500 clk = its measured time per tile, while the MUFU
// (1024 clk), the tensor pipe (~900 clk: an MMA costs >= 64 clk per K16 step whatever N <= 128
// is, the A operand streams) and shared memory were each under half of that. One pass over S
// brings it to 96 KB = 1500 clk per tile (measured ~1700 in steady state). Variants measured on
// the same box (630 us for the two-pass kernel at 512 sequences x 1000 frames, 594 us for this
// one): P through shared memory instead of TMEM 645-660 us and half of P each way 631 us (the
// SS-mode PV MMAs and the P stores contend with the K/V/Q operand traffic in shared memory);
// the keys of a head split over two more slots with 64-key tiles 578-656 us (twice the K16 steps:
// tensor-pipe-bound).
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int AT_TILE = 128 * 64 * 2;  // one 128-row x 64-column 16-bit tile (Q, K or V of one head)
// warp 0 / 1: TMA producer of slot 0 / 1; warp 2 / 3: MMA issuer of slot 0 / 1 (warp 2 also owns the TMEM allocation);
// warps 4-11 softmax of slot 0, 12-19 of slot 1
constexpr int AT_THREADS = 640;
constexpr int AT_OFF_Q = 0;                         // [slot] (one buffer: released by the item's last QK)
constexpr int AT_OFF_K = 2 * AT_TILE;               // [slot][2 stages]
constexpr int AT_OFF_V = AT_OFF_K + 4 * AT_TILE;    // [slot][2 stages]
constexpr int AT_OFF_X = AT_OFF_V + 4 * AT_TILE;    // exchange between the two half-row threads: float [slot][half][128]
constexpr int AT_OFF_BAR = AT_OFF_X + 2 * 2 * 128 * 4;
constexpr int AT_SMEM = AT_OFF_BAR + 320 + 1024 /*alignment slack*/;
// TMEM columns of a slot: S fp32 [0,128), P 16-bit pairs [128,192), O fp32 [192,256)
constexpr int AT_COL_P = 128, AT_COL_O = 192, AT_SLOT_COLS = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(64) AttnParams {
  CUtensorMap tq, tk, tv;  // (256 head*d, T, nseq) 16-bit, SW128; box (64, 128, 1)
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
// k-th work item of pipeline h (= slot) of this CTA; false when it is done.
// The two pipelines of a CTA are independent (own producer, issuer, softmax warps, buffers), so the schedule is over
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
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + AT_OFF_BAR;
  auto q_full = [&](int s) { return bar_base + 8u * s; };
  auto q_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto kv_full = [&](int s, int st) { return bar_base + 8u * (4 + s * 2 + st); };
  auto kv_empty = [&](int s, int st) { return bar_base + 8u * (8 + s * 2 + st); };
  auto s_full = [&](int s) { return bar_base + 8u * (12 + s); };
  auto p_full = [&](int s) { return bar_base + 8u * (14 + s); };
  auto o_final = [&](int s) { return bar_base + 8u * (16 + s); };
  auto pv_done = [&](int s) { return bar_base + 8u * (18 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 20;
  float* slope_s = reinterpret_cast<float*>(smem_gen + AT_OFF_BAR + 8 * 22);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tq);
    prefetch_tmap(&p.tk);
    prefetch_tmap(&p.tv);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      for (int st = 0; st < 2; ++st) {
        mbar_init(kv_full(s, st), 1);
        mbar_init(kv_empty(s, st), 1);
      }
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 256);
      mbar_init(o_final(s), 1);
      mbar_init(pv_done(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 3 && lane < 4) slope_s[lane] = p.slopes[lane];  // weights: not written by the preceding kernel
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp < 2) {
    // ===== TMA producer of slot `warp`
    if (lane == 0) {
      const int s = warp;
      uint32_t kvc = 0;
      auto load_q = [&](const Item& it, uint32_t n) {
        mbar_wait(q_empty(s), (n & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full(s), AT_TILE);
        tma_load_3d(smem_base + AT_OFF_Q + s * AT_TILE, &p.tq, q_full(s), it.head * 64, it.qi * 128, it.seq);
      };
      Item it, nx;
      bool have = next_item(p, 0, s, &it);
      if (have) load_q(it, 0);
      for (uint32_t k = 0; have; ++k) {
        const bool have_next = next_item(p, (int)k + 1, s, &nx);
        const int kvseq = p.cross ? (it.seq + p.nseq / 2) % p.nseq : it.seq;
        const int col = it.head * 64;
        for (int kt = it.qi; kt >= 0; --kt) {
          const int st = kvc & 1;
          mbar_wait(kv_empty(s, st), ((kvc >> 1) & 1u) ^ 1u);
          ++kvc;
          mbar_arrive_expect_tx(kv_full(s, st), 2 * AT_TILE);
          tma_load_3d(smem_base + AT_OFF_K + (s * 2 + st) * AT_TILE, &p.tk, kv_full(s, st), col, kt * 128, kvseq);
          tma_load_3d(smem_base + AT_OFF_V + (s * 2 + st) * AT_TILE, &p.tv, kv_full(s, st), col, kt * 128, kvseq);
        }
        // the Q buffer is released by the item's last QK, about one tile before the item ends: the next Q tile lands
        // behind that tile's softmax and PV
        if (have_next) load_q(nx, k + 1);
        it = nx;
        have = have_next;
      }
    }
  } else if (warp < 4) {
    // ===== MMA issuer of slot `warp - 2`. tcgen05.commit tracks the issuing thread's own MMAs, so the two slots'
    // chains only meet in the tensor pipe's queue.
    if (lane == 0) {
      const int s = warp - 2;
      const uint32_t idesc_qk = make_idesc_16(128, 128, 0, 0, FP16);
      const uint32_t idesc_pv = make_idesc_16(128, 64, 0, 1, FP16);  // B = V is MN-major (d contiguous)
      const uint32_t t_slot = tmem_base + s * AT_SLOT_COLS;
      const uint32_t qa = smem_base + AT_OFF_Q + s * AT_TILE;
      uint32_t n_item = 0, kq = 0, kpv = 0, pc = 0;  // kq / kpv: key tiles whose QK / PV has been issued
      auto issue_qk = [&](long long* stamp = nullptr) {
        const int st = kq & 1;
        mbar_wait(kv_full(s, st), (kq >> 1) & 1u);
        if (stamp) *stamp = clock64();
        ++kq;
        tc_fence_after();
        const uint32_t ka = smem_base + AT_OFF_K + (s * 2 + st) * AT_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(t_slot, make_smem_desc_sw128(qa + k * 32, 0, 1024), make_smem_desc_sw128(ka + k * 32, 0, 1024),
                    idesc_qk, k != 0);
        umma_commit(s_full(s));
      };
      Item it;
      for (int k = 0; next_item(p, k, s, &it); ++k, ++n_item) {
        const int ntiles = it.qi + 1;
        mbar_wait(q_full(s), n_item & 1u);
        issue_qk();
        if (ntiles == 1) umma_commit(q_empty(s));
        for (int n = 0; n < ntiles; ++n) {
          const bool dbg = p.dbg && blockIdx.x == 0 && s == 0 && pc >= 40 && pc < 104;
          if (dbg) p.dbg[(pc - 40) * 8 + 5] = clock64();
          mbar_wait(p_full(s), pc & 1u);
          if (dbg) p.dbg[(pc - 40) * 8 + 6] = clock64();
          ++pc;
          tc_fence_after();
          // QK of the NEXT tile goes first: it only needs S (free: every softmax thread has read it), and the softmax
          // warps can start on it while this tile's PV is still on the tensor pipe. They wait for pv_done before they
          // overwrite P or touch O.
          long long* d2 = dbg ? p.dbg + 512 + (pc - 41) * 8 : nullptr;
          if (dbg) d2[0] = clock64();
          if (n + 1 < ntiles) {
            issue_qk(dbg ? d2 + 1 : nullptr);
            if (n + 2 == ntiles) umma_commit(q_empty(s));  // that was the item's last QK
          }
          if (dbg) d2[2] = clock64();
          const int st = kpv & 1;
          ++kpv;
          const uint32_t va = smem_base + AT_OFF_V + (s * 2 + st) * AT_TILE;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts(t_slot + AT_COL_O, t_slot + AT_COL_P + kk * 8, make_smem_desc_sw128(va + kk * 2048, 1024, 1024),
                         idesc_pv, (n | kk) != 0);
          umma_commit(kv_empty(s, st));
          umma_commit(pv_done(s));
          if (n + 1 == ntiles) umma_commit(o_final(s));
          if (dbg) p.dbg[(pc - 41) * 8 + 7] = clock64();
          if (dbg) d2[3] = clock64();
        }
      }
    }
  } else {
    // ===== softmax: thread = (query row, half of the tile's keys)
    const int s = (warp - 4) >> 3, quad = warp & 3, ch = ((warp - 4) >> 2) & 1;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + s * AT_SLOT_COLS;
    const uint32_t t_s = t_row + ch * 64, t_o = t_row + AT_COL_O + ch * 32;
    float* xs = reinterpret_cast<float*>(smem_gen + AT_OFF_X) + s * 256;  // [half][128]
    float* x_own = xs + ch * 128 + row;
    const float* x_oth = xs + (ch ^ 1) * 128 + row;
    auto slot_bar = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + s) : "memory"); };
    auto slot_any = [&](bool v) {  // OR over the slot's 256 threads (the same named barrier)
      uint32_t r;
      asm volatile(
          "{\n\t.reg .pred pi, po;\n\tsetp.ne.b32 pi, %2, 0;\n\tbar.red.or.pred po, %1, 256, pi;\n\t"
          "selp.u32 %0, 1, 0, po;\n\t}"
          : "=r"(r)
          : "r"(1 + s), "r"((uint32_t)v)
          : "memory");
      return r != 0;
    };
    // PV of the slot's previous tile has completed (P may be overwritten, O may be rescaled); tiles are counted over
    // the whole kernel, the first one has no predecessor
    auto wait_pv = [&](uint32_t tile) {
      if (tile > 0) mbar_wait(pv_done(s), (tile - 1) & 1u);
    };
    const uint32_t t_p = t_row + AT_COL_P + ch * 32;  // column c of P holds keys 2c, 2c + 1
    auto store_p = [&](int ci, const uint32_t (&pk)[16]) { tmem_st16(t_p + ci * 16, pk); };
    constexpr float SC = 0.0625f * kLog2e;
    uint32_t sc_cnt = 0, oc_cnt = 0;
    Item it;
    for (int k = 0; next_item(p, k, s, &it); ++k) {
      const int head = it.head;
      const float slope2 = slope_s[head] * kLog2e;
      const float2 sc2 = make_float2(SC, SC), step2 = make_float2(2.f * slope2, 2.f * slope2);
      float m = -INFINITY, l = 0.f;  // m: the row's reference (both half-row threads hold the same); l: partial sum
      for (int n = 0; n <= it.qi; ++n) {
        const int k0 = (it.qi - n) * 128;
        const bool diag = n == 0;
        // this thread's two 32-key chunks are global chunks 2*ch and 2*ch+1; on the diagonal tile chunk g is
        // visible to this warp's rows iff g <= quad, and chunk g == quad holds the diagonal itself
        const int g0 = 2 * ch;
        const int nvis = diag ? min(max(quad + 1 - g0, 0), 2) : 2;
        const float base = fmaf(slope2, (float)(k0 + 64 * ch), kLog2e);
        const bool dbg = p.dbg && blockIdx.x == 0 && s == 0 && ch == 0 && quad == 0 && lane == 0 && sc_cnt >= 40 && sc_cnt < 104;
        const uint32_t di = (sc_cnt - 40) * 8;
        if (dbg) p.dbg[di + 0] = clock64();
        mbar_wait(s_full(s), sc_cnt & 1u);
        if (dbg) p.dbg[di + 1] = clock64();
        const uint32_t tile = sc_cnt++;
        tc_fence_after();
        float2 ps2 = make_float2(0.f, 0.f);
        bool done = false;
        if (!diag) {
          // ---- one pass against the running reference (finite after the diagonal tile: a row always sees itself)
          const float base_m = base - m;
#pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            uint32_t r[32], pk[16];
            tmem_ld32(t_s + ci * 32, r);
            tmem_ld_wait();
            const float cb = fmaf(slope2, (float)(ci * 32), base_m);
            // bias of the key pair (i, i + 1), stepped by 2 * slope per pair (no per-element constant to materialise)
            float2 bias2 = make_float2(cb, cb + slope2);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 t = __ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, bias2);
              bias2 = __fadd2_rn(bias2, step2);
              const float p0 = ex2_fast(t.x), p1 = ex2_fast(t.y);
              ps2 = __fadd2_rn(ps2, make_float2(p0, p1));
              pk[i >> 1] = pack16(p0, p1, FP16);
            }
            if (ci == 0) {
              if (dbg) p.dbg[512 + di + 4] = clock64();
              wait_pv(tile);
              if (dbg) p.dbg[512 + di + 5] = clock64();
            }
            store_p(ci, pk);
          }
          if (dbg) p.dbg[di + 2] = clock64();
          // The reference still holds if no p exceeds 2^8 (P is a 16-bit float). The p are non-negative, so the sum of
          // this thread's 64 bounds each of them: no separate maximum has to be tracked (one instruction per score pair
          // less in a loop that is partly issue-bound), at the price of taking the exact route a little earlier than
          // needed. Every thread of the slot takes the same route (the exact one has block barriers in it).
          if (!slot_any(!(ps2.x + ps2.y <= 256.0f))) {
            done = true;
          } else {
            ps2 = make_float2(0.f, 0.f);
            tmem_st_wait();  // the exact route rewrites P: its stores must not overtake the ones just issued
          }
          if (dbg) p.dbg[di + 3] = clock64();
        }
        if (!done) {
          // ---- exact route. Pass A: an upper bound of the row maximum of t = SC * s_j + bias_j (log2 domain), taken per
          // group of 16 keys as SC * (largest s of the group) + (bias of the group's last visible key): at most 15 slopes =
          // 5.4 binades (steepest head) above the true maximum. (One bound for the whole tile can sit 46 binades above it
          // when a far key of a steep ALiBi head dominates, and p, a 16-bit float, then loses its precision or
          // underflows.) S is intact (P has its own TMEM columns), so a tile the single pass gave up on is simply redone.
          float mx = -INFINITY;
#pragma unroll 1
          for (int ci = 0; ci < nvis; ++ci) {
            uint32_t r[32];
            tmem_ld32(t_s + ci * 32, r);
            tmem_ld_wait();
            const float cb = fmaf(slope2, (float)(ci * 32), base);  // bias of the chunk's first key
            const bool dchunk = diag && g0 + ci == quad;
#pragma unroll
            for (int gq = 0; gq < 2; ++gq) {
              float mg = -INFINITY;
              if (dchunk) {
#pragma unroll
                for (int i = 16 * gq; i < 16 * gq + 16; ++i) mg = fmaxf(mg, i <= lane ? __uint_as_float(r[i]) : -INFINITY);
              } else {
#pragma unroll
                for (int i = 16 * gq; i < 16 * gq + 16; i += 2)
                  mg = fmaxf(mg, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
              }
              const int lastv = dchunk ? min(lane, 16 * gq + 15) : 16 * gq + 15;
              mx = fmaxf(mx, fmaf(mg, SC, fmaf(slope2, (float)lastv, cb)));  // mg = -inf: the group is not visible
            }
          }
          *x_own = mx;
          slot_bar();
          mx = fmaxf(mx, *x_oth);
          const float m_new = fmaxf(m, mx);
          wait_pv(tile);
          tc_fence_after();
          if (n > 0 && __any_sync(0xffffffffu, m_new > m)) {
            // rescale this thread's half of the accumulator row
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
          // pass B: p = 2^(t - m), partial row sum, P -> shared memory
          const float base_m = base - m;
#pragma unroll 1
          for (int ci = 0; ci < 2; ++ci) {
            uint32_t pk[16];
            if (ci < nvis) {
              uint32_t r[32];
              tmem_ld32(t_s + ci * 32, r);
              tmem_ld_wait();
              const float cb = fmaf(slope2, (float)(ci * 32), base_m);
              float2 bias2 = make_float2(cb, cb + slope2);
              if (diag && g0 + ci == quad) {
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
            store_p(ci, pk);
          }
        }
        l += ps2.x + ps2.y;
        tmem_st_wait();     // P (and the rare accumulator rescale)
        tc_fence_before();  // ... and the S reads, before the MMAs that follow the arrival
        mbar_arrive(p_full(s));
        if (dbg) p.dbg[di + 4] = clock64();
      }
      // epilogue: O / l -> 16-bit -> out[(seq*T + q), head*64 + 32*ch .. +32). o_final also says that every thread of
      // the slot has arrived on the last p_full, i.e. has read the last tile's exchange slots.
      mbar_wait(o_final(s), oc_cnt & 1u);
      ++oc_cnt;
      tc_fence_after();
      *x_own = l;
      slot_bar();
      l += *x_oth;
      const float inv = 1.0f / l;
      const int q = it.qi * 128 + row;
      __nv_bfloat16* dst = p.out + ((long long)it.seq * p.T + q) * kDim + head * 64 + ch * 32;
      {
        uint32_t r[32];
        tmem_ld32(t_o, r);
        tmem_ld_wait();
        tc_fence_before();  // the O reads are ordered before the next item's p_full arrivals
        slot_bar();         // ... and the row sums have been read before the next item's first exchange
        if (q < p.T) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack16(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv, FP16);
            u.y = pack16(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv, FP16);
            u.z = pack16(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv, FP16);
            u.w = pack16(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv, FP16);
            *reinterpret_cast<uint4*>(dst + 8 * i) = u;
          }
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
  if (!mk(&p.tq, q, q_row_stride, 128) || !mk(&p.tk, k, kv_row_stride, 128) || !mk(&p.tv, v, kv_row_stride, 128)) return -1;
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
