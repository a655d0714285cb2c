// BF16 fused causal ALiBi attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
// Reference: vap/modules.py:82-110 (scores, softmax, PV), :169-202 (bias 1 + m_h*j on
// allowed positions, -inf above the diagonal), :52 (scale 1/sqrt(dim) = 1/16).
// Self- and cross-attention share the kernel; cross-attention reads K/V rows of the
// other speaker channel's sequence (vap/modules.py:287-289).
//
// Work item = (sequence, head pair, 128-query tile). One persistent CTA per SM walks the
// items from the longest (last query tile) to the shortest. The two heads of a pair are two
// independent "slots" that ping-pong: while the softmax warps of one slot are on the CUDA
// cores / MUFU, the other slot's QK^T and PV run on the tensor pipe.
//
// Per slot and 128-key tile (keys are visited from the diagonal tile DOWN to key 0: the
// ALiBi bias grows with the key index, so the running row maximum is almost always set by
// the first tile and the accumulator rescale below is rare):
//   S[128 q][128 k]  = Q K^T      tcgen05.mma, Q and K tiles K-major SW128 in smem (TMA)
//   softmax          : one thread per query row (TMEM lane); two passes over S with
//                      tcgen05.ld (row max, then exp2 / row sum); P is written back to
//                      TMEM as packed bf16 pairs
//   O[128 q][64 d]  += P V        tcgen05.mma, A = P from TMEM, B = V tile MN-major SW128
// The T x T score matrix never exists; O stays in TMEM for the whole row of tiles.
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int AT_TILE = 128 * 64 * 2;  // one 128-row x 64-column bf16 tile (Q, K or V of one head)
// warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle; warps 4-11 softmax of slot 0, 12-19 softmax of slot 1. A softmax thread
// owns HALF a query row (64 of the 128 keys of a tile; warps 4 apart share a TMEM lane quadrant), so the
// dependent chain MMA -> softmax -> MMA of a slot is half as long and 16 warps hide each other's TMEM latency.
constexpr int AT_THREADS = 640;
constexpr int AT_OFF_Q = 0;                        // [slot][2 buffers]
constexpr int AT_OFF_K = 4 * AT_TILE;              // [slot][stage]
constexpr int AT_OFF_V = AT_OFF_K + 4 * AT_TILE;   // [slot][stage]
constexpr int AT_OFF_X = AT_OFF_V + 4 * AT_TILE;   // row max / row sum exchange: float [slot][parity][half][128]
constexpr int AT_OFF_BAR = AT_OFF_X + 2 * 2 * 2 * 128 * 4;
constexpr int AT_SMEM = AT_OFF_BAR + 256 + 1024 /*alignment slack*/;
// TMEM columns of a slot: S fp32 [0,128), P bf16x2 [128,192), O fp32 [192,256)
constexpr int AT_COL_P = 128, AT_COL_O = 192, AT_SLOT_COLS = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct alignas(64) AttnParams {
  CUtensorMap tq, tk, tv;  // (256 head*d, T, nseq) bf16, SW128, box (64, 128, 1)
  __nv_bfloat16* out;      // (nseq*T, 256)
  const float* slopes;     // [n_heads]
  int nseq, T, nqt, head_pairs, n_items, cross;
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
  int qi, seq, hp;
};
// k-th work item of this CTA; false when the CTA is done
__device__ __forceinline__ bool next_item(const AttnParams& p, int k, Item* it) {
  const int per_q = p.nseq * p.head_pairs;
  if (p.pair_major) {
    const int pi = blockIdx.x + (k / p.nqt) * gridDim.x;
    if (pi >= per_q) return false;
    it->qi = p.nqt - 1 - k % p.nqt;
    it->seq = pi / p.head_pairs;
    it->hp = pi % p.head_pairs;
    return true;
  }
  const int item = blockIdx.x + k * gridDim.x;
  if (item >= p.n_items) return false;
  it->qi = p.nqt - 1 - item / per_q;
  const int rem = item % per_q;
  it->seq = rem / p.head_pairs;
  it->hp = rem % p.head_pairs;
  return true;
}

// FP16: 16-bit format of Q / K / V / P / out (compile time: a run-time flag made every pack two predicated instructions)
template <int FP16>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + AT_OFF_BAR;
  auto q_full = [&](int s, int b) { return bar_base + 8u * (s * 2 + b); };
  auto q_empty = [&](int s, int b) { return bar_base + 8u * (4 + s * 2 + b); };
  auto kv_full = [&](int s, int st) { return bar_base + 8u * (8 + s * 2 + st); };
  auto kv_empty = [&](int s, int st) { return bar_base + 8u * (12 + s * 2 + st); };
  auto s_full = [&](int s) { return bar_base + 8u * (16 + s); };
  auto p_full = [&](int s) { return bar_base + 8u * (18 + s); };
  auto o_final = [&](int s) { return bar_base + 8u * (20 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 22;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tq);
    prefetch_tmap(&p.tk);
    prefetch_tmap(&p.tv);
    for (int s = 0; s < 2; ++s) {
      for (int st = 0; st < 2; ++st) {
        mbar_init(q_full(s, st), 1);
        mbar_init(q_empty(s, st), 1);
        mbar_init(kv_full(s, st), 1);
        mbar_init(kv_empty(s, st), 1);
      }
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 256);
      mbar_init(o_final(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer. The Q tile of the NEXT item is requested before the K/V tiles of the current one
    // (two Q buffers per slot), so an item boundary costs no load latency.
    if (lane == 0) {
      uint32_t kvc[2] = {0, 0};
      auto load_q = [&](const Item& it, uint32_t n) {
        const int b = n & 1;
        for (int s = 0; s < 2; ++s) {
          mbar_wait(q_empty(s, b), ((n >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(q_full(s, b), AT_TILE);
          tma_load_3d(smem_base + AT_OFF_Q + (s * 2 + b) * AT_TILE, &p.tq, q_full(s, b), (it.hp * 2 + s) * 64,
                      it.qi * 128, it.seq);
        }
      };
      Item it, nx;
      bool have = next_item(p, 0, &it);
      if (have) load_q(it, 0);
      for (uint32_t k = 0; have; ++k) {
        const bool have_next = next_item(p, (int)k + 1, &nx);
        if (have_next) load_q(nx, k + 1);
        const int kvseq = p.cross ? (it.seq + p.nseq / 2) % p.nseq : it.seq;
        for (int kt = it.qi; kt >= 0; --kt) {
          for (int s = 0; s < 2; ++s) {
            const int st = kvc[s] & 1;
            mbar_wait(kv_empty(s, st), ((kvc[s] >> 1) & 1u) ^ 1u);
            ++kvc[s];
            mbar_arrive_expect_tx(kv_full(s, st), 2 * AT_TILE);
            const int col = (it.hp * 2 + s) * 64;
            tma_load_3d(smem_base + AT_OFF_K + (s * 2 + st) * AT_TILE, &p.tk, kv_full(s, st), col, kt * 128, kvseq);
            tma_load_3d(smem_base + AT_OFF_V + (s * 2 + st) * AT_TILE, &p.tv, kv_full(s, st), col, kt * 128, kvseq);
          }
        }
        it = nx;
        have = have_next;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_16(128, 128, 0, 0, FP16);
      const uint32_t idesc_pv = make_idesc_16(128, 64, 0, 1, FP16);  // B = V is MN-major (d contiguous)
      uint32_t n_item = 0, kvc[2] = {0, 0}, pc[2] = {0, 0};
      auto issue_qk = [&](int s) {
        const int st = kvc[s] & 1;
        mbar_wait(kv_full(s, st), (kvc[s] >> 1) & 1u);
        tc_fence_after();
        const uint32_t qa = smem_base + AT_OFF_Q + (s * 2 + (n_item & 1)) * AT_TILE;
        const uint32_t ka = smem_base + AT_OFF_K + (s * 2 + st) * AT_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + s * AT_SLOT_COLS, make_smem_desc_sw128(qa + k * 32, 0, 1024),
                    make_smem_desc_sw128(ka + k * 32, 0, 1024), idesc_qk, k != 0);
        umma_commit(s_full(s));
      };
      Item it;
      for (int k = 0; next_item(p, k, &it); ++k, ++n_item) {
        const int ntiles = it.qi + 1;
        for (int s = 0; s < 2; ++s) {
          mbar_wait(q_full(s, n_item & 1), (n_item >> 1) & 1u);
          issue_qk(s);
        }
        for (int n = 0; n < ntiles; ++n) {
          for (int s = 0; s < 2; ++s) {
            const bool dbg = p.dbg && blockIdx.x == 0 && s == 0 && pc[0] >= 40 && pc[0] < 104;
            if (dbg) p.dbg[(pc[0] - 40) * 8 + 5] = clock64();
            mbar_wait(p_full(s), pc[s] & 1u);
            if (dbg) p.dbg[(pc[0] - 40) * 8 + 6] = clock64();
            ++pc[s];
            tc_fence_after();
            const int st = kvc[s] & 1;
            const uint32_t va = smem_base + AT_OFF_V + (s * 2 + st) * AT_TILE;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_ts(tmem_base + s * AT_SLOT_COLS + AT_COL_O, tmem_base + s * AT_SLOT_COLS + AT_COL_P + kk * 8,
                           make_smem_desc_sw128(va + kk * 2048, 1024, 1024), idesc_pv, (n | kk) != 0);
            umma_commit(kv_empty(s, st));
            ++kvc[s];
            if (dbg) p.dbg[(pc[0] - 41) * 8 + 7] = clock64();
            if (n + 1 < ntiles) {
              issue_qk(s);
            } else {
              umma_commit(q_empty(s, n_item & 1));
              umma_commit(o_final(s));
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== softmax: slot = head of the pair; thread = (query row, half of the tile's keys)
    const int s = (warp - 4) >> 3, quad = warp & 3, ch = ((warp - 4) >> 2) & 1;
    const int row = quad * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + s * AT_SLOT_COLS;
    const uint32_t t_s = t_row + ch * 64, t_p = t_row + AT_COL_P + ch * 32, t_o = t_row + AT_COL_O + ch * 32;
    float* xch = reinterpret_cast<float*>(smem_gen + AT_OFF_X) + s * 512;  // [parity][half][128]
    auto slot_bar = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + s) : "memory"); };
    constexpr float SC = 0.0625f * kLog2e;
    uint32_t sc_cnt = 0, oc_cnt = 0, xpar = 0;
    Item it;
    for (int k = 0; next_item(p, k, &it); ++k) {
      const int head = it.hp * 2 + s;
      const float slope2 = p.slopes[head] * kLog2e;
      float m = -INFINITY, l = 0.f;  // l: this thread's partial row sum
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
        ++sc_cnt;
        tc_fence_after();
        // pass A: upper bound of the row maximum (log2 domain): SC * max_j s_j + bias of the last visible key
        float mx = -INFINITY;
#pragma unroll 1
        for (int ci = 0; ci < nvis; ++ci) {
          uint32_t r[32];
          tmem_ld32(t_s + ci * 32, r);
          tmem_ld_wait();
          if (diag && g0 + ci == quad) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i <= lane ? __uint_as_float(r[i]) : -INFINITY);
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
          }
        }
        if (dbg) p.dbg[di + 2] = clock64();
        xch[xpar * 256 + ch * 128 + row] = mx;
        slot_bar();
        if (dbg) p.dbg[di + 3] = clock64();
        mx = fmaxf(mx, xch[xpar * 256 + (ch ^ 1) * 128 + row]);
        xpar ^= 1;
        const float b_last = fmaf(slope2, (float)(k0 + (diag ? row : 127)), kLog2e);
        const float m_new = fmaxf(m, fmaf(mx, SC, b_last));
        if (n > 0 && __any_sync(0xffffffffu, m_new > m)) {
          // rare: rescale this thread's half of the accumulator row (PV of the previous tile has completed:
          // s_full was committed after it)
          const float alpha = ex2_fast(m - m_new);
          l *= alpha;
          uint32_t r[32];
          tmem_ld32(t_o, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(t_o, r);
        }
        // ALiBi makes far tiles of the steep heads irrelevant: if even the tile's upper bound is more than 40 binades
        // below the running maximum for every row of the warp, every p would be < 2^-40 of the row's largest term
        // (fp32 cannot see it in the row sum or in O), so the exponentials are skipped and P is written as zeros.
        const bool negligible = n > 0 && __all_sync(0xffffffffu, fmaf(mx, SC, b_last) < m - 40.0f);
        m = m_new;
        // pass B: p = 2^(t - m), partial row sum, P -> TMEM as bf16 pairs (column c holds keys 2c, 2c+1)
        const float base_m = base - m;
        // packed fp32 arithmetic (FFMA2 / FADD2: one issue slot per two scores): the softmax warps are issue-bound
        float2 ps2 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(SC, SC), step2 = make_float2(2.f * slope2, 2.f * slope2);
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
          tmem_st16(t_p + ci * 16, pk);
        }
        const float ps0 = ps2.x, ps1 = ps2.y;
        l += ps0 + ps1;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(s));
        if (dbg) p.dbg[di + 4] = clock64();
      }
      // epilogue: O / l -> bf16 -> out[(seq*T + q), head*64 + 32*ch .. +32)
      xch[xpar * 256 + ch * 128 + row] = l;
      slot_bar();
      l += xch[xpar * 256 + (ch ^ 1) * 128 + row];
      xpar ^= 1;
      mbar_wait(o_final(s), oc_cnt & 1u);
      ++oc_cnt;
      tc_fence_after();
      const float inv = 1.0f / l;
      const int q = it.qi * 128 + row;
      __nv_bfloat16* dst = p.out + ((long long)it.seq * p.T + q) * kDim + head * 64 + ch * 32;
      {
        uint32_t r[32];
        tmem_ld32(t_o, r);
        tmem_ld_wait();
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
      tc_fence_before();  // the O reads are ordered before the next item's p_full arrivals
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
  if (n_heads % 2 || n_heads * 64 != kDim) {
    if (err) *err = "attention_tc: needs an even number of heads of 64";
    return -1;
  }
  if (cross && (nseq % 2)) {
    if (err) *err = "attention_tc: cross attention needs both channels";
    return -1;
  }
  AttnParams p{};
  const uint32_t box[3] = {64, 128, 1};
  auto mk = [&](CUtensorMap* m, const void* base, long long rs) {
    const uint64_t dims[3] = {(uint64_t)kDim, (uint64_t)T, (uint64_t)nseq};
    const uint64_t strides[2] = {(uint64_t)rs, (uint64_t)rs * (uint64_t)T};
    return make_tmap(m, base, 2, 3, dims, strides, box, 128, err);
  };
  if (!mk(&p.tq, q, q_row_stride) || !mk(&p.tk, k, kv_row_stride) || !mk(&p.tv, v, kv_row_stride)) return -1;
  p.out = out;
  p.slopes = slopes;
  p.nseq = nseq;
  p.T = T;
  p.nqt = (T + 127) / 128;
  p.head_pairs = n_heads / 2;
  p.n_items = p.nqt * nseq * p.head_pairs;
  p.cross = cross;
  p.dbg = dbg;
  p.fp16 = g_fp16;
  p.pair_major = nseq * p.head_pairs >= 2 * n_sm;
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
  const int grid = p.n_items < n_sm ? p.n_items : n_sm;
  if (g_fp16) launch_pdl(attention_tc_kernel<1>, grid, AT_THREADS, AT_SMEM, st, p);
  else launch_pdl(attention_tc_kernel<0>, grid, AT_THREADS, AT_SMEM, st, p);
  return 1;
}

}  // namespace vapb
