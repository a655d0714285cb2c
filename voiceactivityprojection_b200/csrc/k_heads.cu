// Output heads of VapGPT.probs: bandwidth-bound, one warp per frame.
// Reference: vap/model.py:258-260 (va_classifier on x1, x2, concatenated),
// :189-210 (softmax, sigmoid, H = -sum p log2 p, p_now / p_future),
// vap/objective.py:184-204 (codebook marginalisation, p / (sum + 1e-5)),
// :93-110 (class index -> 8 bits, LSB first; bits 0-3 speaker 0 bins 0-3,
// bits 4-7 speaker 1), :53-72,112-139,209-243 (labels from the model's own
// sigmoid(vad), per-frame cross entropy).
#include "common.cuh"

namespace vapb {

// vad_logits[b,t,c] = dot(x[c*B+b, t, :], w) + bias ; vad = sigmoid
__global__ void __launch_bounds__(256)
vad_head_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int batch,
                int T, float* __restrict__ vad_logits, float* __restrict__ vad_sig) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);  // over (2B * T)
  const int lane = threadIdx.x & 31;
  if (row >= (long long)2 * batch * T) return;
  const float* xr = x + row * kDim;
  const float4 a0 = *reinterpret_cast<const float4*>(xr + lane * 4);
  const float4 a1 = *reinterpret_cast<const float4*>(xr + 128 + lane * 4);
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + lane * 4));
  const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + 128 + lane * 4));
  float s = a0.x * w0.x + a0.y * w0.y + a0.z * w0.z + a0.w * w0.w + a1.x * w1.x + a1.y * w1.y + a1.z * w1.z +
            a1.w * w1.w;
  s = warp_sum(s) + bias[0];
  if (lane == 0) {
    const long long seq = row / T, t = row % T;
    const long long c = seq / batch, b = seq % batch;
    const long long o = (b * T + t) * 2 + c;
    if (vad_logits) vad_logits[o] = s;
    if (vad_sig) vad_sig[o] = 1.0f / (1.0f + expf(-s));
  }
}

// Same head on the row-blocked residual stream [row/128][col/4][row%128][4]: one thread per row, so every
// float4 load is contiguous across the warp.
__global__ void __launch_bounds__(128)
vad_head_blocked_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        int batch, int T, float* __restrict__ vad_logits, float* __restrict__ vad_sig) {
  __shared__ float4 ws[64];
  if (threadIdx.x < 64) ws[threadIdx.x] = reinterpret_cast<const float4*>(w)[threadIdx.x];
  __syncthreads();
  const long long row = (long long)blockIdx.x * 128 + threadIdx.x;
  if (row >= (long long)2 * batch * T) return;
  const float4* xr = reinterpret_cast<const float4*>(x) + (long long)blockIdx.x * 64 * 128 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 16
  for (int c = 0; c < 64; ++c) {
    const float4 a = xr[c * 128];
    const float4 b = ws[c];
    s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
  }
  const float s = (s0 + s1) + (s2 + s3) + bias[0];
  const long long seq = row / T, t = row % T;
  const long long c = seq / batch, b = seq % batch;
  const long long o = (b * T + t) * 2 + c;
  if (vad_logits) vad_logits[o] = s;
  if (vad_sig) vad_sig[o] = 1.0f / (1.0f + expf(-s));
}

int launch_vad_head_blocked(cudaStream_t st, const float* x, const float* w, const float* b, int batch, int T,
                            float* vad_logits, float* vad_sig) {
  const long long rows = (long long)2 * batch * T;
  vad_head_blocked_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(x, w, b, batch, T, vad_logits, vad_sig);
  return 1;
}

int launch_vad_head(cudaStream_t st, const float* x, const float* w, const float* b, int batch, int T,
                    float* vad_logits, float* vad_sig) {
  const long long rows = (long long)2 * batch * T;
  vad_head_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, w, b, batch, T, vad_logits, vad_sig);
  return 1;
}

// exp(d), d <= 0, in 7 instructions instead of expf's ~15 (range checks and branches; it was half of
// zero_shot_kernel's instruction stream): ex2.approx(d * log2e) with the rounding error of that product and of the constant
// carried in a first-order correction, so small terms keep ~2 ulp relative accuracy like the reference's exp.
__device__ __forceinline__ float exp_neg(float d) {
  const float L_HI = 1.44269502162933349609375f, L_LO = 1.925963033500011e-8f, LN2 = 0.693147182464599609375f;
  d = fmaxf(d, -120.0f);  // exp underflows to 0 (ftz) long before; keeps a -inf logit from turning t_lo into NaN
  const float t = d * L_HI;
  const float t_lo = fmaf(d, L_LO, fmaf(d, L_HI, -t));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return fmaf(e * t_lo, LN2, e);
}

// One warp handles PR frames at a time (independent shuffle chains interleave); lane owns classes
// lane*4..+3 and 128+lane*4..+3, whose codebook bit counts are fixed per lane and computed once.
constexpr int PR = 4;
// cnt (nullable): the block's shared-memory counters [258] (arg-max class histogram, active frames per channel)
__device__ __forceinline__ void probs_rows(const float* __restrict__ logits, long long rows, int now_lo, int now_hi,
                                           int fut_lo, int fut_hi, float* __restrict__ probs, float* __restrict__ p_now,
                                           float* __restrict__ p_future, float* __restrict__ H, float* __restrict__ lse,
                                           uint8_t* __restrict__ argmax, unsigned int* cnt,
                                           const float* __restrict__ vad_sig) {
  const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * PR;
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  float wn0[8], wn1[8], wf0[8], wf1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cls = (j < 4 ? 0 : 128) + lane * 4 + (j & 3);
    int n0 = 0, n1 = 0, f0 = 0, f1 = 0;
    for (int b = now_lo; b <= now_hi; ++b) { n0 += (cls >> b) & 1; n1 += (cls >> (4 + b)) & 1; }
    for (int b = fut_lo; b <= fut_hi; ++b) { f0 += (cls >> b) & 1; f1 += (cls >> (4 + b)) & 1; }
    wn0[j] = (float)n0; wn1[j] = (float)n1; wf0[j] = (float)f0; wf1[j] = (float)f1;
  }
  float v[PR][8], mx[PR];
  int bidx[PR];
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    const long long row = row0 + r < rows ? row0 + r : rows - 1;  // tail rows recompute the last row, never stored
    const float* lr = logits + row * kClasses;
    const float4 a0 = *reinterpret_cast<const float4*>(lr + lane * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(lr + 128 + lane * 4);
    v[r][0] = a0.x; v[r][1] = a0.y; v[r][2] = a0.z; v[r][3] = a0.w;
    v[r][4] = a1.x; v[r][5] = a1.y; v[r][6] = a1.z; v[r][7] = a1.w;
  }
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    mx[r] = v[r][0];
    int best = 0;
#pragma unroll
    for (int j = 1; j < 8; ++j)
      if (v[r][j] > mx[r]) { mx[r] = v[r][j]; best = j; }
    bidx[r] = (best < 4 ? 0 : 128) + lane * 4 + (best & 3);
  }
  // warp argmax, first index wins ties (torch.argmax semantics on CPU)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const float om = __shfl_xor_sync(0xffffffffu, mx[r], off);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx[r], off);
      if (om > mx[r] || (om == mx[r] && oi < bidx[r])) { mx[r] = om; bidx[r] = oi; }
    }
  }
  float s[PR];
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[r][j] = expf(v[r][j] - mx[r]);  // exp_neg measured no faster here (heads 0.32 ms either way)
      s[r] += v[r][j];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < PR; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], off);
  float acc[PR][5];  // H, p_now 0/1, p_future 0/1 partial sums
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    const float inv = 1.0f / s[r];
    float h = 0.f, pn0 = 0.f, pn1 = 0.f, pf0 = 0.f, pf1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p = v[r][j] * inv;
      v[r][j] = p;
      // lg2.approx (no .ftz: subnormal p handled; abs error ~2^-22 of a value the H tolerance of 1e-4 bits never sees)
      // instead of the ~12-instruction log2f: the kernel was issue-bound, not HBM-bound. p == 0 still gives
      // 0 * -inf = NaN, like the reference (model.py:201).
      h -= p * __log2f(p);
      pn0 = fmaf(p, wn0[j], pn0); pn1 = fmaf(p, wn1[j], pn1);
      pf0 = fmaf(p, wf0[j], pf0); pf1 = fmaf(p, wf1[j], pf1);
    }
    acc[r][0] = h; acc[r][1] = pn0; acc[r][2] = pn1; acc[r][3] = pf0; acc[r][4] = pf1;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < PR; ++r)
#pragma unroll
      for (int q = 0; q < 5; ++q) acc[r][q] += __shfl_xor_sync(0xffffffffu, acc[r][q], off);
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    const long long row = row0 + r;
    if (row >= rows) break;
    if (probs) {
      float* pr = probs + row * kClasses;
      *reinterpret_cast<float4*>(pr + lane * 4) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
      *reinterpret_cast<float4*>(pr + 128 + lane * 4) = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
    }
    if (lane == 0) {
      if (H) H[row] = acc[r][0];
      if (p_now) {
        const float d = (acc[r][1] + acc[r][2]) + 1e-5f;
        p_now[row * 2] = acc[r][1] / d;
        p_now[row * 2 + 1] = acc[r][2] / d;
      }
      if (p_future) {
        const float d = (acc[r][3] + acc[r][4]) + 1e-5f;
        p_future[row * 2] = acc[r][3] / d;
        p_future[row * 2 + 1] = acc[r][4] / d;
      }
      if (lse) lse[row] = mx[r] + logf(s[r]);
      if (argmax) argmax[row] = (uint8_t)bidx[r];
      if (cnt) {
        atomicAdd(&cnt[bidx[r]], 1u);
        const float2 v = *reinterpret_cast<const float2*>(vad_sig + row * 2);
        if (v.x >= 0.5f) atomicAdd(&cnt[256], 1u);
        if (v.y >= 0.5f) atomicAdd(&cnt[257], 1u);
      }
    }
  }
}

// The bulk driver's counters (class histogram and active-frame counts for the all-reduce) are
// taken here, where the arg-max is already in a register: per-block shared-memory counts, one global atomic per
// non-empty bin and block.
__global__ void __launch_bounds__(256)
probs_kernel(const float* __restrict__ logits, long long rows, int now_lo, int now_hi, int fut_lo, int fut_hi,
             float* __restrict__ probs, float* __restrict__ p_now, float* __restrict__ p_future,
             float* __restrict__ H, float* __restrict__ lse, uint8_t* __restrict__ argmax,
             unsigned long long* __restrict__ counters, const float* __restrict__ vad_sig) {
  __shared__ unsigned int cnt[258];
  if (counters) {
    cnt[threadIdx.x] = 0u;
    if (threadIdx.x < 2) cnt[256 + threadIdx.x] = 0u;
    __syncthreads();
  }
  probs_rows(logits, rows, now_lo, now_hi, fut_lo, fut_hi, probs, p_now, p_future, H, lse, argmax,
             counters ? cnt : nullptr, vad_sig);
  if (counters) {
    __syncthreads();
    if (cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
    if (threadIdx.x < 2 && cnt[256 + threadIdx.x])
      atomicAdd(&counters[256 + threadIdx.x], (unsigned long long)cnt[256 + threadIdx.x]);
  }
}

int launch_probs(cudaStream_t st, const float* logits, long long rows, int now_lo, int now_hi, int fut_lo,
                 int fut_hi, float* probs, float* p_now, float* p_future, float* H, float* lse, uint8_t* argmax,
                 unsigned long long* counters, const float* vad_sig) {
  probs_kernel<<<(unsigned)((rows + 8 * PR - 1) / (8 * PR)), 256, 0, st>>>(logits, rows, now_lo, now_hi, fut_lo, fut_hi, probs,
                                                           p_now, p_future, H, lse, argmax, counters, vad_sig);
  return 1;
}

// int16 PCM -> float32 in [-1, 1): what torchaudio.load's normalisation does on the host (vap/audio.py:47), for the
// callers that cannot hand the PCM to the fused encoder kernel directly (fp32 mode, odd lengths)
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const int16_t* __restrict__ pcm, long long n, float* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(pcm + i) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out + i) & 15) == 0)) {
    const uint4 v = *reinterpret_cast<const uint4*>(pcm + i);
    const int16_t* s = reinterpret_cast<const int16_t*>(&v);
    float4 a, b;
    a.x = s[0] * (1.0f / 32768.0f); a.y = s[1] * (1.0f / 32768.0f); a.z = s[2] * (1.0f / 32768.0f); a.w = s[3] * (1.0f / 32768.0f);
    b.x = s[4] * (1.0f / 32768.0f); b.y = s[5] * (1.0f / 32768.0f); b.z = s[6] * (1.0f / 32768.0f); b.w = s[7] * (1.0f / 32768.0f);
    *reinterpret_cast<float4*>(out + i) = a;
    *reinterpret_cast<float4*>(out + i + 4) = b;
  } else {
    for (long long k = i; k < n && k < i + 8; ++k) out[k] = pcm[k] * (1.0f / 32768.0f);
  }
}

int launch_pcm16_to_f32(cudaStream_t st, const int16_t* pcm, long long n, float* out) {
  const long long threads = (n + 7) / 8;
  pcm16_to_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(pcm, n, out);
  return 1;
}

// loss[b,t] = logsumexp(logits[b,t]) - logits[b,t,label], t < T-100, where the
// label packs 8 bits: bit (4*c + bin) = mean(vad[b, t+1+start : t+1+end, c]) >= 0.5
// over bins of 10/20/30/40 frames.
__global__ void __launch_bounds__(128)
loss_kernel(const float* __restrict__ logits, const float* __restrict__ vad, const float* __restrict__ lse,
            int batch, int T, float* __restrict__ loss) {
  const int n = T - 100;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)batch * n) return;
  const long long b = i / n, t = i % n;
  const float* vb = vad + (b * T + t + 1) * 2;
  const int bins[4] = {10, 20, 30, 40};
  int label = 0, start = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float s0 = 0.f, s1 = 0.f;
    for (int f = start; f < start + bins[k]; ++f) {
      const float2 x = *reinterpret_cast<const float2*>(vb + 2 * f);
      s0 += x.x;
      s1 += x.y;
    }
    if (s0 / (float)bins[k] >= 0.5f) label |= 1 << k;
    if (s1 / (float)bins[k] >= 0.5f) label |= 1 << (4 + k);
    start += bins[k];
  }
  loss[i] = lse[b * T + t] - logits[(b * T + t) * kClasses + label];
}

int launch_loss(cudaStream_t st, const float* logits, const float* vad_sig, const float* lse, int batch, int T,
                float* loss) {
  const long long n = (long long)batch * (T - 100);
  if (n <= 0) return 0;
  loss_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(logits, vad_sig, lse, batch, T, loss);
  return 1;
}

// VapGPT.vad()'s run-length filters (vap/model.py:227-247 -> vap/utils.py:239-272): on a binary (batch, T, 2)
// activity tensor, first every silence run of <= max_fill frames becomes active (vad_fill_silences; runs touching
// either end of the sequence included, as find_island_idx_len reports them), then every activity run of <= max_omit
// frames of the RESULT becomes silence (vad_omit_spikes). One thread per (item, channel) sequence: 2 x T sequential
// steps on 8 KB of data per item is microseconds, and the reference's per-run Python loop is what this replaces.
__global__ void __launch_bounds__(128)
vad_filter_kernel(const float* vad, int nseq, int T, int max_fill, int max_omit, float* out, int from_logits,
                  float cutoff) {  // out may alias vad
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseq) return;
  const float* x = vad + (long long)(s >> 1) * T * 2 + (s & 1);
  float* y = out + (long long)(s >> 1) * T * 2 + (s & 1);
  // from_logits: the input holds VAD logits; active = sigmoid(x) >= cutoff (vap/model.py:237-238, the sigmoid of vad_head_kernel)
  auto on = [&](float v) { return from_logits ? (1.0f / (1.0f + expf(-v))) >= cutoff : v != 0.f; };
  // pass 1: fill short silences (reads x, writes y; in place is fine: a run is rewritten after it has been read)
  int run0 = 0;
  bool cur = T > 0 && on(x[0]);
  for (int t = 1; t <= T; ++t) {
    const bool v = t < T ? on(x[2LL * t]) : !cur;
    if (v != cur) {
      const float w = (cur || t - run0 <= max_fill) ? 1.f : 0.f;
      for (int k = run0; k < t; ++k) y[2LL * k] = w;
      run0 = t;
      cur = v;
    }
  }
  // pass 2: drop short activity runs of the filled sequence
  run0 = 0;
  cur = T > 0 && y[0] != 0.f;
  for (int t = 1; t <= T; ++t) {
    const bool v = t < T ? (y[2LL * t] != 0.f) : !cur;
    if (v != cur) {
      if (cur && t - run0 <= max_omit)
        for (int k = run0; k < t; ++k) y[2LL * k] = 0.f;
      run0 = t;
      cur = v;
    }
  }
}

int launch_vad_filter(cudaStream_t st, const float* vad01, int batch, int T, int max_fill, int max_omit, float* out,
                      int from_logits, float cutoff) {
  const int nseq = batch * 2;
  if (nseq <= 0 || T <= 0) return 0;
  vad_filter_kernel<<<(unsigned)((nseq + 127) / 128), 128, 0, st>>>(vad01, nseq, T, max_fill, max_omit, out, from_logits, cutoff);
  return 1;
}

// ZeroShot marginals (vap/zero_shot.py:159-271). Ten class subsets: 0-1 silence pos[next
// speaker 0/1], 2-3 silence neg, 4-5 active pos, 6-7 active neg, 8-9 backchannel. Per frame: probs = softmax(logits)
// (or the input as is), then
//   p_sil[s] = sum(pos_sil[s]) / (sum(pos_sil[s]) + sum(neg_sil[s])), p_act[s] likewise   (marginal_probs :159-165)
//   p_bc[s]  = sum(bc[s])                                                                (probs_backchannel :173-176)
//   dialog state ds = (long)(2 va1 - va0) + 1 (vap/events.py:70-78): 1 silence -> p = p_sil; 0 only A ->
//   (1 - p_act[1], p_act[1]); 3 only B -> (p_act[0], 1 - p_act[0]); 2 both -> p_act / (p_act[0] + p_act[1]);
//   anything else -> 0                                                                 (probs_next_speaker :226-262)
// HBM-bound (1 KB read, 16-32 B written per frame), so the point is to keep the instruction count and the exposed
// latency per frame low: a warp takes ZR = 8 frames per block lifetime.
// Staging: the eight 1 KB rows go global -> shared memory by cp.async (16 B per lane per copy, all 16 in flight at
// once, no registers held). Phase 1 (lane owns classes lane*4..+3 and 128+lane*4..+3, as probs_kernel): row max,
// exp, row sum by shuffles, four frames at a time; the un-normalised exponentials overwrite the tile. Phase 2 (four
// lanes per frame): every subset is a list of class indices (the subsets are sparse: 4 to 56 of 256 classes, 160
// members in all, padded to fours with a zero slot), lane s of a frame's quad adds members s, s+4, ... from the
// tile; two shuffles per subset finish the sums and the quad's first lane does the divisions and the dialog-state
// switch. Row stride 260 floats: frame r / class c sits in bank (4r + c) mod 32, so the float4 accesses and the
// quad-strided phase-2 reads are conflict-free.
// History on a B=256 x T=1000 batch (tools/zero_shot_probe.py): class-per-lane subset sums (80 masked adds + 50
// shuffles per frame) 205 us; tile + member lists 112; compensated ex2 instead of expf (whose range checks and
// branches were half the instruction stream) 96; list copy behind the loads instead of in front of them 88;
// cp.async staging 80 us = 3.3 TB/s. 217 warp instructions per frame put the issue floor at ~49 us, above the 41 us
// HBM floor; higher occupancy (4 frames per warp) and a 3-instruction exp measured no faster.
constexpr int ZR = 8, ZLPR = 32 / ZR, ZSTRIDE = 260, ZWARPS = 4, ZMAXN = 64;  // ZLPR lanes per frame in phase 2
struct ZeroShotLists {
  uint16_t off[10][ZMAXN];  // byte offsets (4 * class) of each subset's members within a tile row, ascending class,
                            // padded to a multiple of 4 entries with 1024 = the row's zero slot
  int n4[10];               // entries / ZLPR = gather iterations per lane
};

#ifndef ZS_MINBLOCKS
#define ZS_MINBLOCKS 6  // shared memory allows six blocks per SM; keep registers within that
#endif
__global__ void __launch_bounds__(ZWARPS * 32, ZS_MINBLOCKS)
zero_shot_kernel(const float* __restrict__ x, int is_probs, long long rows, int T, const float* __restrict__ va,
                 long long va_T, const __grid_constant__ ZeroShotLists lists, float* __restrict__ p_out,
                 float* __restrict__ p_bc, float* __restrict__ p_sil, float* __restrict__ p_act) {
  __shared__ __align__(16) float tile[ZWARPS][ZR * ZSTRIDE];
  __shared__ float inv_sum[ZWARPS][ZR];
  __shared__ uint16_t members[10][ZMAXN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * ZWARPS + warp) * ZR;
  const bool active = row0 < rows;
  float* const mine = tile[warp];
  const int r = lane / ZLPR, s = lane % ZLPR;  // phase 2: frame and position in the frame's quad
  float2 w = make_float2(0.f, 0.f);            // the frame's voice activity, fetched now (a load at the point of
  if (active && p_out && s == 0 && row0 + r < rows) {  // use stalled every warp for a DRAM round trip at its end)
    const long long row = row0 + r;
    w = *reinterpret_cast<const float2*>(va + ((row / T) * va_T + row % T) * 2);
  }
  if (active) {
  if (lane < ZR) *reinterpret_cast<float4*>(mine + lane * ZSTRIDE + 256) = make_float4(0.f, 0.f, 0.f, 0.f);  // pad slot
  // ---- phase 1: exponentials of ZR frames into the tile, two batches of four frames (independent shuffle chains)
  // all ZR rows leave for the tile at once (16 B per lane per copy, no registers held while they fly)
#pragma unroll
  for (int rr = 0; rr < ZR; ++rr) {
    const long long row = row0 + rr < rows ? row0 + rr : rows - 1;  // tail: recomputed, never stored
    const float* lr = x + row * kClasses + lane * 4;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(mine + rr * ZSTRIDE + lane * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(lr) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512u), "l"(lr + 128) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
#pragma unroll
  for (int half = 0; half < ZR / 4; ++half) {
    float v[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* lr = mine + (half * 4 + r) * ZSTRIDE;
      const float4 a0 = *reinterpret_cast<const float4*>(lr + lane * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(lr + 128 + lane * 4);
      v[r][0] = a0.x; v[r][1] = a0.y; v[r][2] = a0.z; v[r][3] = a0.w;
      v[r][4] = a1.x; v[r][5] = a1.y; v[r][6] = a1.z; v[r][7] = a1.w;
    }
    if (!is_probs) {
      float mx[4], s[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        mx[r] = v[r][0];
#pragma unroll
        for (int j = 1; j < 8; ++j) mx[r] = fmaxf(mx[r], v[r][j]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int r = 0; r < 4; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], off));
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        s[r] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[r][j] = exp_neg(v[r][j] - mx[r]);
          s[r] += v[r][j];
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], off);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) inv_sum[warp][half * 4 + r] = 1.0f / s[r];
      }
    } else if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) inv_sum[warp][half * 4 + r] = 1.0f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float* er = mine + (half * 4 + r) * ZSTRIDE;
      *reinterpret_cast<float4*>(er + lane * 4) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
      *reinterpret_cast<float4*>(er + 128 + lane * 4) = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
    }
  }
  }
  // the member lists travel as kernel parameters; their copy to shared memory (per-thread constant-bank reads, slow)
  // sits here, behind the block's global loads, not in front of them
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&lists.off[0][0]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&members[0][0]);
    for (int i = threadIdx.x; i < 10 * ZMAXN / 2; i += ZWARPS * 32) dst[i] = src[i];
  }
  __syncthreads();
  if (!active) return;
  // ---- phase 2: subset sums, four lanes per frame
  const char* er = reinterpret_cast<const char*>(mine + r * ZSTRIDE);
  float a[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    // every list holds at least one group of four (padded), most hold exactly one: straight-line first gather,
    // then a uniform loop for the rest (no divergence: n4 is the same for every lane)
    const int n4 = lists.n4[q];
    float acc = *reinterpret_cast<const float*>(er + members[q][s]);
    for (int i = 1; i < n4; ++i) acc += *reinterpret_cast<const float*>(er + members[q][i * ZLPR + s]);
    a[q] = acc;
  }
#pragma unroll
  for (int q = 0; q < 10; ++q) {
#pragma unroll
    for (int off = 1; off < ZLPR; off <<= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], off);
  }
  const long long row = row0 + r;
  if (s != 0 || row >= rows) return;
  const float inv = inv_sum[warp][r];
  const float sil0 = a[0] / (a[0] + a[2]), sil1 = a[1] / (a[1] + a[3]);
  const float act0 = a[4] / (a[4] + a[6]), act1 = a[5] / (a[5] + a[7]);
  if (p_sil) *reinterpret_cast<float2*>(p_sil + row * 2) = make_float2(sil0, sil1);
  if (p_act) *reinterpret_cast<float2*>(p_act + row * 2) = make_float2(act0, act1);
  if (p_bc) *reinterpret_cast<float2*>(p_bc + row * 2) = make_float2(a[8] * inv, a[9] * inv);
  if (p_out) {
    const long long ds = (long long)(2.0f * w.y - w.x) + 1;
    float pa = 0.f, pb = 0.f;
    if (ds == 1) { pa = sil0; pb = sil1; }
    else if (ds == 0) { pa = 1.0f - act1; pb = act1; }
    else if (ds == 3) { pa = act0; pb = 1.0f - act0; }
    else if (ds == 2) { const float sum = act0 + act1; pa = act0 / sum; pb = act1 / sum; }
    *reinterpret_cast<float2*>(p_out + row * 2) = make_float2(pa, pb);
  }
}

// sets: host [10][8] uint32, ten 256-bit class sets (bit c%32 of word c/32 = class c). Returns launches or -1
// (a subset of more than ZMAXN classes).
int launch_zero_shot(cudaStream_t st, const float* x, int is_probs, long long batch, int T, const float* va,
                     long long va_T, const uint32_t* sets, float* p, float* p_bc, float* p_sil, float* p_act) {
  const long long rows = batch * T;
  if (rows <= 0) return 0;
  ZeroShotLists zl{};
  for (int q = 0; q < 10; ++q) {
    int n = 0;
    for (int c = 0; c < 256; ++c)
      if ((sets[q * 8 + c / 32] >> (c % 32)) & 1u) {
        if (n == ZMAXN) return -1;
        zl.off[q][n++] = (uint16_t)(c * 4);
      }
    while (n % ZLPR || n == 0) zl.off[q][n++] = 1024;
    zl.n4[q] = n / ZLPR;
  }
  const long long tasks = (rows + ZR - 1) / ZR;
  zero_shot_kernel<<<(unsigned)((tasks + ZWARPS - 1) / ZWARPS), ZWARPS * 32, 0, st>>>(x, is_probs, rows, T, va, va_T,
                                                                                      zl, p, p_bc, p_sil, p_act);
  return 1;
}

}  // namespace vapb
