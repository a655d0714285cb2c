// FP32 (parity mode) fused causal ALiBi attention, flash-style on CUDA cores.
// Reference: vap/modules.py:82-110 (scores, softmax, PV), :169-202 (the bias that
// is actually added: 1.0 + m_h * j on allowed positions, -inf above the diagonal;
// SURVEY.md F8), :52 (scale = 1/sqrt(dim) = 1/16, not 1/sqrt(head_dim); F7).
// Self- and cross-attention share this kernel; cross-attention reads K/V rows of
// the other speaker channel's sequence (vap/modules.py:287-289).
//
// One CTA = 64 queries of one (sequence, head); K/V stream through shared memory
// in 64-key tiles up to the diagonal; the T x T score matrix never exists.
#include "common.cuh"

namespace vapb {

constexpr int AQ = 64, AK = 64, AD = 64, ALD = 68, ATHREADS = 256;
constexpr int ATT_SMEM = (AQ * ALD + AK * ALD + AK * AD + AQ * ALD) * 4;

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename TIO>
__global__ void __launch_bounds__(ATHREADS)
attention_f32_kernel(const TIO* __restrict__ q, long long q_row_stride, const TIO* __restrict__ k,
                     const TIO* __restrict__ v, long long kv_row_stride, TIO* __restrict__ out, int nseq,
                     int T, const float* __restrict__ slopes, int cross) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;               // [AQ][ALD]
  float* Ks = Qs + AQ * ALD;      // [AK][ALD]
  float* Vs = Ks + AK * ALD;      // [AK][AD]
  float* Ps = Vs + AK * AD;       // [AQ][ALD]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int kvseq = cross ? (seq + nseq / 2) % nseq : seq;
  const float slope = slopes[head];
  const TIO* qb = q + ((long long)seq * T) * q_row_stride + head * AD;
  const TIO* kb = k + ((long long)kvseq * T) * kv_row_stride + head * AD;
  const TIO* vb = v + ((long long)kvseq * T) * kv_row_stride + head * AD;
  const int q0 = qt * AQ;

  for (int idx = tid; idx < AQ * AD / 4; idx += ATHREADS) {
    const int r = idx >> 4, d4 = (idx & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) x = load4(qb + (long long)(q0 + r) * q_row_stride + d4);
    *reinterpret_cast<float4*>(&Qs[r * ALD + d4]) = x;
  }

  float o[4][4], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY;
    lrow[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[i][c] = 0.f;
  }

  for (int kt = 0; kt <= qt; ++kt) {
    const int k0 = kt * AK;
    __syncthreads();  // previous tile's readers are done (also orders the Q stores)
    for (int idx = tid; idx < AK * AD / 4; idx += ATHREADS) {
      const int r = idx >> 4, d4 = (idx & 15) * 4;
      float4 kx = make_float4(0.f, 0.f, 0.f, 0.f), vx = kx;
      if (k0 + r < T) {
        kx = load4(kb + (long long)(k0 + r) * kv_row_stride + d4);
        vx = load4(vb + (long long)(k0 + r) * kv_row_stride + d4);
      }
      *reinterpret_cast<float4*>(&Ks[r * ALD + d4]) = kx;
      *reinterpret_cast<float4*>(&Vs[r * AD + d4]) = vx;
    }
    __syncthreads();

    // S[i][j]: query ty*4+i, key tx+16*j
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < AD; d4 += 4) {
      float4 qf[4], kf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qf[i] = *reinterpret_cast<const float4*>(&Qs[(ty * 4 + i) * ALD + d4]);
#pragma unroll
      for (int j = 0; j < 4; ++j) kf[j] = *reinterpret_cast<const float4*>(&Ks[(tx + 16 * j) * ALD + d4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qf[i].x, kf[j].x, s[i][j]);
          s[i][j] = fmaf(qf[i].y, kf[j].y, s[i][j]);
          s[i][j] = fmaf(qf[i].z, kf[j].z, s[i][j]);
          s[i][j] = fmaf(qf[i].w, kf[j].w, s[i][j]);
        }
    }
    // scale, ALiBi bias (1.0 + m*j, the reference's form), causal mask, online softmax
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx + 16 * j;
        const float bias = __fadd_rn(__fmul_rn(slope, (float)kj), 1.0f);  // (m*j) + 1, no FMA contraction
        float x = s[i][j] * 0.0625f + bias;
        if (kj > qi || kj >= T) x = -INFINITY;
        s[i][j] = x;
        mx = fmaxf(mx, x);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(mrow[i], mx);  // finite: key 0 is visible to every query in tile 0
      const float alpha = expf(mrow[i] - mnew);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pij = expf(s[i][j] - mnew);
        ps += pij;
        Ps[(ty * 4 + i) * ALD + tx + 16 * j] = pij;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      lrow[i] = lrow[i] * alpha + ps;
      mrow[i] = mnew;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[i][c] *= alpha;
    }
    __syncthreads();
    // O[i][c] += sum_key P[q][key] * V[key][tx*4+c]
#pragma unroll 4
    for (int k4 = 0; k4 < AK; k4 += 4) {
      float4 pf[4], vf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pf[i] = *reinterpret_cast<const float4*>(&Ps[(ty * 4 + i) * ALD + k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) vf[kk] = *reinterpret_cast<const float4*>(&Vs[(k4 + kk) * AD + tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pv[4] = {pf[i].x, pf[i].y, pf[i].z, pf[i].w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          o[i][0] = fmaf(pv[kk], vf[kk].x, o[i][0]);
          o[i][1] = fmaf(pv[kk], vf[kk].y, o[i][1]);
          o[i][2] = fmaf(pv[kk], vf[kk].z, o[i][2]);
          o[i][3] = fmaf(pv[kk], vf[kk].w, o[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi < T) {
      const float inv = 1.0f / lrow[i];
      store4(out + ((long long)seq * T + qi) * kDim + head * AD + tx * 4,
             make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv));
    }
  }
}

// Attention MAPS for VapGPT.forward(attention=True) (vap/model.py:262-266; MultiHeadAttentionAlibi returns
// `(y, att)`, vap/modules.py:82-110): softmax(q k^T / 16 + 1 + m_h j, -inf above the diagonal) written out as
// fp32 (T, T) per (item, channel, layer, head). A diagnostic output (B x 2 x L x H x T x T floats), so the kernel
// is the plain three-pass form of the fused one above: row max, row sum with that max, then the normalised
// weights; same thread mapping, arithmetic order of the scores identical to attention_f32_kernel.
// maps: [batch][2][n_layers][H][T][T]; sequence seq = c * batch + b goes to item b, channel c.
__global__ void __launch_bounds__(ATHREADS)
attention_map_f32_kernel(const float* __restrict__ q, long long q_row_stride, const float* __restrict__ k,
                         long long kv_row_stride, int nseq, int T, const float* __restrict__ slopes, int cross,
                         float* __restrict__ maps, int batch, int n_layers, int layer) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;           // [AQ][ALD]
  float* Ks = Qs + AQ * ALD;  // [AK][ALD]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z, H = gridDim.y;
  const int kvseq = cross ? (seq + nseq / 2) % nseq : seq;
  const float slope = slopes[head];
  const float* qb = q + ((long long)seq * T) * q_row_stride + head * AD;
  const float* kb = k + ((long long)kvseq * T) * kv_row_stride + head * AD;
  const int q0 = qt * AQ;
  float* out = maps + ((((long long)(seq % batch) * 2 + seq / batch) * n_layers + layer) * H + head) * T * T;
  for (int idx = tid; idx < AQ * AD / 4; idx += ATHREADS) {
    const int r = idx >> 4, d4 = (idx & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) x = load4(qb + (long long)(q0 + r) * q_row_stride + d4);
    *reinterpret_cast<float4*>(&Qs[r * ALD + d4]) = x;
  }
  float mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { mrow[i] = -INFINITY; lrow[i] = 0.f; }
  for (int pass = 0; pass < 3; ++pass) {
    for (int kt = 0; kt <= qt; ++kt) {
      const int k0 = kt * AK;
      __syncthreads();
      for (int idx = tid; idx < AK * AD / 4; idx += ATHREADS) {
        const int r = idx >> 4, d4 = (idx & 15) * 4;
        float4 kx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + r < T) kx = load4(kb + (long long)(k0 + r) * kv_row_stride + d4);
        *reinterpret_cast<float4*>(&Ks[r * ALD + d4]) = kx;
      }
      __syncthreads();
      float s[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
      for (int d4 = 0; d4 < AD; d4 += 4) {
        float4 qf[4], kf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) qf[i] = *reinterpret_cast<const float4*>(&Qs[(ty * 4 + i) * ALD + d4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) kf[j] = *reinterpret_cast<const float4*>(&Ks[(tx + 16 * j) * ALD + d4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s[i][j] = fmaf(qf[i].x, kf[j].x, s[i][j]);
            s[i][j] = fmaf(qf[i].y, kf[j].y, s[i][j]);
            s[i][j] = fmaf(qf[i].z, kf[j].z, s[i][j]);
            s[i][j] = fmaf(qf[i].w, kf[j].w, s[i][j]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qi = q0 + ty * 4 + i;
        float acc = pass == 0 ? -INFINITY : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kj = k0 + tx + 16 * j;
          const float bias = __fadd_rn(__fmul_rn(slope, (float)kj), 1.0f);  // (m*j) + 1, no FMA contraction
          float x = s[i][j] * 0.0625f + bias;
          if (kj > qi || kj >= T) x = -INFINITY;
          if (pass == 0) {
            acc = fmaxf(acc, x);
          } else {
            const float e = expf(x - mrow[i]);
            if (pass == 1) acc += e;
            else if (qi < T && kj < T) out[(long long)qi * T + kj] = e / lrow[i];
          }
        }
        if (pass == 0) {
#pragma unroll
          for (int off = 8; off > 0; off >>= 1) acc = fmaxf(acc, __shfl_xor_sync(0xffffffffu, acc, off));
          mrow[i] = fmaxf(mrow[i], acc);
        } else if (pass == 1) {
#pragma unroll
          for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
          lrow[i] += acc;
        }
      }
    }
  }
  // keys beyond this query tile's last key tile are all above the diagonal
  const int kz = (qt + 1) * AK;
  if (kz < T) {
    const int w = T - kz;
    for (int idx = tid; idx < AQ * w; idx += ATHREADS) {
      const int r = idx / w, c = idx % w;
      if (q0 + r < T) out[(long long)(q0 + r) * T + kz + c] = 0.f;
    }
  }
}

int launch_attention_map_f32(cudaStream_t st, const float* q, long long q_row_stride, const float* k,
                             long long kv_row_stride, int nseq, int T, int n_heads, const float* slopes, int cross,
                             float* maps, int batch, int n_layers, int layer) {
  const int smem = 2 * AQ * ALD * 4;
  dim3 grid((unsigned)((T + AQ - 1) / AQ), (unsigned)n_heads, (unsigned)nseq);
  attention_map_f32_kernel<<<grid, ATHREADS, smem, st>>>(q, q_row_stride, k, kv_row_stride, nseq, T, slopes, cross,
                                                         maps, batch, n_layers, layer);
  return 1;
}

template <typename TIO>
static int launch_attn(cudaStream_t st, const TIO* q, long long q_row_stride, const TIO* k, const TIO* v,
                       long long kv_row_stride, TIO* out, int nseq, int T, int n_heads, const float* slopes,
                       int cross) {
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    cudaFuncSetAttribute(attention_f32_kernel<TIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    configured = true;
  }
  dim3 grid((unsigned)((T + AQ - 1) / AQ), (unsigned)n_heads, (unsigned)nseq);
  attention_f32_kernel<TIO><<<grid, ATHREADS, ATT_SMEM, st>>>(q, q_row_stride, k, v, kv_row_stride, out, nseq, T,
                                                              slopes, cross);
  return 1;
}

int launch_attention_f32(cudaStream_t st, const float* q, long long q_row_stride, const float* k, const float* v,
                         long long kv_row_stride, float* out, int nseq, int T, int n_heads, const float* slopes,
                         int cross) {
  return launch_attn<float>(st, q, q_row_stride, k, v, kv_row_stride, out, nseq, T, n_heads, slopes, cross);
}

// bf16 in / bf16 out, fp32 arithmetic (interim attention of the bf16 path)
int launch_attention_simt_bf16(cudaStream_t st, const __nv_bfloat16* q, long long q_row_stride,
                               const __nv_bfloat16* k, const __nv_bfloat16* v, long long kv_row_stride,
                               __nv_bfloat16* out, int nseq, int T, int n_heads, const float* slopes, int cross) {
  return launch_attn<__nv_bfloat16>(st, q, q_row_stride, k, v, kv_row_stride, out, nseq, T, n_heads, slopes, cross);
}

}  // namespace vapb
