// FP32 (parity mode) recurrence of CPC's autoregressive net gAR.
// Reference: vap/encoder_components.py:140-159 -> nn.LSTM / nn.GRU (256 -> 256,
// batch_first, zero initial state). PyTorch gate order: LSTM i,f,g,o;
// GRU r,z,n with n = tanh(W_in x + b_in + r * (W_hn h + b_hn)).
//
// The input projection x W_ih^T + b_ih (+ b_hh, except GRU's b_hn) is hoisted into
// one GEMM over all time steps (xproj). This kernel runs the T dependent steps:
// a CTA owns NB sequences, thread j owns hidden unit j (all of its gates, so the
// cell update is thread-local); h_t lives in shared memory, W_hh^T ([256][G*256],
// unit-contiguous so a warp's loads coalesce) streams from L2 every step.
#include "common.cuh"

namespace vapb {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int KIND /*0 LSTM, 1 GRU*/, int NB, typename TOut>
__global__ void __launch_bounds__(256)
rnn_f32_kernel(const float* __restrict__ xproj, const float* __restrict__ whh_t, const float* __restrict__ bhn,
               TOut* __restrict__ out, long long out_seq_stride, int nseq, int T) {
  constexpr int G = KIND == 0 ? 4 : 3;
  constexpr int GH = G * kDim;
  __shared__ __align__(16) float hs[2][kDim][NB];
  const int j = threadIdx.x;
  const int seq0 = blockIdx.x * NB;
  float c[NB], hprev[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    c[n] = 0.f;
    hprev[n] = 0.f;
    hs[0][j][n] = 0.f;
  }
  const float b_hn = (KIND == 1) ? bhn[j] : 0.f;
  __syncthreads();
  int cur = 0;
  for (int t = 0; t < T; ++t) {
    float xp[G][NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const int s = min(seq0 + n, nseq - 1);
      const float* x = xproj + ((long long)s * T + t) * GH + j;
#pragma unroll
      for (int g = 0; g < G; ++g) xp[g][n] = __ldg(x + g * kDim);
    }
    float acc[G][NB];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[g][n] = 0.f;
    const float* w = whh_t + j;
#pragma unroll 4
    for (int k = 0; k < kDim; ++k) {
      float wv[G];
#pragma unroll
      for (int g = 0; g < G; ++g) wv[g] = __ldg(w + (long long)k * GH + g * kDim);
      float hv[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) hv[n] = hs[cur][k][n];
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[g][n] = fmaf(wv[g], hv[n], acc[g][n]);
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      float h;
      if (KIND == 0) {
        const float ig = sigmoidf_(xp[0][n] + acc[0][n]);
        const float fg = sigmoidf_(xp[1][n] + acc[1][n]);
        const float gg = tanhf(xp[2][n] + acc[2][n]);
        const float og = sigmoidf_(xp[3][n] + acc[3][n]);
        c[n] = fg * c[n] + ig * gg;
        h = og * tanhf(c[n]);
      } else {
        const float r = sigmoidf_(xp[0][n] + acc[0][n]);
        const float z = sigmoidf_(xp[1][n] + acc[1][n]);
        const float nn = tanhf(xp[2][n] + r * (acc[2][n] + b_hn));
        h = (1.0f - z) * nn + z * hprev[n];
      }
      hprev[n] = h;
      hs[cur ^ 1][j][n] = h;
      if (seq0 + n < nseq) put(out + (long long)(seq0 + n) * out_seq_stride + (long long)t * kDim + j, h);
    }
    __syncthreads();
    cur ^= 1;
  }
}

template <int KIND, typename TOut>
static int launch_kind(cudaStream_t st, const float* xproj, const float* whh_t, const float* bhn, TOut* out,
                       long long out_seq_stride, int nseq, int T, int n_sm) {
  // sequences per CTA: as few as keeps every SM busy, but every CTA streams all of W_hh (1 MB) from L2 each step and
  // the kernel is bound by that aggregate L2 traffic, so large batches trade CTAs for reuse (8 sequences per CTA)
  int nb = 1;
  while (nb < 4 && (nseq + nb - 1) / nb > n_sm) nb *= 2;
  if (nb == 4 && nseq >= 8 * 48) nb = 8;
  const unsigned grid = (unsigned)((nseq + nb - 1) / nb);
  if (nb == 1)
    rnn_f32_kernel<KIND, 1, TOut><<<grid, 256, 0, st>>>(xproj, whh_t, bhn, out, out_seq_stride, nseq, T);
  else if (nb == 2)
    rnn_f32_kernel<KIND, 2, TOut><<<grid, 256, 0, st>>>(xproj, whh_t, bhn, out, out_seq_stride, nseq, T);
  else if (nb == 4)
    rnn_f32_kernel<KIND, 4, TOut><<<grid, 256, 0, st>>>(xproj, whh_t, bhn, out, out_seq_stride, nseq, T);
  else
    rnn_f32_kernel<KIND, 8, TOut><<<grid, 256, 0, st>>>(xproj, whh_t, bhn, out, out_seq_stride, nseq, T);
  return 1;
}

int launch_rnn_f32(cudaStream_t st, int kind, const float* xproj, const float* whh_t, const float* bhn, float* out,
                   long long out_seq_stride, int nseq, int T) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return kind == 0 ? launch_kind<0, float>(st, xproj, whh_t, bhn, out, out_seq_stride, nseq, T, n_sm)
                   : launch_kind<1, float>(st, xproj, whh_t, bhn, out, out_seq_stride, nseq, T, n_sm);
}

// fp32 recurrence, bf16 output (interim recurrence of the bf16 path)
int launch_rnn_f32_bf16out(cudaStream_t st, int kind, const float* xproj, const float* whh_t, const float* bhn,
                           __nv_bfloat16* out, long long out_seq_stride, int nseq, int T) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return kind == 0 ? launch_kind<0, __nv_bfloat16>(st, xproj, whh_t, bhn, out, out_seq_stride, nseq, T, n_sm)
                   : launch_kind<1, __nv_bfloat16>(st, xproj, whh_t, bhn, out, out_seq_stride, nseq, T, n_sm);
}

}  // namespace vapb
