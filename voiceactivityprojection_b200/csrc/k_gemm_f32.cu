// FP32 (parity mode) implicit-GEMM on CUDA cores with the fused row epilogues.
//
//   C[m, n] = sum_k A[row(m), k] * Wt[k, n]
//
// A rows are addressed through a RowMap, so a strided Conv1d over a
// channels-last activation (vap/encoder_components.py:85-92,100-103 and the
// causal stride-2 conv of vap/encoder.py:24-30) is the same kernel as a Linear
// (vap/modules.py:93-95,109; ffn :16-21): for a conv with kernel k and stride s
// the row of output frame t is the contiguous span of k*256 elements starting at
// input frame s*t - pad, which exists physically because the activation buffers
// carry zero rows for the padding.
//
// The CTA tile spans BN = 256 = the whole channel dimension, so the norm over
// channels that follows every conv / precedes every transformer op is a
// warp-shuffle reduction in the epilogue (a warp owns 8 full rows):
//   v = acc + bias -> ChannelNorm (unbiased) | LayerNorm -> ReLU | GELU(erf)
//     -> + residual -> (+= previous out) -> out1 ;  out2 = LayerNorm2(out1 value)
#include "common.cuh"

namespace vapb {

constexpr int GM = 64, GN = 256, GK = 16, GT = 256;
constexpr int AS_LD = GM + 4;

__device__ __forceinline__ void row_norm(float (&v)[8], int kind, const float* __restrict__ g,
                                         const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = warp_sum(s) * (1.0f / kDim);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] -= mean;
    q = fmaf(v[j], v[j], q);
  }
  const float var = warp_sum(q) * (kind == NORM_CHANNEL ? 1.0f / (kDim - 1) : 1.0f / kDim);
  const float rstd = 1.0f / sqrtf(var + kEps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + lane * 4));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + 128 + lane * 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + lane * 4));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + 128 + lane * 4));
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] * rstd, gg[j], bb[j]);
}

__global__ void __launch_bounds__(GT, 2) gemm_f32_kernel(GemmProblem p, Epilogue e) {
  __shared__ __align__(16) float As[2][GK][AS_LD];
  __shared__ __align__(16) float Bs[2][GK][GN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const float* __restrict__ A = reinterpret_cast<const float*>(p.A);
  const float* __restrict__ W = reinterpret_cast<const float*>(p.W);

  // this thread's A-tile load slot: row tid/4, k-chunk (tid%4)*4
  const int lrow = tid >> 2, lkc = (tid & 3) * 4;
  const int lm = m0 + lrow;
  const bool lvalid = lm < p.M;
  const float* arow = A;
  if (lvalid) arow = A + (long long)(lm / p.rows_per_seq) * p.a_map.seq_stride +
                     (long long)(lm % p.rows_per_seq) * p.a_map.row_stride + lkc;
  // B-tile slots: idx = tid + 256*j -> k = idx/64, n4 = idx%64
  const float* wbase = W + n0 + (tid & 63) * 4 + (long long)(tid >> 6) * p.N;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb[4];
  auto gload = [&](int k0) {
    if (lvalid) ra = *reinterpret_cast<const float4*>(arow + k0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      rb[j] = __ldg(reinterpret_cast<const float4*>(wbase + (long long)(k0 + 4 * j) * p.N));
  };
  auto sstore = [&](int buf) {
    As[buf][lkc + 0][lrow] = ra.x;
    As[buf][lkc + 1][lrow] = ra.y;
    As[buf][lkc + 2][lrow] = ra.z;
    As[buf][lkc + 3][lrow] = ra.w;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(&Bs[buf][(tid >> 6) + 4 * j][(tid & 63) * 4]) = rb[j];
  };

  const int nk = p.K / GK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][warp * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][warp * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][lane * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][128 + lane * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: a warp owns rows warp*8..+7 across all 256 columns ----------
  const int c0 = n0 + lane * 4, c1 = n0 + 128 + lane * 4;
  float bias[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (e.bias) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(e.bias + c0));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(e.bias + c1));
    bias[0] = x0.x; bias[1] = x0.y; bias[2] = x0.z; bias[3] = x0.w;
    bias[4] = x1.x; bias[5] = x1.y; bias[6] = x1.z; bias[7] = x1.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + warp * 8 + i;
    if (m >= p.M) continue;  // warp-uniform
    const long long seq = m / p.rows_per_seq, t = m % p.rows_per_seq;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j] + bias[j];
    if (e.norm1 != NORM_NONE) row_norm(v, e.norm1, e.g1, e.b1, lane);
    if (e.act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], e.act);
    }
    if (e.resid) {
      const float* r = e.resid + seq * e.resid_map.seq_stride + t * e.resid_map.row_stride;
      const float4 r0 = *reinterpret_cast<const float4*>(r + c0);
      const float4 r1 = *reinterpret_cast<const float4*>(r + c1);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    }
    float* o = reinterpret_cast<float*>(e.out1) + seq * e.out1_map.seq_stride + t * e.out1_map.row_stride;
    if (e.accumulate) {
      const float4 r0 = *reinterpret_cast<const float4*>(o + c0);
      const float4 r1 = *reinterpret_cast<const float4*>(o + c1);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    }
    *reinterpret_cast<float4*>(o + c0) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + c1) = make_float4(v[4], v[5], v[6], v[7]);
    if (e.norm2 != NORM_NONE) {
      row_norm(v, e.norm2, e.g2, e.b2, lane);
      float* o2 = reinterpret_cast<float*>(e.out2) + seq * e.out2_map.seq_stride + t * e.out2_map.row_stride;
      *reinterpret_cast<float4*>(o2 + c0) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o2 + c1) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

int launch_gemm_f32(cudaStream_t st, const GemmProblem& p, const Epilogue& e) {
  dim3 grid((unsigned)((p.M + GM - 1) / GM), (unsigned)(p.N / GN));
  gemm_f32_kernel<<<grid, GT, 0, st>>>(p, e);
  return 1;
}

}  // namespace vapb
