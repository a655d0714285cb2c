// conv0 of the CPC gEncoder fused with ChannelNorm and ReLU, bf16 output (tensor path).
// Reference: vap/encoder_components.py:83-84,99 (Conv1d(1,256,k=10,s=5,p=3)), :62-70
// (ChannelNorm: unbiased variance over the 256 channels of one time step).
//
// The layer is bound by its 512-byte-per-frame output, so the arithmetic is arranged
// to stay under that: the per-frame statistics are closed forms in the frame's 10
// samples (the conv is linear, so mean_f = wbar.x + bbar and
// sum_c (a_c - mean_f)^2 = x'Gx + 2 h.x + s with G, h, s folded on the host), which
// removes both cross-channel reductions; one thread evaluates them per frame. The
// channel pass then is 10 packed fp32 FMAs (fma.rn.f32x2) per channel pair on
// pre-centred, pre-scaled weights u'_c = g_c (w_c - wbar), plus one for
// rstd_f * (.) + beta_c, a bf16x2 pack and a packed ReLU. A lane owns 8
// consecutive channels, so a warp stores one 512-byte frame row per instruction.
#include "common.cuh"

namespace vapb {

namespace {

constexpr int C2_FRAMES = 512;  // frames per CTA
constexpr int C2_THREADS = 256;
constexpr int C2_SAMPLES = 5 * C2_FRAMES + 5;

__global__ void __launch_bounds__(C2_THREADS)
conv0_v2_kernel(const float* __restrict__ wav, int batch, long long n_samples, int seq0, long long L0,
                const float* __restrict__ u /*[10][256]*/, const float* __restrict__ d /*[256]*/,
                const float* __restrict__ beta /*[256]*/, const __grid_constant__ Conv0Stats cs,
                __nv_bfloat16* __restrict__ out, long long out_seq_stride, int out_pad_rows) {
  __shared__ float2 xs2[C2_SAMPLES];  // every sample duplicated: the f32x2 FMA's broadcast operand
  __shared__ float rs[C2_FRAMES];
  const int lseq = blockIdx.y;
  const int seq = seq0 + lseq;  // channel-major sequence id: c * batch + item
  const int ch = seq / batch, item = seq % batch;
  const float* x = wav + ((long long)item * 2 + ch) * n_samples;
  const long long f0 = (long long)blockIdx.x * C2_FRAMES;
  const long long s0 = 5 * f0 - 3;
  for (int i = threadIdx.x; i < C2_SAMPLES; i += C2_THREADS) {
    const long long s = s0 + i;
    const float v = (s >= 0 && s < n_samples) ? __ldg(x + s) : 0.0f;
    xs2[i] = make_float2(v, v);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this lane's 8 channels: pre-centred, pre-scaled taps, offsets and norm bias
  float2 u2[4][10], d2[4], b2[4];
  {
    const int c0 = lane * 8;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(u + k * kDim + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(u + k * kDim + c0 + 4));
      u2[0][k] = make_float2(a.x, a.y); u2[1][k] = make_float2(a.z, a.w);
      u2[2][k] = make_float2(b.x, b.y); u2[3][k] = make_float2(b.z, b.w);
    }
    const float4 da = __ldg(reinterpret_cast<const float4*>(d + c0)), db = __ldg(reinterpret_cast<const float4*>(d + c0 + 4));
    d2[0] = make_float2(da.x, da.y); d2[1] = make_float2(da.z, da.w);
    d2[2] = make_float2(db.x, db.y); d2[3] = make_float2(db.z, db.w);
    const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + c0)), bb = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    b2[0] = make_float2(ba.x, ba.y); b2[1] = make_float2(ba.z, ba.w);
    b2[2] = make_float2(bb.x, bb.y); b2[3] = make_float2(bb.z, bb.w);
  }
  __syncthreads();
  // per-frame 1/sqrt(var + eps) from the closed form
  for (int fi = threadIdx.x; fi < C2_FRAMES; fi += C2_THREADS) {
    float xv[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) xv[k] = xs2[5 * fi + k].x;
    float ss = cs.s;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      float y = cs.h2[k];  // 2 h_k
#pragma unroll
      for (int l = 0; l < 10; ++l) y = fmaf(cs.G[k][l], xv[l], y);
      ss = fmaf(xv[k], y, ss);
    }
    rs[fi] = rsqrtf(fmaxf(ss, 0.f) * (1.0f / (kDim - 1)) + kEps);
  }
  __syncthreads();
  const long long row0 = (long long)lseq * out_seq_stride + (long long)out_pad_rows * kDim + lane * 8;
#pragma unroll 2
  for (int fi = warp; fi < C2_FRAMES; fi += C2_THREADS / 32) {
    const long long f = f0 + fi;
    if (f >= L0) break;
    const float2* xp = xs2 + 5 * fi;
    float2 acc[4] = {d2[0], d2[1], d2[2], d2[3]};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float2 xk = xp[k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = __ffma2_rn(u2[j][k], xk, acc[j]);
    }
    const float r = rs[fi];
    const float2 r2 = make_float2(r, r);
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] = __ffma2_rn(acc[j], r2, b2[j]);
      __nv_bfloat162 p = __hmax2(__floats2bfloat162_rn(acc[j].x, acc[j].y), __floats2bfloat162_rn(0.f, 0.f));
      pk[j] = *reinterpret_cast<uint32_t*>(&p);
    }
    *reinterpret_cast<uint4*>(out + row0 + f * kDim) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

}  // namespace

// Host folding of the conv0 + ChannelNorm parameters (double precision).
// w: (256, 1, 10) conv weight, bias (256), g / beta: ChannelNorm affine (256).
void conv0_v2_fold(const float* w, const float* bias, const float* g, float* u /*[10][256]*/, float* d /*[256]*/,
                   Conv0Stats* cs) {
  double wbar[10] = {0}, bbar = 0;
  for (int c = 0; c < kDim; ++c) {
    for (int k = 0; k < 10; ++k) wbar[k] += w[c * 10 + k];
    bbar += bias[c];
  }
  for (int k = 0; k < 10; ++k) wbar[k] /= kDim;
  bbar /= kDim;
  double G[10][10] = {{0}}, h[10] = {0}, s = 0;
  for (int c = 0; c < kDim; ++c) {
    double uc[10];
    const double dc = bias[c] - bbar;
    for (int k = 0; k < 10; ++k) uc[k] = w[c * 10 + k] - wbar[k];
    for (int k = 0; k < 10; ++k) {
      for (int l = 0; l < 10; ++l) G[k][l] += uc[k] * uc[l];
      h[k] += dc * uc[k];
      u[k * kDim + c] = (float)(g[c] * uc[k]);
    }
    s += dc * dc;
    d[c] = (float)(g[c] * dc);
  }
  for (int k = 0; k < 10; ++k) {
    for (int l = 0; l < 10; ++l) cs->G[k][l] = (float)G[k][l];
    cs->h2[k] = (float)(2.0 * h[k]);
  }
  cs->s = (float)s;
}

int launch_conv0_v2(cudaStream_t st, const float* wav, int batch, long long n_samples, int seq0, int nseq,
                    long long L0, const float* u, const float* d, const float* beta, const Conv0Stats& cs,
                    __nv_bfloat16* out, long long out_seq_stride, int out_pad_rows) {
  dim3 grid((unsigned)((L0 + C2_FRAMES - 1) / C2_FRAMES), (unsigned)nseq);
  conv0_v2_kernel<<<grid, C2_THREADS, 0, st>>>(wav, batch, n_samples, seq0, L0, u, d, beta, cs, out, out_seq_stride,
                                               out_pad_rows);
  return 1;
}

}  // namespace vapb
