// Shared declarations of the vapb kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace vapb {

constexpr int kDim = 256;       // model dim == CPC hidden (vap/model.py:53)
constexpr int kClasses = 256;   // 2^8 projection-window states (vap/objective.py:84)
constexpr float kEps = 1e-5f;

enum Norm { NORM_NONE = 0, NORM_CHANNEL = 1 /* unbiased var, /255 */, NORM_LAYER = 2 /* biased, /256 */ };
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// Row addressing shared by every GEMM-shaped kernel: logical row m maps to
// (seq, t) = (m / rows_per_seq, m % rows_per_seq) and the row's first element
// lives at base + seq*seq_stride + t*row_stride (element units).
struct RowMap {
  long long seq_stride;
  long long row_stride;
};

// Epilogue of a GEMM whose CTA tile spans the whole N=256 row.
struct Epilogue {
  const float* bias;   // [N] or null
  int norm1;           // Norm applied to (acc + bias)
  const float* g1;     // norm1 affine weight [256]
  const float* b1;     // norm1 affine bias   [256]
  int act;             // Act
  const float* resid;  // fp32 residual rows or null
  RowMap resid_map;
  int accumulate;      // out1 += value (combinator's second branch)
  void* out1;          // fp32 or bf16 (out1_bf16)
  int out1_bf16;
  RowMap out1_map;
  int norm2;           // LayerNorm of the out1 value -> out2 (next op's input)
  const float* g2;
  const float* b2;
  void* out2;
  int out2_bf16;
  RowMap out2_map;
};

struct GemmProblem {
  const void* A;       // fp32 (SIMT path) or bf16 (tensor path), rows addressed by a_map
  RowMap a_map;
  const void* W;       // packed weights, [K][N] fp32 (SIMT path)
  int M;               // logical rows = n_seq * rows_per_seq
  int rows_per_seq;
  int N;
  int K;
};

// Tensor-core GEMMs (k_gemm_lin.cu, k_gemm_2sm.cu): A and W are 16-bit; W is [N][K] (K contiguous).
// The epilogue's out1 goes to out1_f32 and/or out1_bf16 (Epilogue::out1 ignored).
struct TcGemmArgs {
  const void* A;
  RowMap a_map;
  const void* W;
  int nseq, rows_per_seq, N, K;
  Epilogue e;
  float* out1_f32;
  __nv_bfloat16* out1_bf16;
};

// ---- launchers (each returns the number of kernels it launched) -------------
int launch_conv0(cudaStream_t st, const float* wav, int batch, long long n_samples, int seq0, int nseq,
                 long long L0, const float* w /*[10][256]*/, const float* bias, const float* g,
                 const float* b, void* out, int out_bf16, long long out_seq_stride /*elements*/,
                 int out_pad_rows);

// conv0 + ChannelNorm folded for the tensor path (tc_host.cu): centred, pre-scaled taps and closed-form statistics.
struct Conv0Stats {
  float G[10][10];  // sum_c (w_c - wbar)(w_c - wbar)'
  float h2[10];     // 2 sum_c (b_c - bbar)(w_c - wbar)
  float s;          // sum_c (b_c - bbar)^2
};
void conv0_v2_fold(const float* w, const float* bias, const float* g, float* u, float* d, Conv0Stats* cs);

// conv0 on the tensor cores (k_conv0_tc.cu): tf32 GEMM with K = 16 over an im2col tile built in shared memory
int launch_conv0_tc(cudaStream_t st, const float* wav, int batch, long long n_samples, int seq0, int nseq,
                    long long L0, const float* u, const float* d, const float* beta, const Conv0Stats& cs,
                    __nv_bfloat16* out, long long out_seq_stride, int out_pad_rows, int n_sm, std::string* err);

int launch_gemm_f32(cudaStream_t st, const GemmProblem& p, const Epilogue& e);
// The same contraction on the tensor cores with fp32-class accuracy (k_gemm_x3.cu: fp16 hi/lo split of both operands,
// three MMAs per K step). w_packed: device copy of x3_pack_weight's output. Returns launches, or -1 if unsupported.
int launch_gemm_x3(cudaStream_t st, const GemmProblem& p, const Epilogue& e, const void* w_packed, int n_sm,
                   std::string* err);

// Linear-layer GEMM with staged / blocked epilogue I/O (k_gemm_lin.cu). f32_mode: 1 = out1_f32, resid and
// accumulate use the row-blocked fp32 layout [row/128][col/4][row%128][4]; 2 = out1_f32 row-major via TMA.
int launch_gemm_lin(cudaStream_t st, const TcGemmArgs& a, int f32_mode, int n_sm, std::string* err);

// Fused FFN block (k_ffn_fused.cu): x_out = resid + W2 GELU(W1 z); xs = bf16(x_out); zn = LayerNorm(x_out) or null
// vad (optional, last layer only, zn == null): the VAD head Linear(256, 1) taken from the rows of x_out in the same kernel
struct FfnVad {
  const float *w, *b;           // device fp32 [256], [1]
  int batch, T;                 // rows are (channel, batch, frame)-major: row = (c * batch + b) * T + t
  float *logits, *sig;          // (batch, T, 2) fp32, either may be null
};
int launch_ffn_fused(cudaStream_t st, const __nv_bfloat16* z, const __nv_bfloat16* w1, const __nv_bfloat16* w2,
                     const float* resid, float* x_out, __nv_bfloat16* xs, __nv_bfloat16* zn, const float* g2,
                     const float* b2, int M, int n_sm, std::string* err, long long* dbg = nullptr,
                     const FfnVad* vad = nullptr);

// CTA-pair (cta_group::2) GEMM for the K >= 1024 convolutions (k_gemm_2sm.cu); conv epilogue only. Returns launches or -1.
int launch_gemm_2sm(cudaStream_t st, const TcGemmArgs& a, int n_sm, std::string* err);

// conv0 + ChannelNorm + ReLU fused into conv1's operand producer (k_conv01.cu). host_tab / dev_tab: the folded conv0
// table [12][256] = u (10 taps) | d | beta on the host (unused) and on the device. wav: fp32, or int16 PCM (wav_pcm16,
// scaled by 1/32768 on the fly; needs an even n_samples). Returns launches or -1.
int launch_conv01(cudaStream_t st, const void* wav, int wav_pcm16, int batch, long long n_samples, int seq0, int nseq, long long L0,
                  long long L1, const float* host_tab, const float* dev_tab, const Conv0Stats& cs, const void* w1,
                  const float* bias1, const float* g1, const float* b1, void* out, long long out_seq_stride,
                  int out_pad_rows, int n_sm, std::string* err, long long* dbg = nullptr);

int launch_zero_rows(cudaStream_t st, void* buf, int elem_bytes, int nseq, long long seq_stride_elems,
                     long long row0, long long nrows);  // zero rows [row0,row0+nrows) of every sequence

int launch_attention_map_f32(cudaStream_t st, const float* q, long long q_row_stride, const float* k,
                             long long kv_row_stride, int nseq, int T, int n_heads, const float* slopes, int cross,
                             float* maps /*[batch][2][n_layers][H][T][T]*/, int batch, int n_layers, int layer);
int launch_attention_f32(cudaStream_t st, const float* q, long long q_row_stride, const float* k,
                         const float* v, long long kv_row_stride, float* out, int nseq, int T, int n_heads,
                         const float* slopes, int kv_seq_xor_half /* cross: K/V of seq (s+nseq/2)%nseq */);

// fp32-class attention on the tensor cores (k_attn_x3.cu, mode fp32_tc): q / k|v are the contiguous fp32 buffers the
// projections wrote (q_cols / kv_cols wide; k and v start at columns k_off / v_off of the second one), `planes` is
// scratch for their fp16 hi / lo copies (4 bytes per element of both buffers; one buffer if they are the same)
int launch_attention_x3(cudaStream_t st, const float* qbuf, int q_cols, const float* kvbuf, int kv_cols, int k_off,
                        int v_off, void* planes, float* out, int nseq, int T, int n_heads, const float* slopes, int cross,
                        int n_sm, std::string* err);

int launch_rnn_f32(cudaStream_t st, int kind /*0 LSTM 1 GRU*/, const float* xproj /*[nseq][T][G*256]*/,
                   const float* whh_t /*[256][G*256]*/, const float* bhn /*GRU b_hn [256] or null*/,
                   float* out, long long out_seq_stride, int nseq, int T);

int launch_attention_simt_bf16(cudaStream_t st, const __nv_bfloat16* q, long long q_row_stride,
                               const __nv_bfloat16* k, const __nv_bfloat16* v, long long kv_row_stride,
                               __nv_bfloat16* out, int nseq, int T, int n_heads, const float* slopes, int cross);
// tcgen05 fused attention (k_attn_tc.cu): bf16 in/out, rows of 256 = n_heads*64. Returns launches or -1.
int launch_attention_tc(cudaStream_t st, const __nv_bfloat16* q, long long q_row_stride, const __nv_bfloat16* k,
                        const __nv_bfloat16* v, long long kv_row_stride, __nv_bfloat16* out, int nseq, int T,
                        int n_heads, const float* slopes, int cross, int n_sm, std::string* err,
                        long long* dbg = nullptr);
int launch_rnn_f32_bf16out(cudaStream_t st, int kind, const float* xproj, const float* whh_t, const float* bhn,
                           __nv_bfloat16* out, long long out_seq_stride, int nseq, int T);

// Tensor-core recurrence (k_rnn_tc.cu). x rows at x + seq*x_seq_stride + t*x_row_stride (elements, bf16);
// out rows at out + seq*out_seq_stride + t*256. Returns launches or -1.
int launch_rnn_tc(cudaStream_t st, int kind, const __nv_bfloat16* x, long long x_seq_stride, long long x_row_stride,
                  const __nv_bfloat16* w_cat, const float* bias, __nv_bfloat16* out, long long out_seq_stride,
                  int nseq, int T, std::string* err, long long* dbg = nullptr, int groups = 0);
void rnn_tc_pack(int kind, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                 float* w_cat /*[1024][512]*/, float* bias /*[1024]*/);

int launch_vad_head(cudaStream_t st, const float* x /*[2B][T][256] channel-major*/, const float* w,
                    const float* b, int batch, int T, float* vad_logits /*(B,T,2) or null*/,
                    float* vad_sig /*(B,T,2) or null*/);

int launch_resample(cudaStream_t st, const void* x, int x_fmt, long long items, int channels, long long n_in,
                    long long item_stride, long long chan_stride, long long elem_stride, int orig, int new_, int width,
                    const float* bank, float* out, long long n_out, long long out_row_stride, std::string* err);
int launch_vad_filter(cudaStream_t st, const float* vad01, int batch, int T, int max_fill, int max_omit, float* out,
                      int from_logits = 0, float cutoff = 0.5f);
int launch_zero_shot(cudaStream_t st, const float* x, int is_probs, long long batch, int T, const float* va,
                     long long va_T, const uint32_t* sets /* host [10][8] */, float* p, float* p_bc, float* p_sil,
                     float* p_act);
// counters (nullable): unsigned long long [258], ACCUMULATED: [0,256) histogram of the arg-max class over the rows,
// [256 + c] frames of channel c with vad_sig >= 0.5 (vad_sig: (rows, 2) sigmoid outputs, needed with counters)
int launch_probs(cudaStream_t st, const float* logits, long long rows, int now_lo, int now_hi, int fut_lo,
                 int fut_hi, float* probs, float* p_now, float* p_future, float* H, float* lse,
                 uint8_t* argmax, unsigned long long* counters, const float* vad_sig);
int launch_pcm16_to_f32(cudaStream_t st, const int16_t* pcm, long long n, float* out);  // out[i] = pcm[i] / 32768

// vap_head GEMM fused with softmax / entropy / codebook marginals / arg-max / logsumexp / counters (k_head_fused.cu).
// x: 16-bit (rows, 256); w: 16-bit [256][256]; every output optional. Returns launches or -1.
int launch_head_probs(cudaStream_t st, const void* x, const void* w, const float* bias, long long rows, int now_lo,
                      int now_hi, int fut_lo, int fut_hi, float* logits, float* probs, float* p_now, float* p_future,
                      float* H, float* lse, uint8_t* argmax, unsigned long long* counters, const float* vad_sig, int n_sm,
                      std::string* err);

int launch_loss(cudaStream_t st, const float* logits, const float* vad_sig, const float* lse, int batch,
                int T, float* loss);

int launch_to_f32(cudaStream_t st, const void* src, int src_bf16, RowMap src_map, int nseq,
                  int rows_per_seq, float* dst, int src_blocked = 0);
// vad head on the row-blocked fp32 residual stream (k_heads.cu)
int launch_vad_head_blocked(cudaStream_t st, const float* x, const float* w, const float* b, int batch, int T,
                            float* vad_logits, float* vad_sig);

// ---- small device helpers ----------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_erf(float x) {  // nn.GELU() default (exact erf)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_GELU) return gelu_erf(v);
  return v;
}

}  // namespace vapb
