/*
 * vapb.h — C-ABI of the B200-native VAP stereo inference forward path.
 *
 * The reference (ErikEkstedt/VoiceActivityProjection) has no FFI layer: the
 * boundary of this path is the Python nn.Module surface of `VapGPT`
 * (vap/model.py:125-268). This library is what a binding for that surface
 * calls; voiceactivityprojection_b200/model.py is that binding (ctypes), and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative VAPB_E_* code
 * (vapb_last_error() gives the text). No exceptions, no allocation of caller
 * visible memory across the ABI: the caller (PyTorch) owns every device buffer,
 * including the workspace. All device work is enqueued on the caller's stream
 * and is asynchronous. One handle per device; a handle may be used from several
 * streams concurrently when the workspaces differ.
 *
 * There is no CPU implementation behind this interface.
 */
#ifndef VAPB_H_
#define VAPB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct VapbHandle VapbHandle;

enum {
  VAPB_OK = 0,
  VAPB_E_INVALID = -1,   /* bad argument / shape */
  VAPB_E_STATE = -2,     /* state dict incomplete, unexpected key, wrong call order */
  VAPB_E_CUDA = -3,      /* CUDA runtime / driver error */
  VAPB_E_WORKSPACE = -4, /* workspace too small */
  VAPB_E_UNSUPPORTED = -5
};

/* Arithmetic mode of the contractions.
 * FP32: CUDA-core FMA, fp32 activations (parity mode; vap/model.py runs fp32).
 * BF16: tcgen05 tensor cores, bf16 operands, fp32 accumulate / norms / softmax / residual stream.
 * FP16: the same kernels with fp16 operands (3 more mantissa bits; every activation of this model
 *       sits behind a norm, far inside fp16 range): same speed, ~8x lower error than BF16. */
/* FP32: CUDA-core FMA everywhere (bit-exact decisions, probabilities within 1e-5). BF16 / FP16: tcgen05, 16-bit operands.
 * FP32_TC: the fp32 path with every GEMM on the tensor cores at fp32-class accuracy (fp16 hi/lo split of both operands,
 * three MMAs per K step; softmax, recurrence, norms in fp32): 2.7x the FP32 mode, probabilities within 1e-5. */
enum { VAPB_MODE_FP32 = 0, VAPB_MODE_BF16 = 1, VAPB_MODE_FP16 = 2, VAPB_MODE_FP32_TC = 3 };

/* --- construction: replaces VapGPT.__init__ + load_state_dict (run.py:199-201) */

/* Creates an empty model on CUDA device `device`. */
int vapb_create(int device, VapbHandle** out);

/* Offers one state-dict entry (reference key schema), host fp32,
 * contiguous, in the reference's own layout. Copies the data. Unknown keys are
 * an error (strict load), like nn.Module.load_state_dict (run.py:201). */
int vapb_load_tensor(VapbHandle* h, const char* key, const float* data, int ndim,
                     const int64_t* shape);

/* Validates that the state dict is complete, infers the gAR cell (LSTM or GRU)
 * and depth from the shapes (vap/encoder_components.py:384-391 takes them from
 * the CPC checkpoint's config, which a VAP state dict does not carry), repacks
 * the weights into kernel layouts and uploads them. */
int vapb_finalize(VapbHandle* h);

int vapb_destroy(VapbHandle* h);

/* Text of the last error on this handle (or of the last vapb_create failure
 * when h is NULL). Valid until the next call on the handle. */
const char* vapb_last_error(const VapbHandle* h);

/* Model facts after vapb_finalize: ar_kind 0 = LSTM, 1 = GRU. */
int vapb_describe(const VapbHandle* h, int* ar_kind, int* ar_layers, int* channel_layers,
                  int* cross_layers, int* num_heads);

/* --- shapes: the zero-padded conv chain of vap/encoder_components.py:83-91 and
 * the causal stride-2 conv of vap/encoder.py:24-30. frames100 = gAR length,
 * frames50 = output frames T. */
int vapb_frames(int64_t n_samples, int64_t* frames100, int64_t* frames50);

/* Bytes of device workspace vapb_forward/vapb_probs need for this shape. */
int vapb_workspace_bytes(const VapbHandle* h, int batch, int64_t n_samples, int mode,
                         size_t* bytes);

/* --- the hot path ---------------------------------------------------------- */

/* VapGPT.forward(waveform) (vap/model.py:249-268, attention=False).
 * wav:        device fp32 (batch, 2, n_samples), contiguous.
 * logits:     device fp32 (batch, T, 256).
 * vad_logits: device fp32 (batch, T, 2).
 * `stream` is a cudaStream_t. */
int vapb_forward(VapbHandle* h, void* stream, const float* wav, int batch, int64_t n_samples,
                 int mode, void* workspace, size_t workspace_bytes, float* logits,
                 float* vad_logits);

/* VapGPT.forward(waveform, attention=True) (vap/model.py:249-268): the forward of vapb_forward in FP32 mode
 * plus the attention maps the reference returns, softmax(q k^T / 16 + 1 + m_h j) with zeros above the diagonal
 * (vap/modules.py:82-110, 169-202), as device fp32:
 *   self_attn       [batch][2][channel_layers][heads][T][T]  (ar_channel; vap/modules.py:342-358)
 *   cross_attn      [batch][2][cross_layers][heads][T][T]    (stereo layers' cross-attention; :380-408)
 *   cross_self_attn [batch][2][cross_layers][heads][T][T]    (stereo layers' self-attention)
 * Index 1 is the speaker channel. The maps are a diagnostic output (4 T^2 floats per head and layer): every
 * element is written, the caller sizes the batch. Workspace as vapb_workspace_bytes(mode = VAPB_MODE_FP32). */
int vapb_forward_attention(VapbHandle* h, void* stream, const float* wav, int batch, int64_t n_samples,
                           void* workspace, size_t workspace_bytes, float* logits, float* vad_logits,
                           float* self_attn, float* cross_attn, float* cross_self_attn);

/* VapGPT.probs(waveform, now_lims, future_lims) (vap/model.py:180-225), fused
 * with the forward. Outputs, all device fp32:
 *   probs (batch,T,256), vad (batch,T,2) = sigmoid, p_now (batch,T,2),
 *   p_future (batch,T,2), H (batch,T), loss (batch,T-100).
 * logits / vad_logits / loss may be NULL (skipped). `loss` follows the
 * reference quirk that labels come from the model's own sigmoid(vad)
 * (vap/model.py:190,220-224); T must be > 100 when loss is requested.
 * argmax (nullable): device uint8 (batch,T), argmax class of probs. */
int vapb_probs(VapbHandle* h, void* stream, const float* wav, int batch, int64_t n_samples,
               int mode, void* workspace, size_t workspace_bytes, int now_lo, int now_hi,
               int fut_lo, int fut_hi, float* logits, float* vad_logits, float* probs,
               float* vad, float* p_now, float* p_future, float* H, float* loss,
               uint8_t* argmax);

/* vapb_probs with two extras for bulk inference:
 *  - wav_fmt VAPB_WAV_PCM16: `wav` holds int16 PCM in the same (batch, 2, n_samples) layout (what a wav file holds;
 *    the reference converts on the host, vap/audio.py:47). The fused encoder kernel reads it directly and scales by
 *    1/32768 (exact), so the waveform crosses PCIe and HBM at half the bytes. 16-bit modes only, n_samples even,
 *    4-byte aligned buffer; VAPB_E_UNSUPPORTED otherwise (vapb_pcm16_to_f32 converts for those cases).
 *  - counters (nullable): device unsigned long long [258], ACCUMULATED by the heads kernel: [0,256) histogram of the
 *    arg-max projection-window class over the call's frames, [256 + c] frames with vad[..., c] >= 0.5. */
#define VAPB_WAV_F32 0
#define VAPB_WAV_PCM16 1
int vapb_probs_ex(VapbHandle* h, void* stream, const void* wav, int wav_fmt, int batch, int64_t n_samples,
                  int mode, void* workspace, size_t workspace_bytes, int now_lo, int now_hi,
                  int fut_lo, int fut_hi, float* logits, float* vad_logits, float* probs,
                  float* vad, float* p_now, float* p_future, float* H, float* loss,
                  uint8_t* argmax, unsigned long long* counters);

/* cudaMemsetAsync(ptr, 0, bytes) on `stream`: the bulk driver clears its per-step counters with it. */
int vapb_memset_zero(void* stream, void* ptr, size_t bytes);

/* out[i] = pcm[i] / 32768 for n int16 samples (device buffers), on `stream`. */
int vapb_pcm16_to_f32(void* stream, const int16_t* pcm, int64_t n, float* out);

/* ObjectiveVAP.get_probs(logits) (vap/objective.py:249-281) and the post-forward
 * half of VapGPT.probs on logits the caller already has: softmax, p_now,
 * p_future, entropy, argmax over `rows` frames of 256 logits (device fp32).
 * Any output may be NULL. */
int vapb_probs_from_logits(VapbHandle* h, void* stream, const float* logits, int64_t rows, int now_lo,
                           int now_hi, int fut_lo, int fut_hi, float* probs, float* p_now,
                           float* p_future, float* H, uint8_t* argmax);

/* --- diagnostics ----------------------------------------------------------- */

/* Copies an intermediate activation of the LAST forward/probs run on
 * `workspace` (same batch / n_samples / mode) to `out` as fp32.
 * name: "conv" (2B,T100,256: CPC gEncoder output, channels last),
 *       "ar" (2B,T100,256: gAR output), "enc" (2B,T,256: encoder output),
 *       "ch" (2B,T,256: ar_channel output), "ar0".."arN" (2B,T,256: stereo
 *       layer outputs), "comb" (B,T,256: combinator output).
 * Sequence index is channel-major: row c*batch + b holds channel c of item b. */
int vapb_get_stage(VapbHandle* h, void* stream, const char* name, int batch, int64_t n_samples,
                   int mode, void* workspace, size_t workspace_bytes, float* out,
                   size_t out_elems);

/* Per-kernel-family device timing for roofline reports. vapb_profile_begin turns
 * on CUDA-event bracketing of every launch group (events recorded on the launch
 * stream); vapb_profile_end synchronises, returns the summed milliseconds and
 * kernel launches per family and turns it off again. Families (index):
 * 0 conv0, 1 conv implicit-GEMMs (conv1-4, downsample), 2 linear GEMMs
 * (gAR input projection, q/k/v/proj, FFN, combinator, vap_head), 3 attention,
 * 4 gAR recurrence, 5 heads (vad, probs, loss), 6 other (padding, exports).
 * Arrays must hold VAPB_PROFILE_FAMILIES entries. */
#define VAPB_PROFILE_FAMILIES 7
int vapb_profile_begin(VapbHandle* h);
int vapb_profile_end(VapbHandle* h, double* ms, uint64_t* launches);

/* Unit-test hook for the fp32-class tensor-core GEMM of the parity mode (csrc/k_gemm_x3.cu: fp16 hi/lo split of both
 * operands, three MMAs per K step). A: device fp32, row (seq, t) at A + seq*a_seq_stride + t*a_row_stride (elements),
 * K contiguous (rows may overlap: implicit conv). Wt_host: HOST fp32 [K][N] (element (k, n) at k*N + n; split and laid
 * out by the hook). Epilogue as k_gemm_f32.cu, all device fp32, outputs / residual dense (nseq*rows_per_seq, N):
 * bias -> norm1 (1 ChannelNorm, 2 LayerNorm; N == 256) -> act (1 ReLU, 2 GELU) -> + resid -> (+= out1) -> out1;
 * out2 = LayerNorm2(out1 value). N % 256 == 0, K % 32 == 0. */
int vapb_debug_gemm_x3(void* stream, const float* A, int64_t a_seq_stride, int64_t a_row_stride, const float* Wt_host,
                       int nseq, int rows_per_seq, int N, int K, const float* bias, int norm1, const float* g1,
                       const float* b1, int act, const float* resid, int accumulate, float* out1, int norm2,
                       const float* g2, const float* b2, float* out2, char* err, int err_len);

/* Unit-test hook for the CTA-pair (cta_group::2) conv GEMM (csrc/k_gemm_2sm.cu): N = 256,
 * bias -> norm1 -> ReLU (act 1) -> dense bf16 (nseq*rows_per_seq, 256).
 * A: 16-bit, row (seq, t) at A + seq*a_seq_stride + t*a_row_stride (elements), K contiguous elements per row (rows may
 * overlap: implicit conv). W: 16-bit [N][K]. Returns 0 or a negative code; the message is copied to err (if non-NULL). */
int vapb_debug_gemm_2sm(void* stream, const void* A, int64_t a_seq_stride, int64_t a_row_stride, const void* W,
                        int nseq, int rows_per_seq, int K, const float* bias, int norm1, const float* g1,
                        const float* b1, int act, void* out_bf16, char* err, int err_len);

/* Unit-test hook for the stand-alone conv0 + ChannelNorm + ReLU kernel (csrc/k_conv0_tc.cu, the VAPB_CONV01=0 path;
 * vap/encoder_components.py:83-84,99). Operands as vapb_debug_conv01; out: device 16-bit, dense (2*batch, L0, 256),
 * sequences in channel-major order. */
int vapb_debug_conv0_tc(void* stream, const float* wav, int batch, int64_t n_samples, const float* conv0_w,
                        const float* conv0_b, const float* norm0_g, const float* norm0_b, void* out, int fp16,
                        char* err, int err_len);

/* Unit-test hook for the fused conv0 -> conv1 kernel (csrc/k_conv01.cu; vap/encoder_components.py:83-86,99-100).
 * wav: device fp32 (batch, 2, n_samples); every channel of every item is one sequence (channel-major order
 * c*batch + item). conv0_w (256,1,10), conv0_b, norm0_g, norm0_b (256): HOST fp32 parameters of conv0 and its
 * ChannelNorm (folded on the host). w1: device 16-bit [256][8*256] conv1 weight, K index = tap*256 + cin;
 * bias1 / g1 / b1: device fp32 [256]. out: device 16-bit, row t of sequence s at
 * out + s*out_seq_stride + (out_pad_rows + t)*256 (elements); whole 128-row tiles are written (zeros past the
 * L1 = conv1 output length), so a sequence needs out_pad_rows + roundup(L1,128) rows. fp16: 0 = bf16 words. */
int vapb_debug_conv01(void* stream, const float* wav, int batch, int64_t n_samples, const float* conv0_w,
                      const float* conv0_b, const float* norm0_g, const float* norm0_b, const void* w1,
                      const float* bias1, const float* g1, const float* b1, void* out, int64_t out_seq_stride,
                      int out_pad_rows, int fp16, char* err, int err_len,
                      long long* dbg_clocks /* device [4][4][16] SM-clock samples of CTA 0, or NULL */);

/* Unit-test hook for the linear-layer GEMM (csrc/k_gemm_lin.cu): operands as
 * vapb_debug_gemm_2sm plus N; outputs are dense (nseq*rows_per_seq, N). f32_mode 1: out1_f32
 * and resid_blocked use the row-blocked fp32 layout [row/128][col/4][row%128][4]
 * (buffers padded to a multiple of 128 rows); f32_mode 2: out1_f32 is row-major. */
int vapb_debug_gemm_lin(void* stream, const void* A, int64_t a_seq_stride, int64_t a_row_stride, const void* W,
                        int nseq, int rows_per_seq, int N, int K, const float* bias, int norm1, const float* g1,
                        const float* b1, int act, const float* resid_blocked, int accumulate, float* out1_f32,
                        int f32_mode, void* out1_bf16, int norm2, const float* g2, const float* b2, void* out2_bf16,
                        char* err, int err_len);

/* Unit-test hook for the fused FFN block (csrc/k_ffn_fused.cu): x_out = resid + W2 GELU(W1 z).
 * z: device bf16 (M,256); w1: bf16 [768][256]; w2: bf16 [256][768]; resid / x_out: row-blocked
 * fp32 (see vapb_debug_gemm_lin); xs: bf16 (M,256) copy of x_out; zn: bf16 (M,256) =
 * LayerNorm(x_out; g2, b2), or NULL. */
int vapb_debug_ffn_fused(void* stream, const void* z, const void* w1, const void* w2, const float* resid_blocked,
                          float* x_out_blocked, void* xs, void* zn, const float* g2, const float* b2, int M, char* err,
                          int err_len, long long* dbg_clocks /* device [32][16] SM-clock samples or NULL */);

/* Unit-test hooks for the tensor-core gAR recurrence (csrc/k_rnn_tc.cu).
 * vapb_debug_rnn_pack (host only): nn.LSTM / nn.GRU parameters of one layer
 * (weight_ih (G*256,256), weight_hh, bias_ih, bias_hh; kind 0 = LSTM, 1 = GRU) ->
 * the kernel's packed fp32 [1024][512] weight and [1024] bias.
 * vapb_debug_rnn_tc: one launch. x: device bf16, row (seq,t) at
 * x + seq*x_seq_stride + t*x_row_stride; w_cat: device bf16 [1024][512];
 * bias: device fp32 [1024]; out: device bf16 rows at out + seq*out_seq_stride + t*256. */
int vapb_debug_rnn_pack(int kind, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                        float* w_cat, float* bias);
int vapb_debug_rnn_tc(void* stream, int kind, const void* x, int64_t x_seq_stride, int64_t x_row_stride,
                      const void* w_cat, const float* bias, void* out, int64_t out_seq_stride, int nseq, int T,
                      char* err, int err_len, long long* dbg_clocks /* device [32][8] SM-clock samples or NULL */,
                      int groups /* 16-sequence groups per cluster: 2..4, 0 = auto */);

/* Unit-test hook for the tcgen05 fused attention (csrc/k_attn_tc.cu): one launch.
 * q/k/v: device bf16, row (seq,t) at ptr + (seq*T + t)*row_stride, 256 = n_heads*64
 * columns used; out: dense bf16 (nseq*T, 256); slopes: device fp32 [n_heads].
 * cross != 0: K/V of sequence (seq + nseq/2) % nseq (the other speaker channel). With VAPB_DEBUG_FP16=1 in the
 * environment the three inputs and the output are fp16 instead of bf16 (the headline mode's format). */
int vapb_debug_attn_tc(void* stream, const void* q, int64_t q_row_stride, const void* k, const void* v,
                       int64_t kv_row_stride, void* out, int nseq, int T, int n_heads, const float* slopes,
                       int cross, char* err, int err_len,
                       long long* dbg_clocks /* device [64][8] SM-clock samples of CTA 0, or NULL */);

/* Kernel-level test hook: the fp32-class tensor-core attention of mode FP32_TC (csrc/k_attn_x3.cu; reference
 * vap/modules.py:82-110,169-202). qbuf / kvbuf: device fp32, contiguous (nseq*T, q_cols) and (nseq*T, kv_cols) - the
 * same pointer for self-attention (q | k | v in one buffer of 768 columns); K and V start at columns k_off / v_off of
 * kvbuf; planes: device scratch, 4 bytes per element of qbuf plus (when different) kvbuf; out: dense fp32
 * (nseq*T, 256); slopes: device fp32 [4]; cross as in vapb_debug_attn_tc. */
int vapb_debug_attn_x3(void* stream, const float* qbuf, int q_cols, const float* kvbuf, int kv_cols, int k_off,
                       int v_off, void* planes, float* out, int nseq, int T, const float* slopes, int cross, char* err,
                       int err_len);

/* VapGPT.vad()'s post-processing (vap/model.py:240-247 -> vad_fill_silences / vad_omit_spikes,
 * vap/utils.py:239-272) on the device. vad01: device fp32 (batch, T, 2) holding 0/1 (the caller thresholds
 * sigmoid(vad) >= cutoff); silence runs of <= max_fill_frames become 1, then activity runs of <= max_omit_frames
 * of the result become 0. out may alias vad01. `h` may be NULL. */
int vapb_vad_filter(VapbHandle* h, void* stream, const float* vad01, int batch, int64_t T, int max_fill_frames,
                    int max_omit_frames, float* out);
/* The same with the threshold of VapGPT.vad (vap/model.py:237-238) inside the kernel: from_logits != 0 means `vad`
 * holds the model's VAD logits and a frame is active when sigmoid(logit) >= cutoff. */
int vapb_vad_filter_ex(VapbHandle* h, void* stream, const float* vad, int from_logits, float cutoff, int batch,
                       int64_t T, int max_fill_frames, int max_omit_frames, float* out);

/* ZeroShot next-speaker / backchannel marginals: replaces ZeroShot.get_probs,
 * probs_next_speaker, probs_on_silence, probs_on_active and probs_backchannel (vap/zero_shot.py:159-271) with one
 * kernel. x: device fp32 (batch, T, 256) logits (is_probs 0: softmax applied, as get_probs :264-271) or class
 * probabilities (is_probs 1). va: device fp32 (batch, va_T, 2) binary voice activity, va_T >= T, frames [0, T) of
 * every item are used (`va[:, :nmax]`, :268); required only when p is requested. class_sets: HOST uint32 [10][8],
 * ten 256-bit class sets (bit c%32 of word c/32 = class c): silence pos for next speaker 0 and 1, silence neg 0/1,
 * active pos 0/1, active neg 0/1, backchannel 0/1 (ZeroShot.subset_silence, subset_silence_hold, subset_active,
 * subset_active_hold, bc_prediction). Outputs, each device fp32 (batch, T, 2) or NULL: p = next-speaker
 * probabilities by dialog state (vap/events.py:70-78), p_bc, p_sil = probs_on_silence, p_act = probs_on_active.
 * A frame whose subsets carry zero probability gives 0/0 = NaN like the reference. A set may hold at most 64
 * classes (the reference's largest has 56): VAPB_E_UNSUPPORTED otherwise. `h` may be NULL. */
int vapb_zero_shot(VapbHandle* h, void* stream, const float* x, int is_probs, int64_t batch, int64_t T,
                   const float* va, int64_t va_T, const uint32_t* class_sets, float* p, float* p_bc, float* p_sil,
                   float* p_act);

/* Input path: rational polyphase resampling on the device, the arithmetic of
 * torchaudio.functional.resample (sinc_interp_hann) that vap/audio.py:65-68 applies after decoding a file.
 * orig/new are the two rates divided by their gcd. Input: `items` x `channels` rows of n_in samples, float32
 * (x_fmt 0) or int16 PCM scaled by 1/32768 (x_fmt 1); sample i of (item, ch) is at
 * x[item*item_stride + ch*chan_stride + i*elem_stride] (elements), so planar and interleaved buffers both fit.
 * bank: device float32 [new][2*width + orig] windowed-sinc filters (host-computed, see audio.py). Row (item, ch)
 * of the output starts at out + (item*channels + ch)*out_row_stride and holds n_out <= ceil(new*n_in/orig)
 * samples: the (B, 2, n) layout vapb_forward takes. `h` may be NULL (stateless; errors then land in
 * vapb_last_error(NULL)). */
int vapb_resample(VapbHandle* h, void* stream, const void* x, int x_fmt, int64_t items, int channels, int64_t n_in,
                  int64_t item_stride, int64_t chan_stride, int64_t elem_stride, int orig, int new_rate, int width,
                  const float* bank, float* out, int64_t n_out, int64_t out_row_stride);

/* Number of kernels this handle has launched since creation. */
int vapb_launch_count(const VapbHandle* h, uint64_t* launches);

/* Library build facts: "sm_100a" etc. */
const char* vapb_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* VAPB_H_ */
