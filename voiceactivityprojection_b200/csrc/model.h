// Internal model / plan structures shared by api.cu and the two forward paths.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace vapb {

constexpr int kMaxLayers = 16;
constexpr int kFfn = 768;

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

// One dtype's packed weights (device pointers into the arena). T = float or bf16.
struct LayerW {
  const float *ln_sa_g, *ln_sa_b, *ln_ffn_g, *ln_ffn_b, *ln_src_g, *ln_src_b;
  const float *slopes, *slopes_cross;
  const void *wqkv;     // [256][768]  (q | k | v)
  const void *wproj;    // [256][256]
  const void *wq_c;     // [256][256]
  const void *wkv_c;    // [256][512]  (k | v)
  const void *wproj_c;  // [256][256]
  const void *w1;       // [256][768]
  const void *w2;       // [768][256]
};

struct Weights {
  // conv0 (always fp32: it runs on CUDA cores)
  const float *c0_w, *c0_b, *c0_g, *c0_be;
  const void *conv_w[5];                    // [k*256][256], index 1..4
  const float *conv_b[5], *conv_g[5], *conv_be[5];
  const void *rnn_wih[kMaxLayers];          // [256][G*256]
  const float *rnn_bx[kMaxLayers];          // b_ih + b_hh (GRU: n gate gets b_in only)
  const float *rnn_whh_t[kMaxLayers];       // fp32 [256][G*256]
  const float *rnn_bhn[kMaxLayers];         // GRU b_hn
  const void *ds_w;                         // [5*256][256]
  const float *ds_b, *ds_g, *ds_be;
  LayerW chan[kMaxLayers], cross[kMaxLayers];
  const void *comb_a, *comb_b;              // [256][256]
  const float *comb_g, *comb_be;
  const float *va_w, *va_b;
  const void *head_w;                       // [256][256]
  const float *head_b;
};

// Per-kernel-family device timing (CUDA events on the launch stream), used by
// bench.py for the roofline of the dominant kernel. Off unless requested.
enum ProfCat { CAT_CONV0 = 0, CAT_CONV_GEMM, CAT_LINEAR_GEMM, CAT_ATTN, CAT_RNN, CAT_HEADS, CAT_OTHER, CAT_COUNT };
struct ProfRec { int cat; int launches; cudaEvent_t a, b; };

struct Model {
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  int device = 0;
  bool finalized = false;
  int ar_kind = 0, ar_layers = 0, channel_layers = 0, cross_layers = 0, num_heads = 0;
  std::map<std::string, HostTensor> staged;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  Weights w32{};   // fp32 [K][N] packing (SIMT path)
  void* bf16_state = nullptr;  // tensor-path weights / tensor maps (forward_bf16.cu)
  // fp32 mode on the tensor cores (k_gemm_x3.cu): every GEMM weight of the fp32 path, split into fp16 hi / lo parts and
  // laid out for the kernel; keyed by the fp32 weight's device pointer. Used by mode VAPB_MODE_FP32_TC; env
  // VAPB_FP32_TC=0 skips the preparation (the mode is then unavailable)
  int fp32_tc = 1;
  int attn_x3 = 1;        // mode fp32_tc: attention on the tensor cores too (k_attn_x3.cu); env VAPB_ATTN_X3
  void* x3_arena = nullptr;
  std::map<const void*, const void*> x3_w;
  int n_sm = 148;
  int conv_2sm = 1;       // gEncoder convs on CTA pairs (k_gemm_2sm.cu, cta_group::2); env VAPB_CONV_2SM
  int ffn_fused = 1;      // FFN block as one kernel (k_ffn_fused.cu); 0 = two GEMMs (env VAPB_FFN_FUSED)
  int vad_fused = 1;      // VAD head inside the last layer's fused FFN kernel (k_ffn_fused.cu); env VAPB_VAD_FUSED
  int head_fused = 1;     // vap_head GEMM fused with the probs() epilogue (k_head_fused.cu); env VAPB_HEAD_FUSED
  int conv01 = 1;         // conv0 fused into conv1's operand producer (k_conv01.cu); 0 = separate kernels (env VAPB_CONV01)
  int conv0_tc = 1;       // unfused path: conv0 on the tensor cores (k_conv0_tc.cu); the CUDA-core fallback is gone
  int conv0_sms = 0;      // CTAs of the conv0 kernel (0 = one per SM); env VAPB_CONV0_SMS (tuning / overlap experiments)
  long long conv_mb_bytes = 0;  // conv0-output bytes per gEncoder micro-batch (0 = 4 GiB); env VAPB_CONV_MB_MIB (tuning)
  // Item-group pipelining of 16-bit-mode calls (api.cu): a large batch is cut into up to `pipe` groups of items, each
  // on its own stream. The groups' gEncoder phases run one after another (event chain), so one group's gAR recurrence
  // (latency-bound on a fraction of the SMs) and transformer overlap the next group's convolutions. env VAPB_PIPE
  // (1 = off), VAPB_PIPE_MIN (fewest items per group).
  int pipe = 1;
  int pipe_min_items = 16;
  std::vector<cudaStream_t> pipe_st;   // streams of groups 1.. (group 0 runs on the caller's stream)
  std::vector<cudaEvent_t> pipe_conv;  // group k's gEncoder is done
  std::vector<cudaEvent_t> pipe_join;  // group k (>= 1) is done
  cudaEvent_t pipe_fork = nullptr;
  // env VAPB_PIPE_TRACE=1 (diagnostics): timestamps of each group's phases, printed to stderr after a host sync
  int pipe_trace = 0;
  std::vector<std::pair<std::string, cudaEvent_t>> trace_ev;
  void trace(cudaStream_t st, const std::string& tag) {
    if (!pipe_trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    trace_ev.emplace_back(tag, e);
  }
  int trace_group = 0;
  unsigned long long launches = 0;
  std::string err;
};

// RAII: times everything launched on `st` during its lifetime under `cat`.
struct ProfScope {
  Model& m;
  cudaStream_t st;
  int cat;
  unsigned long long l0;
  cudaEvent_t a = nullptr;
  ProfScope(Model& m_, cudaStream_t st_, int cat_) : m(m_), st(st_), cat(cat_), l0(m_.launches) {
    if (!m.profiling) return;
    a = take();
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEvent_t b = take();
    cudaEventRecord(b, st);
    m.prof.push_back(ProfRec{cat, (int)(m.launches - l0), a, b});
  }
  cudaEvent_t take() {
    if (!m.event_pool.empty()) {
      cudaEvent_t e = m.event_pool.back();
      m.event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

// Geometry of one (batch, n_samples) problem.
struct Geometry {
  int batch, nseq;
  long long S, L[5], T;
};
int make_geometry(int batch, long long n_samples, Geometry* g);

// A named intermediate activation of a forward pass (diagnostics).
struct StageRef {
  const void* ptr;
  int is_bf16;
  RowMap map;
  int nseq, rows_per_seq;
  int blocked = 0;  // fp32 row-blocked layout (k_gemm_lin.cu)
};

// ---- FP32 path (forward_fp32.cu) -------------------------------------------
size_t workspace_bytes_fp32(const Model& m, const Geometry& g);
// Optional attention-map outputs of forward(attention=True) (vap/model.py:262-266), fp32 path only:
// self_attn [B][2][channel_layers][H][T][T] (ar_channel), cross_self_attn / cross_attn [B][2][cross_layers][H][T][T]
// (the stereo layers' self- and cross-attention). Pointers address item 0 of the call.
struct AttnMaps { float *self_attn, *cross_attn, *cross_self_attn; };
int forward_fp32(Model& m, cudaStream_t st, const float* wav, const Geometry& g, char* ws, float* logits,
                 float* vad_logits, float* vad_sig, const float** comb_out, const AttnMaps* maps = nullptr,
                 bool tensor_gemms = false /* VAPB_MODE_FP32_TC: k_gemm_x3.cu for every contraction */);
int stage_fp32(const Model& m, const Geometry& g, char* ws, const std::string& name, StageRef* ref);

// Outputs of the fused vap_head + probs kernel for the items of one forward call (pointers address the call's first
// item; any may be null). Passed to forward_bf16 instead of a logits buffer when the caller wants probs().
struct HeadOut {
  int now_lo, now_hi, fut_lo, fut_hi;
  float *logits, *probs, *p_now, *p_future, *H, *lse;
  uint8_t* argmax;
  unsigned long long* counters;
};

// ---- BF16 tensor-core path (forward_bf16.cu) --------------------------------
int bf16_prepare(Model& m);   // pack bf16 weights after the fp32 arena is built
void bf16_release(Model& m);
size_t workspace_bytes_bf16(const Model& m, const Geometry& g);
int forward_bf16(Model& m, cudaStream_t st, const float* wav, const Geometry& g, char* ws, float* logits,
                 float* vad_logits, float* vad_sig, const float** comb_out, int fp16,
                 cudaEvent_t conv_wait = nullptr, cudaEvent_t conv_done = nullptr, int wav_pcm16 = 0,
                 const HeadOut* head = nullptr);
int stage_bf16(const Model& m, const Geometry& g, char* ws, const std::string& name, StageRef* ref);

}  // namespace vapb
