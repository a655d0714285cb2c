// C-ABI of the library (include/vapb.h): handle, strict state-dict loading and
// weight repacking, shape arithmetic, and the forward / probs entry points.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>

#include "../../include/vapb.h"
#include <cuda_fp16.h>

#include "model.h"

using namespace vapb;

namespace vapb { void x3_pack_weight(const float* wt, int K, int N, std::vector<__half>* out); }

namespace vapb { extern thread_local int g_fp16; }  // 16-bit format of this thread's launches (tc_common.cuh)

struct VapbHandle {
  Model m;
};

static thread_local std::string g_create_err;

namespace vapb {

int make_geometry(int batch, long long n, Geometry* g) {
  // vap/encoder_components.py:83-91 (k,s,p) chain, floor division; then the
  // causal k5 s2 conv with 4 left pad frames (vap/encoder.py:24-30).
  static const int ksp[5][3] = {{10, 5, 3}, {8, 4, 2}, {4, 2, 1}, {4, 2, 1}, {4, 2, 1}};
  g->batch = batch;
  g->nseq = 2 * batch;
  g->S = n;
  long long L = n;
  for (int i = 0; i < 5; ++i) {
    const long long num = L + 2 * ksp[i][2] - ksp[i][0];
    if (num < 0) return -1;
    L = num / ksp[i][1] + 1;
    g->L[i] = L;
  }
  g->T = (L - 1) / 2 + 1;
  return (L >= 1) ? 0 : -1;
}

}  // namespace vapb

namespace {

int fail(Model& m, int code, const std::string& msg) {
  m.err = msg;
  return code;
}

#define CUDA_OK(m, call)                                                                           \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail((m), VAPB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
  } while (0)

struct Packer {
  std::vector<char> host;
  size_t add(const void* src, size_t bytes) {
    const size_t off = (host.size() + 255) / 256 * 256;
    host.resize(off + bytes);
    memcpy(host.data() + off, src, bytes);
    return off;
  }
  size_t add_f(const std::vector<float>& v) { return add(v.data(), v.size() * sizeof(float)); }
};

// Linear weight (out, in) -> [in][out]
std::vector<float> transpose_linear(const HostTensor& t) {
  const int64_t n = t.shape[0], k = t.shape[1];
  std::vector<float> o((size_t)n * k);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < k; ++j) o[(size_t)j * n + i] = t.data[(size_t)i * k + j];
  return o;
}
// several Linear weights side by side along N: [in][sum out]
std::vector<float> concat_linear(std::vector<const HostTensor*> ts) {
  const int64_t k = ts[0]->shape[1];
  int64_t ntot = 0;
  for (auto* t : ts) ntot += t->shape[0];
  std::vector<float> o((size_t)ntot * k);
  int64_t n0 = 0;
  for (auto* t : ts) {
    const int64_t n = t->shape[0];
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = 0; j < k; ++j) o[(size_t)j * ntot + n0 + i] = t->data[(size_t)i * k + j];
    n0 += n;
  }
  return o;
}
// Conv1d weight (out, in, k) -> [(tap*in + cin)][out]: the GEMM K index runs over
// a channels-last window of k frames.
std::vector<float> pack_conv(const HostTensor& t) {
  const int64_t co = t.shape[0], ci = t.shape[1], k = t.shape[2];
  std::vector<float> o((size_t)co * ci * k);
  for (int64_t n = 0; n < co; ++n)
    for (int64_t c = 0; c < ci; ++c)
      for (int64_t j = 0; j < k; ++j) o[((size_t)j * ci + c) * co + n] = t.data[((size_t)n * ci + c) * k + j];
  return o;
}

std::string shape_str(const std::vector<int64_t>& s) {
  std::string r = "(";
  for (size_t i = 0; i < s.size(); ++i) r += (i ? "," : "") + std::to_string(s[i]);
  return r + ")";
}

int count_keys(const Model& m, const char* fmt) {
  int n = 0;
  char buf[256];
  for (;;) {
    snprintf(buf, sizeof buf, fmt, n);
    if (!m.staged.count(buf)) return n;
    ++n;
  }
}

}  // namespace

extern "C" {

const char* vapb_build_info(void) { return "vapb sm_100a (cuda " "12.9" ") fp32=simt bf16=tcgen05"; }

int vapb_create(int device, VapbHandle** out) {
  if (!out) return VAPB_E_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n) {
    g_create_err = e != cudaSuccess ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)
                                    : "no such CUDA device " + std::to_string(device);
    return VAPB_E_CUDA;
  }
  VapbHandle* h = new VapbHandle();
  h->m.device = device;
  cudaDeviceGetAttribute(&h->m.n_sm, cudaDevAttrMultiProcessorCount, device);
  if (const char* v = getenv("VAPB_CONV_2SM")) h->m.conv_2sm = atoi(v);
  if (const char* v = getenv("VAPB_FFN_FUSED")) h->m.ffn_fused = atoi(v);
  if (const char* v = getenv("VAPB_CONV0_TC")) h->m.conv0_tc = atoi(v);
  if (const char* v = getenv("VAPB_CONV01")) h->m.conv01 = atoi(v);
  if (const char* v = getenv("VAPB_HEAD_FUSED")) h->m.head_fused = atoi(v);
  if (const char* v = getenv("VAPB_FP32_TC")) h->m.fp32_tc = atoi(v);
  if (const char* v = getenv("VAPB_ATTN_X3")) h->m.attn_x3 = atoi(v);
  if (const char* v = getenv("VAPB_VAD_FUSED")) h->m.vad_fused = atoi(v);
  if (const char* v = getenv("VAPB_CONV0_SMS")) h->m.conv0_sms = atoi(v);
  if (const char* v = getenv("VAPB_CONV_MB_MIB")) h->m.conv_mb_bytes = atoll(v) << 20;
  if (const char* v = getenv("VAPB_PIPE")) h->m.pipe = atoi(v);
  if (const char* v = getenv("VAPB_PIPE_MIN")) h->m.pipe_min_items = atoi(v);
  if (const char* v = getenv("VAPB_PIPE_TRACE")) h->m.pipe_trace = atoi(v);
  *out = h;
  return VAPB_OK;
}

const char* vapb_last_error(const VapbHandle* h) { return h ? h->m.err.c_str() : g_create_err.c_str(); }

int vapb_load_tensor(VapbHandle* h, const char* key, const float* data, int ndim, const int64_t* shape) {
  if (!h || !key || !data || ndim < 0 || ndim > 4) return VAPB_E_INVALID;
  Model& m = h->m;
  if (m.finalized) return fail(m, VAPB_E_STATE, "vapb_load_tensor after vapb_finalize");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(data, data + n);
  m.staged[key] = std::move(t);
  return VAPB_OK;
}

int vapb_finalize(VapbHandle* h) {
  if (!h) return VAPB_E_INVALID;
  Model& m = h->m;
  if (m.finalized) return fail(m, VAPB_E_STATE, "already finalized");
  const std::string AR = "encoder.encoder.gAR.baseNet.";
  const std::string GE = "encoder.encoder.gEncoder.";

  // ---- structure from the keys / shapes
  m.ar_layers = count_keys(m, (AR + "weight_ih_l%d").c_str());
  m.channel_layers = count_keys(m, "ar_channel.layers.%d.ln_self_attn.weight");
  m.cross_layers = count_keys(m, "ar.layers.%d.ln_self_attn.weight");
  if (m.ar_layers < 1) return fail(m, VAPB_E_STATE, "missing key " + AR + "weight_ih_l0");
  if (m.ar_layers > kMaxLayers || m.channel_layers > kMaxLayers || m.cross_layers > kMaxLayers)
    return fail(m, VAPB_E_UNSUPPORTED, "too many layers");
  {
    const auto& s = m.staged[AR + "weight_ih_l0"].shape;
    if (s.size() != 2 || s[1] != kDim || (s[0] != 4 * kDim && s[0] != 3 * kDim))
      return fail(m, VAPB_E_STATE, AR + "weight_ih_l0 has shape " + shape_str(s) +
                                       "; expected (1024,256) [LSTM] or (768,256) [GRU]");
    m.ar_kind = s[0] == 4 * kDim ? 0 : 1;
  }
  const int64_t GH = (m.ar_kind == 0 ? 4 : 3) * kDim;
  {
    auto it = m.staged.find(m.channel_layers ? "ar_channel.layers.0.mha.m" : "ar.layers.0.mha.m");
    m.num_heads = it == m.staged.end() || it->second.shape.size() != 1 ? 0 : (int)it->second.shape[0];
    if (m.num_heads != 4)
      return fail(m, VAPB_E_UNSUPPORTED, "num_heads must be 4 (head_dim 64); state dict has " +
                                             std::to_string(m.num_heads));
  }

  // ---- strict key/shape check (nn.Module.load_state_dict(strict=True), run.py:201)
  std::map<std::string, std::vector<int64_t>> want;
  const int ck[5] = {10, 8, 4, 4, 4};
  for (int i = 0; i < 5; ++i) {
    const std::string n = std::to_string(i);
    want[GE + "conv" + n + ".weight"] = {kDim, i == 0 ? 1 : kDim, ck[i]};
    want[GE + "conv" + n + ".bias"] = {kDim};
    want[GE + "batchNorm" + n + ".weight"] = {1, kDim, 1};
    want[GE + "batchNorm" + n + ".bias"] = {1, kDim, 1};
  }
  for (int l = 0; l < m.ar_layers; ++l) {
    const std::string n = std::to_string(l);
    want[AR + "weight_ih_l" + n] = {GH, kDim};
    want[AR + "weight_hh_l" + n] = {GH, kDim};
    want[AR + "bias_ih_l" + n] = {GH};
    want[AR + "bias_hh_l" + n] = {GH};
  }
  want["encoder.downsample.1.weight"] = {kDim, kDim, 5};
  want["encoder.downsample.1.bias"] = {kDim};
  want["encoder.downsample.2.ln.weight"] = {kDim};
  want["encoder.downsample.2.ln.bias"] = {kDim};
  auto want_layer = [&](const std::string& p, bool cross) {
    for (const char* ln : {"ln_self_attn", "ln_ffnetwork", "ln_src_attn"}) {
      if (!cross && !strcmp(ln, "ln_src_attn")) continue;
      want[p + ln + ".weight"] = {kDim};
      want[p + ln + ".bias"] = {kDim};
    }
    for (const char* a : {"mha", "mha_cross"}) {
      if (!cross && !strcmp(a, "mha_cross")) continue;
      want[p + a + ".m"] = {m.num_heads};
      for (const char* w : {"key", "query", "value", "proj"}) want[p + a + "." + w + ".weight"] = {kDim, kDim};
    }
    want[p + "ffnetwork.0.weight"] = {kFfn, kDim};
    want[p + "ffnetwork.3.weight"] = {kDim, kFfn};
  };
  for (int l = 0; l < m.channel_layers; ++l) want_layer("ar_channel.layers." + std::to_string(l) + ".", false);
  for (int l = 0; l < m.cross_layers; ++l) want_layer("ar.layers." + std::to_string(l) + ".", true);
  want["ar.combinator.h0_a.weight"] = {kDim, kDim};
  want["ar.combinator.h0_b.weight"] = {kDim, kDim};
  want["ar.combinator.ln.weight"] = {kDim};
  want["ar.combinator.ln.bias"] = {kDim};
  want["objective.codebook.emb.weight"] = {kClasses, 8};
  want["va_classifier.weight"] = {1, kDim};
  want["va_classifier.bias"] = {1};
  want["vap_head.weight"] = {kClasses, kDim};
  want["vap_head.bias"] = {kClasses};

  std::string missing, unexpected, badshape;
  for (auto& kv : want) {
    auto it = m.staged.find(kv.first);
    if (it == m.staged.end()) missing += " " + kv.first;
    else if (it->second.shape != kv.second)
      badshape += " " + kv.first + shape_str(it->second.shape) + "!=" + shape_str(kv.second);
  }
  for (auto& kv : m.staged)
    if (!want.count(kv.first)) unexpected += " " + kv.first;
  if (!missing.empty() || !unexpected.empty() || !badshape.empty()) {
    std::string msg = "Error(s) in loading state_dict for VapGPT:";
    if (!missing.empty()) msg += " Missing key(s):" + missing + ".";
    if (!unexpected.empty()) msg += " Unexpected key(s):" + unexpected + ".";
    if (!badshape.empty()) msg += " size mismatch:" + badshape + ".";
    return fail(m, VAPB_E_STATE, msg);
  }
  {  // the kernels hard-code the codebook's bit semantics (vap/objective.py:93-110)
    const auto& cb = m.staged["objective.codebook.emb.weight"].data;
    for (int c = 0; c < kClasses; ++c)
      for (int b = 0; b < 8; ++b)
        if (cb[c * 8 + b] != (float)((c >> b) & 1))
          return fail(m, VAPB_E_UNSUPPORTED, "objective.codebook.emb.weight is not the LSB-first bit table");
  }

  // ---- repack (fp32 path) into one arena
  auto T = [&](const std::string& k) -> const HostTensor& { return m.staged[k]; };
  Packer pk;
  struct Fix { const void** slot; size_t off; };
  std::vector<Fix> fixes;
  struct GemmW { const void** slot; size_t off; int K, N; };  // every fp32 GEMM weight [K][N]: split copy for k_gemm_x3.cu
  std::vector<GemmW> gemm_ws;
  auto put = [&](const void** slot, const std::vector<float>& v, int K, int N) {
    const size_t off = pk.add_f(v);
    fixes.push_back({slot, off});
    gemm_ws.push_back({slot, off, K, N});
  };
  auto putf = [&](const float** slot, const std::vector<float>& v) {
    fixes.push_back({reinterpret_cast<const void**>(slot), pk.add_f(v)});
  };
  Weights& w = m.w32;
  {
    const HostTensor& c0 = T(GE + "conv0.weight");  // (256,1,10) -> [10][256]
    std::vector<float> v(10 * kDim);
    for (int c = 0; c < kDim; ++c)
      for (int k = 0; k < 10; ++k) v[k * kDim + c] = c0.data[c * 10 + k];
    putf(&w.c0_w, v);
    putf(&w.c0_b, T(GE + "conv0.bias").data);
    putf(&w.c0_g, T(GE + "batchNorm0.weight").data);
    putf(&w.c0_be, T(GE + "batchNorm0.bias").data);
  }
  for (int i = 1; i < 5; ++i) {
    const std::string n = std::to_string(i);
    put(&w.conv_w[i], pack_conv(T(GE + "conv" + n + ".weight")), ck[i] * kDim, kDim);
    putf(&w.conv_b[i], T(GE + "conv" + n + ".bias").data);
    putf(&w.conv_g[i], T(GE + "batchNorm" + n + ".weight").data);
    putf(&w.conv_be[i], T(GE + "batchNorm" + n + ".bias").data);
  }
  for (int l = 0; l < m.ar_layers; ++l) {
    const std::string n = std::to_string(l);
    put(&w.rnn_wih[l], transpose_linear(T(AR + "weight_ih_l" + n)), kDim, (int)GH);
    putf(&w.rnn_whh_t[l], transpose_linear(T(AR + "weight_hh_l" + n)));
    const auto& bi = T(AR + "bias_ih_l" + n).data;
    const auto& bh = T(AR + "bias_hh_l" + n).data;
    std::vector<float> bx(GH), bhn(kDim, 0.f);
    for (int64_t i = 0; i < GH; ++i) bx[i] = bi[i] + bh[i];
    if (m.ar_kind == 1)
      for (int j = 0; j < kDim; ++j) {
        bx[2 * kDim + j] = bi[2 * kDim + j];  // n gate: b_hn stays inside r*(...)
        bhn[j] = bh[2 * kDim + j];
      }
    putf(&w.rnn_bx[l], bx);
    putf(&w.rnn_bhn[l], bhn);
  }
  put(&w.ds_w, pack_conv(T("encoder.downsample.1.weight")), 5 * kDim, kDim);
  putf(&w.ds_b, T("encoder.downsample.1.bias").data);
  putf(&w.ds_g, T("encoder.downsample.2.ln.weight").data);
  putf(&w.ds_be, T("encoder.downsample.2.ln.bias").data);
  auto pack_layer = [&](const std::string& p, bool cross, LayerW& lw) {
    putf(&lw.ln_sa_g, T(p + "ln_self_attn.weight").data);
    putf(&lw.ln_sa_b, T(p + "ln_self_attn.bias").data);
    putf(&lw.ln_ffn_g, T(p + "ln_ffnetwork.weight").data);
    putf(&lw.ln_ffn_b, T(p + "ln_ffnetwork.bias").data);
    putf(&lw.slopes, T(p + "mha.m").data);
    put(&lw.wqkv, concat_linear({&T(p + "mha.query.weight"), &T(p + "mha.key.weight"), &T(p + "mha.value.weight")}), kDim, 3 * kDim);
    put(&lw.wproj, transpose_linear(T(p + "mha.proj.weight")), kDim, kDim);
    if (cross) {
      putf(&lw.ln_src_g, T(p + "ln_src_attn.weight").data);
      putf(&lw.ln_src_b, T(p + "ln_src_attn.bias").data);
      putf(&lw.slopes_cross, T(p + "mha_cross.m").data);
      put(&lw.wq_c, transpose_linear(T(p + "mha_cross.query.weight")), kDim, kDim);
      put(&lw.wkv_c, concat_linear({&T(p + "mha_cross.key.weight"), &T(p + "mha_cross.value.weight")}), kDim, 2 * kDim);
      put(&lw.wproj_c, transpose_linear(T(p + "mha_cross.proj.weight")), kDim, kDim);
    }
    put(&lw.w1, transpose_linear(T(p + "ffnetwork.0.weight")), kDim, kFfn);
    put(&lw.w2, transpose_linear(T(p + "ffnetwork.3.weight")), kFfn, kDim);
  };
  for (int l = 0; l < m.channel_layers; ++l) pack_layer("ar_channel.layers." + std::to_string(l) + ".", false, w.chan[l]);
  for (int l = 0; l < m.cross_layers; ++l) pack_layer("ar.layers." + std::to_string(l) + ".", true, w.cross[l]);
  put(&w.comb_a, transpose_linear(T("ar.combinator.h0_a.weight")), kDim, kDim);
  put(&w.comb_b, transpose_linear(T("ar.combinator.h0_b.weight")), kDim, kDim);
  putf(&w.comb_g, T("ar.combinator.ln.weight").data);
  putf(&w.comb_be, T("ar.combinator.ln.bias").data);
  putf(&w.va_w, T("va_classifier.weight").data);
  putf(&w.va_b, T("va_classifier.bias").data);
  put(&w.head_w, transpose_linear(T("vap_head.weight")), kDim, kClasses);
  putf(&w.head_b, T("vap_head.bias").data);

  CUDA_OK(m, cudaSetDevice(m.device));
  m.arena_bytes = pk.host.size();
  CUDA_OK(m, cudaMalloc(&m.arena, m.arena_bytes));
  CUDA_OK(m, cudaMemcpy(m.arena, pk.host.data(), m.arena_bytes, cudaMemcpyHostToDevice));
  for (auto& f : fixes) *f.slot = static_cast<char*>(m.arena) + f.off;
  if (m.fp32_tc) {  // fp16 hi / lo images of every GEMM weight for the tensor-core parity GEMM
    std::vector<__half> all, one;
    std::vector<size_t> offs;
    for (auto& gw : gemm_ws) {
      x3_pack_weight(reinterpret_cast<const float*>(pk.host.data() + gw.off), gw.K, gw.N, &one);
      offs.push_back(all.size());
      all.insert(all.end(), one.begin(), one.end());
    }
    CUDA_OK(m, cudaMalloc(&m.x3_arena, all.size() * sizeof(__half)));
    CUDA_OK(m, cudaMemcpy(m.x3_arena, all.data(), all.size() * sizeof(__half), cudaMemcpyHostToDevice));
    for (size_t i = 0; i < gemm_ws.size(); ++i)
      m.x3_w[*gemm_ws[i].slot] = static_cast<const __half*>(m.x3_arena) + offs[i];
  }

  const int rc = bf16_prepare(m);
  if (rc != 0) return rc;
  m.finalized = true;
  m.staged.clear();
  return VAPB_OK;
}

int vapb_destroy(VapbHandle* h) {
  if (!h) return VAPB_OK;
  for (auto& r : h->m.prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->m.event_pool) cudaEventDestroy(e);
  for (auto e : h->m.pipe_conv) cudaEventDestroy(e);
  for (auto e : h->m.pipe_join) cudaEventDestroy(e);
  if (h->m.pipe_fork) cudaEventDestroy(h->m.pipe_fork);
  for (auto st : h->m.pipe_st) cudaStreamDestroy(st);
  bf16_release(h->m);
  if (h->m.arena) cudaFree(h->m.arena);
  if (h->m.x3_arena) cudaFree(h->m.x3_arena);
  delete h;
  return VAPB_OK;
}

int vapb_describe(const VapbHandle* h, int* ar_kind, int* ar_layers, int* channel_layers, int* cross_layers,
                  int* num_heads) {
  if (!h || !h->m.finalized) return VAPB_E_STATE;
  if (ar_kind) *ar_kind = h->m.ar_kind;
  if (ar_layers) *ar_layers = h->m.ar_layers;
  if (channel_layers) *channel_layers = h->m.channel_layers;
  if (cross_layers) *cross_layers = h->m.cross_layers;
  if (num_heads) *num_heads = h->m.num_heads;
  return VAPB_OK;
}

int vapb_frames(int64_t n_samples, int64_t* frames100, int64_t* frames50) {
  Geometry g;
  if (n_samples < 1 || make_geometry(1, n_samples, &g) != 0) return VAPB_E_INVALID;
  if (frames100) *frames100 = g.L[4];
  if (frames50) *frames50 = g.T;
  return VAPB_OK;
}

}  // extern "C"

namespace {

struct Aux {  // api-level scratch appended to the path workspace
  size_t logits, vad_sig, lse, bytes;
};
struct Group {  // items [b0, b0 + g.batch) of a pipelined call and where their path workspace starts
  int b0;
  Geometry g;
  size_t off;
};
struct CallPlan {
  Geometry g;
  std::vector<Group> groups;  // size 1: not pipelined
  Aux aux;
};

int n_groups_for(const Model& m, int batch, int mode) {
  if (mode == VAPB_MODE_FP32 || mode == VAPB_MODE_FP32_TC || m.pipe <= 1) return 1;
  int n = batch / (m.pipe_min_items > 0 ? m.pipe_min_items : 1);
  if (n > m.pipe) n = m.pipe;
  return n < 1 ? 1 : n;
}

int plan_all(const Model& m, int batch, int64_t n_samples, int mode, CallPlan* cp, std::string* err) {
  Geometry* g = &cp->g;
  Aux* aux = &cp->aux;
  if (batch < 1 || batch > 16384) { *err = "batch out of range"; return VAPB_E_INVALID; }
  if (mode != VAPB_MODE_FP32 && mode != VAPB_MODE_BF16 && mode != VAPB_MODE_FP16 && mode != VAPB_MODE_FP32_TC) { *err = "unknown mode"; return VAPB_E_INVALID; }
  if (mode == VAPB_MODE_FP32_TC && m.x3_w.empty()) { *err = "mode FP32_TC is switched off (VAPB_FP32_TC=0)"; return VAPB_E_UNSUPPORTED; }
  if (n_samples < 1 || make_geometry(batch, n_samples, g) != 0 || g->T < 1) {
    *err = "n_samples too small for the conv chain";
    return VAPB_E_INVALID;
  }
  if ((long long)g->nseq * g->L[1] > 2000000000LL) { *err = "batch * n_samples too large for one call"; return VAPB_E_INVALID; }
  const bool f32path = mode == VAPB_MODE_FP32 || mode == VAPB_MODE_FP32_TC;
  size_t path_bytes = f32path ? workspace_bytes_fp32(m, *g) : workspace_bytes_bf16(m, *g);
  if (path_bytes == 0) { *err = "mode not available in this build"; return VAPB_E_UNSUPPORTED; }
  const int ng = n_groups_for(m, batch, mode);
  cp->groups.clear();
  if (ng <= 1) {
    cp->groups.push_back(Group{0, *g, 0});
  } else {
    size_t off = 0;
    for (int k = 0; k < ng; ++k) {
      Group gr{};
      gr.b0 = (int)((long long)batch * k / ng);
      const int b1 = (int)((long long)batch * (k + 1) / ng);
      if (make_geometry(b1 - gr.b0, n_samples, &gr.g) != 0) { *err = "bad group geometry"; return VAPB_E_INVALID; }
      gr.off = off;
      off = (off + workspace_bytes_bf16(m, gr.g) + 1023) / 1024 * 1024;
      cp->groups.push_back(gr);
    }
    if (off > path_bytes) path_bytes = off;  // the unsplit layout (used while profiling) fits too
  }
  size_t off = (path_bytes + 1023) / 1024 * 1024;
  const size_t rows = (size_t)batch * g->T;
  aux->logits = off;  off += (rows * kClasses * 4 + 1023) / 1024 * 1024;
  aux->vad_sig = off; off += (rows * 2 * 4 + 1023) / 1024 * 1024;
  aux->lse = off;     off += (rows * 4 + 1023) / 1024 * 1024;
  aux->bytes = off;
  return 0;
}

int run_forward(Model& m, cudaStream_t st, const float* wav, const Geometry& g, int mode, char* ws, float* logits,
                float* vad_logits, float* vad_sig, cudaEvent_t conv_wait = nullptr, cudaEvent_t conv_done = nullptr,
                int wav_pcm16 = 0, const HeadOut* head = nullptr) {
  const float* comb = nullptr;
  if (mode == VAPB_MODE_FP32 || mode == VAPB_MODE_FP32_TC) {
    if (wav_pcm16) { m.err = "int16 PCM input is read by the 16-bit modes only (vapb_pcm16_to_f32 converts for fp32)"; return VAPB_E_UNSUPPORTED; }
    return forward_fp32(m, st, wav, g, ws, logits, vad_logits, vad_sig, &comb, nullptr, mode == VAPB_MODE_FP32_TC);
  }
  const int rc = forward_bf16(m, st, wav, g, ws, logits, vad_logits, vad_sig, &comb, mode == VAPB_MODE_FP16, conv_wait,
                              conv_done, wav_pcm16, head);
  return rc == -4 ? VAPB_E_UNSUPPORTED : rc;
}

// Streams and events of the item-group pipeline (created on first use, owned by the handle).
int ensure_pipe(Model& m, int ng) {
  if (!m.pipe_fork && cudaEventCreateWithFlags(&m.pipe_fork, cudaEventDisableTiming) != cudaSuccess) return -1;
  while ((int)m.pipe_st.size() < ng - 1) {
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return -1;
    m.pipe_st.push_back(s);
  }
  while ((int)m.pipe_conv.size() < ng) {
    cudaEvent_t a, b;
    if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess) return -1;
    m.pipe_conv.push_back(a);
    if (cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) return -1;
    m.pipe_join.push_back(b);
  }
  return 0;
}

struct ProbsOut {  // vapb_probs's optional outputs; every pointer addresses item 0 of the call
  float *logits, *vad_logits, *probs, *vad, *p_now, *p_future, *H, *loss;
  uint8_t* argmax;
  int now_lo, now_hi, fut_lo, fut_hi;
  bool want_probs;  // false: vapb_forward (logits + vad logits only)
  unsigned long long* counters = nullptr;  // [258] accumulated: arg-max class histogram, active frames per channel
  int wav_pcm16 = 0;                       // the waveform buffer holds int16 PCM
};

// Forward (+ the probs()/loss kernels) of items [b0, b0 + g.batch) on `st`.
int run_items(Model& m, cudaStream_t st, const float* wav, int b0, const Geometry& g, int mode, char* ws_path,
              char* ws_all, const Aux& aux, const ProbsOut& o, cudaEvent_t conv_wait, cudaEvent_t conv_done) {
  const long long T = g.T, r0 = (long long)b0 * T;
  auto at = [](float* p, long long n) { return p ? p + n : nullptr; };
  float* lg = o.logits ? o.logits + r0 * kClasses : reinterpret_cast<float*>(ws_all + aux.logits) + r0 * kClasses;
  float* vs = nullptr;
  if (o.want_probs) vs = o.vad ? o.vad + r0 * 2 : reinterpret_cast<float*>(ws_all + aux.vad_sig) + r0 * 2;
  float* lse = reinterpret_cast<float*>(ws_all + aux.lse) + r0;
  const float* wav_b0 = o.wav_pcm16 ? reinterpret_cast<const float*>(reinterpret_cast<const int16_t*>(wav) + (long long)b0 * 2 * g.S)
                                    : wav + (long long)b0 * 2 * g.S;
  // 16-bit modes: the head GEMM produces the probs() outputs itself (k_head_fused.cu); logits reach memory only for
  // the caller (vapb_probs' logits argument) or the loss kernel
  const bool fused = o.want_probs && m.head_fused && (mode == VAPB_MODE_BF16 || mode == VAPB_MODE_FP16);
  HeadOut ho{o.now_lo, o.now_hi, o.fut_lo, o.fut_hi, (o.logits || o.loss) ? lg : nullptr, at(o.probs, r0 * kClasses),
             at(o.p_now, r0 * 2), at(o.p_future, r0 * 2), at(o.H, r0), o.loss ? lse : nullptr,
             o.argmax ? o.argmax + r0 : nullptr, o.counters};
  const int rc = run_forward(m, st, wav_b0, g, mode, ws_path, lg, at(o.vad_logits, r0 * 2), vs, conv_wait, conv_done,
                             o.wav_pcm16, fused ? &ho : nullptr);
  if (rc || !o.want_probs) return rc;
  const long long rows = (long long)g.batch * T;
  ProfScope ps(m, st, CAT_HEADS);
  if (fused) {
    if (o.loss) m.launches += launch_loss(st, lg, vs, lse, g.batch, (int)T, o.loss + (long long)b0 * (T - 100));
    return 0;
  }
  m.launches += launch_probs(st, lg, rows, o.now_lo, o.now_hi, o.fut_lo, o.fut_hi, at(o.probs, r0 * kClasses),
                             at(o.p_now, r0 * 2), at(o.p_future, r0 * 2), at(o.H, r0), o.loss ? lse : nullptr,
                             o.argmax ? o.argmax + r0 : nullptr, o.counters, o.counters ? vs : nullptr);
  if (o.loss) m.launches += launch_loss(st, lg, vs, lse, g.batch, (int)T, o.loss + (long long)b0 * (T - 100));
  return 0;
}

// One call = one group on the caller's stream, or the item-group pipeline (model.h) forked from / joined to it.
int run_call(Model& m, cudaStream_t st, const float* wav, const CallPlan& cp, int mode, char* ws, const ProbsOut& o) {
  const int ng = (int)cp.groups.size();
  if (ng <= 1 || m.profiling)  // per-family event timing needs one stream
    return run_items(m, st, wav, 0, cp.g, mode, ws, ws, cp.aux, o, nullptr, nullptr);
  if (ensure_pipe(m, ng) != 0) { m.err = "cannot create the pipeline's streams/events"; return VAPB_E_CUDA; }
  cudaEventRecord(m.pipe_fork, st);
  m.trace(st, "fork");
  int rc = 0;
  for (int k = 0; k < ng; ++k) {
    const Group& gr = cp.groups[k];
    cudaStream_t sk = k == 0 ? st : m.pipe_st[k - 1];
    if (k) cudaStreamWaitEvent(sk, m.pipe_fork, 0);
    m.trace_group = k;
    if (!rc)
      rc = run_items(m, sk, wav, gr.b0, gr.g, mode, ws + gr.off, ws, cp.aux, o, k ? m.pipe_conv[k - 1] : nullptr,
                     m.pipe_conv[k]);
    m.trace(sk, "g" + std::to_string(k) + " end");
    if (k) cudaEventRecord(m.pipe_join[k], sk);
  }
  for (int k = 1; k < ng; ++k) cudaStreamWaitEvent(st, m.pipe_join[k], 0);  // always join what was forked
  if (m.pipe_trace && !m.trace_ev.empty()) {
    cudaStreamSynchronize(st);
    for (auto& te : m.trace_ev) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, m.trace_ev[0].second, te.second);
      fprintf(stderr, "[pipe] %-16s %8.3f ms\n", te.first.c_str(), ms);
    }
    for (auto& te : m.trace_ev) cudaEventDestroy(te.second);
    m.trace_ev.clear();
  }
  return rc;
}

}  // namespace

extern "C" {

int vapb_workspace_bytes(const VapbHandle* h, int batch, int64_t n_samples, int mode, size_t* bytes) {
  if (!h || !bytes) return VAPB_E_INVALID;
  Model& m = const_cast<Model&>(h->m);
  if (!m.finalized) return fail(m, VAPB_E_STATE, "vapb_finalize has not been called");
  CallPlan cp;
  std::string err;
  const int rc = plan_all(m, batch, n_samples, mode, &cp, &err);
  if (rc) return fail(m, rc, err);
  *bytes = cp.aux.bytes;
  return VAPB_OK;
}

int vapb_forward(VapbHandle* h, void* stream, const float* wav, int batch, int64_t n_samples, int mode,
                 void* workspace, size_t workspace_bytes, float* logits, float* vad_logits) {
  if (!h || !wav || !workspace || !logits || !vad_logits) return VAPB_E_INVALID;
  Model& m = h->m;
  if (!m.finalized) return fail(m, VAPB_E_STATE, "vapb_finalize has not been called");
  CallPlan cp;
  std::string err;
  int rc = plan_all(m, batch, n_samples, mode, &cp, &err);
  if (rc) return fail(m, rc, err);
  if (workspace_bytes < cp.aux.bytes) return fail(m, VAPB_E_WORKSPACE, "workspace too small");
  CUDA_OK(m, cudaSetDevice(m.device));
  ProbsOut o{};
  o.logits = logits;
  o.vad_logits = vad_logits;
  rc = run_call(m, (cudaStream_t)stream, wav, cp, mode, (char*)workspace, o);
  if (rc) return rc;
  CUDA_OK(m, cudaPeekAtLastError());
  return VAPB_OK;
}

int vapb_forward_attention(VapbHandle* h, void* stream, const float* wav, int batch, int64_t n_samples,
                           void* workspace, size_t workspace_bytes, float* logits, float* vad_logits,
                           float* self_attn, float* cross_attn, float* cross_self_attn) {
  if (!h || !wav || !workspace || !logits || !vad_logits || !self_attn || !cross_attn || !cross_self_attn)
    return VAPB_E_INVALID;
  Model& m = h->m;
  if (!m.finalized) return fail(m, VAPB_E_STATE, "vapb_finalize has not been called");
  CallPlan cp;
  std::string err;
  int rc = plan_all(m, batch, n_samples, VAPB_MODE_FP32, &cp, &err);
  if (rc) return fail(m, rc, err);
  if (workspace_bytes < cp.aux.bytes) return fail(m, VAPB_E_WORKSPACE, "workspace too small");
  CUDA_OK(m, cudaSetDevice(m.device));
  const AttnMaps maps{self_attn, cross_attn, cross_self_attn};
  const float* comb = nullptr;
  rc = forward_fp32(m, (cudaStream_t)stream, wav, cp.g, (char*)workspace, logits, vad_logits, nullptr, &comb, &maps);
  if (rc) return rc;
  CUDA_OK(m, cudaPeekAtLastError());
  return VAPB_OK;
}

int vapb_probs(VapbHandle* h, void* stream, const float* wav, int batch, int64_t n_samples, int mode,
               void* workspace, size_t workspace_bytes, int now_lo, int now_hi, int fut_lo, int fut_hi,
               float* logits, float* vad_logits, float* probs, float* vad, float* p_now, float* p_future, float* H,
               float* loss, uint8_t* argmax) {
  return vapb_probs_ex(h, stream, wav, VAPB_WAV_F32, batch, n_samples, mode, workspace, workspace_bytes, now_lo, now_hi,
                       fut_lo, fut_hi, logits, vad_logits, probs, vad, p_now, p_future, H, loss, argmax, nullptr);
}

int vapb_probs_ex(VapbHandle* h, void* stream, const void* wav_any, int wav_fmt, int batch, int64_t n_samples, int mode,
                  void* workspace, size_t workspace_bytes, int now_lo, int now_hi, int fut_lo, int fut_hi,
                  float* logits, float* vad_logits, float* probs, float* vad, float* p_now, float* p_future, float* H,
                  float* loss, uint8_t* argmax, unsigned long long* counters) {
  const float* wav = static_cast<const float*>(wav_any);
  if (!h || !wav || !workspace) return VAPB_E_INVALID;
  if (wav_fmt != VAPB_WAV_F32 && wav_fmt != VAPB_WAV_PCM16) return VAPB_E_INVALID;
  Model& m = h->m;
  if (!m.finalized) return fail(m, VAPB_E_STATE, "vapb_finalize has not been called");
  if (now_lo < 0 || now_hi > 3 || now_lo > now_hi || fut_lo < 0 || fut_hi > 3 || fut_lo > fut_hi)
    return fail(m, VAPB_E_INVALID, "bin limits must satisfy 0 <= lo <= hi <= 3");
  CallPlan cp;
  std::string err;
  int rc = plan_all(m, batch, n_samples, mode, &cp, &err);
  if (rc) return fail(m, rc, err);
  if (workspace_bytes < cp.aux.bytes) return fail(m, VAPB_E_WORKSPACE, "workspace too small");
  if (loss && cp.g.T <= 100)
    return fail(m, VAPB_E_INVALID, "loss needs more than 100 frames (maximum size for tensor at dimension 1 is " +
                                       std::to_string(cp.g.T - 1) + " but size is 100)");
  CUDA_OK(m, cudaSetDevice(m.device));
  ProbsOut o{logits, vad_logits, probs, vad, p_now, p_future, H, loss, argmax, now_lo, now_hi, fut_lo, fut_hi, true};
  o.counters = counters;
  o.wav_pcm16 = wav_fmt == VAPB_WAV_PCM16;
  rc = run_call(m, (cudaStream_t)stream, wav, cp, mode, (char*)workspace, o);
  if (rc) return rc;
  CUDA_OK(m, cudaPeekAtLastError());
  return VAPB_OK;
}

int vapb_probs_from_logits(VapbHandle* h, void* stream, const float* logits, int64_t rows, int now_lo, int now_hi,
                           int fut_lo, int fut_hi, float* probs, float* p_now, float* p_future, float* H,
                           uint8_t* argmax) {
  if (!h || !logits || rows < 0) return VAPB_E_INVALID;
  Model& m = h->m;
  if (now_lo < 0 || now_hi > 3 || now_lo > now_hi || fut_lo < 0 || fut_hi > 3 || fut_lo > fut_hi)
    return fail(m, VAPB_E_INVALID, "bin limits must satisfy 0 <= lo <= hi <= 3");
  if (rows == 0) return VAPB_OK;
  CUDA_OK(m, cudaSetDevice(m.device));
  m.launches += launch_probs((cudaStream_t)stream, logits, rows, now_lo, now_hi, fut_lo, fut_hi, probs, p_now,
                             p_future, H, nullptr, argmax, nullptr, nullptr);
  CUDA_OK(m, cudaPeekAtLastError());
  return VAPB_OK;
}

int vapb_get_stage(VapbHandle* h, void* stream, const char* name, int batch, int64_t n_samples, int mode,
                   void* workspace, size_t workspace_bytes, float* out, size_t out_elems) {
  if (!h || !name || !workspace || !out) return VAPB_E_INVALID;
  Model& m = h->m;
  if (!m.finalized) return fail(m, VAPB_E_STATE, "vapb_finalize has not been called");
  CallPlan cp;
  std::string err;
  int rc = plan_all(m, batch, n_samples, mode, &cp, &err);
  if (rc) return fail(m, rc, err);
  const Geometry& g = cp.g;
  if (workspace_bytes < cp.aux.bytes) return fail(m, VAPB_E_WORKSPACE, "workspace too small");
  if (cp.groups.size() > 1)
    return fail(m, VAPB_E_UNSUPPORTED, "stage export is not available for a pipelined call (batch >= " +
                                           std::to_string(2 * m.pipe_min_items) + "); use a smaller batch or VAPB_PIPE=1");
  StageRef ref{};
  rc = (mode == VAPB_MODE_FP32 || mode == VAPB_MODE_FP32_TC) ? stage_fp32(m, g, (char*)workspace, name, &ref)
                              : stage_bf16(m, g, (char*)workspace, name, &ref);
  if (rc) return fail(m, VAPB_E_INVALID, std::string("unknown stage ") + name);
  if (out_elems < (size_t)ref.nseq * ref.rows_per_seq * kDim) return fail(m, VAPB_E_INVALID, "stage output too small");
  CUDA_OK(m, cudaSetDevice(m.device));
  m.launches += launch_to_f32((cudaStream_t)stream, ref.ptr, ref.is_bf16 ? (mode == VAPB_MODE_FP16 ? 2 : 1) : 0, ref.map, ref.nseq, ref.rows_per_seq, out, ref.blocked);
  CUDA_OK(m, cudaPeekAtLastError());
  return VAPB_OK;
}

int vapb_profile_begin(VapbHandle* h) {
  if (!h) return VAPB_E_INVALID;
  Model& m = h->m;
  for (auto& r : m.prof) { m.event_pool.push_back(r.a); m.event_pool.push_back(r.b); }
  m.prof.clear();
  m.profiling = true;
  return VAPB_OK;
}

int vapb_profile_end(VapbHandle* h, double* ms, uint64_t* launches) {
  if (!h || !ms || !launches) return VAPB_E_INVALID;
  Model& m = h->m;
  m.profiling = false;
  for (int i = 0; i < CAT_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
  CUDA_OK(m, cudaSetDevice(m.device));
  for (auto& r : m.prof) {
    CUDA_OK(m, cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_OK(m, cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cat] += t;
    launches[r.cat] += r.launches;
    m.event_pool.push_back(r.a);
    m.event_pool.push_back(r.b);
  }
  m.prof.clear();
  return VAPB_OK;
}

int vapb_debug_gemm_x3(void* stream, const float* A, int64_t a_seq_stride, int64_t a_row_stride, const float* Wt_host,
                       int nseq, int rows_per_seq, int N, int K, const float* bias, int norm1, const float* g1,
                       const float* b1, int act, const float* resid, int accumulate, float* out1, int norm2,
                       const float* g2, const float* b2, float* out2, char* err, int err_len) {
  std::string msg;
  int rc = -1;
  void* dw = nullptr;
  if (!A || !Wt_host || !out1 || N % 256 || K % 32) {
    msg = "gemm_x3: invalid argument";
  } else {
    std::vector<__half> packed;
    x3_pack_weight(Wt_host, K, N, &packed);
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaMalloc(&dw, packed.size() * sizeof(__half)) != cudaSuccess ||
        cudaMemcpy(dw, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) {
      msg = "gemm_x3: cannot stage the weight";
    } else {
      GemmProblem p{A, RowMap{a_seq_stride, a_row_stride}, nullptr, nseq * rows_per_seq, rows_per_seq, N, K};
      const RowMap dense{(long long)rows_per_seq * N, N};
      Epilogue e{};
      e.bias = bias;
      e.norm1 = norm1; e.g1 = g1; e.b1 = b1;
      e.act = act;
      e.resid = resid; e.resid_map = dense;
      e.accumulate = accumulate;
      e.out1 = out1; e.out1_map = dense;
      e.norm2 = norm2; e.g2 = g2; e.b2 = b2;
      e.out2 = out2; e.out2_map = dense;
      rc = launch_gemm_x3((cudaStream_t)stream, p, e, dw, n_sm, &msg);
      if (rc >= 0) {
        cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
        if (ce != cudaSuccess) { msg = cudaGetErrorString(ce); rc = -1; }
      }
    }
  }
  if (dw) cudaFree(dw);
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_gemm_2sm(void* stream, const void* A, int64_t a_seq_stride, int64_t a_row_stride, const void* W, int nseq,
                        int rows_per_seq, int K, const float* bias, int norm1, const float* g1, const float* b1, int act,
                        void* out_bf16, char* err, int err_len) {
  TcGemmArgs a{};
  a.A = A;
  a.a_map = RowMap{a_seq_stride, a_row_stride};
  a.W = W;
  a.nseq = nseq;
  a.rows_per_seq = rows_per_seq;
  a.N = 256;
  a.K = K;
  a.e.bias = bias;
  a.e.norm1 = norm1; a.e.g1 = g1; a.e.b1 = b1;
  a.e.act = act;
  a.e.out1_map = RowMap{(long long)rows_per_seq * 256, 256};
  a.out1_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  std::string msg;
  int rc = launch_gemm_2sm((cudaStream_t)stream, a, n_sm, &msg);
  if (rc >= 0) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
  }
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_conv0_tc(void* stream, const float* wav, int batch, int64_t n_samples, const float* conv0_w,
                        const float* conv0_b, const float* norm0_g, const float* norm0_b, void* out, int fp16, char* err,
                        int err_len) {
  Geometry g;
  std::string msg;
  int rc = -1;
  float* dev_tab = nullptr;
  if (!wav || !conv0_w || !conv0_b || !norm0_g || !norm0_b || !out || batch < 1 || make_geometry(batch, n_samples, &g) != 0) {
    msg = "conv0_tc: invalid argument";
  } else {
    std::vector<float> tab(12 * kDim);
    Conv0Stats cs;
    conv0_v2_fold(conv0_w, conv0_b, norm0_g, tab.data(), tab.data() + 10 * kDim, &cs);
    memcpy(tab.data() + 11 * kDim, norm0_b, kDim * sizeof(float));
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaMalloc(&dev_tab, tab.size() * 4) != cudaSuccess ||
        cudaMemcpy(dev_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      msg = "conv0_tc: cannot stage the conv0 table";
    } else {
      const int prev = g_fp16;
      g_fp16 = fp16 ? 1 : 0;
      rc = launch_conv0_tc((cudaStream_t)stream, wav, batch, n_samples, 0, 2 * batch, g.L[0], dev_tab, dev_tab + 10 * kDim,
                           dev_tab + 11 * kDim, cs, reinterpret_cast<__nv_bfloat16*>(out), g.L[0] * kDim, 0, n_sm, &msg);
      g_fp16 = prev;
      if (rc >= 0) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
      }
    }
  }
  if (dev_tab) cudaFree(dev_tab);
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_conv01(void* stream, const float* wav, int batch, int64_t n_samples, const float* conv0_w,
                      const float* conv0_b, const float* norm0_g, const float* norm0_b, const void* w1,
                      const float* bias1, const float* g1, const float* b1, void* out, int64_t out_seq_stride,
                      int out_pad_rows, int fp16, char* err, int err_len, long long* dbg_clocks) {
  Geometry g;
  std::string msg;
  int rc = -1;
  float* dev_tab = nullptr;
  if (!wav || !conv0_w || !conv0_b || !norm0_g || !norm0_b || !w1 || !bias1 || !g1 || !b1 || !out || batch < 1 ||
      make_geometry(batch, n_samples, &g) != 0) {
    msg = "conv01: invalid argument";
  } else {
    std::vector<float> tab(12 * kDim);
    Conv0Stats cs;
    conv0_v2_fold(conv0_w, conv0_b, norm0_g, tab.data(), tab.data() + 10 * kDim, &cs);
    memcpy(tab.data() + 11 * kDim, norm0_b, kDim * sizeof(float));
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaMalloc(&dev_tab, tab.size() * 4) != cudaSuccess ||
        cudaMemcpy(dev_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      msg = "conv01: cannot stage the conv0 table";
    } else {
      const int prev = g_fp16;
      g_fp16 = fp16 ? 1 : 0;
      rc = launch_conv01((cudaStream_t)stream, wav, 0, batch, n_samples, 0, 2 * batch, g.L[0], g.L[1], tab.data(), dev_tab,
                         cs, w1, bias1, g1, b1, out, out_seq_stride, out_pad_rows, n_sm, &msg, dbg_clocks);
      g_fp16 = prev;
      if (rc >= 0) {
        cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
      }
    }
  }
  if (dev_tab) cudaFree(dev_tab);
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_gemm_lin(void* stream, const void* A, int64_t a_seq_stride, int64_t a_row_stride, const void* W, int nseq,
                        int rows_per_seq, int N, int K, const float* bias, int norm1, const float* g1, const float* b1,
                        int act, const float* resid_blocked, int accumulate, float* out1_f32, int f32_mode,
                        void* out1_bf16, int norm2, const float* g2, const float* b2, void* out2_bf16, char* err,
                        int err_len) {
  TcGemmArgs a{};
  a.A = A;
  a.a_map = RowMap{a_seq_stride, a_row_stride};
  a.W = W;
  a.nseq = nseq;
  a.rows_per_seq = rows_per_seq;
  a.N = N;
  a.K = K;
  const RowMap dense{(long long)rows_per_seq * N, N};
  a.e.bias = bias;
  a.e.norm1 = norm1; a.e.g1 = g1; a.e.b1 = b1;
  a.e.act = act;
  a.e.resid = resid_blocked;
  a.e.accumulate = accumulate;
  a.e.out1_map = dense;
  a.e.norm2 = norm2; a.e.g2 = g2; a.e.b2 = b2;
  a.e.out2 = out2_bf16; a.e.out2_map = dense;
  a.out1_f32 = out1_f32;
  a.out1_bf16 = reinterpret_cast<__nv_bfloat16*>(out1_bf16);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  std::string msg;
  int rc = launch_gemm_lin((cudaStream_t)stream, a, f32_mode, n_sm, &msg);
  if (rc >= 0) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
  }
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_ffn_fused(void* stream, const void* z, const void* w1, const void* w2, const float* resid_blocked,
                          float* x_out_blocked, void* xs, void* zn, const float* g2, const float* b2, int M, char* err,
                          int err_len, long long* dbg_clocks) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  std::string msg;
  typedef const __nv_bfloat16* bp;
  int rc = launch_ffn_fused((cudaStream_t)stream, (bp)z, (bp)w1, (bp)w2, resid_blocked, x_out_blocked,
                            reinterpret_cast<__nv_bfloat16*>(xs), reinterpret_cast<__nv_bfloat16*>(zn), g2, b2, M, n_sm,
                            &msg, dbg_clocks);
  if (rc >= 0) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
  }
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_rnn_pack(int kind, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                        float* w_cat, float* bias) {
  if (!w_ih || !w_hh || !b_ih || !b_hh || !w_cat || !bias || (kind != 0 && kind != 1)) return VAPB_E_INVALID;
  rnn_tc_pack(kind, w_ih, w_hh, b_ih, b_hh, w_cat, bias);
  return VAPB_OK;
}

int vapb_debug_rnn_tc(void* stream, int kind, const void* x, int64_t x_seq_stride, int64_t x_row_stride,
                      const void* w_cat, const float* bias, void* out, int64_t out_seq_stride, int nseq, int T,
                      char* err, int err_len, long long* dbg_clocks, int groups) {
  std::string msg;
  int rc = launch_rnn_tc((cudaStream_t)stream, kind, reinterpret_cast<const __nv_bfloat16*>(x), x_seq_stride,
                         x_row_stride, reinterpret_cast<const __nv_bfloat16*>(w_cat), bias,
                         reinterpret_cast<__nv_bfloat16*>(out), out_seq_stride, nseq, T, &msg, dbg_clocks, groups);
  if (rc >= 0) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
  }
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_attn_tc(void* stream, const void* q, int64_t q_row_stride, const void* k, const void* v,
                       int64_t kv_row_stride, void* out, int nseq, int T, int n_heads, const float* slopes,
                       int cross, char* err, int err_len, long long* dbg_clocks) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  std::string msg;
  typedef const __nv_bfloat16* bp;
  const int prev = g_fp16;
  if (const char* e = getenv("VAPB_DEBUG_FP16")) g_fp16 = atoi(e) ? 1 : 0;  // the buffers then hold fp16 instead of bf16
  int rc = launch_attention_tc((cudaStream_t)stream, (bp)q, q_row_stride, (bp)k, (bp)v, kv_row_stride,
                               reinterpret_cast<__nv_bfloat16*>(out), nseq, T, n_heads, slopes, cross, n_sm, &msg, dbg_clocks);
  g_fp16 = prev;
  if (rc >= 0) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
  }
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_debug_attn_x3(void* stream, const float* qbuf, int q_cols, const float* kvbuf, int kv_cols, int k_off,
                       int v_off, void* planes, float* out, int nseq, int T, const float* slopes, int cross, char* err,
                       int err_len) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  std::string msg;
  int rc = launch_attention_x3((cudaStream_t)stream, qbuf, q_cols, kvbuf, kv_cols, k_off, v_off, planes, out, nseq, T, 4,
                               slopes, cross, n_sm, &msg);
  if (rc >= 0) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { msg = cudaGetErrorString(e); rc = -1; }
  }
  if (rc < 0) {
    if (err && err_len > 0) snprintf(err, err_len, "%s", msg.c_str());
    return VAPB_E_CUDA;
  }
  return VAPB_OK;
}

int vapb_vad_filter(VapbHandle* h, void* stream, const float* vad01, int batch, int64_t T, int max_fill_frames,
                    int max_omit_frames, float* out) {
  return vapb_vad_filter_ex(h, stream, vad01, 0, 0.5f, batch, T, max_fill_frames, max_omit_frames, out);
}

int vapb_vad_filter_ex(VapbHandle* h, void* stream, const float* vad01, int from_logits, float cutoff, int batch, int64_t T,
                       int max_fill_frames, int max_omit_frames, float* out) {
  if (!vad01 || !out || batch < 0 || T < 0 || T > 0x3fffffff || max_fill_frames < 0 || max_omit_frames < 0) {
    if (h) h->m.err = "vad_filter: invalid argument"; else g_create_err = "vad_filter: invalid argument";
    return VAPB_E_INVALID;
  }
  if (h) cudaSetDevice(h->m.device);
  const int n = launch_vad_filter((cudaStream_t)stream, vad01, batch, (int)T, max_fill_frames, max_omit_frames, out,
                                  from_logits, cutoff);
  if (h) h->m.launches += n;
  return cudaPeekAtLastError() == cudaSuccess ? VAPB_OK : VAPB_E_CUDA;
}

int vapb_zero_shot(VapbHandle* h, void* stream, const float* x, int is_probs, int64_t batch, int64_t T,
                   const float* va, int64_t va_T, const uint32_t* class_sets, float* p, float* p_bc, float* p_sil,
                   float* p_act) {
  const char* e = nullptr;
  if (!x || !class_sets) e = "zero_shot: NULL input";
  else if (batch < 0 || T < 0 || T > 0x3fffffff) e = "zero_shot: invalid shape";
  else if (p && (!va || va_T < T)) e = "zero_shot: p needs voice activity of at least T frames per item";
  if (e) {
    if (h) h->m.err = e; else g_create_err = e;
    return VAPB_E_INVALID;
  }
  if (h) cudaSetDevice(h->m.device);
  const int n = launch_zero_shot((cudaStream_t)stream, x, is_probs, batch, (int)T, va, va_T, class_sets, p, p_bc,
                                 p_sil, p_act);
  if (n < 0) {
    e = "zero_shot: a class set holds more than 64 classes";
    if (h) h->m.err = e; else g_create_err = e;
    return VAPB_E_UNSUPPORTED;
  }
  if (h) h->m.launches += n;
  return cudaPeekAtLastError() == cudaSuccess ? VAPB_OK : VAPB_E_CUDA;
}

int vapb_resample(VapbHandle* h, void* stream, const void* x, int x_fmt, int64_t items, int channels, int64_t n_in,
                  int64_t item_stride, int64_t chan_stride, int64_t elem_stride, int orig, int new_rate, int width,
                  const float* bank, float* out, int64_t n_out, int64_t out_row_stride) {
  std::string err;
  int rc = VAPB_OK;
  if (!x || !bank || !out) {
    err = "resample: NULL buffer";
    rc = VAPB_E_INVALID;
  } else {
    if (h) cudaSetDevice(h->m.device);
    const int n = launch_resample((cudaStream_t)stream, x, x_fmt, items, channels, n_in, item_stride, chan_stride,
                                  elem_stride, orig, new_rate, width, bank, out, n_out, out_row_stride, &err);
    if (n < 0) rc = VAPB_E_INVALID;
    else if (h) h->m.launches += n;
  }
  if (rc) {
    if (h) h->m.err = err; else g_create_err = err;
  }
  return rc;
}

int vapb_memset_zero(void* stream, void* ptr, size_t bytes) {
  if (!ptr) return VAPB_E_INVALID;
  return cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream) == cudaSuccess ? VAPB_OK : VAPB_E_CUDA;
}

int vapb_pcm16_to_f32(void* stream, const int16_t* pcm, int64_t n, float* out) {
  if (!pcm || !out || n < 0) return VAPB_E_INVALID;
  if (n) launch_pcm16_to_f32((cudaStream_t)stream, pcm, n, out);
  return cudaPeekAtLastError() == cudaSuccess ? VAPB_OK : VAPB_E_CUDA;
}

int vapb_launch_count(const VapbHandle* h, uint64_t* launches) {
  if (!h || !launches) return VAPB_E_INVALID;
  *launches = h->m.launches;
  return VAPB_OK;
}

}  // extern "C"
