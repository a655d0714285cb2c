// BF16 tensor-core orchestration of VapGPT.forward (same call stack as
// forward_fp32.cu; reference vap/model.py:249-268).
//
// Every contraction runs on tcgen05 (k_conv01.cu, k_gemm_2sm.cu, k_gemm_lin.cu, ...) with 16-bit operands and fp32
// accumulation in TMEM; norms, softmax, residual stream and the recurrence state
// stay fp32. Activation buffers that feed a GEMM are bf16 channels-last; the
// residual stream x is fp32 with a bf16 shadow where an un-normalised copy is a
// GEMM operand (cross-attention K/V source, combinator input).
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

#include "model.h"
#include "tc_common.cuh"

namespace vapb {

namespace {

typedef __nv_bfloat16 bf16;

struct Conv { int k, s, p; };
const Conv kConv[5] = {{10, 5, 3}, {8, 4, 2}, {4, 2, 1}, {4, 2, 1}, {4, 2, 1}};

size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

struct LayerW16 {
  const bf16 *wqkv, *wproj, *wq_c, *wkv_c, *wproj_c, *w1, *w2;
};
struct State16 {
  void* arena = nullptr;
  const bf16* conv_w[5];
  const bf16* rnn_wcat[kMaxLayers];   // [1024][512] packed [W_ih | W_hh] (k_rnn_tc.cu)
  const float* rnn_bias[kMaxLayers];  // [1024]
  const bf16* ds_w;
  LayerW16 chan[kMaxLayers], cross[kMaxLayers];
  const bf16 *comb_a, *comb_b, *head_w;
  const float *c0_u, *c0_d;  // folded conv0 + ChannelNorm parameters (tc_host.cu: conv0_v2_fold)
  Conv0Stats c0_stats;
  const float* c0_tab;             // [12][256] = u | d | beta on the device, and the same on the host: the fused
  std::vector<float> c0_tab_host;  // conv0 -> conv1 kernel takes the table as kernel parameters (k_conv01.cu)
};

struct Plan16 {
  int mb;
  long long lo[4], lpad[4], rnn_lpad;
  size_t act[4], act4, rnn[2], stage[2 * kMaxLayers + 1], xs, xa, xb, z, qkv, kvc, qc, y, h, comb, combb,
      bytes;
};

Plan16 make_plan(const Model& m, const Geometry& g) {
  Plan16 p{};
  const bool fused01 = m.conv01 != 0;  // conv0 lives inside conv1's producer: no act[0], act[1] holds whole 128-row tiles
  const long long l1_tiles = (g.L[1] + 127) / 128 * 128;
  const long long per_seq0 = (fused01 ? l1_tiles + 16 : g.L[0] + 16) * kDim * 2;
  long long mb = (m.conv_mb_bytes > 0 ? m.conv_mb_bytes : (4LL << 30)) / per_seq0;
  if (mb < 1) mb = 1;
  if (mb > g.nseq) mb = g.nseq;
  p.mb = (int)mb;
  size_t off = 0;
  for (int i = 0; i < 4; ++i) {
    const Conv& nx = kConv[i + 1];
    p.lo[i] = nx.p;
    const long long need = (long long)nx.s * (g.L[i + 1] - 1) + nx.k;
    long long lp = need > nx.p + g.L[i] ? need : nx.p + g.L[i];
    if (fused01 && i == 1 && lp < nx.p + l1_tiles) lp = nx.p + l1_tiles;
    lp = (lp + nx.s - 1) / nx.s * nx.s;
    p.lpad[i] = lp;
    p.act[i] = off;
    if (!(fused01 && i == 0)) off = align_up(off + (size_t)mb * lp * kDim * 2);
  }
  p.act4 = off;  off = align_up(off + (size_t)g.nseq * g.L[4] * kDim * 2);
  p.rnn_lpad = 4 + g.L[4] + (g.L[4] & 1);  // even, so the stride-2 tensor map strides nest
  for (int i = 0; i < 2; ++i) {
    p.rnn[i] = off;
    if (i == 0 || m.ar_layers > 1) off = align_up(off + (size_t)g.nseq * p.rnn_lpad * kDim * 2);
  }
  const size_t rows = (size_t)g.nseq * g.T;
  const size_t rows_blk = (rows + 127) / 128 * 128;  // fp32 residual-stream buffers use the row-blocked layout
  const int n_stage = 1 + m.channel_layers + m.cross_layers;
  for (int i = 0; i < n_stage; ++i) { p.stage[i] = off; off = align_up(off + rows_blk * kDim * 4); }
  p.xs = off;  off = align_up(off + rows * kDim * 2);   // bf16 shadow of the current layer input
  p.xa = off;  off = align_up(off + rows_blk * kDim * 4);
  p.xb = off;  off = align_up(off + rows_blk * kDim * 4);
  p.z = off;   off = align_up(off + rows * kDim * 2);
  p.qkv = off; off = align_up(off + rows * 3 * kDim * 2);
  p.kvc = off; off = align_up(off + rows * 2 * kDim * 2);
  p.qc = off;  off = align_up(off + rows * kDim * 2);
  p.y = off;   off = align_up(off + rows * kDim * 2);
  p.h = off;   off = align_up(off + rows * kFfn * 2);
  p.comb = off;  off = align_up(off + (rows / 2 + 127) / 128 * 128 * kDim * 4);
  p.combb = off; off = align_up(off + rows / 2 * kDim * 2);
  p.bytes = off;
  return p;
}

RowMap dense(long long rows, long long n) { return RowMap{rows * n, n}; }

struct Ctx {
  Model& m;
  cudaStream_t st;
  int rc = 0;
  // row-local layers: staged / blocked epilogue I/O (k_gemm_lin.cu); fp32 rows are blocked unless f32_mode == 2
  void lin(const bf16* A, RowMap amap, const bf16* W, int nseq, int rps, int N, int K, const Epilogue& e,
           float* out_f32, bf16* out_bf16, int f32_mode = 1, int cat = CAT_LINEAR_GEMM) {
    if (rc) return;
    TcGemmArgs a{};
    a.A = A; a.a_map = amap; a.W = W;
    a.nseq = nseq; a.rows_per_seq = rps; a.N = N; a.K = K;
    a.e = e;
    a.out1_f32 = out_f32; a.out1_bf16 = out_bf16;
    ProfScope ps(m, st, cat);
    std::string err;
    const int n = launch_gemm_lin(st, a, f32_mode, m.n_sm, &err);
    if (n < 0) { m.err = err; rc = -3; return; }
    m.launches += n;
  }
};

// 16-bit image of a weight vector: bf16 (fmt 0) or fp16 (fmt 1); both travel as 2-byte words
std::vector<bf16> to_16(const std::vector<float>& v, int fp16) {
  std::vector<bf16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    if (fp16) {
      const __half h = __float2half_rn(v[i]);
      memcpy(&o[i], &h, 2);
    } else {
      o[i] = __float2bfloat16_rn(v[i]);
    }
  }
  return o;
}
struct Fp16Scope {  // the 16-bit format of every launch issued by this thread while it lives
  int prev;
  explicit Fp16Scope(int v) : prev(g_fp16) { g_fp16 = v; }
  ~Fp16Scope() { g_fp16 = prev; }
};

}  // namespace

static int prepare_fmt(Model& m, State16* s, int fp16);

int bf16_prepare(Model& m) {
  State16* s = new State16[2];  // [0] bf16, [1] fp16 images of the same weights
  for (int f = 0; f < 2; ++f) {
    const int rc = prepare_fmt(m, &s[f], f);
    if (rc) {
      for (int k = 0; k < f; ++k) cudaFree(s[k].arena);
      delete[] s;
      return rc;
    }
  }
  m.bf16_state = s;
  return 0;
}

static int prepare_fmt(Model& m, State16* s, int fp16) {
  std::vector<char> host;
  struct Fix { const bf16** slot; size_t off; };
  std::vector<Fix> fixes;
  auto put = [&](const bf16** slot, const std::vector<float>& v) {
    const size_t off = (host.size() + 1023) / 1024 * 1024;
    std::vector<bf16> b = to_16(v, fp16);
    host.resize(off + b.size() * 2);
    memcpy(host.data() + off, b.data(), b.size() * 2);
    fixes.push_back({slot, off});
  };
  auto put_f32 = [&](const float** slot, const std::vector<float>& v) {
    const size_t off = (host.size() + 1023) / 1024 * 1024;
    host.resize(off + v.size() * 4);
    memcpy(host.data() + off, v.data(), v.size() * 4);
    fixes.push_back({reinterpret_cast<const bf16**>(slot), off});
  };
  auto T = [&](const std::string& k) -> const HostTensor& { return m.staged[k]; };
  auto conv_nk = [&](const HostTensor& t) {  // (out,in,k) -> [out][tap*in + cin]
    const int64_t co = t.shape[0], ci = t.shape[1], k = t.shape[2];
    std::vector<float> o((size_t)co * ci * k);
    for (int64_t n = 0; n < co; ++n)
      for (int64_t c = 0; c < ci; ++c)
        for (int64_t j = 0; j < k; ++j) o[((size_t)n * k + j) * ci + c] = t.data[((size_t)n * ci + c) * k + j];
    return o;
  };
  auto cat_rows = [&](std::vector<const HostTensor*> ts) {
    std::vector<float> o;
    for (auto* t : ts) o.insert(o.end(), t->data.begin(), t->data.end());
    return o;
  };
  const std::string GE = "encoder.encoder.gEncoder.", AR = "encoder.encoder.gAR.baseNet.";
  for (int i = 1; i < 5; ++i) put(&s->conv_w[i], conv_nk(T(GE + "conv" + std::to_string(i) + ".weight")));
  for (int l = 0; l < m.ar_layers; ++l) {
    const std::string n = std::to_string(l);
    std::vector<float> wcat((size_t)1024 * 512), bias(1024);
    rnn_tc_pack(m.ar_kind, T(AR + "weight_ih_l" + n).data.data(), T(AR + "weight_hh_l" + n).data.data(),
                T(AR + "bias_ih_l" + n).data.data(), T(AR + "bias_hh_l" + n).data.data(), wcat.data(), bias.data());
    put(&s->rnn_wcat[l], wcat);
    put_f32(&s->rnn_bias[l], bias);
  }
  put(&s->ds_w, conv_nk(T("encoder.downsample.1.weight")));
  auto layer = [&](const std::string& p, bool cross, LayerW16& lw) {
    put(&lw.wqkv, cat_rows({&T(p + "mha.query.weight"), &T(p + "mha.key.weight"), &T(p + "mha.value.weight")}));
    put(&lw.wproj, T(p + "mha.proj.weight").data);
    if (cross) {
      put(&lw.wq_c, T(p + "mha_cross.query.weight").data);
      put(&lw.wkv_c, cat_rows({&T(p + "mha_cross.key.weight"), &T(p + "mha_cross.value.weight")}));
      put(&lw.wproj_c, T(p + "mha_cross.proj.weight").data);
    }
    put(&lw.w1, T(p + "ffnetwork.0.weight").data);
    put(&lw.w2, T(p + "ffnetwork.3.weight").data);
  };
  for (int l = 0; l < m.channel_layers; ++l) layer("ar_channel.layers." + std::to_string(l) + ".", false, s->chan[l]);
  for (int l = 0; l < m.cross_layers; ++l) layer("ar.layers." + std::to_string(l) + ".", true, s->cross[l]);
  put(&s->comb_a, T("ar.combinator.h0_a.weight").data);
  put(&s->comb_b, T("ar.combinator.h0_b.weight").data);
  put(&s->head_w, T("vap_head.weight").data);
  {
    std::vector<float> u(10 * kDim), d(kDim);
    conv0_v2_fold(T(GE + "conv0.weight").data.data(), T(GE + "conv0.bias").data.data(),
                  T(GE + "batchNorm0.weight").data.data(), u.data(), d.data(), &s->c0_stats);
    put_f32(&s->c0_u, u);
    put_f32(&s->c0_d, d);
    s->c0_tab_host = u;
    s->c0_tab_host.insert(s->c0_tab_host.end(), d.begin(), d.end());
    const auto& be = T(GE + "batchNorm0.bias").data;
    s->c0_tab_host.insert(s->c0_tab_host.end(), be.begin(), be.end());
    put_f32(&s->c0_tab, s->c0_tab_host);
  }
  if (cudaMalloc(&s->arena, host.size()) != cudaSuccess ||
      cudaMemcpy(s->arena, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    m.err = "16-bit weight arena: CUDA allocation/copy failed";
    return -3;
  }
  for (auto& f : fixes) *f.slot = reinterpret_cast<const bf16*>(static_cast<char*>(s->arena) + f.off);
  return 0;
}

void bf16_release(Model& m) {
  State16* s = static_cast<State16*>(m.bf16_state);
  if (!s) return;
  for (int f = 0; f < 2; ++f)
    if (s[f].arena) cudaFree(s[f].arena);
  delete[] s;
  m.bf16_state = nullptr;
}

size_t workspace_bytes_bf16(const Model& m, const Geometry& g) { return make_plan(m, g).bytes; }

int forward_bf16(Model& m, cudaStream_t st, const float* wav_f32, const Geometry& g, char* ws, float* logits,
                 float* vad_logits, float* vad_sig, const float**, int fp16, cudaEvent_t conv_wait,
                 cudaEvent_t conv_done, int wav_pcm16, const HeadOut* head) {
  // wav_pcm16: the buffer holds int16 PCM (same (B, 2, S) layout); only the fused conv0 -> conv1 kernel reads it
  if (wav_pcm16 && (!m.conv01 || (g.S & 1))) {
    m.err = "int16 PCM input needs the fused conv0/conv1 path (VAPB_CONV01=1) and an even n_samples";
    return -4;
  }
  const float* wav = wav_f32;
  const State16& s = static_cast<const State16*>(m.bf16_state)[fp16 ? 1 : 0];
  Fp16Scope fmt_scope(fp16 ? 1 : 0);
  const Plan16 p = make_plan(m, g);
  const Weights& w = m.w32;  // fp32 vectors (biases, norm affine, slopes, va head) and conv0
  Ctx cx{m, st};
  const int nseq = g.nseq;
  const long long T = g.T, L4 = g.L[4];
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto H = [&](size_t off) { return reinterpret_cast<bf16*>(ws + off); };

  {
    ProfScope ps(m, st, CAT_OTHER);
    for (int i = m.conv01 ? 1 : 0; i < 4; ++i) {
      m.launches += launch_zero_rows(st, H(p.act[i]), 2, p.mb, p.lpad[i] * kDim, 0, p.lo[i]);
      m.launches += launch_zero_rows(st, H(p.act[i]), 2, p.mb, p.lpad[i] * kDim, p.lo[i] + g.L[i],
                                     p.lpad[i] - p.lo[i] - g.L[i]);
    }
    for (int i = 0; i < (m.ar_layers > 1 ? 2 : 1); ++i) {
      m.launches += launch_zero_rows(st, H(p.rnn[i]), 2, nseq, p.rnn_lpad * kDim, 0, 4);
      m.launches += launch_zero_rows(st, H(p.rnn[i]), 2, nseq, p.rnn_lpad * kDim, 4 + L4, p.rnn_lpad - 4 - L4);
    }
  }

  // ---- CPC gEncoder (pipelined calls: after the previous item group's, see api.cu)
  if (conv_wait) cudaStreamWaitEvent(st, conv_wait, 0);
  m.trace(st, "g" + std::to_string(m.trace_group) + " conv_begin");
  for (int s0 = 0; s0 < nseq; s0 += p.mb) {
    const int n = (nseq - s0 < p.mb) ? nseq - s0 : p.mb;
    if (m.conv01) {
      // conv0 + conv1 in one kernel (k_conv01.cu): the first layer's 512 B per frame never reach HBM
      ProfScope ps(m, st, CAT_CONV_GEMM);
      std::string err;
      const int k = launch_conv01(st, wav, wav_pcm16, g.batch, g.S, s0, n, g.L[0], g.L[1], s.c0_tab_host.data(), s.c0_tab, s.c0_stats,
                                  s.conv_w[1], w.conv_b[1], w.conv_g[1], w.conv_be[1], H(p.act[1]), p.lpad[1] * kDim,
                                  (int)p.lo[1], m.n_sm, &err);
      if (k < 0) { m.err = err; return -3; }
      m.launches += k;
    } else {
      ProfScope ps(m, st, CAT_CONV0);
      if (m.conv0_tc) {
        std::string err;
        const int k = launch_conv0_tc(st, wav, g.batch, g.S, s0, n, g.L[0], s.c0_u, s.c0_d, w.c0_be, s.c0_stats,
                                      H(p.act[0]), p.lpad[0] * kDim, (int)p.lo[0],
                                      m.conv0_sms > 0 ? m.conv0_sms : m.n_sm, &err);
        if (k < 0) { m.err = err; return -3; }
        m.launches += k;
      } else {
        m.err = "VAPB_CONV0_TC=0: the CUDA-core conv0 kernel was removed in round 2 (k_conv0_tc.cu or the fused k_conv01.cu)";
        return -3;
      }
    }
    for (int i = m.conv01 ? 2 : 1; i <= 4; ++i) {
      const Conv& c = kConv[i];
      Epilogue e{};
      e.bias = w.conv_b[i];
      e.norm1 = NORM_CHANNEL;
      e.g1 = w.conv_g[i];
      e.b1 = w.conv_be[i];
      e.act = ACT_RELU;
      bf16* out;
      if (i < 4) {
        out = H(p.act[i]) + p.lo[i] * kDim;
        e.out1_map = RowMap{p.lpad[i] * kDim, kDim};
      } else {
        out = H(p.act4) + (long long)s0 * L4 * kDim;
        e.out1_map = RowMap{L4 * kDim, kDim};
      }
      if (m.conv_2sm) {
        if (cx.rc) return cx.rc;
        TcGemmArgs a{};
        a.A = H(p.act[i - 1]); a.a_map = RowMap{p.lpad[i - 1] * kDim, (long long)c.s * kDim}; a.W = s.conv_w[i];
        a.nseq = n; a.rows_per_seq = (int)g.L[i]; a.N = kDim; a.K = c.k * kDim;
        a.e = e; a.out1_bf16 = out;
        ProfScope ps(m, st, CAT_CONV_GEMM);
        std::string err;
        const int k = launch_gemm_2sm(st, a, m.n_sm, &err);
        if (k < 0) { m.err = err; return -3; }
        m.launches += k;
      } else {
        cx.lin(H(p.act[i - 1]), RowMap{p.lpad[i - 1] * kDim, (long long)c.s * kDim}, s.conv_w[i], n, (int)g.L[i],
               kDim, c.k * kDim, e, nullptr, out, 1, CAT_CONV_GEMM);
      }
    }
  }

  if (cx.rc) return cx.rc;
  if (conv_done) cudaEventRecord(conv_done, st);
  m.trace(st, "g" + std::to_string(m.trace_group) + " conv_end");

  // ---- gAR
  const bf16* rnn_in = H(p.act4);
  RowMap rnn_in_map{L4 * kDim, kDim};
  for (int l = 0; l < m.ar_layers; ++l) {
    bf16* rnn_out = H(p.rnn[l & 1]) + 4 * kDim;
    {
      ProfScope ps(m, st, CAT_RNN);
      std::string err;
      const int n = launch_rnn_tc(st, m.ar_kind, rnn_in, rnn_in_map.seq_stride, rnn_in_map.row_stride, s.rnn_wcat[l],
                                  s.rnn_bias[l], rnn_out, p.rnn_lpad * kDim, nseq, (int)L4, &err);
      if (n < 0) { m.err = err; return -3; }
      m.launches += n;
    }
    rnn_in = rnn_out;
    rnn_in_map = RowMap{p.rnn_lpad * kDim, kDim};
  }

  m.trace(st, "g" + std::to_string(m.trace_group) + " rnn_end");

  // ---- downsample
  const int n_layers = m.channel_layers + m.cross_layers;
  auto first_ln = [&](int li, const float** gg, const float** bb) {
    if (li >= n_layers) { *gg = nullptr; *bb = nullptr; return; }
    const LayerW& lw = li < m.channel_layers ? w.chan[li] : w.cross[li - m.channel_layers];
    *gg = lw.ln_sa_g;
    *bb = lw.ln_sa_b;
  };
  const int MT = (int)(nseq * T);
  {
    Epilogue e{};
    e.bias = w.ds_b;
    e.norm1 = NORM_LAYER;
    e.g1 = w.ds_g;
    e.b1 = w.ds_be;
    e.act = ACT_GELU;
    e.out1_map = dense(T, kDim);
    first_ln(0, &e.g2, &e.b2);
    if (e.g2) { e.norm2 = NORM_LAYER; e.out2 = H(p.z); e.out2_map = dense(T, kDim); }
    cx.lin(H(p.rnn[(m.ar_layers - 1) & 1]), RowMap{p.rnn_lpad * kDim, 2 * kDim}, s.ds_w, nseq, (int)T, kDim,
            5 * kDim, e, F(p.stage[0]), H(p.xs), 1, CAT_CONV_GEMM);
  }

  bool vad_done = false;  // the VAD head has been taken inside the last layer's FFN kernel
  // ---- transformer layers (rows are dense (nseq*T, .) from here on)
  const RowMap d256 = dense(MT, kDim);
  for (int li = 0; li < n_layers; ++li) {
    const bool cross = li >= m.channel_layers;
    const LayerW& lw = cross ? w.cross[li - m.channel_layers] : w.chan[li];
    const LayerW16& lh = cross ? s.cross[li - m.channel_layers] : s.chan[li];
    const float* x_in = F(p.stage[li]);
    {
      Epilogue e{};
      e.out1_map = dense(MT, 3 * kDim);
      cx.lin(H(p.z), d256, lh.wqkv, 1, MT, 3 * kDim, kDim, e, nullptr, H(p.qkv));
    }
    if (cross) {
      Epilogue e{};
      e.out1_map = dense(MT, 2 * kDim);
      cx.lin(H(p.xs), d256, lh.wkv_c, 1, MT, 2 * kDim, kDim, e, nullptr, H(p.kvc));
    }
    if (cx.rc) return cx.rc;
    {
      ProfScope ps(m, st, CAT_ATTN);
      std::string err;
      const int n = launch_attention_tc(st, H(p.qkv), 3 * kDim, H(p.qkv) + kDim, H(p.qkv) + 2 * kDim, 3 * kDim, H(p.y),
                                        nseq, (int)T, m.num_heads, lw.slopes, 0, m.n_sm, &err);
      if (n < 0) { m.err = err; return -3; }
      m.launches += n;
    }
    float* x_mid = cross ? F(p.xa) : F(p.xb);
    {
      Epilogue e{};
      e.resid = x_in;
      e.resid_map = d256;
      e.out1_map = d256;
      e.norm2 = NORM_LAYER;
      e.g2 = cross ? lw.ln_src_g : lw.ln_ffn_g;
      e.b2 = cross ? lw.ln_src_b : lw.ln_ffn_b;
      e.out2 = H(p.z);
      e.out2_map = d256;
      cx.lin(H(p.y), d256, lh.wproj, 1, MT, kDim, kDim, e, x_mid, nullptr);
    }
    if (cross) {
      {
        Epilogue e{};
        e.out1_map = d256;
        cx.lin(H(p.z), d256, lh.wq_c, 1, MT, kDim, kDim, e, nullptr, H(p.qc));
      }
      if (cx.rc) return cx.rc;
      {
        ProfScope ps(m, st, CAT_ATTN);
        std::string err;
        const int n = launch_attention_tc(st, H(p.qc), kDim, H(p.kvc), H(p.kvc) + kDim, 2 * kDim, H(p.y), nseq, (int)T,
                                          m.num_heads, lw.slopes_cross, 1, m.n_sm, &err);
        if (n < 0) { m.err = err; return -3; }
        m.launches += n;
      }
      Epilogue e{};
      e.resid = x_mid;
      e.resid_map = d256;
      e.out1_map = d256;
      e.norm2 = NORM_LAYER;
      e.g2 = lw.ln_ffn_g;
      e.b2 = lw.ln_ffn_b;
      e.out2 = H(p.z);
      e.out2_map = d256;
      cx.lin(H(p.y), d256, lh.wproj_c, 1, MT, kDim, kDim, e, F(p.xb), nullptr);
    }
    const float *gn = nullptr, *bn = nullptr;
    first_ln(li + 1, &gn, &bn);
    if (m.ffn_fused) {
      // FFN in one kernel: the 768-wide hidden activation stays on the SM. zn overwrites z in place (a tile's z is in
      // shared memory long before its rows are rewritten).
      if (cx.rc) return cx.rc;
      ProfScope ps(m, st, CAT_LINEAR_GEMM);
      std::string err;
      // last layer: the VAD head rides on the final epilogue (its rows are the head's input)
      FfnVad vad{w.va_w, w.va_b, (int)g.batch, (int)T, vad_logits, vad_sig};
      const bool with_vad = m.vad_fused && li == n_layers - 1 && !gn && (vad_logits || vad_sig);
      vad_done = vad_done || with_vad;
      const int n = launch_ffn_fused(st, H(p.z), lh.w1, lh.w2, F(p.xb), F(p.stage[li + 1]), H(p.xs),
                                     gn ? H(p.z) : nullptr, gn, bn, MT, m.n_sm, &err, nullptr, with_vad ? &vad : nullptr);
      if (n < 0) { m.err = err; return -3; }
      m.launches += n;
    } else {
      {
        Epilogue e{};
        e.act = ACT_GELU;
        e.out1_map = dense(MT, kFfn);
        cx.lin(H(p.z), d256, lh.w1, 1, MT, kFfn, kDim, e, nullptr, H(p.h));
      }
      {
        Epilogue e{};
        e.resid = F(p.xb);
        e.resid_map = d256;
        e.out1_map = d256;
        if (gn) { e.norm2 = NORM_LAYER; e.g2 = gn; e.b2 = bn; e.out2 = H(p.z); e.out2_map = d256; }
        // bf16 shadow of the new residual stream: next cross layer's K/V source, or the combinator input
        cx.lin(H(p.h), dense(MT, kFfn), lh.w2, 1, MT, kDim, kFfn, e, F(p.stage[li + 1]), H(p.xs));
      }
    }
  }

  // ---- combinator + heads
  const float* x_last = F(p.stage[n_layers]);
  const int MB = (int)(g.batch * T);
  for (int c = 0; c < 2; ++c) {
    Epilogue e{};
    e.norm1 = NORM_LAYER;
    e.g1 = w.comb_g;
    e.b1 = w.comb_be;
    e.act = ACT_GELU;
    e.accumulate = c;
    e.out1_map = dense(MB, kDim);
    cx.lin(H(p.xs) + (long long)c * MB * kDim, dense(MB, kDim), c == 0 ? s.comb_a : s.comb_b, 1, MB, kDim, kDim, e,
            F(p.comb), c == 1 ? H(p.combb) : nullptr);
  }
  if ((vad_logits || vad_sig) && !vad_done) {
    ProfScope ps(m, st, CAT_HEADS);
    m.launches += launch_vad_head_blocked(st, x_last, w.va_w, w.va_b, g.batch, (int)T, vad_logits, vad_sig);
  }
  if (head) {
    // vap_head fused with softmax, entropy, marginals, arg-max and the counters: the logits stay in TMEM unless asked for
    if (cx.rc) return cx.rc;
    ProfScope ps(m, st, CAT_HEADS);
    std::string err;
    const int n = launch_head_probs(st, H(p.combb), s.head_w, w.head_b, MB, head->now_lo, head->now_hi, head->fut_lo,
                                    head->fut_hi, head->logits, head->probs, head->p_now, head->p_future, head->H,
                                    head->lse, head->argmax, head->counters, vad_sig, m.n_sm, &err);
    if (n < 0) { m.err = err; return -3; }
    m.launches += n;
  } else {
    Epilogue e{};
    e.bias = w.head_b;
    e.out1_map = dense(MB, kClasses);
    cx.lin(H(p.combb), dense(MB, kDim), s.head_w, 1, MB, kClasses, kDim, e, logits, nullptr, 2);
  }
  return cx.rc;
}

int stage_bf16(const Model& m, const Geometry& g, char* ws, const std::string& name, StageRef* ref) {
  const Plan16 p = make_plan(m, g);
  ref->is_bf16 = 0;
  ref->nseq = g.nseq;
  ref->rows_per_seq = (int)g.T;
  ref->map = RowMap{g.T * kDim, kDim};
  if (name == "conv") {
    ref->ptr = ws + p.act4;
    ref->is_bf16 = 1;  // 16-bit buffer; api.cu turns this into the mode's format
    ref->rows_per_seq = (int)g.L[4];
    ref->map = RowMap{g.L[4] * kDim, kDim};
  } else if (name == "ar") {
    ref->ptr = reinterpret_cast<const bf16*>(ws + p.rnn[(m.ar_layers - 1) & 1]) + 4 * kDim;
    ref->is_bf16 = 1;  // 16-bit buffer; api.cu turns this into the mode's format
    ref->rows_per_seq = (int)g.L[4];
    ref->map = RowMap{p.rnn_lpad * kDim, kDim};
  } else if (name == "enc") {
    ref->ptr = ws + p.stage[0];
    ref->blocked = 1;
  } else if (name == "ch") {
    ref->ptr = ws + p.stage[m.channel_layers];
    ref->blocked = 1;
  } else if (name.size() >= 3 && name.compare(0, 2, "ar") == 0) {
    const int l = atoi(name.c_str() + 2);
    if (l < 0 || l >= m.cross_layers) return -1;
    ref->ptr = ws + p.stage[m.channel_layers + l + 1];
    ref->blocked = 1;
  } else if (name == "comb") {
    ref->ptr = ws + p.comb;
    ref->nseq = g.batch;
    ref->blocked = 1;
  } else {
    return -1;
  }
  return 0;
}

}  // namespace vapb
