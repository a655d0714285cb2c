// FP32 (parity mode) orchestration of VapGPT.forward on CUDA cores.
// Reference call stack: vap/model.py:249-268 -> vap/encoder.py:49-66 ->
// vap/modules.py:342-358 (GPT), :380-408 (GPTStereo), :434-449 (Combinator).
//
// Sequences are channel-major: row c*B + b is channel c of batch item b, so the
// encoder, ar_channel and both directions of the stereo layers (shared weights)
// run as one batch of 2B sequences, the "other" channel of sequence s is
// (s + B) mod 2B, and the combinator's x1 / x2 are the two halves.
#include <cstdio>

#include "model.h"

namespace vapb {

namespace {

struct Conv { int k, s, p; };
const Conv kConv[5] = {{10, 5, 3}, {8, 4, 2}, {4, 2, 1}, {4, 2, 1}, {4, 2, 1}};

size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

struct Plan32 {
  int mb;
  long long lo[4], lpad[4];  // act0..act3: leading zero rows and padded length
  long long rnn_lpad;
  size_t act[4], act4, xproj, rnn[2], stage[2 * kMaxLayers + 1], xa, xb, z, qkv, kvc, qc, y, h, comb, bytes;
  int n_stage;
};

Plan32 make_plan(const Model& m, const Geometry& g) {
  Plan32 p{};
  const int G = m.ar_kind == 0 ? 4 : 3;
  // micro-batch so that act0 stays under ~4 GiB
  const long long per_seq0 = (g.L[0] + 16) * kDim * 4;
  long long mb = (4LL << 30) / per_seq0;
  if (mb < 1) mb = 1;
  if (mb > g.nseq) mb = g.nseq;
  p.mb = (int)mb;
  size_t off = 0;
  for (int i = 0; i < 4; ++i) {
    const Conv& nx = kConv[i + 1];
    p.lo[i] = nx.p;
    const long long need = (long long)nx.s * (g.L[i + 1] - 1) + nx.k;  // padded rows the next conv reads
    long long lp = need > nx.p + g.L[i] ? need : nx.p + g.L[i];
    lp = (lp + nx.s - 1) / nx.s * nx.s;
    p.lpad[i] = lp;
    p.act[i] = off;
    off = align_up(off + (size_t)mb * lp * kDim * 4);
  }
  p.act4 = off;  off = align_up(off + (size_t)g.nseq * g.L[4] * kDim * 4);
  p.xproj = off; off = align_up(off + (size_t)g.nseq * g.L[4] * G * kDim * 4);
  p.rnn_lpad = 4 + g.L[4];
  for (int i = 0; i < 2; ++i) {
    p.rnn[i] = off;
    if (i == 0 || m.ar_layers > 1) off = align_up(off + (size_t)g.nseq * p.rnn_lpad * kDim * 4);
  }
  const size_t xbytes = (size_t)g.nseq * g.T * kDim * 4;
  p.n_stage = 1 + m.channel_layers + m.cross_layers;
  for (int i = 0; i < p.n_stage; ++i) { p.stage[i] = off; off = align_up(off + xbytes); }
  p.xa = off;  off = align_up(off + xbytes);
  p.xb = off;  off = align_up(off + xbytes);
  p.z = off;   off = align_up(off + xbytes);
  p.qkv = off; off = align_up(off + 3 * xbytes);
  p.kvc = off; off = align_up(off + 2 * xbytes);
  p.qc = off;  off = align_up(off + xbytes);
  p.y = off;   off = align_up(off + xbytes);
  p.h = off;   off = align_up(off + 3 * xbytes);
  p.comb = off; off = align_up(off + xbytes / 2);
  p.bytes = off;
  return p;
}

RowMap dense(long long n) { return RowMap{0, n}; }

struct Ctx {
  Model& m;
  cudaStream_t st;
  bool x3 = false;  // VAPB_MODE_FP32_TC
  void gemm(const void* A, RowMap amap, const void* W, int M, int rps, int N, int K, const Epilogue& e,
            int cat = CAT_LINEAR_GEMM) {
    GemmProblem p{A, amap, W, M, rps, N, K};
    ProfScope ps(m, st, cat);
    if (x3) {  // split-fp16 tensor-core GEMM, same operands and epilogue (k_gemm_x3.cu)
      auto it = m.x3_w.find(W);
      if (it != m.x3_w.end()) {
        std::string err;
        const int n = launch_gemm_x3(st, p, e, it->second, m.n_sm, &err);
        if (n >= 0) { m.launches += n; return; }
      }
    }
    m.launches += launch_gemm_f32(st, p, e);
  }
};

Epilogue epi_plain(float* out, int N) {
  Epilogue e{};
  e.out1 = out;
  e.out1_map = dense(N);
  return e;
}

}  // namespace

size_t workspace_bytes_fp32(const Model& m, const Geometry& g) { return make_plan(m, g).bytes; }

int forward_fp32(Model& m, cudaStream_t st, const float* wav, const Geometry& g, char* ws, float* logits,
                 float* vad_logits, float* vad_sig, const float**, const AttnMaps* maps, bool tensor_gemms) {
  const Plan32 p = make_plan(m, g);
  const Weights& w = m.w32;
  Ctx cx{m, st, tensor_gemms};
  const int G = m.ar_kind == 0 ? 4 : 3;
  const int nseq = g.nseq;
  const long long T = g.T, L4 = g.L[4];
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };

  // physical zero padding of the conv inputs (interior rows are fully rewritten)
  {
  ProfScope ps(m, st, CAT_OTHER);
  for (int i = 0; i < 4; ++i) {
    m.launches += launch_zero_rows(st, F(p.act[i]), 4, p.mb, p.lpad[i] * kDim, 0, p.lo[i]);
    m.launches += launch_zero_rows(st, F(p.act[i]), 4, p.mb, p.lpad[i] * kDim, p.lo[i] + g.L[i],
                                   p.lpad[i] - p.lo[i] - g.L[i]);
  }
  for (int i = 0; i < (m.ar_layers > 1 ? 2 : 1); ++i)
    m.launches += launch_zero_rows(st, F(p.rnn[i]), 4, nseq, p.rnn_lpad * kDim, 0, 4);
  }

  // ---- CPC gEncoder: conv0 (CUDA cores) + conv1..4 (implicit GEMM), micro-batched
  for (int s0 = 0; s0 < nseq; s0 += p.mb) {
    const int n = (nseq - s0 < p.mb) ? nseq - s0 : p.mb;
    { ProfScope ps(m, st, CAT_CONV0);
    m.launches += launch_conv0(st, wav, g.batch, g.S, s0, n, g.L[0], w.c0_w, w.c0_b, w.c0_g, w.c0_be,
                               F(p.act[0]), 0, p.lpad[0] * kDim, (int)p.lo[0]); }
    for (int i = 1; i <= 4; ++i) {
      const Conv& c = kConv[i];
      Epilogue e{};
      e.bias = w.conv_b[i];
      e.norm1 = NORM_CHANNEL;
      e.g1 = w.conv_g[i];
      e.b1 = w.conv_be[i];
      e.act = ACT_RELU;
      if (i < 4) {
        e.out1 = F(p.act[i]) + p.lo[i] * kDim;
        e.out1_map = RowMap{p.lpad[i] * kDim, kDim};
      } else {
        e.out1 = F(p.act4) + (long long)s0 * L4 * kDim;
        e.out1_map = RowMap{L4 * kDim, kDim};
      }
      // input frame s*t - p sits at padded row s*t (lo == p)
      cx.gemm(F(p.act[i - 1]), RowMap{p.lpad[i - 1] * kDim, (long long)c.s * kDim}, w.conv_w[i],
              (int)(n * g.L[i]), (int)g.L[i], kDim, c.k * kDim, e, CAT_CONV_GEMM);
    }
  }

  // ---- gAR: hoisted input projection + recurrence, per layer
  const float* rnn_in = F(p.act4);
  RowMap rnn_in_map{L4 * kDim, kDim};
  float* rnn_out = nullptr;
  for (int l = 0; l < m.ar_layers; ++l) {
    Epilogue e = epi_plain(F(p.xproj), G * kDim);
    e.out1_map = RowMap{L4 * G * kDim, (long long)G * kDim};  // rows_per_seq == L4 in this GEMM
    e.bias = w.rnn_bx[l];
    cx.gemm(rnn_in, rnn_in_map, w.rnn_wih[l], (int)(nseq * L4), (int)L4, G * kDim, kDim, e);
    rnn_out = F(p.rnn[l & 1]) + 4 * kDim;
    ProfScope ps(m, st, CAT_RNN);
    m.launches += launch_rnn_f32(st, m.ar_kind, F(p.xproj), w.rnn_whh_t[l], w.rnn_bhn[l], rnn_out,
                                 p.rnn_lpad * kDim, nseq, (int)L4);
    rnn_in = rnn_out;
    rnn_in_map = RowMap{p.rnn_lpad * kDim, kDim};
  }

  // ---- downsample: causal conv k5 s2 (left pad 4) + LayerNorm + GELU
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
    e.out1 = F(p.stage[0]);
    e.out1_map = RowMap{T * kDim, kDim};  // rows_per_seq == T in this GEMM
    first_ln(0, &e.g2, &e.b2);
    if (e.g2) { e.norm2 = NORM_LAYER; e.out2 = F(p.z); e.out2_map = RowMap{T * kDim, kDim}; }
    const float* base = F(p.rnn[(m.ar_layers - 1) & 1]);  // frame 2t-4 == padded row 2t
    cx.gemm(base, RowMap{p.rnn_lpad * kDim, 2 * kDim}, w.ds_w, MT, (int)T, kDim, 5 * kDim, e, CAT_CONV_GEMM);
  }

  // ---- transformer layers
  for (int li = 0; li < n_layers; ++li) {
    const bool cross = li >= m.channel_layers;
    const LayerW& lw = cross ? w.cross[li - m.channel_layers] : w.chan[li];
    const float* x_in = F(p.stage[li]);
    float* x_out = F(p.stage[li + 1]);
    // z == LN_self_attn(x_in)
    cx.gemm(F(p.z), dense(kDim), lw.wqkv, MT, MT, 3 * kDim, kDim, epi_plain(F(p.qkv), 3 * kDim));
    if (cross) cx.gemm(x_in, dense(kDim), lw.wkv_c, MT, MT, 2 * kDim, kDim, epi_plain(F(p.kvc), 2 * kDim));
    { ProfScope ps(m, st, CAT_ATTN);
    if (tensor_gemms && m.attn_x3) {
      // the FFN's hidden buffer is idle here and is exactly as large as the hi / lo copies of q | k | v
      std::string err;
      const int n = launch_attention_x3(st, F(p.qkv), 3 * kDim, F(p.qkv), 3 * kDim, kDim, 2 * kDim, F(p.h), F(p.y), nseq,
                                        (int)T, m.num_heads, lw.slopes, 0, m.n_sm, &err);
      if (n < 0) { m.err = err; return -3; }
      m.launches += n;
    } else {
      m.launches += launch_attention_f32(st, F(p.qkv), 3 * kDim, F(p.qkv) + kDim, F(p.qkv) + 2 * kDim, 3 * kDim,
                                         F(p.y), nseq, (int)T, m.num_heads, lw.slopes, 0);
    }
    if (maps)
      m.launches += launch_attention_map_f32(st, F(p.qkv), 3 * kDim, F(p.qkv) + kDim, 3 * kDim, nseq, (int)T,
                                             m.num_heads, lw.slopes, 0,
                                             cross ? maps->cross_self_attn : maps->self_attn, g.batch,
                                             cross ? m.cross_layers : m.channel_layers,
                                             cross ? li - m.channel_layers : li); }
    const float* x_mid = nullptr;
    {
      Epilogue e{};
      e.resid = x_in;
      e.resid_map = dense(kDim);
      e.out1 = cross ? F(p.xa) : F(p.xb);
      e.out1_map = dense(kDim);
      e.norm2 = NORM_LAYER;
      e.g2 = cross ? lw.ln_src_g : lw.ln_ffn_g;
      e.b2 = cross ? lw.ln_src_b : lw.ln_ffn_b;
      e.out2 = F(p.z);
      e.out2_map = dense(kDim);
      cx.gemm(F(p.y), dense(kDim), lw.wproj, MT, MT, kDim, kDim, e);
      x_mid = reinterpret_cast<const float*>(e.out1);
    }
    if (cross) {
      cx.gemm(F(p.z), dense(kDim), lw.wq_c, MT, MT, kDim, kDim, epi_plain(F(p.qc), kDim));
      { ProfScope ps(m, st, CAT_ATTN);
      if (tensor_gemms && m.attn_x3) {
        std::string err;
        const int n = launch_attention_x3(st, F(p.qc), kDim, F(p.kvc), 2 * kDim, 0, kDim, F(p.h), F(p.y), nseq, (int)T,
                                          m.num_heads, lw.slopes_cross, 1, m.n_sm, &err);
        if (n < 0) { m.err = err; return -3; }
        m.launches += n;
      } else {
        m.launches += launch_attention_f32(st, F(p.qc), kDim, F(p.kvc), F(p.kvc) + kDim, 2 * kDim, F(p.y), nseq,
                                           (int)T, m.num_heads, lw.slopes_cross, 1);
      }
      if (maps)
        m.launches += launch_attention_map_f32(st, F(p.qc), kDim, F(p.kvc), 2 * kDim, nseq, (int)T, m.num_heads,
                                               lw.slopes_cross, 1, maps->cross_attn, g.batch, m.cross_layers,
                                               li - m.channel_layers); }
      Epilogue e{};
      e.resid = x_mid;
      e.resid_map = dense(kDim);
      e.out1 = F(p.xb);
      e.out1_map = dense(kDim);
      e.norm2 = NORM_LAYER;
      e.g2 = lw.ln_ffn_g;
      e.b2 = lw.ln_ffn_b;
      e.out2 = F(p.z);
      e.out2_map = dense(kDim);
      cx.gemm(F(p.y), dense(kDim), lw.wproj_c, MT, MT, kDim, kDim, e);
    }
    {
      Epilogue e = epi_plain(F(p.h), kFfn);
      e.act = ACT_GELU;
      cx.gemm(F(p.z), dense(kDim), lw.w1, MT, MT, kFfn, kDim, e);
    }
    {
      Epilogue e{};
      e.resid = F(p.xb);
      e.resid_map = dense(kDim);
      e.out1 = x_out;
      e.out1_map = dense(kDim);
      first_ln(li + 1, &e.g2, &e.b2);
      if (e.g2) { e.norm2 = NORM_LAYER; e.out2 = F(p.z); e.out2_map = dense(kDim); }
      cx.gemm(F(p.h), dense(kFfn), lw.w2, MT, MT, kDim, kFfn, e);
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
    e.out1 = F(p.comb);
    e.out1_map = dense(kDim);
    cx.gemm(x_last + (long long)c * MB * kDim, dense(kDim), c == 0 ? w.comb_a : w.comb_b, MB, MB, kDim, kDim, e);
  }
  if (vad_logits || vad_sig) {
    ProfScope ps(m, st, CAT_HEADS);
    m.launches += launch_vad_head(st, x_last, w.va_w, w.va_b, g.batch, (int)T, vad_logits, vad_sig);
  }
  {
    Epilogue e = epi_plain(logits, kClasses);
    e.bias = w.head_b;
    cx.gemm(F(p.comb), dense(kDim), w.head_w, MB, MB, kClasses, kDim, e);
  }
  return 0;
}

int stage_fp32(const Model& m, const Geometry& g, char* ws, const std::string& name, StageRef* ref) {
  const Plan32 p = make_plan(m, g);
  auto F = [&](size_t off) { return reinterpret_cast<const float*>(ws + off); };
  ref->is_bf16 = 0;
  ref->nseq = g.nseq;
  ref->rows_per_seq = (int)g.T;
  ref->map = RowMap{g.T * kDim, kDim};
  if (name == "conv") {
    ref->ptr = F(p.act4);
    ref->rows_per_seq = (int)g.L[4];
    ref->map = RowMap{g.L[4] * kDim, kDim};
  } else if (name == "ar") {
    ref->ptr = F(p.rnn[(m.ar_layers - 1) & 1]) + 4 * kDim;
    ref->rows_per_seq = (int)g.L[4];
    ref->map = RowMap{p.rnn_lpad * kDim, kDim};
  } else if (name == "enc") {
    ref->ptr = F(p.stage[0]);
  } else if (name == "ch") {
    ref->ptr = F(p.stage[m.channel_layers]);
  } else if (name.size() >= 3 && name.compare(0, 2, "ar") == 0) {
    const int l = atoi(name.c_str() + 2);
    if (l < 0 || l >= m.cross_layers) return -1;
    ref->ptr = F(p.stage[m.channel_layers + l + 1]);
  } else if (name == "comb") {
    ref->ptr = F(p.comb);
    ref->nseq = g.batch;
  } else {
    return -1;
  }
  return 0;
}

}  // namespace vapb
