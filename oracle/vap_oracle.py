"""CPU oracle: a self-contained restatement of the reference's stereo inference
forward path, `VapGPT.forward` / `VapGPT.probs` (TEST INFRASTRUCTURE).

It exists because the reference is Python and cannot travel to the GPU box
(gpurun ships only /root/repo). It follows the reference's arithmetic forms
line by line, in fp32 (or fp64 when the state dict / input are cast), using
plain torch functional ops on the CPU and nothing from the product package.

Pinning: the reference's own tests hold no golden vector for this path
(SURVEY.md §0 F11), so the oracle is pinned against outputs of the UNMODIFIED
reference run in the authoring container: oracle/make_golden.py writes them to
tests/golden/*.npz and tests/test_oracle.py checks this file against them (and
directly against the imported reference when /root/reference is present).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module. It is never a fallback for the CUDA path.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BIN_FRAMES = [10, 20, 30, 40]  # vap/objective.py:10-11 on bin_times .2/.4/.6/.8 @50 Hz


# --------------------------------------------------------------------------- #
# encoder                                                                     #
# --------------------------------------------------------------------------- #
def channel_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """vap/encoder_components.py:62-70. x is (B, C, T); unbiased variance over C."""
    mean = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, keepdim=True)  # unbiased: /(C-1)
    x = (x - mean) * torch.rsqrt(var + eps)
    return x * weight + bias


_CONV_GEOM = [(5, 3), (4, 2), (2, 1), (2, 1), (2, 1)]  # (stride, padding), :83-91


def cpc_conv_stack(sd: Dict[str, Tensor], wav: Tensor, upto: int = 5) -> Tensor:
    """vap/encoder_components.py:98-104. wav (B, 1, S) -> (B, 256, S/160)."""
    p = "encoder.encoder.gEncoder."
    x = wav
    for i, (s, pad) in enumerate(_CONV_GEOM[:upto]):
        x = F.conv1d(x, sd[f"{p}conv{i}.weight"], sd[f"{p}conv{i}.bias"], stride=s, padding=pad)
        x = F.relu(channel_norm(x, sd[f"{p}batchNorm{i}.weight"], sd[f"{p}batchNorm{i}.bias"]))
    return x


def ar_kind(sd: Dict[str, Tensor]):
    """Cell type and depth from the state-dict shapes."""
    p = "encoder.encoder.gAR.baseNet."
    rows = sd[f"{p}weight_ih_l0"].shape[0]
    hidden = sd[f"{p}weight_hh_l0"].shape[1]
    kind = {4: "LSTM", 3: "GRU"}[rows // hidden]
    layers = 0
    while f"{p}weight_ih_l{layers}" in sd:
        layers += 1
    return kind, layers


def ar_net(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """vap/encoder_components.py:140-159 with keepHidden=False, reverse=False:
    nn.LSTM / nn.GRU (batch_first, zero initial state). z (B, T, 256).
    Uses the same ATen op the reference's nn.LSTM/nn.GRU dispatches to."""
    kind, layers = ar_kind(sd)
    p = "encoder.encoder.gAR.baseNet."
    flat = []
    for l in range(layers):
        flat += [sd[f"{p}weight_ih_l{l}"], sd[f"{p}weight_hh_l{l}"],
                 sd[f"{p}bias_ih_l{l}"], sd[f"{p}bias_hh_l{l}"]]
    B, H = z.shape[0], sd[f"{p}weight_hh_l0"].shape[1]
    h0 = z.new_zeros(layers, B, H)
    if kind == "LSTM":
        out = torch._VF.lstm(z, (h0, h0.clone()), flat, True, layers, 0.0, False, False, True)
    else:
        out = torch._VF.gru(z, h0, flat, True, layers, 0.0, False, False, True)
    return out[0]


def ar_net_loop(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """The same recurrence written out step by step (PyTorch gate order:
    LSTM i,f,g,o; GRU r,z,n with n = tanh(W_in x + b_in + r*(W_hn h + b_hn))).
    Slow; used by the tests on short inputs to document the cell equations the
    CUDA recurrence kernel implements."""
    kind, layers = ar_kind(sd)
    p = "encoder.encoder.gAR.baseNet."
    x = z
    for l in range(layers):
        wih, whh = sd[f"{p}weight_ih_l{l}"], sd[f"{p}weight_hh_l{l}"]
        bih, bhh = sd[f"{p}bias_ih_l{l}"], sd[f"{p}bias_hh_l{l}"]
        B, T, _ = x.shape
        H = whh.shape[1]
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        ys = []
        for t in range(T):
            gi = x[:, t] @ wih.T + bih
            gh = h @ whh.T + bhh
            if kind == "LSTM":
                i, f, g, o = (gi + gh).chunk(4, dim=-1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
            else:
                ir, iz, in_ = gi.chunk(3, dim=-1)
                hr, hz, hn = gh.chunk(3, dim=-1)
                r = torch.sigmoid(ir + hr)
                u = torch.sigmoid(iz + hz)
                n = torch.tanh(in_ + r * hn)
                h = (1 - u) * n + u * h
            ys.append(h)
        x = torch.stack(ys, dim=1)
    return x


def downsample(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """vap/encoder.py:24-30,65; CConv1d (encoder_components.py:454-482: left pad
    k-1=4 zeros), LayerNorm over channels (:414-425), GELU(erf) (:497).
    z (B, T100, 256) -> (B, T50, 256)."""
    x = z.transpose(1, 2)
    x = F.pad(x, (4, 0))
    x = F.conv1d(x, sd["encoder.downsample.1.weight"], sd["encoder.downsample.1.bias"], stride=2)
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (x.shape[-1],), sd["encoder.downsample.2.ln.weight"],
                     sd["encoder.downsample.2.ln.bias"], 1e-5)
    # The reference applies GELU on the (b d t) view of the LayerNorm output
    # (a strided tensor; ATen's scalar erf path) and rearranges back afterwards;
    # doing the same keeps the oracle bit-identical to the reference on CPU.
    return F.gelu(x.transpose(1, 2)).transpose(1, 2)


def encoder(sd: Dict[str, Tensor], wav: Tensor, stages: Optional[dict] = None) -> Tensor:
    """vap/encoder.py:49-66. wav (B, 1, S) -> (B, T, 256)."""
    z = cpc_conv_stack(sd, wav)
    z = z.transpose(1, 2)  # b c n -> b n c (:63)
    if stages is not None:
        stages["conv"] = z
    z = ar_net(sd, z)
    if stages is not None:
        stages["ar"] = z
    z = downsample(sd, z)
    if stages is not None:
        stages["enc"] = z
    return z


# --------------------------------------------------------------------------- #
# transformer                                                                 #
# --------------------------------------------------------------------------- #
def alibi_mask(m: Tensor, T: int) -> Tensor:
    """vap/modules.py:169-187. (1, H, T, T): 1.0 + m_h*j on/below the diagonal,
    -inf above (the tril's 1.0 entries survive the masked_fill)."""
    H = m.shape[0]
    rel = torch.arange(T, dtype=m.dtype, device=m.device).view(1, 1, -1).expand(1, H, -1)
    alibi = rel * m.unsqueeze(0).unsqueeze(-1)
    mask = torch.tril(torch.ones((T, T), dtype=m.dtype, device=m.device)).view(1, 1, T, T).repeat(1, H, 1, 1)
    mask.masked_fill_(mask == 0, float("-inf"))
    return alibi.unsqueeze(-2) + mask


def mha_alibi(sd, prefix: str, Q: Tensor, K: Tensor, V: Tensor, n_heads: int = 4,
              maps: Optional[list] = None) -> Tensor:
    """vap/modules.py:82-110 + 189-202. No biases; scale = 1/sqrt(dim) (F7). `maps`: list that receives the
    (B, H, T, T) attention weights the reference returns next to y (:104-110)."""
    B, T, D = Q.shape
    hd = D // n_heads

    def heads(x):
        return x.view(B, -1, n_heads, hd).transpose(1, 2)

    k = heads(F.linear(K, sd[prefix + "key.weight"]))
    q = heads(F.linear(Q, sd[prefix + "query.weight"]))
    v = heads(F.linear(V, sd[prefix + "value.weight"]))
    att = torch.einsum("bhid,bhjd->bhij", q, k) * (1.0 / (D ** 0.5))
    att = att + alibi_mask(sd[prefix + "m"].to(att.dtype), T)
    att = F.softmax(att, dim=-1)
    if maps is not None:
        maps.append(att)
    y = (att @ v).transpose(1, 2).reshape(B, T, D)
    return F.linear(y, sd[prefix + "proj.weight"])


def _ln(sd, prefix: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], 1e-5)


def transformer_layer(sd, prefix: str, x: Tensor, src: Optional[Tensor] = None,
                      n_heads: int = 4, self_maps: Optional[list] = None,
                      cross_maps: Optional[list] = None) -> Tensor:
    """vap/modules.py:246-275 (pre-LN; cross-attention K/V = un-normalised src)."""
    z = _ln(sd, prefix + "ln_self_attn.", x)
    x = x + mha_alibi(sd, prefix + "mha.", z, z, z, n_heads, self_maps)
    if src is not None:
        z = _ln(sd, prefix + "ln_src_attn.", x)
        x = x + mha_alibi(sd, prefix + "mha_cross.", z, src, src, n_heads, cross_maps)
    h = F.linear(_ln(sd, prefix + "ln_ffnetwork.", x), sd[prefix + "ffnetwork.0.weight"])
    x = x + F.linear(F.gelu(h), sd[prefix + "ffnetwork.3.weight"])
    return x


def _count(sd, fmt: str) -> int:
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def gpt(sd, x: Tensor, n_heads: int = 4, maps: Optional[list] = None) -> Tensor:
    """vap/modules.py:342-358 (ar_channel); `maps` receives one (B, H, T, T) per layer."""
    for l in range(_count(sd, "ar_channel.layers.{}.ln_self_attn.weight")):
        x = transformer_layer(sd, f"ar_channel.layers.{l}.", x, None, n_heads, maps)
    return x


def gpt_stereo(sd, x1: Tensor, x2: Tensor, n_heads: int = 4, stages=None, maps: Optional[dict] = None):
    """vap/modules.py:380-408, 287-289 (both directions read the layer INPUT)
    and Combinator 446-449 (one shared LayerNorm). `maps`: dict of four lists
    (self_a, cross_a, self_b, cross_b; :385-395), one entry per layer."""
    m = maps if maps is not None else {"self_a": None, "cross_a": None, "self_b": None, "cross_b": None}
    for l in range(_count(sd, "ar.layers.{}.ln_self_attn.weight")):
        p = f"ar.layers.{l}."
        z1 = transformer_layer(sd, p, x1, x2, n_heads, m["self_a"], m["cross_a"])
        z2 = transformer_layer(sd, p, x2, x1, n_heads, m["self_b"], m["cross_b"])
        x1, x2 = z1, z2
        if stages is not None:
            stages[f"ar{l}_x1"], stages[f"ar{l}_x2"] = x1, x2
    ha = F.gelu(_ln(sd, "ar.combinator.ln.", F.linear(x1, sd["ar.combinator.h0_a.weight"])))
    hb = F.gelu(_ln(sd, "ar.combinator.ln.", F.linear(x2, sd["ar.combinator.h0_b.weight"])))
    return ha + hb, x1, x2


# --------------------------------------------------------------------------- #
# model facade                                                                #
# --------------------------------------------------------------------------- #
def forward(sd: Dict[str, Tensor], waveform: Tensor, n_heads: int = 4,
            stages: Optional[dict] = None, attention: bool = False) -> Dict[str, Tensor]:
    """vap/model.py:249-268. waveform (B, 2, S). attention=True adds self_attn (B, 2, Lc, H, T, T),
    cross_attn and cross_self_attn (B, 2, Lx, H, T, T) (:262-266, modules.py:397-406)."""
    assert waveform.shape[1] == 2, f"audio VAP ENCODER: {waveform.shape} != (B, 2, n_samples)"
    s1 = {} if stages is not None else None
    x1 = encoder(sd, waveform[:, :1], s1)
    x2 = encoder(sd, waveform[:, 1:])
    if stages is not None:
        stages.update({k + "_1": v for k, v in s1.items()})
        stages["enc_2"] = x2
    a1, a2 = ([], []) if attention else (None, None)
    sm = {"self_a": [], "cross_a": [], "self_b": [], "cross_b": []} if attention else None
    o1 = gpt(sd, x1, n_heads, a1)
    o2 = gpt(sd, x2, n_heads, a2)
    if stages is not None:
        stages["ch_1"], stages["ch_2"] = o1, o2
    x, x1, x2 = gpt_stereo(sd, o1, o2, n_heads, stages, sm)
    if stages is not None:
        stages["comb"] = x
    v1 = F.linear(x1, sd["va_classifier.weight"], sd["va_classifier.bias"])
    v2 = F.linear(x2, sd["va_classifier.weight"], sd["va_classifier.bias"])
    vad = torch.cat((v1, v2), dim=-1)
    logits = F.linear(x, sd["vap_head.weight"], sd["vap_head.bias"])
    ret = {"logits": logits, "vad": vad}
    if attention:
        ret["self_attn"] = torch.stack([torch.stack(a1, dim=1), torch.stack(a2, dim=1)], dim=1)
        ret["cross_attn"] = torch.stack([torch.stack(sm["cross_a"], dim=1), torch.stack(sm["cross_b"], dim=1)], dim=1)
        ret["cross_self_attn"] = torch.stack([torch.stack(sm["self_a"], dim=1), torch.stack(sm["self_b"], dim=1)],
                                             dim=1)
    return ret


def code_vectors(total_bins: int = 8) -> Tensor:
    """vap/objective.py:93-110 (bit i of class idx, LSB first)."""
    idx = torch.arange(2 ** total_bins)
    return torch.stack([(idx >> i) & 1 for i in range(total_bins)], dim=-1).float()


def probs_next_speaker_aggregate(probs: Tensor, from_bin: int, to_bin: int) -> Tensor:
    """vap/objective.py:184-204."""
    states = code_vectors(8).to(probs.dtype).to(probs.device).view(256, 2, 4)  # decode: (c b) -> c b
    abp = states[:, :, from_bin: to_bin + 1].sum(-1)
    p_all = torch.einsum("bid,dc->bic", probs, abp)
    p_all = p_all / (p_all.sum(-1, keepdim=True) + 1e-5)
    return p_all


def get_labels(va: Tensor) -> Tensor:
    """vap/objective.py:209-212 -> ProjectionWindow :53-72 -> Codebook.encode
    :112-139. va (B, T, 2) -> (B, T-100) class indices."""
    horizon = sum(BIN_FRAMES)
    win = va[..., 1:, :].unfold(dimension=-2, size=horizon, step=1)  # (B, N, 2, 100)
    start, bins = 0, []
    for b in BIN_FRAMES:
        m = win[..., start: start + b].sum(dim=-1) / b
        bins.append((m >= 0.5).to(va.dtype))
        start += b
    pw = torch.stack(bins, dim=-1)  # (B, N, 2, 4)
    flat = pw.reshape(-1, 8)
    embed = code_vectors(8).to(va.dtype).to(va.device).T
    dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ embed + embed.pow(2).sum(0, keepdim=True))
    return dist.max(dim=-1).indices.view(*pw.shape[:-2])


def loss_vap(logits: Tensor, labels: Tensor) -> Tensor:
    """vap/objective.py:220-243 with reduction='none'."""
    n = labels.shape[1]
    lg = logits[:, :n]
    loss = F.cross_entropy(lg.reshape(-1, lg.shape[-1]), labels.reshape(-1), reduction="none")
    return loss.view(-1, n)


def probs(sd: Dict[str, Tensor], waveform: Tensor, now_lims: List[int] = [0, 1],
          future_lims: List[int] = [2, 3], n_heads: int = 4) -> Dict[str, Tensor]:
    """vap/model.py:180-225, including the always-present `loss`:
    labels come from the model's own sigmoid(vad)."""
    with torch.no_grad():
        out = forward(sd, waveform, n_heads)
        p = out["logits"].softmax(dim=-1)
        vad = out["vad"].sigmoid()
        H = (-p * p.log2()).sum(dim=-1)
        ret = {
            "probs": p,
            "vad": vad,
            "p_now": probs_next_speaker_aggregate(p, now_lims[0], now_lims[-1]),
            "p_future": probs_next_speaker_aggregate(p, future_lims[0], future_lims[1]),
            "H": H,
        }
        labels = get_labels(vad)
        ret["loss"] = loss_vap(out["logits"], labels)
    return ret


def n_frames(n_samples: int):
    """Frame counts of the zero-padded conv chain:
    returns (L0..L4, T)."""
    L, out = n_samples, []
    for (k, s, p) in [(10, 5, 3), (8, 4, 2), (4, 2, 1), (4, 2, 1), (4, 2, 1)]:
        L = (L + 2 * p - k) // s + 1
        out.append(L)
    out.append((L - 1) // 2 + 1)
    return out


def step_extraction(sd, waveform: Tensor, sample_rate=16000, frame_hz=50,
                    context_time=20, step_time=5) -> Dict[str, Tensor]:
    """run.py:23-131: 25 s windows, 5 s hop, later windows contribute their last
    250 frames, right-aligned tail window for the remainder; `loss` from fold 0."""
    n_samples = waveform.shape[-1]
    duration = round(n_samples / sample_rate, 2)
    chunk_time = context_time + step_time
    step_samples = int(step_time * sample_rate)
    chunk_samples = int(chunk_time * sample_rate)
    step_frames = int(step_time * frame_hz)
    folds = waveform.unfold(dimension=-1, size=chunk_samples, step=step_samples).permute(2, 0, 1, 3)
    expected_frames = round(duration * frame_hz)
    out = probs(sd, folds[0])
    keys = ["vad", "p_now", "p_future", "probs", "H"]
    for w in folds[1:]:
        o = probs(sd, w)
        for k in keys:
            out[k] = torch.cat([out[k], o[k][:, -step_frames:]], dim=1)
    processed = out["p_now"].shape[1]
    if expected_frames != processed:
        omitted = expected_frames - processed
        o = probs(sd, waveform[..., -chunk_samples:])
        for k in keys:
            out[k] = torch.cat([out[k], o[k][:, -omitted:]], dim=1)
    return out


# --------------------------------------------------------------------------- #
# VapGPT.vad() post-processing                                                 #
# --------------------------------------------------------------------------- #
def _runs(x: Tensor):
    """vap/utils.py:21-49 find_island_idx_len as a plain loop: (start, length, value) of every run."""
    out, n, s = [], len(x), 0
    for t in range(1, n + 1):
        if t == n or x[t] != x[s]:
            out.append((s, t - s, float(x[s])))
            s = t
    return out


def vad_filter(vad: Tensor, max_fill_time: float = 0.02, max_omit_time: float = 0.02, frame_hz: float = 50) -> Tensor:
    """vap/model.py:240-247: per item vad_fill_silences (vap/utils.py:239-254) then vad_omit_spikes (:257-272).
    vad (B, T, 2) binary float; returns a new tensor."""
    v = vad.clone()
    fill, omit = round(max_fill_time * frame_hz), round(max_omit_time * frame_hz)
    for b in range(v.shape[0]):
        for ch in range(2):
            for s, d, val in _runs(v[b, :, ch]):
                if val == 0 and d <= fill:
                    v[b, s: s + d, ch] = 1.0
        for ch in range(2):
            for s, d, val in _runs(v[b, :, ch]):
                if val == 1 and d <= omit:
                    v[b, s: s + d, ch] = 0.0
    return v
