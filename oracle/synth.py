"""Seeded synthetic VapGPT state dicts and waveforms (TEST INFRASTRUCTURE).

Every checkpoint of the reference is absent from /root/reference
(.MISSING_LARGE_BLOBS), so parity is demonstrated on synthetic weights that have
the reference's exact key/shape schema (reference
vap/model.py:125-163, vap/encoder_components.py:83-159, vap/modules.py:205-449).

The values come from numpy's PCG64 stream (stable across machines), not from
torch's RNG, so the authoring container and the GPU box build bit-identical
weights from a seed without shipping 23 MB files.

An input generator, not an implementation of the path: tests/, tools/ (diagnostic probes), __graft_entry__.smoke()
and bench.py import it for weights and waveforms; nothing under voiceactivityprojection_b200/ does.
"""
from __future__ import annotations

import math

import numpy as np
import torch

DIM = 256
FFN = 768
N_CLASSES = 256


def alibi_slopes(n: int):
    """Reference vap/modules.py:125-156 (power-of-two branch and the workaround)."""

    def pow2(n):
        start = 2 ** (-(2 ** -(math.log2(n) - 3)))
        return [start * start**i for i in range(n)]

    if math.log2(n).is_integer():
        return pow2(n)
    c = 2 ** math.floor(math.log2(n))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: n - c]


def code_vectors(total_bins: int = 8) -> torch.Tensor:
    """Reference vap/objective.py:93-110: class idx -> bits, LSB first."""
    idx = torch.arange(2**total_bins)
    return torch.stack([(idx >> i) & 1 for i in range(total_bins)], dim=-1).float()


class _Rng:
    def __init__(self, seed):
        self.g = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, shape, bound):
        return torch.from_numpy(
            self.g.uniform(-bound, bound, size=shape).astype(np.float32)
        )

    def normal(self, shape, std, mean=0.0):
        return torch.from_numpy(
            (self.g.standard_normal(size=shape) * std + mean).astype(np.float32)
        )


def make_state_dict(
    seed: int = 0,
    ar_mode: str = "LSTM",
    ar_layers: int = 1,
    gain: float = 1.0,
    channel_layers: int = 1,
    cross_layers: int = 3,
    num_heads: int = 4,
):
    """A state dict with the reference schema.

    `gain` scales the transformer/head Linear weights (reference init is
    N(0, 0.02), vap/modules.py:328-335, which gives almost flat softmaxes);
    gain>1 makes attention and class posteriors peaked like a trained model.
    Norm affine parameters are perturbed away from (1, 0) so that a kernel that
    forgot them fails the parity test.
    """
    r = _Rng(seed)
    sd = {}
    enc = "encoder.encoder.gEncoder."
    convs = [(1, 10), (DIM, 8), (DIM, 4), (DIM, 4), (DIM, 4)]
    for i, (cin, k) in enumerate(convs):
        b = 1.0 / math.sqrt(cin * k)
        sd[f"{enc}conv{i}.weight"] = r.uniform((DIM, cin, k), b)
        sd[f"{enc}conv{i}.bias"] = r.uniform((DIM,), b)
        sd[f"{enc}batchNorm{i}.weight"] = r.normal((1, DIM, 1), 0.1, 1.0)
        sd[f"{enc}batchNorm{i}.bias"] = r.normal((1, DIM, 1), 0.1)
    ng = {"LSTM": 4, "GRU": 3}[ar_mode]
    ar = "encoder.encoder.gAR.baseNet."
    b = 1.0 / math.sqrt(DIM)
    for l in range(ar_layers):
        sd[f"{ar}weight_ih_l{l}"] = r.uniform((ng * DIM, DIM), b)
        sd[f"{ar}weight_hh_l{l}"] = r.uniform((ng * DIM, DIM), b)
        sd[f"{ar}bias_ih_l{l}"] = r.uniform((ng * DIM,), b)
        sd[f"{ar}bias_hh_l{l}"] = r.uniform((ng * DIM,), b)
    b = 1.0 / math.sqrt(DIM * 5)
    sd["encoder.downsample.1.weight"] = r.uniform((DIM, DIM, 5), b)
    sd["encoder.downsample.1.bias"] = r.uniform((DIM,), b)
    sd["encoder.downsample.2.ln.weight"] = r.normal((DIM,), 0.1, 1.0)
    sd["encoder.downsample.2.ln.bias"] = r.normal((DIM,), 0.1)

    slopes = torch.tensor(alibi_slopes(num_heads))
    std = 0.02 * gain

    def layer(prefix, cross):
        for ln in ["ln_self_attn", "ln_ffnetwork"] + (["ln_src_attn"] if cross else []):
            sd[f"{prefix}{ln}.weight"] = r.normal((DIM,), 0.1, 1.0)
            sd[f"{prefix}{ln}.bias"] = r.normal((DIM,), 0.1)
        for mha in ["mha"] + (["mha_cross"] if cross else []):
            sd[f"{prefix}{mha}.m"] = slopes.clone()
            for w in ["key", "query", "value", "proj"]:
                sd[f"{prefix}{mha}.{w}.weight"] = r.normal((DIM, DIM), std)
        sd[f"{prefix}ffnetwork.0.weight"] = r.normal((FFN, DIM), std)
        sd[f"{prefix}ffnetwork.3.weight"] = r.normal((DIM, FFN), std)

    for l in range(channel_layers):
        layer(f"ar_channel.layers.{l}.", cross=False)
    for l in range(cross_layers):
        layer(f"ar.layers.{l}.", cross=True)
    sd["ar.combinator.h0_a.weight"] = r.normal((DIM, DIM), std)
    sd["ar.combinator.h0_b.weight"] = r.normal((DIM, DIM), std)
    sd["ar.combinator.ln.weight"] = r.normal((DIM,), 0.1, 1.0)
    sd["ar.combinator.ln.bias"] = r.normal((DIM,), 0.1)
    sd["objective.codebook.emb.weight"] = code_vectors(8)
    sd["va_classifier.weight"] = r.normal((1, DIM), std * 4)
    sd["va_classifier.bias"] = r.normal((1,), 0.1)
    sd["vap_head.weight"] = r.normal((N_CLASSES, DIM), std * 2)
    sd["vap_head.bias"] = r.normal((N_CLASSES,), 0.2)
    return sd


def make_waveform(batch: int, n_samples: int, seed: int = 0, kind: str = "noise"):
    """(B, 2, n_samples) float32 test audio.

    noise: 0.05 * N(0,1)                                  
    turns: per channel, 1-3 s on/off amplitude gates over noise plus a tone, so
           VAD / argmax outputs are not degenerate        (config 2 variant)
    mono : channel 1 identically zero (what run.py:219-220 builds; config 5)
    """
    g = np.random.Generator(np.random.PCG64(1000 + seed))
    x = (g.standard_normal(size=(batch, 2, n_samples)) * 0.05).astype(np.float32)
    if kind == "turns":
        t = np.arange(n_samples, dtype=np.float32) / 16000.0
        for b in range(batch):
            for c in range(2):
                gate = np.zeros(n_samples, dtype=np.float32)
                pos, on = 0, bool(g.integers(0, 2))
                while pos < n_samples:
                    d = int(g.uniform(1.0, 3.0) * 16000)
                    if on:
                        gate[pos : pos + d] = 1.0
                    pos, on = pos + d, not on
                f0 = g.uniform(90.0, 250.0)
                x[b, c] = gate * (x[b, c] * 2 + 0.1 * np.sin(2 * np.pi * f0 * t)) + (
                    1 - gate
                ) * x[b, c] * 0.02
    elif kind == "mono":
        x[:, 1] = 0.0
    elif kind != "noise":
        raise ValueError(kind)
    return torch.from_numpy(x)
