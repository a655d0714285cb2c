"""Long-session extraction: the reference's `run.py:step_extraction` (:23-131)
with the windows BATCHED instead of looped.

Reference semantics (kept bit-for-bit in which frames are taken from which
window): 25 s windows (20 s context + 5 s step), hop 5 s; window 0 contributes
all its 1250 frames, every later window its last 250 frames; if frames are
missing at the end, a right-aligned window supplies them; `loss` comes from
window 0 only. The reference runs one B=1 forward per window; here all windows
of the session are gathered into micro-batches of `max_batch` windows on the
device and stitched with slices, which is what a B200 wants (independent
windows = batch dimension).
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import Tensor

STITCH_KEYS = ["vad", "p_now", "p_future", "probs", "H"]


def window_plan(n_samples: int, sample_rate: int = 16000, frame_hz: int = 50, context_time: float = 20,
                step_time: float = 5):
    """Start sample of every window and how many trailing frames each contributes."""
    duration = round(n_samples / sample_rate, 2)
    chunk_samples = int((context_time + step_time) * sample_rate)
    step_samples = int(step_time * sample_rate)
    step_frames = int(step_time * frame_hz)
    if n_samples < chunk_samples:
        raise RuntimeError(
            f"maximum size for tensor at dimension 2 is {n_samples} but size is {chunk_samples}")
    n_folds = (n_samples - chunk_samples) // step_samples + 1
    starts = [i * step_samples for i in range(n_folds)]
    return dict(duration=duration, chunk_samples=chunk_samples, step_frames=step_frames, starts=starts,
                expected_frames=round(duration * frame_hz))


@torch.no_grad()
def step_extraction(waveform: Tensor, model, device="cuda", context_time: float = 20, step_time: float = 5,
                    max_batch: int = 64, precision=None, to_cpu: bool = True, **_ignored) -> Dict[str, Tensor]:
    """waveform (B, 2, n_samples) -> stitched {probs, vad, p_now, p_future, H, loss}."""
    assert waveform.ndim == 3 and waveform.shape[1] == 2
    plan = window_plan(waveform.shape[-1], model.sample_rate, model.frame_hz, context_time, step_time)
    B = waveform.shape[0]
    cs, sf = plan["chunk_samples"], plan["step_frames"]
    wav = waveform.to(device)
    # (n_folds, B, 2, chunk) as a strided view; micro-batches are materialised contiguously
    folds = wav.unfold(dimension=-1, size=cs, step=int(step_time * model.sample_rate)).permute(2, 0, 1, 3)
    nf = folds.shape[0]
    kw = {} if precision is None else {"precision": precision}
    out = {k: v.clone() for k, v in model.probs(folds[0].contiguous(), **kw).items()}
    tails = {k: [] for k in STITCH_KEYS}
    per = max(1, max_batch // B)
    for i in range(1, nf, per):
        j = min(nf, i + per)
        o = model.probs(folds[i:j].reshape((j - i) * B, 2, cs), **kw)
        for k in STITCH_KEYS:
            v = o[k][:, -sf:]
            v = v.reshape(j - i, B, *v.shape[1:]).transpose(0, 1)  # (B, windows, step_frames, ...)
            tails[k].append(v.reshape(B, (j - i) * sf, *v.shape[3:]))
    for k in STITCH_KEYS:
        out[k] = torch.cat([out[k]] + tails[k], dim=1)
    processed = out["p_now"].shape[1]
    if plan["expected_frames"] != processed:
        omitted = plan["expected_frames"] - processed
        o = model.probs(wav[..., -cs:].contiguous(), **kw)
        for k in STITCH_KEYS:
            out[k] = torch.cat([out[k], o[k][:, -omitted:]], dim=1)
    if to_cpu:
        out = {k: v.cpu() for k, v in out.items()}
    return out
