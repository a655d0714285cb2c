"""CPU oracle for the device resampler (TEST INFRASTRUCTURE; never imported by the product path).

The reference resamples in `vap/audio.py:65-68` by calling `torchaudio.functional.resample(x, sr, 16000)` — a
third-party dependency that is not vendored in /root/reference (requirements.txt:1-9 pins no version; this image has
torchaudio 2.11.0). This file restates its published algorithm (functional.py: `_get_sinc_resample_kernel`,
sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99, float32; `_apply_sinc_resample_kernel`: zero padding
(width, width + orig), a stride-`orig` correlation with `new` phase filters, cut to ceil(new * n / orig)) in numpy.
Pinned by tests/test_oracle.py against torchaudio itself (when importable) and against
tests/golden/resample_example_24k_16k.npz, which oracle/make_golden_resample.py wrote with torchaudio.
"""
import math

import numpy as np


def bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    f = np.float32
    idx = np.arange(-width, width + orig, dtype=f)[None] / f(orig)
    t = np.arange(0, -new, -1, dtype=f)[:, None] / f(new) + idx
    t = t * f(base)
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width).astype(f)
    window = np.cos(t * f(math.pi) / f(lowpass_filter_width) / f(2)) ** 2
    t = t * f(math.pi)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, f(1.0), np.sin(t) / t)
    k = (k * (window * f(base / orig))).astype(f)
    return k, width, orig, new


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """x (..., n) float32 (or int16 PCM, scaled by 1/32768) -> (..., ceil(new * n / orig)) float32."""
    if x.dtype == np.int16:
        x = x.astype(np.float32) / np.float32(32768.0)
    k, width, orig, new = bank(orig_freq, new_freq)
    K = k.shape[1]
    n = x.shape[-1]
    rows = x.reshape(-1, n).astype(np.float32)
    pad = np.pad(rows, ((0, 0), (width, width + orig)))
    n_frames = (pad.shape[1] - K) // orig + 1
    # frames[r, j, :] = pad[r, j*orig : j*orig + K]
    sl = np.lib.stride_tricks.sliding_window_view(pad, K, axis=1)[:, ::orig][:, :n_frames]
    out = np.einsum("rjk,pk->rjp", sl, k, dtype=np.float32, optimize=False).reshape(rows.shape[0], -1)
    n_out = -(-new * n // orig)
    return out[:, :n_out].reshape(x.shape[:-1] + (n_out,)).astype(np.float32)
