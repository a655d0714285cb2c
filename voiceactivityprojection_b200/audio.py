"""Waveform loading for the inference CLI, mirroring the reference's
`vap/audio.py:load_waveform` (:39-69): returns (channels, n_samples) float32 in
[-1, 1) at `sample_rate`, optionally mono / normalised / cropped.

The reference decodes with `torchaudio.load`; that backend (TorchCodec) is not
part of this image, so PCM wav files are decoded with the standard library /
scipy and scaled exactly like torchaudio's normalisation (int16 / 32768), then
resampled with `torchaudio.functional.resample` as the reference does (:65-68).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor


def _read_wav(path: str) -> Tuple[Tensor, int]:
    from scipy.io import wavfile

    sr, data = wavfile.read(path)
    if data.ndim == 1:
        data = data[:, None]
    if data.dtype == np.int16:
        x = torch.from_numpy(data.astype(np.float32)) / 32768.0
    elif data.dtype == np.int32:
        x = torch.from_numpy(data.astype(np.float64) / 2147483648.0).float()
    elif data.dtype == np.uint8:
        x = (torch.from_numpy(data.astype(np.float32)) - 128.0) / 128.0
    else:
        x = torch.from_numpy(data.astype(np.float32))
    return x.t().contiguous(), int(sr)


def time_to_samples(t: float, sample_rate: int) -> int:
    return int(t * sample_rate)


def time_to_frames(t: float, hop_time: float) -> int:
    return int(t / hop_time)  # vap/audio.py:18-19


def sample_to_time(n_samples: int, sample_rate: int) -> float:
    return n_samples / sample_rate  # vap/audio.py:22-23


def get_audio_info(audio_path: str) -> Dict[str, Any]:
    """vap/audio.py:26-36 for PCM wav files, from the file header (no decode). Same keys; `num_channels` is the
    channel count (the reference stores `bits_per_sample` under that key, :34)."""
    import wave

    with wave.open(audio_path, "rb") as f:
        n, sr, width, ch = f.getnframes(), f.getframerate(), f.getsampwidth(), f.getnchannels()
    return {"name": audio_path, "duration": sample_to_time(n, sr), "sample_rate": sr, "num_frames": n,
            "bits_per_sample": 8 * width, "num_channels": ch, "encoding": "PCM_U" if width == 1 else "PCM_S"}


def load_waveform(path: str, sample_rate: Optional[int] = 16000, start_time: Optional[float] = None,
                  end_time: Optional[float] = None, mono: bool = False, *, normalize: bool = False,
                  audio_normalize_threshold: float = 0.05, device=None) -> Tuple[Tensor, int]:
    """vap/audio.py:39-69: same positional parameters and the same default (resample to 16 kHz; sample_rate=None keeps
    the file's rate). normalize / audio_normalize_threshold / device are keyword-only extras."""
    x, sr = _read_wav(path)
    if start_time is not None or end_time is not None:
        s = time_to_samples(start_time, sr) if start_time is not None else 0
        e = time_to_samples(end_time, sr) if end_time is not None else x.shape[-1]
        x = x[:, s:e]
    if normalize and x.shape[0] > 1:
        if x.abs().max() > audio_normalize_threshold:
            x = x / x.abs().max()
    if mono and x.shape[0] > 1:
        x = x.mean(dim=0, keepdim=True)
    if device is not None:
        x = x.to(device)
    if sample_rate is not None and sr != sample_rate:
        if x.is_cuda:  # device=...: the file's samples are already on the GPU; resample there with our own kernel
            x = resample_device(x, sr, sample_rate)
        else:
            import torchaudio.functional as AF

            x = AF.resample(x, orig_freq=sr, new_freq=sample_rate)
        sr = sample_rate
    return x, sr


# --------------------------------------------------------------------------- #
# device resampler (vapb_resample, csrc/k_resample.cu)                         #
# --------------------------------------------------------------------------- #
def sinc_resample_bank(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6,
                       rolloff: float = 0.99) -> Tuple[Tensor, int, int, int]:
    """The windowed-sinc polyphase bank of `torchaudio.functional.resample` (torchaudio 2.11 functional.py
    `_get_sinc_resample_kernel`, sinc_interp_hann, evaluated in float32 as `resample` does for a float32
    waveform). Returns (bank (new, 2*width + orig) float32 CPU, width, orig, new) with orig/new reduced by
    their gcd. Host-side weight preparation for vapb_resample."""
    import math

    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("frequencies must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float32)[None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float32)[:, None] / new + idx
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    bank = torch.where(t == 0, torch.tensor(1.0), t.sin() / t)
    bank *= window * (base / orig)
    return bank.contiguous(), width, orig, new


_BANKS = {}


def resample_device(x: Tensor, orig_freq: int, new_freq: int, interleaved: bool = False,
                    out: Optional[Tensor] = None) -> Tensor:
    """`torchaudio.functional.resample(x, orig_freq, new_freq)` on the GPU through vapb_resample.
    x: CUDA float32 or int16 (PCM, scaled by 1/32768 on the fly). Planar (..., n) -> (..., n_out); with
    interleaved=True x is (items, n, channels) as a wav file or a sound card delivers it and the result is
    planar (items, channels, n_out) float32, the layout VapGPT takes. n_out = ceil(new * n / orig)."""
    import ctypes as C

    from . import _lib

    if x.device.type != "cuda":
        raise RuntimeError("resample_device needs a CUDA tensor (no CPU fallback); use torchaudio on the host")
    if x.dtype not in (torch.float32, torch.int16):
        raise TypeError(f"expected float32 or int16 samples, got {x.dtype}")
    key = (int(orig_freq), int(new_freq), x.device)
    if key not in _BANKS:
        bank, width, orig, new = sinc_resample_bank(orig_freq, new_freq)
        _BANKS[key] = (bank.to(x.device), width, orig, new)
    bank, width, orig, new = _BANKS[key]
    x = x.contiguous()
    if interleaved:
        if x.ndim != 3:
            raise ValueError("interleaved input must be (items, n_samples, channels)")
        items, n, ch = x.shape
        strides = (n * ch, 1, ch)
        shape = (items, ch)
    else:
        n = x.shape[-1]
        items, ch = (x.numel() // n if n else 0), 1
        strides = (n, 0, 1)
        shape = tuple(x.shape[:-1])
    n_out = -(-new * n // orig)
    if out is None:
        out = torch.empty(shape + (n_out,), dtype=torch.float32, device=x.device)
    elif (tuple(out.shape[:-1]) != shape or out.shape[-1] > n_out or out.dtype != torch.float32
          or not out.is_contiguous()):
        raise ValueError(f"out must be contiguous float32 {shape + (n_out,)} (or shorter in the last dimension)")
    n_out = out.shape[-1]
    lib = _lib.load()
    if out.numel() == 0:  # empty input: nothing to launch (empty tensors have no device pointer)
        return out
    st = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        rc = lib.vapb_resample(None, st, x.data_ptr(), 1 if x.dtype == torch.int16 else 0, items, ch, n, strides[0],
                               strides[1], strides[2], orig, new, width, bank.data_ptr(), out.data_ptr(), n_out, n_out)
    if rc != 0:
        msg = lib.vapb_last_error(None)
        raise _lib.VapbError(f"vapb error {rc}: {msg.decode() if msg else ''}")
    return out
