"""Waveform loading for the inference CLI, mirroring the reference's
`vap/audio.py:load_waveform` (:39-69): returns (channels, n_samples) float32 in
[-1, 1) at `sample_rate`, optionally mono / normalised / cropped.

The reference decodes with `torchaudio.load`; that backend (TorchCodec) is not
part of this image, so PCM wav files are decoded with the standard library /
scipy and scaled exactly like torchaudio's normalisation (int16 / 32768), then
resampled with `torchaudio.functional.resample` as the reference does (:65-68).
"""
from __future__ import annotations

from typing import Optional, Tuple

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


def load_waveform(path: str, sample_rate: Optional[int] = None, start_time: Optional[float] = None,
                  end_time: Optional[float] = None, normalize: bool = False, mono: bool = False,
                  audio_normalize_threshold: float = 0.05, device=None) -> Tuple[Tensor, int]:
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
        import torchaudio.functional as AF

        x = AF.resample(x, orig_freq=sr, new_freq=sample_rate)
        sr = sample_rate
    return x, sr
