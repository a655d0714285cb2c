"""Host-side helpers of the inference path, mirroring the reference's vap/utils.py
names and behaviour: `batch_to_device` (:106-113), `tensor_dict_to_json`
(:116-124), `write_json` / `read_json` (:287-295), and the run-length VAD
clean-up used by `VapGPT.vad` (`find_island_idx_len` :21-49,
`vad_fill_silences` :239-254, `vad_omit_spikes` :257-272)."""
from __future__ import annotations

import json
from typing import Tuple

import torch
from torch import Tensor


def batch_to_device(batch, device="cuda"):
    return {k: (v.to(device) if isinstance(v, Tensor) else v) for k, v in batch.items()}


def tensor_dict_to_json(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, Tensor):
            v = v.tolist()
        elif isinstance(v, dict):
            v = tensor_dict_to_json(v)
        out[k] = v
    return out


def write_json(data, filename):
    with open(filename, "w", encoding="utf-8") as f:
        json.dump(data, f, ensure_ascii=False)


def read_json(path, encoding="utf8"):
    with open(path, "r", encoding=encoding) as f:
        return json.loads(f.read())


def find_island_idx_len(x: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Run-length encoding of a 1-D tensor: (start index, length, value) per run."""
    assert x.ndim == 1
    n = len(x)
    change = torch.where(x[1:] != x[:-1])[0]
    last = torch.cat((change, torch.tensor([n - 1], device=x.device))).long()  # last index of each run
    bounds = torch.cat((torch.tensor([-1], device=x.device), last))
    dur = bounds[1:] - bounds[:-1]
    start = bounds[:-1] + 1
    return start, dur, x[last]


def _rewrite_short_runs(vad: Tensor, run_value: int, new_value: float, max_frames: int) -> Tensor:
    assert vad.ndim == 2 and vad.shape[-1] == 2, f"Expects (N_FRAMES, 2) got {tuple(vad.shape)}"
    for ch in range(2):
        # runs are taken from the column as it was on entry, like the reference
        start, dur, val = find_island_idx_len(vad[:, ch])
        sel = (val == run_value) & (dur <= max_frames)
        for s, d in zip(start[sel].tolist(), dur[sel].tolist()):
            vad[s : s + d, ch] = new_value
    return vad


def vad_fill_silences(vad: Tensor, max_fill_time: float = 0.02, frame_hz: float = 50) -> Tensor:
    """Silences of at most max_fill_time become active (in place)."""
    return _rewrite_short_runs(vad, 0, 1.0, round(max_fill_time * frame_hz))


def vad_omit_spikes(vad: Tensor, max_omit_time: float = 0.02, frame_hz: float = 50) -> Tensor:
    """Active runs of at most max_omit_time become silence (in place)."""
    return _rewrite_short_runs(vad, 1, 0.0, round(max_omit_time * frame_hz))
