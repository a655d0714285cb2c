"""Host-side helpers of the inference path, mirroring the reference's vap/utils.py
names and behaviour: `batch_to_device` (:106-113), `tensor_dict_to_json`
(:116-124), `write_json` / `read_json` (:287-295), and the run-length VAD
clean-up used by `VapGPT.vad` (`find_island_idx_len` :21-49,
`vad_fill_silences` :239-254, `vad_omit_spikes` :257-272), plus the voice-activity
data formats either side of the path: `add_zero_channel` (:15-18),
`get_dialog_states` (:130-138), `get_vad_list_subset` (:141-167),
`vad_list_to_onehot` (:170-195) and `vad_onehot_to_vad_list` (:198-236) — the
`[[start, end], ...]` per-speaker lists of `example/*_vad_list.json`.
Host code on small tensors; the device work is in csrc/."""
from __future__ import annotations

import json
from typing import List, Tuple

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


VAD_LIST = List[List[List[float]]]


def add_zero_channel(w: Tensor) -> Tensor:
    """A silent second speaker under a mono waveform (..., 1, n) -> (..., 2, n)."""
    return torch.cat((w, torch.zeros_like(w)), dim=-2)


def get_dialog_states(vad: Tensor) -> Tensor:
    """(..., 2) binary activity -> 0: only speaker 0, 1: silence, 2: both, 3: only speaker 1."""
    assert vad.ndim >= 1
    return (2 * vad[..., 1] - vad[..., 0]).long() + 1


def _frames(t: float, hop_time: float) -> int:
    return int(t / hop_time)  # vap/audio.py:18-19 (truncation, not rounding)


def get_vad_list_subset(vad_list: VAD_LIST, start_time: float, end_time: float) -> VAD_LIST:
    """Segments of each speaker that touch [start_time, end_time], clipped to it and made relative to its start
    (two decimals). Segment lists are sorted by start, so a channel stops at the first segment past the end."""
    span = end_time - start_time
    subset: VAD_LIST = [[], []]
    for ch, segments in enumerate(vad_list):
        for s, e in segments:
            if e < start_time:
                continue
            if s > end_time:
                break
            if s >= start_time and e <= end_time:      # inside
                subset[ch].append([round(s - start_time, 2), round(e - start_time, 2)])
            elif s <= start_time:                      # began before (or exactly at) the window
                subset[ch].append([0, round(e - start_time, 2) if e < end_time else span])
            elif s < end_time:                         # begins inside, runs past the end
                subset[ch].append([round(s - start_time, 2), span])
    return subset


def vad_list_to_onehot(vad_list: VAD_LIST, duration: float, hop_time: float = 0, frame_hz: float = 0,
                       channel_first: bool = False) -> Tensor:
    assert hop_time > 0 or frame_hz > 0, "vad_list_to_onehot requires `frame_hz` or `hop_time`"
    if frame_hz > 0:
        hop_time = 1 / frame_hz
    onehot = torch.zeros((_frames(duration, hop_time), 2))
    for ch, segments in enumerate(vad_list):
        for seg in segments:
            onehot[_frames(seg[0], hop_time): _frames(seg[1], hop_time), ch] = 1.0
    return onehot.permute(1, 0) if channel_first else onehot


def vad_onehot_to_vad_list(vad: Tensor, frame_hz: int = 50, ipu_thresh_time: float = 0.1) -> List[VAD_LIST]:
    """(B, n_frames, 2) binary activity -> per item, per speaker, [start, end] times (two decimals); activity runs
    closer than `ipu_thresh_time` to the end of the previous one are merged into it."""
    assert vad.ndim == 3, f"Expects vad with batch-dim of shape (B, n_frames, 2) but got {vad.shape}"
    batch = []
    for item in vad:
        per_speaker = []
        for ch in range(2):
            start, dur, val = find_island_idx_len(item[:, ch])
            on = val == 1
            # the reference divides torch tensors (float32) and rounds the Python floats of .tolist()
            begins = (start[on] / frame_hz).tolist()
            ends = (start[on] / frame_hz + dur[on] / frame_hz).tolist()
            segments: List[List[float]] = []
            for s, e in zip(begins, ends):
                s, e = round(s, 2), round(e, 2)
                if segments and s - segments[-1][1] < ipu_thresh_time:
                    segments[-1][1] = e
                else:
                    segments.append([s, e])
            per_speaker.append(segments)
        batch.append(per_speaker)
    return batch
