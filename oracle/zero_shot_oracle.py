"""CPU restatement (numpy) of the reference's ZeroShot marginals. TEST INFRASTRUCTURE: only tests/ may import it.

Follows vap/zero_shot.py: subset construction :101-158 (with the helpers :9-90), `marginal_probs` :159-165,
`probs_backchannel` :173-176, `probs_next_speaker` :226-262, `get_probs` :264-271; dialog states
vap/events.py:70-78; `Codebook.encode` on binary windows vap/objective.py:93-143 (class index = sum of
bit(c, b) << (4c + b)).  Pinned to the reference itself by tests/golden/zero_shot.npz (oracle/make_golden_zeroshot.py
runs the unmodified reference class): subsets bit-exact, probabilities within 1e-6.
"""
from __future__ import annotations

import itertools

import numpy as np

N_BINS = 4


def _encode(windows: np.ndarray) -> np.ndarray:
    """(..., 2, 4) binary projection windows -> class index (objective.py:112-139 for exact code vectors)."""
    w = 2 ** np.arange(2 * N_BINS)
    return (windows.reshape(*windows.shape[:-2], 2 * N_BINS).astype(np.int64) * w).sum(-1)


def _active_at_end(min_active: int = 2) -> np.ndarray:
    """on_activity_change_mono (:32-60): last `min_active` bins on, every pattern of the others."""
    free = N_BINS - min_active
    rows = [list(bits) + [1] * min_active for bits in itertools.product((0, 1), repeat=free)]
    return np.array(rows, dtype=np.int64)


def _end_of_segment(mx: int = 2) -> np.ndarray:
    """end_of_segment_mono (:9-19): 0000, 1000, 1100, ... (mx + 1 rows)."""
    v = np.zeros((mx + 1, N_BINS), dtype=np.int64)
    for i in range(mx):
        v[i + 1, : i + 1] = 1
    return v


def _pairs(x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """combine_speakers(mirror=True) (:63-76): [0] = (speaker 0 from x1, speaker 1 from x2), [1] = swapped."""
    vad = np.array([[a, b] for a in x1 for b in x2])
    return np.stack([vad, vad[:, ::-1]])


def subsets():
    zero = np.zeros((1, N_BINS), dtype=np.int64)
    nav = _active_at_end(2)
    sil = np.sort(_encode(_pairs(nav, zero)), axis=-1)             # :101-123
    act = np.sort(_encode(_pairs(nav, _end_of_segment(2))), axis=-1)  # :125-132
    act_hold = np.sort(_encode(_pairs(zero, nav)), axis=-1)        # :134-140
    three = np.array(list(itertools.product((0, 1), repeat=3)), dtype=np.int64)
    bc_speaker = np.concatenate([three[1:], np.zeros((7, 1), dtype=np.int64)], -1)  # :146-151
    current = np.concatenate([three, np.ones((8, 1), dtype=np.int64)], -1)          # :153-155
    bc = _encode(_pairs(bc_speaker, current))                       # :157-158
    return {"subset_silence": sil, "subset_silence_hold": sil[::-1].copy(), "subset_active": act,
            "subset_active_hold": act_hold, "bc_prediction": bc}


def _softmax(x: np.ndarray) -> np.ndarray:
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def marginal(probs: np.ndarray, pos: np.ndarray, neg: np.ndarray) -> np.ndarray:
    out = []
    for s in (0, 1):
        joint = np.concatenate([pos[s], neg[s]])
        out.append(probs[..., pos[s]].sum(-1) / probs[..., joint].sum(-1))
    return np.stack(out, -1)


def get_probs(logits: np.ndarray, va: np.ndarray, is_probs: bool = False):
    """-> dict p, p_bc, p_sil, p_act, each (B, T, 2) in the input's float type."""
    ss = subsets()
    probs = logits if is_probs else _softmax(logits)
    T = probs.shape[-2]
    va = va[:, :T]
    sil = marginal(probs, ss["subset_silence"], ss["subset_silence_hold"])
    act = marginal(probs, ss["subset_active"], ss["subset_active_hold"])
    bc = np.stack([probs[..., ss["bc_prediction"][s]].sum(-1) for s in (0, 1)], -1)
    ds = np.trunc(2 * va[..., 1] - va[..., 0]).astype(np.int64) + 1
    pa = np.zeros(ds.shape, dtype=probs.dtype)
    pb = np.zeros(ds.shape, dtype=probs.dtype)
    m = ds == 1
    pa[m], pb[m] = sil[m][:, 0], sil[m][:, 1]
    m = ds == 0
    pa[m], pb[m] = 1 - act[m][:, 1], act[m][:, 1]
    m = ds == 3
    pa[m], pb[m] = act[m][:, 0], 1 - act[m][:, 0]
    m = ds == 2
    tot = act[m][:, 0] + act[m][:, 1]
    pa[m], pb[m] = act[m][:, 0] / tot, act[m][:, 1] / tot
    return {"p": np.stack([pa, pb], -1), "p_bc": bc, "p_sil": sil, "p_act": act}
