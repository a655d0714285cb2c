"""Writes tests/golden/zero_shot.npz with the UNMODIFIED reference's ZeroShot (vap/zero_shot.py:94-271):
the class-index subsets it builds (:101-158) and `get_probs(logits, va)` / `probs_on_silence` / `probs_on_active`
(:159-271) on seeded logits and binary voice activity covering all four dialog states (vap/events.py:70-78).
TEST INFRASTRUCTURE. Run in the authoring container: python oracle/make_golden_zeroshot.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from vap.zero_shot import ZeroShot  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
zs = ZeroShot(bin_times=[0.2, 0.4, 0.6, 0.8], frame_hz=50)
out = {
    "subset_silence": zs.subset_silence.numpy(),
    "subset_silence_hold": zs.subset_silence_hold.numpy(),
    "subset_active": zs.subset_active.numpy(),
    "subset_active_hold": zs.subset_active_hold.numpy(),
    "bc_prediction": zs.bc_prediction.numpy(),
}
for k, v in out.items():
    print(k, v.shape, v.tolist() if v.size <= 24 else "...")
g = torch.Generator().manual_seed(0)
for name, (B, T, Tva, scale) in {
    "flat": (3, 64, 64, 1.0),       # near-uniform class distribution
    "peaked": (2, 250, 300, 6.0),   # peaked like a trained model; va longer than logits (va[:, :nmax], :268)
    "odd": (5, 37, 37, 3.0),        # row count not a multiple of the kernel's rows per warp
}.items():
    logits = torch.randn((B, T, 256), generator=g) * scale
    flips = (torch.rand((B, Tva, 2), generator=g) < 0.15).long()
    va = (flips.cumsum(1) % 2).float()
    with torch.no_grad():
        r = zs.get_probs(logits, va)
        probs = logits.softmax(-1)
        sil = zs.probs_on_silence(probs)
        act = zs.probs_on_active(probs)
    out[name + "_logits"] = logits.numpy()
    out[name + "_va"] = va.numpy()
    out[name + "_p"] = r["p"].numpy()
    out[name + "_p_bc"] = r["p_bc"].numpy()
    out[name + "_p_sil"] = sil.numpy()
    out[name + "_p_act"] = act.numpy()
    ds = (2 * va[:, :T, 1] - va[:, :T, 0]).long() + 1
    print(name, tuple(logits.shape), "dialog states", torch.bincount(ds.flatten(), minlength=4).tolist())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "zero_shot.npz"), **out)
