"""Writes tests/golden/vad_filter.npz with the UNMODIFIED reference's vad_fill_silences / vad_omit_spikes
(vap/utils.py:239-272), applied in VapGPT.vad()'s order (vap/model.py:240-247), on seeded binary activity with short
and long runs, including runs that touch both ends. Run in the authoring container: python oracle/make_golden_vadfilter.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from vap.utils import vad_fill_silences, vad_omit_spikes  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = torch.Generator().manual_seed(0)
cases = {}
for name, (B, T, p_flip, fill, omit) in {
    "short_runs": (6, 200, 0.35, 0.02, 0.02),
    "long_runs": (4, 1000, 0.04, 0.02, 0.02),
    "wide_filters": (5, 300, 0.2, 0.06, 0.1),
    "no_filter": (2, 50, 0.3, 0.0, 0.0),
    "tiny": (3, 2, 0.5, 0.02, 0.02),
}.items():
    flips = (torch.rand((B, T, 2), generator=g) < p_flip).long()
    v = (flips.cumsum(1) % 2).float()
    v[0, :, 0] = 0.0  # all silent
    if B > 1:
        v[1, :, 1] = 1.0  # all active
    ref = v.clone()
    for b in range(B):
        ref[b] = vad_fill_silences(ref[b], max_fill_time=fill, frame_hz=50)
        ref[b] = vad_omit_spikes(ref[b], max_omit_time=omit, frame_hz=50)
    cases[name + "_in"] = v.numpy()
    cases[name + "_out"] = ref.numpy()
    cases[name + "_par"] = np.array([fill, omit])
    print(name, v.shape, int((v != ref).sum()), "frames changed")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vad_filter.npz"), **cases)
