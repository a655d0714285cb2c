"""Writes tests/golden/attention_maps_T70.npz with the UNMODIFIED reference's `VapGPT.forward(waveform,
attention=True)` (vap/model.py:249-268; maps from vap/modules.py:82-110, stacked :342-358, :380-408) on CPU fp32:
seeded synthetic weights (oracle.synth, the shipped checkpoints are absent) and a seeded 1.4 s stereo waveform,
T = 70 frames (two 64-query tiles in the kernel). TEST INFRASTRUCTURE. Run in the authoring container:
    python -m oracle.make_golden_attention"""
import os

import numpy as np
import torch

from oracle import ref_import, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECIPE = dict(seed=11, ar_mode="LSTM", ar_layers=1, gain=2.0, batch=1, n_samples=22400, wav_seed=5, kind="turns")

if __name__ == "__main__":
    sd = synth.make_state_dict(RECIPE["seed"], RECIPE["ar_mode"], RECIPE["ar_layers"], RECIPE["gain"])
    wav = synth.make_waveform(RECIPE["batch"], RECIPE["n_samples"], RECIPE["wav_seed"], RECIPE["kind"])
    ref = ref_import.build_reference(sd, RECIPE["ar_mode"], RECIPE["ar_layers"])
    with torch.no_grad():
        out = ref(wav, attention=True)
    arrays = {k: v.numpy() for k, v in out.items()}
    for k, v in arrays.items():
        print(k, v.shape, float(np.abs(v).max()))
    # the maps are exact zeros above the diagonal and rows sum to one
    for k in ("self_attn", "cross_attn", "cross_self_attn"):
        a = arrays[k]
        assert np.all(np.triu(a, 1) == 0) and np.abs(a.sum(-1) - 1).max() < 1e-5
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "attention_maps_T70.npz"), recipe=repr(RECIPE), **arrays)
