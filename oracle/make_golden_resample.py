"""Writes tests/golden/resample_example_24k_16k.npz with torchaudio.functional.resample (the function the reference
calls in vap/audio.py:65-68) on the first 1.5 s of the reference's example wav (24 kHz mono int16) and on seeded
noise at 48 kHz, 44.1 kHz and 8 kHz. Run in the authoring container:  python oracle/make_golden_resample.py"""
import os

import numpy as np
import torch
import torchaudio.functional as AF
from scipy.io import wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sr, d = wavfile.read("/root/reference/example/student_long_female_en-US-Wavenet-G.wav")
assert sr == 24000 and d.dtype == np.int16
pcm = d[: 36000].copy()
out = {"pcm_24k": pcm, "y_24k": AF.resample(torch.from_numpy(pcm).float().div(32768)[None], 24000, 16000)[0].numpy()}
g = torch.Generator().manual_seed(0)
for rate, n in ((48000, 9601), (44100, 4411), (8000, 2001)):
    x = torch.rand((2, n), generator=g) * 2 - 1
    out[f"x_{rate}"] = x.numpy()
    out[f"y_{rate}"] = AF.resample(x, rate, 16000).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resample_example_24k_16k.npz"), **out)
print({k: v.shape for k, v in out.items()})
