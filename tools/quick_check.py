#!/usr/bin/env python
"""Ten-second sanity run of the default path in every mode on one small batch (no oracle): finite outputs and the
modes agreeing with each other. For a last look at a build when GPU time is short; the parity tests are the real check."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

sd = synth.make_state_dict(7, "LSTM", 1, 2.0)
wav = synth.make_waveform(2, 40000, 7, "turns").cuda()
m = VapGPT(VapConfig()).to("cuda:0")
m.load_state_dict(sd)
ref = None
for prec in ("fp32", "bf16", "fp16"):
    o = m.probs(wav, precision=prec)
    torch.cuda.synchronize()
    ok = all(bool(torch.isfinite(v).all()) for v in o.values())
    ref = o if ref is None else ref
    print(prec, "finite" if ok else "NOT FINITE", "max |p_now - fp32| =", float((o["p_now"] - ref["p_now"]).abs().max()),
          flush=True)
print("launches", m.launch_count())
