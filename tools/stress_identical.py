#!/usr/bin/env python
"""Run-to-run determinism under concurrency: N calls at B=256 x 20 s, unsplit and as concurrent item groups
(VAPB_PIPE), every call compared bit for bit with the first unsplit one.
    python tools/stress_identical.py [bf16|fp16] [calls]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 40
B = 256
sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05
ref = None
for pipe in (1, 4, 3, 8):
    os.environ["VAPB_PIPE"] = str(pipe)
    m = VapGPT(VapConfig(), precision=prec).to("cuda")
    m.load_state_dict(sd)
    out = m.alloc_outputs(B, 1000, "cuda", argmax=True)
    bad = 0
    for i in range(calls):
        m.probs(w, out=out)
        if ref is None:
            ref = {k: v.clone() for k, v in out.items()}
        elif not all(torch.equal(out[k], ref[k]) for k in ref):
            bad += 1
            items = (out["vad"] != ref["vad"]).flatten(1).any(1).nonzero().flatten().tolist()
            print(f"  pipe={pipe} call {i}: items {items[:10]}", flush=True)
    print(f"pipe={pipe}: {bad} of {calls} calls differ", flush=True)
    del m, out
    torch.cuda.empty_cache()
