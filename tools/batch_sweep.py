#!/usr/bin/env python
"""Throughput of VapGPT.probs against batch size (20 s stereo chunks), device-timed. Needs a B200.
    python tools/batch_sweep.py [bf16|fp16|fp32] > profiles/<name>.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
m = VapGPT(VapConfig(), precision=prec).to("cuda")
m.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))
g = torch.Generator(device="cuda").manual_seed(0)
print(f"precision {prec}, 20 s stereo chunks, median of 5 CUDA-event timings after 2 warm-ups\n")
print("| B | ms | audio-s/s |")
print("|---:|---:|---:|")
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 384):
    w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05
    out = m.alloc_outputs(B, 1000, "cuda")
    for _ in range(2):
        m.probs(w, out=out)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.probs(w, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"| {B} | {ts[2]:.2f} | {B * 20 / ts[2] * 1e3:.0f} |")
    del w, out
