#!/usr/bin/env python
"""Long single chunks (run.py feeds up to 160 s = 8000 frames in one call): the tensor modes against the fp32 mode
and the fp32 mode against the CPU oracle.
    python tools/long_probe.py [seconds] [oracle:0|1]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, vap_oracle as O  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 160.0
with_oracle = len(sys.argv) > 2 and sys.argv[2] == "1"
sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
S = int(secs * 16000)
w = synth.make_waveform(1, 2, S) if hasattr(synth, "make_waveform") else None
if w is None or w.shape[0] != 2:
    g = torch.Generator().manual_seed(1)
    w = torch.randn((2, 2, S), generator=g) * 0.05
outs = {}
for prec in ("fp32", "bf16", "fp16"):
    m = VapGPT(VapConfig(), precision=prec).to("cuda")
    m.load_state_dict(sd)
    t0 = time.time()
    o = m.probs(w.cuda())
    torch.cuda.synchronize()
    outs[prec] = {k: v.cpu() for k, v in o.items()}
    print(prec, "T =", o["probs"].shape[1], f"{time.time() - t0:.2f} s (first call)", "nan:", any(torch.isnan(v).any().item() for v in o.values()), flush=True)
    del m
for prec in ("bf16", "fp16"):
    print(prec, "vs fp32:", {k: round((outs[prec][k] - outs["fp32"][k]).abs().max().item(), 6) for k in ("probs", "vad", "p_now", "p_future", "H")},
          "argmax agree", round((outs[prec]["probs"].argmax(-1) == outs["fp32"]["probs"].argmax(-1)).float().mean().item(), 4), flush=True)
if with_oracle:
    t0 = time.time()
    with torch.no_grad():
        ref = O.probs(sd, w)
    print(f"oracle {time.time() - t0:.1f} s")
    print("fp32 vs oracle:", {k: float(f"{(outs['fp32'][k] - ref[k]).abs().max().item():.3g}") for k in ("probs", "vad", "p_now", "p_future", "H")},
          "argmax equal", torch.equal(outs["fp32"]["probs"].argmax(-1), ref["probs"].argmax(-1)),
          "vad decisions equal", torch.equal(outs["fp32"]["vad"] >= 0.5, ref["vad"] >= 0.5), flush=True)
