#!/usr/bin/env python
"""Hunts timing-dependent differences: runs one model's forward on a stream while a second model keeps the GPU busy
on another stream, and compares the first model's intermediate stages with a quiet run, bit for bit.
    python tools/race_probe.py [bf16|fp16] [B] [iters]"""
import os
import sys

import torch

os.environ["VAPB_PIPE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
ma, mb = [VapGPT(VapConfig(), precision=prec).to("cuda") for _ in range(2)]
ma.load_state_dict(sd)
mb.load_state_dict(sd)
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05
wn = torch.randn((96, 2, 320000), device="cuda", generator=g) * 0.05
STAGES = ["conv", "ar", "enc", "ch"]
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
with torch.cuda.stream(sa):
    ref = {s: ma.stage(s, w).clone() for s in STAGES}
    ref["logits"] = ma(w)["logits"].clone()
torch.cuda.synchronize()
nbad = 0
for it in range(iters):
    with torch.cuda.stream(sb):
        for _ in range(2):
            mb.probs(wn)
    with torch.cuda.stream(sa):
        name = (STAGES + ["logits"])[it % 5]
        got = ma(w)["logits"] if name == "logits" else ma.stage(name, w)
    torch.cuda.synchronize()
    if not torch.equal(got, ref[name]):
        nbad += 1
        d = (got != ref[name])
        seqs = d.flatten(1).any(1).nonzero().flatten().tolist()
        msg = f"iter {it} stage {name}: seqs {seqs[:8]}"
        for sq in seqs[:3]:
            rows = d[sq].any(-1).nonzero().flatten().tolist()
            cols = d[sq].any(0).nonzero().flatten().tolist()
            msg += f"\n    seq {sq}: {len(rows)} rows differ, first {rows[0]}, last {rows[-1]}; cols {len(cols)} first {cols[0]} last {cols[-1]}; max |d| {(got[sq] - ref[name][sq]).abs().max().item():.4g}"
        print(msg, flush=True)
print(f"{nbad} of {iters} noisy runs differed", flush=True)
