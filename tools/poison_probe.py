#!/usr/bin/env python
"""Does any kernel read scratch memory it did not write in the same call? Fills the workspace with 0xFF bytes (NaN in
every format) before a call and compares the outputs with an unpoisoned call, bit for bit.
    [VAPB_PIPE=6] python tools/poison_probe.py [bf16|fp16|fp32] [B] [n_samples]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = int(sys.argv[3]) if len(sys.argv) > 3 else 320000
m = VapGPT(VapConfig(), precision=prec).to("cuda")
m.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, S), device="cuda", generator=g) * 0.05
ref = {k: v.clone() for k, v in m.probs(w).items()}
for fill in (0xFF, 0x00, 0x7F):
    for ws in m._ws.values():
        ws.fill_(fill)
    out = m.probs(w)
    torch.cuda.synchronize()
    bad = {}
    for k in ref:
        if not torch.equal(out[k], ref[k]):
            d = (out[k] - ref[k]).abs().flatten(1)
            items = (d.max(1).values != 0).nonzero().flatten().tolist()
            nan = torch.isnan(out[k]).sum().item()
            bad[k] = (d.nan_to_num(1e9).max().item(), nan, items[:8], len(items))
    print(f"fill 0x{fill:02X}: " + ("identical" if not bad else f"DIFF {bad}"), flush=True)
