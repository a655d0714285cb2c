#!/usr/bin/env python
"""Are an item's outputs independent of the batch it is computed in? Compares sub-batches of several sizes and offsets
against the same items inside one B=256 call, bit for bit.
    python tools/partition_probe.py [bf16|fp16|fp32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = 256
m = VapGPT(VapConfig(), precision=prec).to("cuda")
m.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05
full = {k: v.clone() for k, v in m.probs(w).items()}
fl = m(w)["logits"].clone()
for b0, nb in [(0, 42), (42, 43), (85, 43), (0, 43), (0, 64), (0, 1), (255, 1), (100, 7), (128, 128), (0, 85), (3, 33)]:
    sub = m.probs(w[b0:b0 + nb])
    sl = m(w[b0:b0 + nb])["logits"]
    bad = {k: (sub[k] - full[k][b0:b0 + nb]).abs().max().item() for k in full if not torch.equal(sub[k], full[k][b0:b0 + nb])}
    dl = (sl - fl[b0:b0 + nb]).abs()
    items = dl.flatten(1).max(1).values.nonzero().flatten().tolist()
    print(f"items [{b0}, {b0 + nb}): " + ("identical" if not bad and not items else f"DIFF {bad} logits {dl.max().item():.3g} items {items[:10]}"),
          flush=True)
