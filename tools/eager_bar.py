#!/usr/bin/env python
"""The library-kernel bar: the torch restatement of the reference forward (oracle/vap_oracle.py, the
same ATen ops the reference calls: cuDNN conv1d / LSTM, cuBLAS matmuls, eager softmax) run on the B200 in fp32, TF32
and bf16 autocast, against this library on the same input. A measurement tool, not part of the product path.
    python tools/eager_bar.py [B] > profiles/<name>.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, vap_oracle as O  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
sd_dev = {k: v.cuda() for k, v in sd.items()}
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05


def timed(fn, it=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


rows = []
with torch.no_grad():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = O.probs(sd_dev, w)
    rows.append(("torch eager fp32 (cuDNN/cuBLAS, no TF32)", timed(lambda: O.probs(sd_dev, w)), None))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    rows.append(("torch eager TF32", timed(lambda: O.probs(sd_dev, w)), None))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.probs(sd_dev, w)
        err = (out["probs"].float() - ref["probs"]).abs().max().item()
        rows.append(("torch eager bf16 autocast", timed(lambda: O.probs(sd_dev, w)), err))
for prec in ("fp32", "bf16", "fp16"):
    m = VapGPT(VapConfig(), precision=prec).to("cuda")
    m.load_state_dict(sd)
    out = m.probs(w)
    err = (out["probs"] - ref["probs"]).abs().max().item()
    rows.append((f"this library, {prec}", timed(lambda: m.probs(w), 5), err))
    del m
print(f"B = {B} stereo chunks of 20 s, synthetic weights/input, one B200; probs() end to end, device-timed\n")
print("| path | ms | audio-s/s | max abs probs error vs eager fp32 |")
print("|---|---:|---:|---:|")
for name, ms, err in rows:
    print(f"| {name} | {ms:.2f} | {B * 20 / ms * 1e3:.0f} | {'' if err is None else f'{err:.2e}'} |")
