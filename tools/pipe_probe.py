#!/usr/bin/env python
"""Item-group pipelining (VAPB_PIPE, model.h): step time at B=256 x 20 s against the number of groups, and that the
outputs equal the unsplit call bit for bit.
    python tools/pipe_probe.py [bf16|fp16] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05
ref = None
for pipe in [int(x) for x in os.environ.get("PIPES", "1,2,3,4,6,8").split(",")]:
    os.environ["VAPB_PIPE"] = str(pipe)
    m = VapGPT(VapConfig(), precision=prec).to("cuda")
    m.load_state_dict(sd)
    out = m.alloc_outputs(B, 1000, "cuda", argmax=True)
    for _ in range(3):
        m.probs(w, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            m.probs(w, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 4)
    ts.sort()
    same = ""
    if ref is None:
        ref = {k: v.clone() for k, v in out.items()}
    else:
        bad = {}
        for k in ref:
            if not torch.equal(out[k], ref[k]):
                d = (out[k].float() - ref[k].float()).abs().flatten(1).max(1).values
                bad[k] = (round(d.max().item(), 6), d.nonzero().flatten().tolist()[:12])
        same = " identical" if not bad else f" DIFF {bad}"
        if bad:
            dv = (out["vad"] != ref["vad"])  # (B, T, 2)
            for it in dv.flatten(1).any(1).nonzero().flatten().tolist():
                for c in range(2):
                    fr = dv[it, :, c].nonzero().flatten().tolist()
                    if fr:
                        same += f"\n    item {it} ch{c}: {len(fr)} frames differ, first {fr[0]}, last {fr[-1]}"
    print(f"pipe={pipe}: {ts[1]:.2f} ms ({B * 20 / ts[1] * 1e3:.0f} audio-s/s){same}", flush=True)
    del m, out
    torch.cuda.empty_cache()
