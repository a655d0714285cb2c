#!/usr/bin/env python
"""Timeline of one pipelined call (VAPB_PIPE_TRACE=1 prints each group's phase boundaries to stderr).
    VAPB_PIPE=2 python tools/pipe_trace.py [bf16] [B]"""
import os
import sys

import torch

os.environ["VAPB_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = VapGPT(VapConfig(), precision=prec).to("cuda")
m.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))
w = torch.randn((B, 2, 320000), device="cuda") * 0.05
out = m.alloc_outputs(B, 1000, "cuda")
for i in range(3):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    m.probs(w, out=out)
    torch.cuda.synchronize()
