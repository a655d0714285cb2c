#!/usr/bin/env python
"""Step time of VapGPT.probs at B=256 x 20 s: eager launches against a CUDA-graph replay of the same call.
    python tools/graph_probe.py [bf16|fp16|fp32]"""
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
for B in (1, 16, 256):
    w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05
    out = m.alloc_outputs(B, 1000, "cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            m.probs(w, out=out)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            m.probs(w, out=out)

        def timed(fn, it=8):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(it):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / it

        te, tg = timed(lambda: m.probs(w, out=out)), timed(graph.replay)
    print(f"B={B}: eager {te:.3f} ms, graph replay {tg:.3f} ms")
    del w, out, graph
