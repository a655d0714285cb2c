#!/usr/bin/env python
"""Does splitting a B=256 step into micro-batches on concurrent streams hide the gAR recurrence (64 of 148 SMs busy,
chain-bound)? Times K micro-batches of 256/K chunks, each on its own stream and model handle, against one B=256 call.
    python tools/overlap_probe.py [bf16|fp16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
B, IT = 256, 6
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 2, 320000), device="cuda", generator=g) * 0.05


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(IT):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / IT


for K in (1, 2, 4):
    ms = [VapGPT(VapConfig(), precision=prec).to("cuda") for _ in range(K)]
    for m in ms:
        m.load_state_dict(sd)
    streams = [torch.cuda.Stream() for _ in range(K)]
    parts = list(w.chunk(K))
    outs = [m.alloc_outputs(B // K, 1000, "cuda") for m in ms]
    main = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event()
        ev.record(main)
        for m, s, p, o in zip(ms, streams, parts, outs):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                m.probs(p, out=o)
            e2 = torch.cuda.Event()
            e2.record(s)
            main.wait_event(e2)

    def step_seq():
        for m, p, o in zip(ms, parts, outs):
            m.probs(p, out=o)

    t_c, t_s = timed(step), timed(step_seq)
    print(f"K={K}: concurrent {t_c:.2f} ms ({B * 20 / t_c * 1e3:.0f} audio-s/s), sequential {t_s:.2f} ms")
    del ms, outs
