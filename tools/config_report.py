#!/usr/bin/env python
"""Measurements for BASELINE.json configs[2] (10-minute session as overlapping windows: throughput and
streaming latency), configs[4] (mono input with a silent second channel, 10 s chunks) and the B=1
full-window latency the reference's realtime loop pays every poll (sds/run_sds.py:241).

    python tools/config_report.py [bf16|fp32] > profiles/<name>.md      (needs a B200)
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402
from voiceactivityprojection_b200.session import step_extraction, window_plan  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    m = VapGPT(VapConfig(), precision=prec).to("cuda")
    m.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))
    g = torch.Generator(device="cuda").manual_seed(2)
    print(f"precision {prec}; synthetic weights (LSTMx1); times are medians of CUDA-event timings\n")
    print("| case | shape | ms | audio-s/s |")
    print("|---|---|---:|---:|")
    # B=1 full-window latency (20 s and the 25 s window run.py uses)
    for S in (320000, 400000):
        w = torch.randn((1, 2, S), device="cuda", generator=g) * 0.05
        ms = timed(lambda: m.probs(w))
        print(f"| B=1 window latency | (1,2,{S}) | {ms:.2f} | {S / 16000 / ms * 1e3:.0f} |")
    # configs[2]: one 10-minute session, reference window plan (25 s windows, 5 s hop), windows batched
    S = 9_600_000
    sess = torch.randn((1, 2, S), device="cuda", generator=g) * 0.05
    plan = window_plan(S)
    for mb in (116, 32):
        ms = timed(lambda: step_extraction(sess, m, "cuda", max_batch=mb, to_cpu=False), n=3, warm=1)
        print(f"| 10-min session, {len(plan['starts'])} windows of 25 s, micro-batch {mb} | (1,2,{S}) | {ms:.1f} | "
              f"{600.0 / ms * 1e3:.0f} |")
    # streaming step: newest 5 s arrive -> one 25 s window forward (what each later window of run.py costs)
    w = sess[..., :400000].contiguous()
    ms = timed(lambda: m.probs(w))
    print(f"| streaming step (5 s hop, 25 s window, B=1) | (1,2,400000) | {ms:.2f} | {5.0 / ms * 1e3:.0f} (new audio) |")
    # configs[4]: mono input, silent second channel, 10 s chunks, B=256
    mono = torch.randn((256, 1, 160000), device="cuda", generator=g) * 0.05
    w = torch.cat((mono, torch.zeros_like(mono)), dim=1)
    ms = timed(lambda: m.probs(w))
    print(f"| mono + silent channel, 10 s chunks | (256,2,160000) | {ms:.2f} | {256 * 10 / ms * 1e3:.0f} |")
    w = torch.randn((256, 2, 320000), device="cuda", generator=g) * 0.05
    ms = timed(lambda: m.probs(w))
    print(f"| configs[1] for scale | (256,2,320000) | {ms:.2f} | {256 * 20 / ms * 1e3:.0f} |")


if __name__ == "__main__":
    main()
