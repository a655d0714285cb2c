#!/usr/bin/env python
"""Times the ZeroShot marginals kernel (k_heads.cu zero_shot_kernel through ZeroShot.get_probs) on one B=256 x
T=1000 batch of logits (262 MB read, 4 MB written) and prints achieved GB/s of that algorithmic traffic next to
the same computation as torch indexing ops on the GPU (what the reference's vap/zero_shot.py:159-271 executes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voiceactivityprojection_b200.zero_shot import ZeroShot  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
zs = ZeroShot(bin_times=[0.2, 0.4, 0.6, 0.8], frame_hz=50)
B, T = 256, 1000
g = torch.Generator(device="cuda").manual_seed(0)
logits = torch.randn((B, T, 256), device="cuda", generator=g) * 4
va = (torch.rand((B, T, 2), device="cuda", generator=g) < 0.5).float()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def eager():
    probs = logits.softmax(-1)
    out = []
    for pos, neg in ((zs.subset_silence, zs.subset_silence_hold), (zs.subset_active, zs.subset_active_hold)):
        p = []
        for s in (0, 1):
            joint = torch.cat((pos[s], neg[s])).cuda()
            p.append(probs[..., pos[s].cuda()].sum(-1) / probs[..., joint].sum(-1))
        out.append(torch.stack(p, -1))
    bc = torch.stack([probs[..., zs.bc_prediction[s].cuda()].sum(-1) for s in (0, 1)], -1)
    return out, bc


def timed(fn):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


from voiceactivityprojection_b200 import _lib  # noqa: E402

lib = _lib.load()
p_out, p_bc = (torch.empty((B, T, 2), device="cuda") for _ in range(2))
st = torch.cuda.current_stream().cuda_stream


def direct(n=20):
    """n back-to-back launches through the C-ABI (the 262 MB input is larger than L2, so every launch reads HBM)."""
    for _ in range(n):
        rc = lib.vapb_zero_shot(None, st, logits.data_ptr(), 0, B, T, va.data_ptr(), T, zs._sets, p_out.data_ptr(),
                                p_bc.data_ptr(), None, None)
        assert rc == 0


direct(3)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
direct(20)
b.record()
torch.cuda.synchronize()
t = a.elapsed_time(b) / 20
gb = (logits.numel() * 4 + va.numel() * 4 + 2 * B * T * 2 * 4) / 1e9
print(f"zero_shot_kernel: {t * 1e3:.1f} us per launch (20 back to back), {gb / t * 1e3:.0f} GB/s algorithmic "
      f"({gb:.3f} GB)", flush=True)
tw = timed(lambda: zs.get_probs(logits, va))
print(f"ZeroShot.get_probs (wrapper + kernel, L2 flushed): {tw * 1e3:.1f} us", flush=True)
te = timed(eager)
print(f"torch eager marginals (softmax + gathers, no dialog-state switch): {te * 1e3:.1f} us ({te / t:.1f}x)", flush=True)
