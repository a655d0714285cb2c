#!/usr/bin/env python
"""Timeline of the attention kernel's per-tile dependency chain (SM clocks of CTA 0, slot 0):
softmax thread: s_full wait start / end, after the max pass, after the exchange barrier, after p_full arrive;
MMA thread: p_full wait start / end, after issuing PV. Needs a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voiceactivityprojection_b200 import _lib  # noqa: E402

lib = _lib.load()
nseq, T = 512, 1000
g = torch.Generator(device="cuda").manual_seed(0)
buf = torch.randn((nseq, T, 768), device="cuda", generator=g).bfloat16()
q, k, v = buf[..., :256], buf[..., 256:512], buf[..., 512:]
out = torch.empty((nseq, T, 256), device="cuda", dtype=torch.bfloat16)
slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625], device="cuda")
dbg = torch.zeros((128, 8), device="cuda", dtype=torch.int64)
err = C.create_string_buffer(512)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    rc = lib.vapb_debug_attn_tc(st, q.data_ptr(), 768, k.data_ptr(), v.data_ptr(), 768, out.data_ptr(), nseq, T, 4,
                                slopes.data_ptr(), 0, err, 512, dbg.data_ptr())
    assert rc == 0, err.value.decode()
torch.cuda.synchronize()
d = dbg.cpu()
t0 = int(d[0, 0])
print("tile | sm: wait_s_full  single pass  exchange  (exact route +) arrive | total | mma: wait_p_full  issue | softmax start->next start")
for i in range(1, 40):
    r = [int(x) for x in d[i]]
    nxt = int(d[i + 1, 0])
    print(f"{i:3d} | {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:6d} {r[4]-r[3]:6d} | {r[4]-r[0]:6d} | {r[6]-r[5]:6d} {r[7]-r[6]:6d} | {nxt-r[0]:6d}"
          f" | abs sm {r[0]-t0:7d} mma {r[5]-t0:7d}")
print("mma thread: kv wait  QK issue  PV issue | softmax: chunk0 compute (s_full -> before pv wait)  pv wait")
for i in range(1, 40):
    a = [int(x) for x in d[64 + i]]
    r = [int(x) for x in d[i]]
    print(f"{i:3d} | {a[1]-a[0]:6d} {a[2]-a[1]:6d} {a[3]-a[2]:6d} | {a[4]-r[1]:6d} {a[5]-a[4]:6d}")
# kernel time without the probe (CUDA events on the launching stream)
null = 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    lib.vapb_debug_attn_tc(st, q.data_ptr(), 768, k.data_ptr(), v.data_ptr(), 768, out.data_ptr(), nseq, T, 4,
                           slopes.data_ptr(), 0, err, 512, null)
ev[0].record()
for _ in range(10):
    lib.vapb_debug_attn_tc(st, q.data_ptr(), 768, k.data_ptr(), v.data_ptr(), 768, out.data_ptr(), nseq, T, 4,
                           slopes.data_ptr(), 0, err, 512, null)
ev[1].record()
torch.cuda.synchronize()
print("kernel us (self, nseq=512, T=1000):", ev[0].elapsed_time(ev[1]) * 100)
