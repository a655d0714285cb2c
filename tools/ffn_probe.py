#!/usr/bin/env python
"""Per-tile timeline of the fused FFN kernel's epilogue (SM clocks, CTA 0, first epilogue thread)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voiceactivityprojection_b200 import _lib  # noqa: E402

lib = _lib.load()
M = 512000
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn((M, 256), device="cuda", generator=g).bfloat16()
w1 = (torch.randn((768, 256), device="cuda", generator=g) * 0.06).bfloat16()
w2 = (torch.randn((256, 768), device="cuda", generator=g) * 0.04).bfloat16()
rb = torch.randn((M // 128, 64, 128, 4), device="cuda", generator=g)
xo = torch.empty_like(rb)
xs = torch.empty((M, 256), device="cuda", dtype=torch.bfloat16)
zn = torch.empty((M, 256), device="cuda", dtype=torch.bfloat16)
g2 = torch.ones(256, device="cuda")
b2 = torch.zeros(256, device="cuda")
dbg = torch.zeros((32, 16), device="cuda", dtype=torch.int64)
err = C.create_string_buffer(512)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    rc = lib.vapb_debug_ffn_fused(st, z.data_ptr(), w1.data_ptr(), w2.data_ptr(), rb.data_ptr(), xo.data_ptr(), xs.data_ptr(),
                                  zn.data_ptr(), g2.data_ptr(), b2.data_ptr(), M, err, 512, dbg.data_ptr())
    assert rc == 0, err.value.decode()
torch.cuda.synchronize()
d = dbg.cpu()
print("tile | wait_acch0 | chunk0..5 (each: until h_full arrive) | wait_acco | final pass1 | LN pass + arrive | tile total")
for i in range(1, 27):
    r = [int(x) for x in d[i]]
    ch = [r[2] - r[1]] + [r[2 + j] - r[1 + j] for j in range(1, 6)]
    print(f"{i:3d} | {r[1]-r[0]:6d} | " + " ".join(f"{c:5d}" for c in ch) + f" | {r[8]-r[7]:6d} | {r[9]-r[8]:6d} | {r[10]-r[9]:6d} | {r[10]-r[0]:6d}")

print("MMA thread per tile: waiting for weights | waiting for the epilogue (h_full) | tile period")
for i in range(2, 14):
    r = [int(x) for x in d[i]]
    print(f"{i:3d} | {r[11]:6d} | {r[12]:6d} | {r[13]-int(d[i-1][13]):6d}")
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib.vapb_debug_ffn_fused(st, z.data_ptr(), w1.data_ptr(), w2.data_ptr(), rb.data_ptr(), xo.data_ptr(), xs.data_ptr(),
                             zn.data_ptr(), g2.data_ptr(), b2.data_ptr(), M, err, 512, None)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print(f"kernel time (median of 10): {ts[5] * 1e3:.1f} us")
