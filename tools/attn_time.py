#!/usr/bin/env python
"""Time of the attention kernel alone (CUDA events, self and cross, 512 sequences x 1000 frames) for the library
named by VAPB_LIB: same-box comparison of kernel variants. Needs a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voiceactivityprojection_b200 import _lib  # noqa: E402

lib = _lib.load()
nseq, T = 512, 1000
g = torch.Generator(device="cuda").manual_seed(0)
buf = (torch.randn((nseq, T, 768), device="cuda", generator=g) * float(os.environ.get("SCALE", "1"))).bfloat16()
q, k, v = buf[..., :256], buf[..., 256:512], buf[..., 512:]
out = torch.empty((nseq, T, 256), device="cuda", dtype=torch.bfloat16)
slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625], device="cuda")
err = C.create_string_buffer(512)
st = torch.cuda.current_stream().cuda_stream
res = []
for cross in (0, 1):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        rc = lib.vapb_debug_attn_tc(st, q.data_ptr(), 768, k.data_ptr(), v.data_ptr(), 768, out.data_ptr(), nseq, T, 4,
                                    slopes.data_ptr(), cross, err, 512, 0)
        assert rc == 0, err.value.decode()
    ev[0].record()
    for _ in range(20):
        lib.vapb_debug_attn_tc(st, q.data_ptr(), 768, k.data_ptr(), v.data_ptr(), 768, out.data_ptr(), nseq, T, 4,
                               slopes.data_ptr(), cross, err, 512, 0)
    ev[1].record()
    torch.cuda.synchronize()
    res.append(ev[0].elapsed_time(ev[1]) * 50)
print(os.environ.get("VAPB_LIB", "default"), "self us %.1f cross us %.1f" % tuple(res), "checksum %.6f" % float(out.float().abs().mean()))
