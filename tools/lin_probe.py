#!/usr/bin/env python
"""Times the plain bf16-out linear GEMMs of a B=256 step in isolation (k_gemm_lin.cu through vapb_debug_gemm_lin):
QKV (N=768), cross K|V (N=512), cross Q (N=256), all K=256 over M = 512 000 rows; prints achieved GB/s of the
algorithmic traffic. Run under ncu for the stall picture:
    ncu --set full -k regex:gemm_lin -c 3 -o gpurun_out/lin python tools/lin_probe.py 1"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voiceactivityprojection_b200 import _lib  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
lib = _lib.load()
M, K = 512000, 256
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
err = C.create_string_buffer(512)
st = torch.cuda.current_stream().cuda_stream
for N in (768, 512, 256):
    W = (torch.randn((N, K), device="cuda", generator=g) * 0.05).bfloat16()
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)

    def run():
        rc = lib.vapb_debug_gemm_lin(st, A.data_ptr(), 0, K, W.data_ptr(), 1, M, N, K, None, 0, None, None, 0, None, 0,
                                     None, 1, out.data_ptr(), 0, None, None, None, err, 512)
        assert rc == 0, err.value.decode()

    run()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    gb = (M * K * 2 + M * N * 2) / 1e9
    print(f"N={N}: {t * 1e3:.1f} us, {gb / t * 1e3:.0f} GB/s algorithmic ({gb:.2f} GB)", flush=True)

# the residual + LayerNorm2 form (attention out-projection): y (bf16) W^T + x (fp32 blocked) -> x' (fp32 blocked), LN(x') (bf16)
N = 256
W = (torch.randn((N, K), device="cuda", generator=g) * 0.05).bfloat16()
rb = torch.randn((M // 128, 64, 128, 4), device="cuda", generator=g)
xo = torch.empty_like(rb)
zn = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
g2 = torch.ones(N, device="cuda")
b2 = torch.zeros(N, device="cuda")


def run_res():
    rc = lib.vapb_debug_gemm_lin(st, A.data_ptr(), 0, K, W.data_ptr(), 1, M, N, K, None, 0, None, None, 0, rb.data_ptr(), 0,
                                 xo.data_ptr(), 1, None, 2, g2.data_ptr(), b2.data_ptr(), zn.data_ptr(), err, 512)
    assert rc == 0, err.value.decode()


run_res()
ts = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run_res()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
t = ts[len(ts) // 2]
gb = (M * K * 2 + 2 * M * N * 4 + M * N * 2) / 1e9
print(f"N=256 + residual + LN2: {t * 1e3:.1f} us, {gb / t * 1e3:.0f} GB/s algorithmic ({gb:.2f} GB)", flush=True)
