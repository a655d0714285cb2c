#!/usr/bin/env python
"""Error of the fp32-class tensor-core attention (csrc/k_attn_x3.cu) against a float64 reference, next to the error
of a plain fp32 torch evaluation of the same formula. Needs a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from voiceactivityprojection_b200 import _lib  # noqa: E402
from test_gpu_kernels import _attn_ref  # noqa: E402

lib = _lib.load()
slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625], device="cuda")
for nseq, T, cross, scale in [(2, 117, 0, 1.0), (4, 128, 1, 1.0), (2, 1000, 0, 1.0), (6, 500, 1, 3.0), (40, 300, 0, 6.0),
                              (2, 1250, 1, 1.0), (6, 500, 0, 3.0)]:
    g = torch.Generator(device="cuda").manual_seed(nseq * 1000 + T + 7)
    if not cross:
        buf = (torch.randn((nseq, T, 768), device="cuda", generator=g) * scale).contiguous()
        q, k, v = buf[..., :256], buf[..., 256:512], buf[..., 512:]
        qb, kvb, qc, kvc, ko, vo = buf, buf, 768, 768, 256, 512
        planes = torch.empty(buf.numel() * 4, device="cuda", dtype=torch.uint8)
    else:
        qb = (torch.randn((nseq, T, 256), device="cuda", generator=g) * scale).contiguous()
        kvb = (torch.randn((nseq, T, 512), device="cuda", generator=g) * scale).contiguous()
        q, k, v = qb, kvb[..., :256], kvb[..., 256:]
        qc, kvc, ko, vo = 256, 512, 0, 256
        planes = torch.empty((qb.numel() + kvb.numel()) * 4, device="cuda", dtype=torch.uint8)
    out = torch.full((nseq, T, 256), float("nan"), device="cuda")
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_attn_x3(st, qb.data_ptr(), qc, kvb.data_ptr(), kvc, ko, vo, planes.data_ptr(), out.data_ptr(),
                                nseq, T, slopes.data_ptr(), cross, err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, slopes, cross, dtype=torch.float64)
    r32 = _attn_ref(q, k, v, slopes, cross, dtype=torch.float32).double()
    d = (out.double() - ref).abs()
    d32 = (r32 - ref).abs()
    # the 16-bit kernel on the same (bf16-rounded) inputs, against float64 on those inputs
    qr, kr, vr = q.bfloat16(), k.bfloat16(), v.bfloat16()
    if not cross:
        b16 = buf.bfloat16().contiguous()
        q16, k16, v16, qs, ks = b16[..., :256], b16[..., 256:512], b16[..., 512:], 768, 768
    else:
        q16 = qb.bfloat16().contiguous()
        kv16 = kvb.bfloat16().contiguous()
        k16, v16, qs, ks = kv16[..., :256], kv16[..., 256:], 256, 512
    o16 = torch.full((nseq, T, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.vapb_debug_attn_tc(st, q16.data_ptr(), qs, k16.data_ptr(), v16.data_ptr(), ks, o16.data_ptr(), nseq, T, 4,
                                slopes.data_ptr(), cross, err, 512, None)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    ref16 = _attn_ref(qr, kr, vr, slopes, cross, dtype=torch.float64)
    d16 = (o16.double() - ref16).abs()
    print(f"   16-bit kernel: max {d16.max().item():.3e} mean {d16.mean().item():.3e} (bf16 ulp of |ref| max: {ref16.abs().max().item() * 2**-8:.3e})")
    print(f"nseq {nseq} T {T} cross {cross} scale {scale}: x3 max {d.max().item():.3e} mean {d.mean().item():.3e} | "
          f"torch fp32 max {d32.max().item():.3e} mean {d32.mean().item():.3e} | |ref| max {ref.abs().max().item():.2f} "
          f"finite {bool(torch.isfinite(out).all())}")

# where are the large errors? (last case: nseq 6, T 500, self, scale 3)
e = (out.double() - ref).abs()            # (nseq, T, 256)
bad = (e > 1e-3).nonzero()
print("bad elements", bad.shape[0], "of", e.numel())
if bad.shape[0]:
    seqs = torch.unique(bad[:, 0]).tolist()
    ts = torch.unique(bad[:, 1]).tolist()
    heads = torch.unique(bad[:, 2] // 64).tolist()
    print("seqs", seqs[:10], "heads", heads, "rows", ts[:40], "n rows", len(ts))
    # emulate the split arithmetic in float64
    def split(x):
        hi = x.half().float()
        lo = (x - hi).half().float()
        return hi.double(), lo.double()
    sq, tq = int(bad[0, 0]), int(bad[0, 1])
    hd = int(bad[0, 2]) // 64
    qv = q[sq, tq, hd * 64:(hd + 1) * 64]
    kk = k[sq, :tq + 1, hd * 64:(hd + 1) * 64]
    qh, ql = split(qv); kh, kl = split(kk)
    s_true = (kk.double() @ qv.double())
    s_emu = kh @ qh + kl @ qh + kh @ ql
    print("row", tq, "head", hd, "max |s|", s_true.abs().max().item(), "max split error", (s_true - s_emu).abs().max().item())
    w = torch.softmax(s_true / 16 + 1 + slopes[hd].double() * torch.arange(tq + 1, device="cuda").double(), 0)
    top = torch.topk(w, 3)
    print("top weights", top.values.tolist(), "at keys", top.indices.tolist())
    print("out  ", out[sq, tq, hd * 64:hd * 64 + 6].tolist())
    print("ref  ", ref[sq, tq, hd * 64:hd * 64 + 6].tolist())
