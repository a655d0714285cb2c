"""Diagnostics for the fused conv0 -> conv1 kernel (csrc/k_conv01.cu): where errors sit (row position inside the
128-row tile, channel block), for both store modes (run with VAPB_CONV01_STORE=0 / 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_kernels import _conv01_case  # noqa: E402

for B, S, fp16 in [(1, 2000, 0), (1, 40000, 1), (3, 320000, 1)]:
    try:
        out, ref, pad, L1 = _conv01_case(B, S, fp16, seed=5)
    except AssertionError as e:
        print(f"B={B} S={S} fp16={fp16}: launch failed: {e}")
        continue
    got = out[:, pad:pad + L1].float()
    nan = torch.isnan(got)
    err = torch.where(nan, torch.full_like(got, 99.0), (got - ref).abs())
    print(f"B={B} S={S} fp16={fp16} store_mode={os.environ.get('VAPB_CONV01_STORE', 'auto')}: L1={L1} max {err.max().item():.4g} "
          f"mean {err.mean().item():.4g} nan {int(nan.sum())} ref_absmax {ref.abs().max().item():.3g}")
    if err.max().item() > 0.05:
        rows = torch.arange(L1, device=err.device)
        per_row = err.amax(dim=(0, 2))
        bad = (per_row > 0.05)
        print("  bad rows:", int(bad.sum()), "of", L1, "first:", rows[bad][:24].tolist())
        pos = rows % 128
        hist = torch.zeros(128, device=err.device).index_add_(0, pos[bad], torch.ones(int(bad.sum()), device=err.device))
        print("  bad rows by (t mod 128):", [int(v) for v in hist.tolist()])
        per_cb = err.reshape(err.shape[0], L1, 4, 64).amax(dim=(0, 1, 3))
        print("  max err per output channel block:", [round(v, 4) for v in per_cb.tolist()])
        per_seq = err.amax(dim=(1, 2))
        print("  max err per sequence:", [round(v, 4) for v in per_seq.tolist()])

# timeline of the first tiles (SM clocks of cluster 0 / CTA 0): producer warp 12 and the MMA thread
if os.environ.get("CONV01_TIMELINE", "1") != "0":
    dbg = torch.zeros(4 * 4 * 16 + 32, dtype=torch.int64, device="cuda")
    _conv01_case(64, 320000, 1, seed=5, dbg=dbg)
    tm = dbg[256:264].cpu().tolist()
    d = dbg[:256].reshape(4, 4, 16).cpu()
    et = dbg[264:268].cpu().tolist()
    if et[3]:
        print(f"epilogue warp 4 of CTA 0, cycles per tile: accumulator ready -> statistics {et[0] / et[3]:.0f}, -> released "
              f"{et[1] / et[3]:.0f}, -> tile done {et[2] / et[3]:.0f}  ({et[3]} tiles)")
    t00 = int(d[d > 0].min())
    pn = ["first block (ld + st) before X", "next tap's X (+ window barrier at tile end)", "wait d0_full", "load_pack (TMEM -> regs)",
          "wait stage empty", "STS", "fence.proxy.async", "syncwarp + arrive"]
    tot = sum(tm)
    print("producer warp 12 of CTA 0, cycles per phase over the whole kernel:")
    for n, v in zip(pn, tm):
        print(f"  {n:40s} {v:12d}  {100.0 * v / max(tot, 1):5.1f} %")
    names = ["X.begin", "xa.arrive", "d0_full", "ld3", "st0", "st1", "d0_free", "st2", "st3", "Z.enter", "Z.xa_full",
             "Z.issued", "C0", "C1", "C2", "C3"]
    print("event clocks relative to the first stamp (tile iteration, tap):")
    for it in range(4):
        for j in range(4):
            row = d[it, j]
            print(f"  it{it} j{j}: " + "  ".join(f"{n}={int(v) - t00 if v > 0 else -1}" for n, v in zip(names, row.tolist())))
