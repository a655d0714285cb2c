"""Prints max-abs errors of the CUDA path against the golden fixtures (and the
oracle for stages) — a debugging aid for the GPU box; the gating checks live in
tests/test_gpu_parity.py.

    python tools/parity_report.py [fp32|bf16] [case ...]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CASE_NAMES, golden_inputs, load_golden  # noqa: E402
from oracle import vap_oracle as O  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402


def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    if a.shape != b.shape:
        return f"SHAPE {tuple(a.shape)} vs {tuple(b.shape)}"
    d = (a - b).abs()
    d = d[~torch.isnan(d)]
    return f"{d.max().item():.3e} (rel {(d.max() / (b.abs().max() + 1e-30)).item():.1e})"


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    cases = sys.argv[2:] or CASE_NAMES
    for name in cases:
        recipe, g = load_golden(name)
        sd, wav = golden_inputs(recipe, g)
        model = VapGPT(VapConfig(), precision=prec).to("cuda")
        model.load_state_dict(sd)
        print(f"== {name} [{prec}] {model.describe()}")
        x = wav.cuda()
        B = x.shape[0]
        if recipe.get("stages"):
            st = {}
            O.forward(sd, wav, stages=st)
            ref_stage = {
                "conv": torch.cat([st["conv_1"], O.cpc_conv_stack(sd, wav[:, 1:]).transpose(1, 2)]),
                "enc": torch.cat([st["enc_1"], st["enc_2"]]),
                "ch": torch.cat([st["ch_1"], st["ch_2"]]),
                "ar0": torch.cat([st["ar0_x1"], st["ar0_x2"]]),
                "ar1": torch.cat([st["ar1_x1"], st["ar1_x2"]]),
                "ar2": torch.cat([st["ar2_x1"], st["ar2_x2"]]),
                "comb": st["comb"],
            }
            ref_stage["ar"] = torch.cat([st["ar_1"], O.ar_net(sd, ref_stage["conv"][B:])])
            for k in ["conv", "ar", "enc", "ch", "ar0", "ar1", "ar2", "comb"]:
                print(f"   stage {k:5s} {err(model.stage(k, x), ref_stage[k])}")
        t0 = time.time()
        fwd = model(x)
        out = model.probs(x)
        torch.cuda.synchronize()
        print(f"   forward+probs wall {time.time() - t0:.3f}s")
        print(f"   logits   {err(fwd['logits'], g['logits'])}")
        print(f"   vad_lg   {err(fwd['vad'], g['vad_logits'])}")
        for k in ["probs", "vad", "p_now", "p_future", "H", "loss"]:
            if k in g:
                print(f"   {k:8s} {err(out[k], g[k])}")
        am = (fwd["logits"].argmax(-1).cpu() != g["logits"].argmax(-1)).sum().item()
        vd = ((out["vad"].cpu() >= 0.5) != (g["vad"] >= 0.5)).sum().item()
        print(f"   argmax mismatches {am} / {g['logits'].shape[0] * g['logits'].shape[1]}, vad-threshold mismatches {vd}")


if __name__ == "__main__":
    main()
