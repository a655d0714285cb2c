"""Writes tests/golden/objective.npz with the UNMODIFIED reference's ObjectiveVAP helpers (vap/objective.py:
ProjectionWindow :14-76, Codebook :79-146, get_da_labels :214-218, loss_vad :245-247) on seeded inputs.
TEST INFRASTRUCTURE. Run in the authoring container: python oracle/make_golden_objective.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from vap.objective import ObjectiveVAP  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
r = ObjectiveVAP()
g = torch.Generator().manual_seed(0)
flips = (torch.rand((3, 260, 2), generator=g) < 0.05).long()
va = (flips.cumsum(1) % 2).float()
idx, ds = r.get_da_labels(va)
wins = r.projection_window_extractor(va)
soft = torch.rand((4, 9, 2, 4), generator=g)
soft[0, 0] = 0.5  # exact ties go to the lower index
some = torch.randint(0, 256, (5, 7), generator=g)
vo, v = torch.randn(2, 50, 2, generator=g), torch.randint(0, 2, (2, 60, 2), generator=g).float()
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "objective.npz"),
    va=va.numpy().astype(np.uint8), labels=idx.numpy(), dialog_states=ds.numpy(), windows=wins.numpy().astype(np.uint8),
    soft=soft.numpy(), soft_idx=r.codebook.encode(soft).numpy(), some_idx=some.numpy(),
    some_windows=r.codebook.decode(some).numpy(), code_vectors=r.codebook.emb.weight.numpy(),
    vad_logits=vo.numpy(), vad_target=v.numpy(), loss_vad=r.loss_vad(vo, v).numpy(),
    repr_pw=repr(r.projection_window_extractor))
print("labels", tuple(idx.shape), "dialog states", torch.bincount(ds.flatten()).tolist())
