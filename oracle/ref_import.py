"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Works only in the authoring container (the reference is not shipped to the GPU
box). Used by oracle/make_golden.py and by the CPU tests that are skipped when
/root/reference is absent. Nothing here is on a product path.

The reference cannot be constructed as-is: load_CPC() reads
the CPC checkpoint's `config` (vap/encoder_components.py:370-380), the file is
missing and there is no network. We hand it the config it would have read.
"""
from __future__ import annotations

import os
import sys

import torch

REF_ROOT = os.environ.get("VAP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "vap"))


def import_reference(ar_mode: str = "LSTM", ar_layers: int = 1):
    """Returns the reference `vap.model` module with load_CPC stubbed for
    (ar_mode, ar_layers). torch is put into deterministic mode as a side effect
    of the reference import (vap/model.py:21)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torch.hub

    cfg = {
        "hiddenEncoder": 256,
        "hiddenGar": 256,
        "arMode": ar_mode,
        "nLevelsGRU": ar_layers,
        "normMode": "layerNorm",
        "samplingType": "samespeaker",
        "cpc_mode": None,
    }
    torch.hub.load_state_dict_from_url = lambda url, **kw: {"config": cfg, "weights": {}}
    import vap.encoder_components as ec

    ec.makedirs = lambda *a, **k: None  # :378 mkdir inside the read-only tree
    import vap.model as vm

    return vm


def build_reference(sd, ar_mode="LSTM", ar_layers=1):
    """Reference VapGPT, eval mode, with `sd` loaded strictly (run.py:199-201)."""
    vm = import_reference(ar_mode, ar_layers)
    import contextlib
    import io

    _save = torch.save
    torch.save = lambda *a, **k: None  # encoder_components.py:379 writes the ckpt
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            model = vm.VapGPT(vm.VapConfig(load_pretrained=0))
    finally:
        torch.save = _save
    model.load_state_dict(sd)
    return model.eval()
