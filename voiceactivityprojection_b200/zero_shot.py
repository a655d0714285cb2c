"""Host-side mirror of the inference part of the reference's vap/zero_shot.py: `ZeroShot(ObjectiveVAP)` with the
class-index subsets it builds (`subset_silence`, `subset_silence_hold`, `subset_active`, `subset_active_hold`,
`bc_prediction`; :101-158) and `get_probs(logits, va)`, `probs_next_speaker`, `probs_on_silence`,
`probs_on_active`, `probs_backchannel` (:159-271).

All five run as one CUDA kernel (csrc/k_heads.cu zero_shot_kernel through vapb_zero_shot): softmax, the subset sums
as 256-bit set memberships per lane, and the dialog-state switch of vap/events.py:70-78. The event-metric side
(`extract_prediction_and_targets`, :274-) is evaluation code and out of scope. No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .objective import ObjectiveVAP


def _mono(bits: Tuple[int, ...]) -> int:
    """One speaker's 4-bin window as its nibble of the class index: bin b -> bit b (objective.py:93-110)."""
    return sum(v << b for b, v in enumerate(bits))


def _class(spk0: int, spk1: int) -> int:
    return spk0 | (spk1 << 4)


class ZeroShot(ObjectiveVAP):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.n_bins != 4:
            raise NotImplementedError("Not implemented for bin-size != 4")
        # one speaker's windows, as nibbles
        late = [_mono((a, b, 1, 1)) for a in (0, 1) for b in (0, 1)]     # active in the last two bins (:32-60)
        ending = [_mono(v) for v in ((0, 0, 0, 0), (1, 0, 0, 0), (1, 1, 0, 0))]  # end of segment, max 2 (:9-19)
        early = [_mono((a, b, c, 0)) for a in (0, 1) for b in (0, 1) for c in (0, 1)][1:]   # some of bins 0-2, not 3
        through = [_mono((a, b, c, 1)) for a in (0, 1) for b in (0, 1) for c in (0, 1)]     # active in bin 3

        def both_ways(first: List[int], second: List[int], sort: bool = True) -> Tensor:
            """row 0: speaker 0 takes `first`, speaker 1 `second`; row 1: the mirror image (:63-76)."""
            r0 = [_class(a, b) for a in first for b in second]
            r1 = [_class(b, a) for a in first for b in second]
            if sort:
                r0, r1 = sorted(r0), sorted(r1)
            return torch.tensor([r0, r1], dtype=torch.long)

        self.subset_silence = both_ways(late, [0])               # :101-123
        self.subset_silence_hold = self.subset_silence.flip(0)
        self.subset_active = both_ways(late, ending)             # :125-132
        self.subset_active_hold = both_ways([0], late)           # :134-140
        self.bc_prediction = both_ways(early, through, sort=False)  # :142-158
        self._sets = self._pack_sets()

    def _pack_sets(self):
        """Ten 256-bit class sets in vapb_zero_shot's order. A class may appear once per list and not in both the
        pos and the neg list of a marginal: `probs[..., cat(pos, neg)].sum()` would count it twice (:159-165)."""
        groups = [self.subset_silence, self.subset_silence_hold, self.subset_active, self.subset_active_hold,
                  self.bc_prediction]
        for pos, neg in ((groups[0], groups[1]), (groups[2], groups[3])):
            for s in (0, 1):
                joint = pos[s].tolist() + neg[s].tolist()
                assert len(set(joint)) == len(joint), "zero-shot subsets overlap"
        words = (C.c_uint32 * 80)()
        for g, idx in enumerate(groups):
            for s in (0, 1):
                assert len(set(idx[s].tolist())) == idx.shape[1]
                for c in idx[s].tolist():
                    words[(2 * g + s) * 8 + c // 32] |= 1 << (c % 32)
        return words

    # ------------------------------------------------------------------ the kernel call
    def _run(self, x: Tensor, is_probs: bool, va: Optional[Tensor], want: Tuple[str, ...]) -> Dict[str, Tensor]:
        if x.device.type != "cuda":
            raise RuntimeError("ZeroShot runs on CUDA tensors (no CPU fallback)")
        assert x.shape[-1] == self.n_classes, (
            f"Logits have wrong shape. {x.shape} != (..., {self.n_classes}) that is (B, N_FRAMES, N_CLASSES)")
        x = x.to(torch.float32).contiguous()
        if x.ndim == 2:
            x = x[None]
        B, T = x.shape[0], x.shape[1]
        va_ptr, va_T = None, 0
        if "p" in want:
            assert va is not None and va.ndim == 3 and va.shape[0] == B and va.shape[-1] == 2 and va.shape[1] >= T
            va = va.to(device=x.device, dtype=torch.float32).contiguous()
            va_ptr, va_T = va.data_ptr(), va.shape[1]
        out = {k: torch.empty((B, T, 2), dtype=torch.float32, device=x.device) for k in want}
        if x.numel() == 0:
            return out
        ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        lib = _lib.load()
        h = self._owner._ensure_handle() if self._owner is not None else None
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib, h, lib.vapb_zero_shot(h, st, x.data_ptr(), int(is_probs), B, T, va_ptr, va_T, self._sets,
                                                  ptr("p"), ptr("p_bc"), ptr("p_sil"), ptr("p_act")))
        return out

    # ------------------------------------------------------------------ reference API
    def probs_on_silence(self, probs: Tensor) -> Tensor:
        return self._run(probs, True, None, ("p_sil",))["p_sil"]

    def probs_on_active(self, probs: Tensor) -> Tensor:
        return self._run(probs, True, None, ("p_act",))["p_act"]

    def probs_backchannel(self, probs: Tensor) -> Tensor:
        return self._run(probs, True, None, ("p_bc",))["p_bc"]

    def probs_next_speaker(self, probs: Tensor, va: Tensor) -> Tensor:
        return self._run(probs, True, va, ("p",))["p"]

    def get_probs(self, logits: Tensor, va: Tensor) -> Dict[str, Tensor]:  # noqa: D102  (:264-271)
        return self._run(logits, False, va, ("p", "p_bc"))
