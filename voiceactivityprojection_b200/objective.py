"""Host-side mirror of the inference part of the reference's vap/objective.py:
`ObjectiveVAP` with `n_classes`, `n_bins`, `horizon`, `horizon_time`,
`probs_next_speaker_aggregate` (:184-204), `get_labels` (:209-212), `loss_vap`
(:220-243) and `get_probs` (:249-281), plus `Codebook` / `ProjectionWindow`
semantics (:14-146).

Inside `VapGPT.probs` all of this is fused into the CUDA heads kernels
(csrc/k_heads.cu); the methods here serve callers that hold logits / probs
tensors themselves (e.g. vap/phrases/dataset.py:214-215). `get_probs` on CUDA
logits runs the same kernel through vapb_probs_from_logits.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
from torch import Tensor


def bin_times_to_frames(bin_times: List[float], frame_hz: int) -> List[int]:
    return (torch.tensor(bin_times) * frame_hz).long().tolist()


def code_vectors(total_bins: int = 8) -> Tensor:
    """(2^n, n) table: bit i of the class index, LSB first."""
    idx = torch.arange(2 ** total_bins)
    return torch.stack([(idx >> i) & 1 for i in range(total_bins)], dim=-1).float()


class ObjectiveVAP:
    def __init__(self, bin_times: List[float] = [0.2, 0.4, 0.6, 0.8], frame_hz: int = 50,
                 threshold_ratio: float = 0.5):
        self.frame_hz = frame_hz
        self.bin_times = bin_times
        self.bin_frames: List[int] = bin_times_to_frames(bin_times, frame_hz)
        self.horizon = sum(self.bin_frames)
        self.horizon_time = sum(bin_times)
        self.threshold_ratio = threshold_ratio
        self.n_bins = len(self.bin_frames)
        self.total_bins = 2 * self.n_bins
        self.n_classes = 2 ** self.total_bins
        self._codes = code_vectors(self.total_bins)
        self._owner = None  # the VapGPT whose library handle get_probs uses

    def __repr__(self):
        return (f"ObjectiveVAP(bin_times={self.bin_times}, bin_frames={self.bin_frames}, "
                f"frame_hz={self.frame_hz}, thresh={self.threshold_ratio})")

    # ---- codebook
    def decode(self, idx: Tensor) -> Tensor:
        return self._codes.to(idx.device)[idx].unflatten(-1, (2, self.n_bins))

    def probs_next_speaker_aggregate(self, probs: Tensor, from_bin: int = 0, to_bin: int = 3,
                                     scale_with_bins: bool = False) -> Tensor:
        assert probs.ndim == 3, f"Expected probs of shape (B, n_frames, n_classes) but got {probs.shape}"
        states = self._codes.to(probs.device, probs.dtype).view(self.n_classes, 2, self.n_bins)
        if scale_with_bins:
            states = states * torch.tensor(self.bin_frames, device=probs.device, dtype=probs.dtype)
        abp = states[:, :, from_bin: to_bin + 1].sum(-1)
        p_all = torch.einsum("bid,dc->bic", probs, abp)
        return p_all / (p_all.sum(-1, keepdim=True) + 1e-5)

    # ---- labels / loss
    def get_labels(self, va: Tensor) -> Tensor:
        """va (B, T, 2) -> class index of the next-100-frame projection window, (B, T-100)."""
        win = va[..., 1:, :].unfold(dimension=-2, size=self.horizon, step=1)
        start, bits = 0, []
        for b in self.bin_frames:
            bits.append(win[..., start: start + b].sum(dim=-1) / b >= self.threshold_ratio)
            start += b
        bits = torch.stack(bits, dim=-1).flatten(-2).long()  # (B, N, 8), order (c bin)
        weights = (2 ** torch.arange(self.total_bins, device=va.device)).long()
        return (bits * weights).sum(-1)

    def loss_vap(self, logits: Tensor, labels: Tensor, reduction: str = "mean") -> Tensor:
        assert logits.ndim == 3 and labels.ndim == 2
        n = labels.shape[1]
        lg = logits[:, :n]
        loss = torch.nn.functional.cross_entropy(lg.reshape(-1, lg.shape[-1]), labels.reshape(-1),
                                                 reduction=reduction)
        return loss.view(-1, n) if reduction == "none" else loss

    # ---- probabilities from logits
    def get_probs(self, logits: Tensor) -> Dict[str, Tensor]:
        assert logits.shape[-1] == self.n_classes, (
            f"Logits have wrong shape. {logits.shape} != (..., {self.n_classes}) that is (B, N_FRAMES, N_CLASSES)"
        )
        if logits.device.type != "cuda" or self._owner is None:
            raise RuntimeError("get_probs runs on CUDA logits of a VapGPT that lives on the GPU (no CPU fallback)")
        from . import _lib

        lib, h = _lib.load(), self._owner._ensure_handle()
        lg = logits.to(torch.float32).contiguous()
        rows = lg.numel() // self.n_classes
        lead = lg.shape[:-1]
        f = dict(dtype=torch.float32, device=lg.device)
        probs = torch.empty(lg.shape, **f)
        p_now, p_fut, p_tot = (torch.empty((*lead, 2), **f) for _ in range(3))
        st = torch.cuda.current_stream(lg.device).cuda_stream
        _lib.check(lib, h, lib.vapb_probs_from_logits(h, st, lg.data_ptr(), rows, 0, 1, 2, 3, probs.data_ptr(),
                                                      p_now.data_ptr(), p_fut.data_ptr(), None, None))
        _lib.check(lib, h, lib.vapb_probs_from_logits(h, st, lg.data_ptr(), rows, 0, 3, 0, 3, None,
                                                      p_tot.data_ptr(), None, None, None))
        return {"probs": probs, "p_now": p_now, "p_future": p_fut, "p_tot": p_tot}
