"""Host-side mirror of the inference part of the reference's vap/objective.py:
`ObjectiveVAP` with `n_classes`, `n_bins`, `horizon`, `horizon_time`,
`probs_next_speaker_aggregate` (:184-204), `get_labels` (:209-212), `get_da_labels`
(:214-218), `loss_vap` (:220-243), `loss_vad` (:245-247) and `get_probs` (:249-281),
plus the `ProjectionWindow` (:14-76) and `Codebook` (:79-146) classes behind
`objective.projection_window_extractor` / `objective.codebook`.

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


class ProjectionWindow:
    """vap/objective.py:14-76: the next `horizon` frames of every frame as activity bins per speaker."""

    def __init__(self, bin_times: List[float] = [0.2, 0.4, 0.6, 0.8], frame_hz: int = 50,
                 threshold_ratio: float = 0.5):
        self.bin_times, self.frame_hz, self.threshold_ratio = bin_times, frame_hz, threshold_ratio
        self.bin_frames: List[int] = bin_times_to_frames(bin_times, frame_hz)
        self.n_bins = len(self.bin_frames)
        self.total_bins = 2 * self.n_bins
        self.horizon = sum(self.bin_frames)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(\n  bin_times: {self.bin_times}\n  bin_frames: {self.bin_frames}\n"
                f"  frame_hz: {self.frame_hz}\n  thresh: {self.threshold_ratio}\n)\n")

    def projection(self, va: Tensor) -> Tensor:
        """(B, N, C) -> (B, N - horizon, C, horizon): frames t+1 .. t+horizon for every t that has them all."""
        return va[..., 1:, :].unfold(dimension=-2, size=self.horizon, step=1)

    def projection_bins(self, projection_window: Tensor) -> Tensor:
        """(..., C, horizon) -> (..., C, n_bins) of 0/1: a bin is active when its mean activity reaches the threshold."""
        edges = [0]
        for b in self.bin_frames:
            edges.append(edges[-1] + b)
        bins = [(projection_window[..., lo:hi].sum(dim=-1) / (hi - lo) >= self.threshold_ratio).float()
                for lo, hi in zip(edges[:-1], edges[1:])]
        return torch.stack(bins, dim=-1)

    def __call__(self, va: Tensor) -> Tensor:
        return self.projection_bins(self.projection(va))


class _Embedding:
    """What callers read of the reference's `nn.Embedding` code table: `.weight` and indexing by call."""

    def __init__(self, weight: Tensor):
        self.weight = weight

    def __call__(self, idx: Tensor) -> Tensor:
        return self.weight.to(idx.device)[idx]


class Codebook:
    """vap/objective.py:79-146: class index <-> (2, n_bins) binary window; bit i of the index is entry i of the
    flattened (speaker, bin) window, speaker 0 first."""

    def __init__(self, bin_frames: List[int]):
        self.bin_frames = bin_frames
        self.n_bins = len(bin_frames)
        self.total_bins = 2 * self.n_bins
        self.n_classes = 2 ** self.total_bins
        self.emb = _Embedding(self.create_code_vectors(self.total_bins))

    def single_idx_to_onehot(self, idx: int, d: int = 8) -> Tensor:
        assert idx < 2 ** d, "must be possible with {d} binary digits"
        return torch.tensor([float((idx >> i) & 1) for i in range(d)])

    def create_code_vectors(self, n_bins: int) -> Tensor:
        return code_vectors(n_bins)

    def encode(self, x: Tensor) -> Tensor:
        """(*, 2, n_bins) -> (*): index of the nearest code vector. Per coordinate the nearest of {0, 1} is 1 above
        0.5; the reference's arg-max over negated distances (:126-139) resolves a tie at exactly 0.5 to the lower
        index, i.e. to 0, like the strict comparison here."""
        assert x.shape[-2:] == (2, self.n_bins), f"Codebook expects (..., 2, {self.n_bins}) got {x.shape}"
        bits = (x.flatten(-2) > 0.5).long()
        return (bits << torch.arange(self.total_bins, device=x.device)).sum(-1)

    def decode(self, idx: Tensor) -> Tensor:
        return self.emb(idx).unflatten(-1, (2, self.n_bins))

    def forward(self, projection_windows: Tensor) -> Tensor:
        return self.encode(projection_windows)

    __call__ = forward


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
        self.projection_window_extractor = ProjectionWindow(bin_times, frame_hz, threshold_ratio)
        self.codebook = Codebook(self.bin_frames)

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

    def window_to_win_dialog_states(self, wins: Tensor) -> Tensor:
        """(..., 2, n_bins) -> number of speakers active anywhere in the window (vap/objective.py:206-207)."""
        return (wins.sum(-1) > 0).sum(-1)

    def get_da_labels(self, va: Tensor):
        """vap/objective.py:214-218 -> (class index, dialog state of the window), each (B, T - horizon)."""
        wins = self.projection_window_extractor(va).type(va.dtype)
        return self.codebook(wins), self.window_to_win_dialog_states(wins)

    def loss_vad(self, vad_output: Tensor, vad: Tensor) -> Tensor:
        """vap/objective.py:245-247."""
        n = vad_output.shape[-2]
        return torch.nn.functional.binary_cross_entropy_with_logits(vad_output, vad[:, :n])

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
