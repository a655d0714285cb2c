"""Bulk inference over many independent chunks / sessions (BASELINE.json configs[3]).

The reference has no bulk or multi-GPU inference driver (every entry point is a
single-process loop around `model.probs`, run.py:236-241, vap/extraction.py:262-270);
this is the B200-side caller that keeps the device busy:

* `BulkRunner.run` pipelines host batches through three CUDA streams —
  host->device copy of batch i+1, the forward of batch i and the device->host
  copy of batch i-1 overlap (double-buffered device and pinned host buffers).
* Chunks shard by index across ranks (`shard_range`), one process per GPU; the
  path has no data-path exchange. `gather_compact` all-gathers
  the compact per-chunk outputs and `BulkStats.all_reduce` sums the counters
  (class histogram, VAD-active frames, frames, chunks) — the only collectives.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, Optional, Sequence, Tuple

import torch
from torch import Tensor

ALL_KEYS = ("probs", "vad", "p_now", "p_future", "H", "loss")
COMPACT_KEYS = ("vad", "p_now", "p_future", "H", "argmax")


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of item ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class BulkStats:
    chunks: int = 0
    frames: int = 0
    class_hist: Tensor = field(default_factory=lambda: torch.zeros(256, dtype=torch.int64))
    vad_active: Tensor = field(default_factory=lambda: torch.zeros(2, dtype=torch.int64))

    def as_tensor(self, device) -> Tensor:
        head = torch.tensor([self.chunks, self.frames], dtype=torch.int64, device=device)
        return torch.cat([head, self.class_hist.to(device), self.vad_active.to(device)])

    @staticmethod
    def from_tensor(t: Tensor) -> "BulkStats":
        t = t.cpu()
        return BulkStats(int(t[0]), int(t[1]), t[2:258].clone(), t[258:260].clone())

    def all_reduce(self, device=None, group=None) -> "BulkStats":
        """Sum over ranks (NCCL on GPUs, gloo in the CPU tests)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return self
        t = self.as_tensor(device if device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return BulkStats.from_tensor(t)


def gather_compact(local: Dict[str, Tensor], group=None) -> Dict[str, Tensor]:
    """All-gather per-chunk outputs along dim 0 in rank order; per-rank chunk counts may differ."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    any_t = next(iter(local.values()))
    n = torch.tensor([any_t.shape[0]], dtype=torch.int64, device=any_t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    out = {}
    for k, v in local.items():
        pad = v if v.shape[0] == nmax else torch.cat([v, v.new_zeros((nmax - v.shape[0],) + tuple(v.shape[1:]))])
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad.contiguous(), group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out


class BulkRunner:
    """Pipelined `model.probs` over a stream of pinned host batches of shape (b <= batch, 2, n_samples)."""

    def __init__(self, model, batch: int, n_samples: int, precision: Optional[str] = None,
                 keys: Sequence[str] = ALL_KEYS, stats: bool = True, depth: int = 2, pcm16: bool = False,
                 input_rate: Optional[int] = None):
        """pcm16=True: host batches are int16 PCM (what wav files hold; the reference converts to float on the CPU,
        vap/audio.py:47); they cross PCIe at half the bytes and are scaled by 1/32768 on the device.
        input_rate: sample rate of the host batches when it is not the model's 16 kHz; they are then
        (b, 2, n_in) with ceil(16000 * n_in / input_rate) == n_samples and are resampled on the device
        (audio.resample_device; the reference resamples on the CPU, vap/audio.py:65-68)."""
        from . import _lib

        if model._device.type != "cuda":
            raise RuntimeError("BulkRunner needs the model on a CUDA device (no CPU fallback)")
        self.model, self.batch, self.n_samples, self.precision = model, batch, n_samples, precision
        self.keys, self.stats_on, self.depth = tuple(keys), stats, depth
        self.dev = model._device
        _, self.T = _lib.frames(n_samples)
        want_argmax = stats or "argmax" in self.keys
        self.pcm16 = pcm16
        in_dtype = torch.int16 if pcm16 else torch.float32
        self.input_rate = None if input_rate in (None, model.sample_rate) else int(input_rate)
        self.n_in = n_samples
        if self.input_rate:
            # the same duration at the input rate (its resampled length, ceil(16000 * n_in / rate), is >= n_samples;
            # the resampler writes the first n_samples)
            self.n_in = -(-n_samples * self.input_rate // model.sample_rate)
        self.din = [torch.empty((batch, 2, self.n_in), dtype=in_dtype, device=self.dev) for _ in range(depth)]
        self.dwav = (torch.empty((batch, 2, n_samples), dtype=torch.float32, device=self.dev)
                     if pcm16 or self.input_rate else None)
        self.dout = [model.alloc_outputs(batch, self.T, self.dev, argmax=want_argmax) for _ in range(depth)]
        host = model.alloc_outputs(batch, self.T, "cpu", argmax=True, pin_memory=True)
        self.hout = [{k: torch.empty_like(host[k], pin_memory=True) for k in self.keys} for _ in range(depth)]
        self.s_h2d, self.s_cmp, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.ev_h2d, self.ev_cmp, self.ev_d2h = mk(), mk(), mk()
        self.h2d_bytes = self.d2h_bytes = 0
        self._hist = torch.zeros(256, dtype=torch.int64, device=self.dev)
        self._ones = torch.ones(batch * self.T, dtype=torch.int64, device=self.dev)
        self._vact = torch.zeros(2, dtype=torch.int64, device=self.dev)

    def run(self, batches: Iterable[Tensor], sink: Optional[Callable[[int, int, Dict[str, Tensor]], None]] = None
            ) -> BulkStats:
        """sink(batch_index, b, host_outputs) is called once per batch, in order; the pinned
        host tensors it receives are valid until `depth` more batches have been issued."""
        st = BulkStats()
        self._hist.zero_()
        self._vact.zero_()
        pending = []  # (index, slot, b)
        kw = {} if self.precision is None else {"precision": self.precision}
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            s.wait_stream(cur)
        for i, hb in enumerate(batches):
            b = hb.shape[0]
            if (hb.device.type != "cpu" or b > self.batch or tuple(hb.shape[1:]) != (2, self.n_in)
                    or hb.dtype != self.din[0].dtype):
                raise ValueError(f"batch {i}: expected a CPU {self.din[0].dtype} tensor (<= {self.batch}, 2, "
                                 f"{self.n_in}), got {hb.dtype} {tuple(hb.shape)} on {hb.device}")
            slot = i % self.depth
            with torch.cuda.stream(self.s_h2d):
                self.s_h2d.wait_event(self.ev_cmp[slot])      # the forward that read din[slot] has finished
                self.din[slot][:b].copy_(hb, non_blocking=True)
                self.ev_h2d[slot].record(self.s_h2d)
            self.h2d_bytes += hb.numel() * hb.element_size()
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.ev_h2d[slot])
                self.s_cmp.wait_event(self.ev_d2h[slot])      # dout[slot] has been copied out
                o = {k: v[:b] for k, v in self.dout[slot].items()}
                if self.input_rate:
                    from .audio import resample_device

                    wav = resample_device(self.din[slot][:b], self.input_rate, self.model.sample_rate, out=self.dwav[:b])
                elif self.pcm16:
                    wav = self.dwav[:b]
                    torch.mul(self.din[slot][:b], 1.0 / 32768.0, out=wav)  # one pass: int16 -> float32, exact scaling
                else:
                    wav = self.din[slot][:b]
                self.model.probs(wav, out=o, **kw)
                if self.stats_on:
                    # index_add_, not bincount: bincount reads its maximum back to the host and would stall the pipeline
                    self._hist.index_add_(0, o["argmax"].reshape(-1).to(torch.int64), self._ones[: b * self.T])
                    self._vact += (o["vad"] >= 0.5).sum(dim=(0, 1))
                self.ev_cmp[slot].record(self.s_cmp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(self.ev_cmp[slot])
                for k in self.keys:
                    self.hout[slot][k][:b].copy_(self.dout[slot][k][:b], non_blocking=True)
                    self.d2h_bytes += self.dout[slot][k][:b].numel() * self.dout[slot][k].element_size()
                self.ev_d2h[slot].record(self.s_d2h)
            pending.append((i, slot, b))
            st.chunks += b
            st.frames += b * self.T
            if len(pending) >= self.depth:
                self._finish(pending.pop(0), sink)
        while pending:
            self._finish(pending.pop(0), sink)
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            cur.wait_stream(s)
        if self.stats_on:
            st.class_hist, st.vad_active = self._hist.cpu(), self._vact.cpu()
        return st

    def _finish(self, item, sink):
        i, slot, b = item
        self.ev_d2h[slot].synchronize()
        if sink is not None:
            sink(i, b, {k: v[:b] for k, v in self.hout[slot].items()})
