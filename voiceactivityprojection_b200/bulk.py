"""Bulk inference over many independent chunks / sessions (BASELINE.json configs[3]).

The reference has no bulk or multi-GPU inference driver (every entry point is a
single-process loop around `model.probs`, run.py:236-241, vap/extraction.py:262-270);
this is the B200-side caller that keeps the device busy:

* `BulkRunner.run` pipelines host batches through CUDA streams — host->device copy
  of batch i+1, the forward of batch i, the collectives and the device->host copy of
  batch i-1 overlap (double-buffered device and pinned host buffers).
* Defaults follow SURVEY.md §8e/§8f: host batches may be int16 PCM (what wav files
  hold; the fused encoder kernel reads it directly, half the PCIe bytes of float32),
  and what comes back is the COMPACT per-chunk set — vad, p_now, p_future, H and the
  uint8 arg-max class, 29 KB per 20 s chunk instead of 1.05 MB — laid out in ONE
  contiguous device buffer per batch, so the device->host copy and the all-gather
  are one transfer each. Full `probs` / `loss` on request (`keys=ALL_KEYS`).
* Chunks shard by index across ranks (`shard_range`), one process per GPU; the
  path has no data-path exchange. With `gather=True` every step issues, on a side
  stream, ONE all-gather of the compact buffer and ONE all-reduce of the step's
  counters (class histogram, VAD-active frames) — NCCL on GPUs, the only
  collectives of the path. The counters are taken inside the heads kernel
  (`vapb_probs_ex`), not by torch ops.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, Optional, Sequence, Tuple

import torch
from torch import Tensor

ALL_KEYS = ("probs", "vad", "p_now", "p_future", "H", "loss")
COMPACT_KEYS = ("vad", "p_now", "p_future", "H", "argmax")
N_COUNTERS = 258  # 256 arg-max classes + active frames of the two channels


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


class CompactLayout:
    """The compact per-chunk outputs of a batch of capacity B in one byte buffer:
    vad | p_now | p_future (B,T,2) float32, H (B,T) float32, argmax (B,T) uint8, back to back."""

    def __init__(self, batch: int, T: int):
        self.batch, self.T = batch, T
        f2, f1 = batch * T * 2 * 4, batch * T * 4
        self.offsets = {"vad": 0, "p_now": f2, "p_future": 2 * f2, "H": 3 * f2, "argmax": 3 * f2 + f1}
        self.nbytes = 3 * f2 + f1 + batch * T
        self.bytes_per_chunk = self.nbytes // batch

    def views(self, buf: Tensor) -> Dict[str, Tensor]:
        """Tensors over a (nbytes,) uint8 buffer (device or pinned host); nothing is copied."""
        B, T, o = self.batch, self.T, self.offsets
        f = lambda key, shape: buf[o[key]: o[key] + 4 * B * T * (2 if len(shape) == 3 else 1)].view(torch.float32).view(shape)
        return {"vad": f("vad", (B, T, 2)), "p_now": f("p_now", (B, T, 2)), "p_future": f("p_future", (B, T, 2)),
                "H": f("H", (B, T)), "argmax": buf[o["argmax"]: o["argmax"] + B * T].view(B, T)}


class BulkRunner:
    """Pipelined `model.probs` over a stream of pinned host batches of shape (b <= batch, 2, n_samples).

    Host batches are float32 or int16 PCM (decided by the first batch, or `pcm16=`). A batch must stay untouched
    until its `sink` callback has run (the host->device copy is asynchronous; no host-side staging copy is made).
    """

    def __init__(self, model, batch: int, n_samples: int, precision: Optional[str] = None,
                 keys: Sequence[str] = COMPACT_KEYS, stats: bool = True, depth: int = 2, pcm16: Optional[bool] = None,
                 input_rate: Optional[int] = None, gather: bool = False, group=None):
        """keys: what reaches the host per batch — any of probs, vad, p_now, p_future, H, loss, argmax. The compact
        five travel as one buffer; `probs` and `loss` are computed only when asked for.
        pcm16: host batches are int16 PCM (the reference converts to float on the CPU, vap/audio.py:47); the fused
        encoder kernel reads them directly and scales by 1/32768. None = follow the first batch's dtype.
        input_rate: sample rate of the host batches when it is not the model's 16 kHz; they are then
        (b, 2, n_in) with ceil(16000 * n_in / input_rate) == n_samples and are resampled on the device
        (audio.resample_device; the reference resamples on the CPU, vap/audio.py:65-68).
        gather: every step all-gathers the compact buffer and all-reduces the step's counters over `group` on a side
        stream (torch.distributed must be initialised); rank 0's sink then also receives `gathered` (world, ...)."""
        from . import _lib

        if model._device.type != "cuda":
            raise RuntimeError("BulkRunner needs the model on a CUDA device (no CPU fallback)")
        self.model, self.batch, self.n_samples, self.precision = model, batch, n_samples, precision
        self.keys, self.stats_on, self.depth = tuple(keys), stats, depth
        bad = [k for k in self.keys if k not in ALL_KEYS + ("argmax",)]
        if bad:
            raise ValueError(f"unknown output keys {bad}")
        self.dev = model._device
        _, self.T = _lib.frames(n_samples)
        self.pcm16 = pcm16
        self.input_rate = None if input_rate in (None, model.sample_rate) else int(input_rate)
        self.n_in = n_samples
        if self.input_rate:
            # the same duration at the input rate (its resampled length, ceil(16000 * n_in / rate), is >= n_samples;
            # the resampler writes the first n_samples)
            self.n_in = -(-n_samples * self.input_rate // model.sample_rate)
        self.din = None  # allocated with the first batch (dtype)
        self.dwav = None
        self.layout = CompactLayout(batch, self.T)
        self.cbuf = [torch.empty(self.layout.nbytes, dtype=torch.uint8, device=self.dev) for _ in range(depth)]
        self.dout = [self.layout.views(b) for b in self.cbuf]
        self.want_probs, self.want_loss = "probs" in self.keys, "loss" in self.keys
        f = dict(dtype=torch.float32, device=self.dev)
        for o in self.dout:
            if self.want_probs:
                o["probs"] = torch.empty((batch, self.T, 256), **f)
            if self.want_loss:
                o["loss"] = torch.empty((batch, max(self.T - 100, 0)), **f)
        self.hbuf = [torch.empty(self.layout.nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
        self.hout = [self.layout.views(b) for b in self.hbuf]
        for ho, do in zip(self.hout, self.dout):
            for k in ("probs", "loss"):
                if k in do:
                    ho[k] = torch.empty(do[k].shape, dtype=torch.float32, pin_memory=True)
        # per-step counters, taken by the heads kernel (uint64 on the device; int64 views here)
        self.cnt = [torch.zeros(N_COUNTERS, dtype=torch.int64, device=self.dev) for _ in range(depth)]
        self.hcnt = [torch.zeros(N_COUNTERS, dtype=torch.int64, pin_memory=True) for _ in range(depth)]
        self.gather = bool(gather)
        self.group = group
        self.world = self.rank = None
        self.gath = self.hgath = None
        if self.gather:
            import torch.distributed as dist

            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("BulkRunner(gather=True) needs an initialised torch.distributed process group")
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            self.gath = [torch.empty((self.world, self.layout.nbytes), dtype=torch.uint8, device=self.dev)
                         for _ in range(depth)]
            if self.rank == 0:
                self.hgath = [torch.empty((self.world, self.layout.nbytes), dtype=torch.uint8, pin_memory=True)
                              for _ in range(depth)]
        self.s_h2d, self.s_cmp, self.s_coll, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(4))
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]
        self.ev_h2d, self.ev_cmp, self.ev_coll, self.ev_d2h = mk(), mk(), mk(), mk()
        self.h2d_bytes = self.d2h_bytes = self.coll_bytes = 0

    # ------------------------------------------------------------------
    def _alloc_inputs(self, dtype):
        if dtype not in (torch.float32, torch.int16):
            raise ValueError(f"host batches must be float32 or int16 PCM, got {dtype}")
        if self.pcm16 is None:
            self.pcm16 = dtype == torch.int16
        want = torch.int16 if self.pcm16 else torch.float32
        if dtype != want:
            raise ValueError(f"expected {want} host batches, got {dtype}")
        self.din = [torch.empty((self.batch, 2, self.n_in), dtype=want, device=self.dev) for _ in range(self.depth)]
        if self.input_rate:
            self.dwav = torch.empty((self.batch, 2, self.n_samples), dtype=torch.float32, device=self.dev)

    def run(self, batches: Iterable[Tensor], sink: Optional[Callable[[int, int, Dict[str, Tensor]], None]] = None
            ) -> BulkStats:
        """sink(batch_index, b, host_outputs) is called once per batch, in order; the pinned host tensors it
        receives are valid until `depth` more batches have been issued. With gather=True rank 0's dict also holds
        `gathered`: one dict of compact outputs (capacity `batch`) per rank, views of that step's all-gathered buffer."""
        from . import _lib

        lib = _lib.load()
        st = BulkStats()
        self._tot = torch.zeros(N_COUNTERS, dtype=torch.int64)
        pending = []  # (index, slot, b)
        kw = {} if self.precision is None else {"precision": self.precision}
        cur = torch.cuda.current_stream(self.dev)
        streams = (self.s_h2d, self.s_cmp, self.s_coll, self.s_d2h)
        for s in streams:
            s.wait_stream(cur)
        for i, hb in enumerate(batches):
            if self.din is None:
                self._alloc_inputs(hb.dtype)
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
                self.s_cmp.wait_event(self.ev_d2h[slot])      # dout[slot] / cnt[slot] have been copied out
                o = {k: v[:b] for k, v in self.dout[slot].items()}
                if self.input_rate:
                    from .audio import resample_device

                    wav = resample_device(self.din[slot][:b], self.input_rate, self.model.sample_rate, out=self.dwav[:b])
                else:
                    wav = self.din[slot][:b]  # float32, or int16 PCM read by the encoder kernel itself
                cnt = None
                if self.stats_on:
                    cnt = self.cnt[slot]
                    _lib.check(lib, None, lib.vapb_memset_zero(self.s_cmp.cuda_stream, cnt.data_ptr(), cnt.numel() * 8))
                self.model.probs(wav, out=o, counters=cnt, want_probs=self.want_probs, want_loss=self.want_loss, **kw)
                self.ev_cmp[slot].record(self.s_cmp)
            last = self.ev_cmp[slot]
            if self.gather:
                import torch.distributed as dist

                with torch.cuda.stream(self.s_coll):
                    self.s_coll.wait_event(self.ev_cmp[slot])
                    dist.all_gather_into_tensor(self.gath[slot].view(-1), self.cbuf[slot], group=self.group)
                    self.coll_bytes += self.world * self.layout.nbytes
                    if self.stats_on:
                        dist.all_reduce(self.cnt[slot], op=dist.ReduceOp.SUM, group=self.group)
                        self.coll_bytes += N_COUNTERS * 8
                    self.ev_coll[slot].record(self.s_coll)
                last = self.ev_coll[slot]
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(last)
                if self.gather and self.rank == 0:
                    self.hgath[slot].copy_(self.gath[slot], non_blocking=True)
                    self.d2h_bytes += self.gath[slot].numel()
                else:
                    n = self.layout.nbytes if b == self.batch else None
                    if n is not None:
                        self.hbuf[slot].copy_(self.cbuf[slot], non_blocking=True)
                        self.d2h_bytes += n
                    else:  # ragged last batch: the b items of every segment
                        for k in COMPACT_KEYS:
                            self.hout[slot][k][:b].copy_(self.dout[slot][k][:b], non_blocking=True)
                            self.d2h_bytes += self.dout[slot][k][:b].numel() * self.dout[slot][k].element_size()
                for k in ("probs", "loss"):
                    if k in self.dout[slot]:
                        self.hout[slot][k][:b].copy_(self.dout[slot][k][:b], non_blocking=True)
                        self.d2h_bytes += self.dout[slot][k][:b].numel() * 4
                if self.stats_on:
                    self.hcnt[slot].copy_(self.cnt[slot], non_blocking=True)
                    self.d2h_bytes += N_COUNTERS * 8
                self.ev_d2h[slot].record(self.s_d2h)
            pending.append((i, slot, b))
            st.chunks += b
            st.frames += b * self.T
            if len(pending) >= self.depth:
                self._finish(pending.pop(0), sink)
        while pending:
            self._finish(pending.pop(0), sink)
        for s in streams:
            cur.wait_stream(s)
        if self.stats_on:
            # with gather=True the per-step counters were all-reduced on the device: these are global totals
            st.class_hist, st.vad_active = self._tot[:256].clone(), self._tot[256:258].clone()
        return st

    def _finish(self, item, sink):
        i, slot, b = item
        self.ev_d2h[slot].synchronize()
        if self.stats_on:
            self._tot += self.hcnt[slot]
        if sink is None:
            return
        if self.gather and self.rank == 0:
            g = [self.layout.views(self.hgath[slot][r]) for r in range(self.world)]  # views, nothing is copied
            out = {k: g[0][k][:b] for k in COMPACT_KEYS}
            out["gathered"] = g
        else:
            out = {k: self.hout[slot][k][:b] for k in COMPACT_KEYS}
        for k in ("probs", "loss"):
            if k in self.hout[slot]:
                out[k] = self.hout[slot][k][:b]
        sink(i, b, {k: v for k, v in out.items() if k in self.keys or k in ("gathered", "argmax")})
