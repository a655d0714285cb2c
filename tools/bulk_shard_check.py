#!/usr/bin/env python
"""BASELINE configs[3] / SURVEY §8(d) config 4, the small full run: N chunks of synthetic stereo audio (seeded by global
chunk id) sharded by contiguous blocks over the ranks, each rank running its shard through BulkRunner; the compact
outputs are all-gathered over NCCL and the counters all-reduced. Rank 0 then computes ALL chunks alone and checks the
gathered result is bit-identical and the counters equal. Launch with torchrun:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bulk_shard_check.py [chunks] [secs]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from voiceactivityprojection_b200 import VapConfig, VapGPT  # noqa: E402
from voiceactivityprojection_b200.bulk import COMPACT_KEYS, BulkRunner, BulkStats, gather_compact, shard_range  # noqa: E402

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
S, B = int(secs * 16000), 96
m = VapGPT(VapConfig(), precision="bf16").to(dev)
m.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))


def chunk(i):  # seeded by the GLOBAL chunk id, so every partition sees the same audio
    g = torch.Generator().manual_seed(1000 + i)
    return torch.randn((2, S), generator=g) * 0.05


def run(lo, hi):
    runner = BulkRunner(m, B, S, keys=COMPACT_KEYS[:-1] + ("argmax",), stats=True)
    outs = []
    batches = (torch.stack([chunk(i) for i in range(a, min(a + B, hi))]).pin_memory() for a in range(lo, hi, B))
    stats = runner.run(batches, sink=lambda i, b, o: outs.append({k: v[:b].clone() for k, v in o.items()}))
    return {k: torch.cat([o[k] for o in outs]).to(dev) for k in outs[0]}, stats


lo, hi = shard_range(n_chunks, rank, world)
local, stats = run(lo, hi)
allout = gather_compact(local)
tot = stats.all_reduce(device=dev)
if rank == 0:
    ref, rstats = run(0, n_chunks)
    same = {k: bool(torch.equal(allout[k], ref[k])) for k in ref}
    print(f"world {world}: {n_chunks} chunks of {secs} s, shard sizes {[shard_range(n_chunks, r, world)[1] - shard_range(n_chunks, r, world)[0] for r in range(world)]}")
    print("gathered == single-GPU:", same)
    print("counters:", tot.chunks, tot.frames, tot.vad_active.tolist(), "== single-GPU:",
          tot.chunks == rstats.chunks and tot.frames == rstats.frames and torch.equal(tot.class_hist, rstats.class_hist)
          and torch.equal(tot.vad_active, rstats.vad_active))
    assert all(same.values())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
