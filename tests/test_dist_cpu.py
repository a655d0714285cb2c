"""World-size-2 gloo tests (CPU) of the multi-GPU host logic of the bulk driver:
chunk sharding, all-gather of ragged per-rank outputs, all-reduce of counters.
On the GPU box the same code runs over NCCL (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chunks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from voiceactivityprojection_b200.bulk import BulkStats, gather_compact, shard_range

        lo, hi = shard_range(n_chunks, rank, world)
        ids = torch.arange(lo, hi)
        # per-chunk "outputs" that are functions of the global chunk id
        local = {"p_now": ids.float()[:, None, None].repeat(1, 3, 2), "argmax": (ids % 256).to(torch.uint8)[:, None].repeat(1, 3)}
        full = gather_compact(local)
        st = BulkStats(chunks=hi - lo, frames=3 * (hi - lo))
        st.class_hist = torch.bincount(local["argmax"].reshape(-1).long(), minlength=256)
        st.vad_active = torch.tensor([hi - lo, rank])
        tot = st.all_reduce()
        if rank == 0:
            q.put((full["p_now"][:, 0, 0].tolist(), full["argmax"][:, 0].tolist(), tot.chunks, tot.frames,
                   tot.class_hist.tolist(), tot.vad_active.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [7, 8, 1])
def test_shard_gather_reduce_world2(n_chunks):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chunks, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    p_now, amax, chunks, frames, hist, vact = q.get()
    assert p_now == [float(i) for i in range(n_chunks)]          # rank order == global chunk order
    assert amax == [i % 256 for i in range(n_chunks)]
    assert chunks == n_chunks and frames == 3 * n_chunks
    assert sum(hist) == 3 * n_chunks and hist[0] == 3
    assert vact == [n_chunks, 1]


def test_shard_range_partitions_exactly():
    from voiceactivityprojection_b200.bulk import shard_range

    for n in [0, 1, 5, 8, 1_800_000]:
        for world in [1, 2, 3, 8]:
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
