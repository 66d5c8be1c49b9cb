"""The N>1 path on CPU: world_size-2 gloo processes shard a clip list, agree on
a max-over-ranks time and reassemble results in input order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meyda_b200.sharding import aggregate_throughput, shard_by_frames, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 20000):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_frames_balances_ragged_clips():
    rng = np.random.default_rng(0)
    f = rng.integers(0, 3000, 500)
    for world in (2, 3, 8):
        parts = shard_by_frames(f, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(f)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        loads = [int(f[a:b].sum()) for a, b in parts]
        assert max(loads) - min(loads) <= 2 * 3000
    assert shard_by_frames([0, 0, 0], 2) == [(0, 0), (0, 3)]
    assert shard_by_frames([], 2) == [(0, 0), (0, 0)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_clips, fpc = 11, 5
    c0, c1 = shard_range(n_clips, world, rank)
    # stand-in for the per-rank extract: frame g of clip c gets value 1000*c + g
    mine = torch.tensor([[1000.0 * c + g for g in range(fpc)] for c in range(c0, c1)]).reshape(-1)
    ms = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks, as bench.py does
    frames = torch.tensor([mine.numel()])
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tolist())
    if rank == 0:
        q.put((float(ms.item()), int(frames.item()), [v for part in gathered for v in part]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, frames, values = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 20.0 and frames == 55
    assert values == [1000.0 * c + g for c in range(11) for g in range(5)]  # input order preserved
    assert aggregate_throughput(frames, 3, ms) == 55 * 3 / 0.020
