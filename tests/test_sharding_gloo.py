"""N>1 host logic on CPU: contiguous frame shards over ranks, no data-path collective, host gather of
per-frame counts and max-over-ranks timing -- exercised with a world_size-2 gloo group. The per-frame
work here is the CPU oracle (the checker), standing in for a GPU so the plumbing can run in CI."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_every_unit_once():
    import cuda_surf_b200 as sb
    for n in (1, 7, 512, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sb.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sb.shard_range(n, world, r)[1] - sb.shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sb.shard_range(1024, 8, 3) == (384, 512)
    with pytest.raises(ValueError):
        sb.shard_range(10, 2, 2)


def _worker(rank, world, port, nframes, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cuda_surf_b200 as sb
    from cuda_surf_b200.sharding import gather_counts
    import oracle_lib as ol
    lo, hi = sb.shard_range(nframes, world, rank)
    orc = ol.Oracle(2, 4.0, False, 9, 2, True, False, 4)
    counts = []
    for f in range(lo, hi):
        pts, _ = orc.detect_and_compute(sb.synth_frame(160, 120, 1000 + f), desc=False)
        counts.append(len(pts))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)  # stand-in for this rank's elapsed time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    allc = gather_counts(counts, nframes, world, rank, dist)
    if rank == 0:
        q.put((allc, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    import cuda_surf_b200 as sb
    import oracle_lib as ol
    nframes, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    allc, tmax = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    orc = ol.Oracle(2, 4.0, False, 9, 2, True, False, 4)
    want = [len(orc.detect_and_compute(sb.synth_frame(160, 120, 1000 + f), desc=False)[0]) for f in range(nframes)]
    assert allc == want          # frame order preserved by the gather
    assert tmax == float(world)  # max over ranks
