"""world_size-2 gloo test (CPU) of the multi-GPU plumbing in bench.py: contiguous batch shards with
seeds taken from the global sample index, and the max/sum statistics reduction that is the only
collective on the path (NCCL on the GPUs)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
    import bench
    from oracle import ref_port as P
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 7
    lo, hi = bench.shard_range(total, rank, world)
    # every rank generates its own samples from the GLOBAL index: results do not depend on world size
    sums = [float(P.synthetic_volume(i, (1, 6, 5, 4)).double().sum()) for i in range(lo, hi)]
    stats = torch.tensor([10.0 + rank, 20.0 - rank, float(hi - lo), float(np.sum(sums))], dtype=torch.float64)
    mx, sm = bench.aggregate(stats, world)
    q.put((rank, lo, hi, mx.tolist(), sm.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    import bench
    for total in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [bench.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_stats_reduction_gloo():
    from oracle import ref_port as P
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, mx0, sm0), (r1, lo1, hi1, mx1, sm1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)
    assert mx0 == mx1 and sm0 == sm1
    assert mx0[0] == 11.0 and mx0[1] == 20.0 and sm0[2] == 7.0
    want = sum(float(P.synthetic_volume(i, (1, 6, 5, 4)).double().sum()) for i in range(7))
    assert abs(sm0[3] - want) < 1e-9
