"""View-sharded aggregation over 2 processes (gloo, CPU).  The CUDA kernels cannot run here, so the
per-rank partial is computed by the oracle (test stand-in); what is under test is the host logic of
mvster_b200/sharding.py: view partitioning, the single fused [acc|wsum] all-reduce per stage, the
one-time normalisation - the result must equal the unsharded oracle aggregation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import oracle

from mvster_b200 import sharding, synth


def test_partition_views():
    assert sharding.partition_views(9, 4) == [(1, 3), (4, 2), (6, 2), (8, 2)]
    assert sharding.partition_views(4, 2) == [(1, 2), (3, 2)]
    assert sharding.partition_views(2, 4) == [(1, 1), (2, 1), (3, 0), (3, 0)]
    for n, p in [(4, 1), (10, 8), (7, 3)]:
        parts = sharding.partition_views(n, p)
        assert sum(c for _, c in parts) == n and parts[0][0] == 1
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(p - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, view_parallel, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        B, C, G, D, H, W = 1, 16, 4, 4, 16, 24
        g = torch.Generator().manual_seed(5)
        feats = [torch.randn(B, C, H, W, generator=g) for _ in range(n_views)]
        cams = synth.stage_projections(synth.arc_cameras(n_views, H, W, 2.0), B, num_stage=1)["stage1"]
        hypo = oracle.hypo_init_inverse(torch.tensor([[425.0, 935.0]]), D, H, W)
        shard = sharding.make_view_shard(n_views - 1, view_parallel)

        def partial(acc, wsum):  # oracle stand-in for mvster_et_fuse_f32(..., MVSTER_ET_PARTIAL)
            a, w = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0, views=shard.views, partial=True)
            acc.copy_(a.permute(0, 2, 3, 4, 1))
            wsum.copy_(w)

        def normalize(acc, wsum):
            return acc / (1e-8 + wsum).unsqueeze(-1)

        cost = sharding.sharded_aggregate(partial, normalize, (B, D, H, W, G), shard, torch.device("cpu"))
        full = oracle.et_aggregate(feats, cams, hypo, True, G, 2.0).permute(0, 2, 3, 4, 1)
        err = (cost - full).abs().max().item() / full.abs().max().item()
        q.put((rank, shard.first_view, shard.count, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,view_parallel,n_views", [(2, 2, 5), (2, 2, 2), (2, 1, 3)])
def test_view_sharded_aggregate_equals_unsharded(world, view_parallel, n_views):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, view_parallel, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    owned = 0
    for rank, first, count, err in res:
        assert err < 2e-6, f"rank {rank}: sharded != unsharded ({err:.2e})"
        owned += count
    assert owned == (n_views - 1) * (world // view_parallel)
