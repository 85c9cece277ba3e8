"""CPU, world_size 2, gloo: the host-side logic of the ray-sharded step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loner_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) normalisers
        counts = torch.tensor([100 + rank, 90 + 3 * rank], dtype=torch.int32)
        parallel.allreduce_counts(counts)
        # (2) flat gradient exchange: per-rank partial sums of a "loss" over this rank's ray shard
        g = torch.Generator().manual_seed(0)
        rays = torch.randn(1001, 7, generator=g)            # the same global ray set on every rank
        lo, hi = parallel.shard_slice(rays.shape[0], rank, world)
        w = torch.linspace(-1, 1, 7)
        local_grad = (rays[lo:hi] * w).sum(0)               # d/dw-like partial gradient
        local_pose = rays[lo:hi, :6].sum(0)
        local_acc = torch.tensor([float(hi - lo), 1.0, 2.0, 3.0])
        ex = parallel.FlatGrads(7, 1, "cpu")                 # the buffer the engine's kernels accumulate into
        ex.zero_()
        ex.d_params += local_grad
        ex.d_poses12[0, :6] += local_pose
        ex.loss_acc += local_acc
        ex.allreduce()
        assert ex.flat.numel() == 7 + 12 + 4 and ex.d_params.data_ptr() == ex.flat.data_ptr()      # views, no copies
        q.put((rank, counts.tolist(), ex.d_params.clone(), ex.d_poses12[0, :6].clone(), ex.loss_acc.clone(), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    rays = torch.randn(1001, 7, generator=g)
    w = torch.linspace(-1, 1, 7)
    full_grad, full_pose = (rays * w).sum(0), rays[:, :6].sum(0)
    covered = []
    for rank, counts, grad, pose, acc, (lo, hi) in res:
        assert counts == [201, 183]
        assert torch.allclose(grad, full_grad, rtol=1e-5, atol=1e-4)
        assert torch.allclose(pose, full_pose, rtol=1e-5, atol=1e-4)
        assert acc.tolist() == [1001.0, 2.0, 4.0, 6.0]
        covered.append((lo, hi))
    assert covered == [(0, 501), (501, 1001)]          # disjoint, complete, remainder on the first rank


def test_shard_slice_edge_cases():
    assert parallel.shard_slice(0, 0, 4) == (0, 0)
    assert [parallel.shard_slice(5, r, 8) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    assert parallel.shard_slice(8192, 3, 4) == (6144, 8192)
