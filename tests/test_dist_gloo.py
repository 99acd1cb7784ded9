"""Env sharding + pose all-gather with world_size 2 and 3 on CPU (gloo): the host-side logic of the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgbmanip_b200.dist import estimate_sharded, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_range(n, r, world)
                assert 0 <= hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _fake_estimate(K, tag):
    """Stand-in for the device estimator: a deterministic function of the env's inputs only."""
    n = K.shape[0]
    base = K.reshape(n, -1).sum(1, keepdim=True).double()
    return (base[:, :, None] + torch.arange(24, dtype=torch.float64).reshape(1, 8, 3) * tag[:, None, None]).contiguous()


def _worker(rank, world, port, n, q, pass_device=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        K = torch.rand((n, 3, 3), generator=g, dtype=torch.float64)
        tag = torch.arange(n, dtype=torch.float64) + 1
        out = estimate_sharded(_fake_estimate, (K, tag), n, device=torch.device("cpu") if pass_device else None)
        want = _fake_estimate(K, tag)
        q.put((rank, bool(torch.equal(out, want))))
    finally:
        dist.destroy_process_group()


def test_single_process_empty_batch_returns_empty_tensor():
    out = estimate_sharded(_fake_estimate, (torch.zeros((0, 3, 3), dtype=torch.float64), torch.zeros(0, dtype=torch.float64)), 0)
    assert tuple(out.shape) == (0, 8, 3) and out.dtype == torch.float64


@pytest.mark.parametrize("world,n,pass_device", [(2, 10, True), (2, 7, True), (3, 8, True), (3, 2, False)],
                         ids=["w2n10", "w2n7", "w3n8", "w3n2_empty_shard_device_from_backend"])
def test_sharded_estimate_is_identical_to_single_process(world, n, pass_device):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q, pass_device)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(r, True) for r in range(world)]
