"""CPU test of the N>1 path: world_size-2 gloo run of the omega-point sharding + final gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynfam_b200 import shard


def _fake_solve(om):
    """Deterministic stand-in for the per-point result (2 'strengths' per point)."""
    return np.stack([1.0 / (om - 3.0), om ** 2], axis=1)


def _worker(rank, world, port, npts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    om = 5.0 + 5.0 * np.exp(1j * np.linspace(0.1, 6.0, npts))
    mine = shard.partition(om, world)[rank]
    full = shard.gather_strengths(mine, _fake_solve(om[mine]), npts, dist=dist, device="cpu")
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_every_point_once_and_balances():
    om = 5.0 + 5.0 * np.exp(1j * np.linspace(0.1, 6.0, 37))
    for world in (1, 2, 4, 8):
        parts = shard.partition(om, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(37))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        # the most expensive points (smallest |Im omega|) are spread over different ranks
        hard = set(np.argsort(np.abs(om.imag))[:world])
        assert all(len(hard & set(p.tolist())) == 1 for p in parts)


def test_two_rank_gloo_gather_reassembles_all_points():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    npts = 11   # ragged: 6 + 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, npts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    om = 5.0 + 5.0 * np.exp(1j * np.linspace(0.1, 6.0, npts))
    want = _fake_solve(om)
    for rank, full in res:
        assert np.allclose(full, want, rtol=0, atol=0)


def test_single_process_gather_is_a_reorder():
    om = np.array([1 + 1j, 2 + 0.1j, 3 + 2j])
    idx = shard.partition(om, 1)[0]
    full = shard.gather_strengths(idx, _fake_solve(om[idx]), 3)
    assert np.array_equal(full, _fake_solve(om))
