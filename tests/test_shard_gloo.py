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
        # the most expensive points (largest expected cost) are spread over different ranks and the estimated loads agree
        cost = shard.expected_cost(om)
        hard = set(np.argsort(-cost, kind="stable")[:world].tolist())
        assert all(len(hard & set(p.tolist())) == 1 for p in parts)
        loads = [cost[p].sum() for p in parts]
        assert max(loads) - min(loads) <= cost.max() + 1e-9


def test_partition_balances_the_measured_iteration_counts_of_the_bench_sweep():
    """tests/golden/sweep_iters_1024.json: iteration counts of the 1024-point bench sweep (scripts/sweep_iters.py, one B200).
    Simulated slot scheduler (64 slots, a lock step costs a constant plus the active points): the slowest of 8 ranks stays
    within 2 % of the mean, the work per rank within 1 %."""
    import json
    import os
    d = np.array(json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sweep_iters_1024.json"))))
    om, it = d[:, 0] + 1j * d[:, 1], d[:, 2].astype(int)
    assert np.corrcoef(shard.expected_cost(om), it)[0, 1] > 0.95

    def makespan(idx, slots=64, c0=8.0):
        queue = sorted(idx, key=lambda i: abs(om[i].imag))
        live, t = [], 0.0
        while queue or live:
            while queue and len(live) < slots:
                live.append(it[queue.pop(0)])
            t += c0 + len(live)
            live = [r - 1 for r in live if r > 1]
        return t
    parts = shard.partition(om, 8)
    assert sorted(set(len(p) for p in parts)) == [128]
    t = np.array([makespan(p.tolist()) for p in parts])
    w = np.array([it[p].sum() for p in parts])
    assert t.max() / t.mean() < 1.02 and w.max() / w.mean() < 1.01


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


# ---- full-contour driver over ranks (strength.run_contours_sharded) -------------------------------------------------
def _stub_solve(fs, prob, ctx, idx):
    """Deterministic stand-in for the batched GPU solve: 'strengths' that depend on (operator, K, omega) only."""
    z = np.asarray(fs.contour.ctr_z)[idx]
    n = 1 + prob.iscalar("nxterms")
    s = np.stack([(j + 1 + fs.k) / (z - (3.0 + len(fs.op))) for j in range(n)], axis=1)
    return dict(strength=s, conv=np.ones(len(idx), int), iters=np.full(len(idx), 7), minutes=np.full(len(idx), 0.01),
                labels=[prob.label(i) for i in range(n)])


OPS = [("GT-", 0), ("GT-", 1), ("RS0-", 0)]


def _stage(wd):
    from conftest import stage_point
    stage_point("S40_SKOP_6sh", "GT-K0", 0, wd)


def _contour_worker(rank, world, port, wd, dest):
    from pynfam_b200.strength import famContour, run_contours_sharded
    d = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        d = dist
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036, "nr_points": 14})
    fss = run_contours_sharded(wd, "x.in", OPS, c, dest=dest, dist=d, solve_points=_stub_solve)
    assert [f.opname for f in fss] == ["GT-K0", "GT-K1", "RS0-K0"] and all(f.meta["Conv"] == "Yes" for f in fss)
    if world > 1:
        dist.destroy_process_group()


def test_partition_tasks_covers_every_task_once():
    om = 5.0 + 5.0 * np.exp(1j * np.linspace(np.pi, 2 * np.pi, 30))
    for nops, world in ((14, 8), (3, 8), (1, 2), (2, 1), (5, 4)):
        parts = shard.partition_tasks(nops, om, world)
        cover = np.zeros((nops, 30), int)
        for tasks in parts:
            for o, idx in tasks:
                cover[o, idx] += 1
        assert (cover == 1).all()
        load = [sum(len(i) for _, i in t) for t in parts]
        assert max(load) - min(load) <= 1 and max(len(t) for t in parts) <= 3
    # fewer tasks than ranks: some ranks (possibly rank 0) hold nothing, every task is still owned exactly once
    parts = shard.partition_tasks(1, om[:3], 8)
    assert sorted(int(i) for t in parts for _, idx in t for i in idx) == [0, 1, 2] and sum(1 for t in parts if not t) == 5


def test_two_rank_contour_driver_writes_the_same_files_as_one_rank(tmp_path):
    """world_size 2 (gloo): tasks sharded, strengths all-reduced, rank 0 writes OP.out / OP.out.ctr -- byte-identical
    .ctr files to the single-process run (the host set-up is real, the GPU solve is a stub: no GPU here)."""
    wd1, wd2 = str(tmp_path / "w1"), str(tmp_path / "w2")
    d1, d2 = str(tmp_path / "o1"), str(tmp_path / "o2")
    for d in (d1, d2):
        os.makedirs(d)
    _stage(wd1)
    _stage(wd2)
    _contour_worker(0, 1, 0, wd1, d1)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_contour_worker, args=(r, 2, port, wd2, d2)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    for op in ("GT-K0", "GT-K1", "RS0-K0"):
        a = open(os.path.join(d1, op + ".out.ctr"), "rb").read()
        assert a == open(os.path.join(d2, op + ".out.ctr"), "rb").read() and len(a) > 500
        assert os.path.isfile(os.path.join(d2, op + ".out"))


def _tbc_worker(rank, world, port, wd, dest):
    from pynfam_b200.strength import famContour, run_contours_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["PNFAM_B200_SETUP_TIMING"] = "1"          # the generator reports on stderr when it runs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.476036, "nr_points": 14})
    import contextlib
    import io
    import sys
    log = os.path.join(dest, "stderr.%d" % rank)
    fd = os.open(log, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 2)
    fss = run_contours_sharded(wd, "x.in", [("GT-", 0), ("GT-", 1)], c, dest=dest, dist=dist, solve_points=_stub_solve)
    assert [f.opname for f in fss] == ["GT-K0", "GT-K1"]
    dist.destroy_process_group()


def test_two_ranks_generate_each_missing_two_body_current_file_once(tmp_path):
    """Full-FAM two-body currents in a sharded run whose run directory has no .tbc files: the operators' points are split
    over both ranks, yet each file is computed by exactly one rank (round robin) and read by the other; the files equal the
    reference's (tests/golden/S40_All_GT2bc)."""
    import shutil
    from conftest import GOLDEN, load_points
    from test_tbc_generator import compare_files
    wd, dest = str(tmp_path / "w"), str(tmp_path / "o")
    os.makedirs(wd)
    os.makedirs(dest)
    src = os.path.join(GOLDEN, "S40_All_GT2bc")
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(src, f), wd)
    open(os.path.join(wd, "x.in"), "w").write(load_points("S40_All_GT2bc")["GT-K0"][0]["namelist"])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_tbc_worker, args=(r, 2, port, wd, dest)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    runs = [open(os.path.join(dest, "stderr.%d" % r)).read().count("2BC 1D tables") for r in range(2)]
    assert runs == [1, 1], runs                         # two files, one generation on each rank, none repeated
    for op in ("GT-K0", "GT-K1"):
        assert compare_files(os.path.join(wd, op + ".tbc"), os.path.join(src, op + ".tbc")) < 1e-12
        assert os.path.isfile(os.path.join(dest, op + ".out.ctr"))
