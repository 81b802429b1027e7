"""Sharding of independent (operator, omega) FAM problems over the GPUs of one box.

Every omega point is an independent solve (that is how pynfam farms them out,
pynfam/strength/fam_strength.py:204-249, pynfam/utilities/mpi_utils.py:61-226), so the partition has NO
data-path collective: each rank solves whole points; the only exchange is one gather of
2*(1+nxterms) doubles per point at the end.  The points of a contour differ in cost by a factor of three (18 to 54
iterations on the bench sweep), so the static partition balances an ESTIMATE of the iteration counts.
"""
import numpy as np


def expected_cost(omegas):
    """Heuristic iteration count of a FAM solve at each omega -- only its ordering and rough proportions matter.

    Measured on the bench sweep (162Gd, 16 shells, GT- K=0, 1024 points, scripts/sweep_iters.py): points below the
    threshold (left end of the interval) converge in 18 iterations whatever Im omega is; the count grows towards the
    upper end of the interval and towards the real axis (54 at Re omega = 10, |Im omega| = 0.4).  The fit
    18 + 40 s(Re omega) exp(-|Im omega| / 2.5), s rising linearly over the upper 70 % of the interval of the given points,
    correlates 0.98 with the measured counts."""
    om = np.asarray(omegas, dtype=complex)
    if len(om) == 0:
        return np.zeros(0)
    lo, hi = om.real.min(), om.real.max()
    span = max(hi - lo, 1e-12)
    s = np.clip((om.real - lo - 0.14 * span) / (0.7 * span), 0.0, 1.0) if hi > lo else np.ones(len(om))
    return 18.0 + 40.0 * s * np.exp(-np.abs(om.imag) / 2.5)


def partition(omegas, world):
    """Return a list of index arrays, one per rank: longest-processing-time-first on expected_cost with equal point counts
    (+-1) -- points by decreasing estimate, each to the rank with the least accumulated cost that still has room.  On the
    1024-point sweep over 8 ranks the simulated makespan is 0.9 % above the mean (round robin over |Im omega|: 4.2 %)."""
    om = np.asarray(omegas, dtype=complex)
    n = len(om)
    cost = expected_cost(om)
    order = np.lexsort((np.arange(n), np.abs(om.imag), -cost))      # cost descending, then |Im| ascending, then index
    cap = -(-n // world) if world > 0 else n
    full = n - (cap - 1) * world if n % world else world          # ranks that may take `cap` points (the others cap - 1)
    load = np.zeros(world)
    count = np.zeros(world, dtype=int)
    parts = [[] for _ in range(world)]
    for i in order:
        best = None
        for r in range(world):
            room = cap if r < full else cap - 1
            if count[r] < room and (best is None or load[r] < load[best] - 1e-12):
                best = r
        parts[best].append(int(i))
        load[best] += cost[i]
        count[best] += 1
    return [np.array(sorted(p), dtype=int) for p in parts]


def partition_tasks(n_operators, omegas, world):
    """(operator, omega point) tasks of a full contour run -> per rank a list of (operator index, point indices).

    The flattened operator-major task list is cut into `world` contiguous runs of equal length, so a rank holds whole
    operators plus at most two partial ones (large batches per solve), and the totals differ by at most one point.
    Inside an operator the points are listed cost-interleaved (stride permutation of the order by expected_cost), so
    every run gets the same mix of cheap and expensive points."""
    om = np.asarray(omegas, dtype=complex)
    npts = len(om)
    order = np.argsort(-expected_cost(om), kind="stable")       # most expensive first
    stride = max(1, -(-world // max(1, n_operators))) + 1
    perm = np.concatenate([order[s::stride] for s in range(stride)]) if npts else order
    total = n_operators * npts
    out = []
    for r in range(world):
        lo, hi = (total * r) // world, (total * (r + 1)) // world
        tasks = []
        for o in range(lo // npts if npts else 0, n_operators):
            a, b = max(lo, o * npts), min(hi, (o + 1) * npts)
            if a >= b:
                break
            tasks.append((o, np.sort(perm[a - o * npts:b - o * npts])))
        out.append(tasks)
    return out


def gather_strengths(local_idx, local_strength, npoints, dist=None, device=None):
    """All-gather the per-rank results into the original point order.

    local_strength: complex array [n_local, nstr].  With dist=None (single process) this is a reorder only.
    Uses torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    nstr = local_strength.shape[1]
    out = np.zeros((npoints, nstr), complex)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out[np.asarray(local_idx)] = local_strength
        return out
    world = dist.get_world_size()
    nmax = (npoints + world - 1) // world
    buf = torch.full((nmax, 1 + 2 * nstr), -1.0, dtype=torch.float64, device=device)
    n = len(local_idx)
    if n:
        buf[:n, 0] = torch.as_tensor(np.asarray(local_idx, dtype=np.float64), device=device)
        buf[:n, 1:1 + nstr] = torch.as_tensor(np.ascontiguousarray(local_strength.real), device=device)
        buf[:n, 1 + nstr:] = torch.as_tensor(np.ascontiguousarray(local_strength.imag), device=device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    for p in parts:
        p = p.cpu().numpy()
        for row in p:
            if row[0] >= 0:
                out[int(row[0])] = row[1:1 + nstr] + 1j * row[1 + nstr:]
    return out
