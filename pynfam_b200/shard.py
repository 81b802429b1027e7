"""Sharding of independent (operator, omega) FAM problems over the GPUs of one box.

Every omega point is an independent solve (that is how pynfam farms them out,
pynfam/strength/fam_strength.py:204-249, pynfam/utilities/mpi_utils.py:61-226), so the partition has NO
data-path collective: each rank solves whole points; the only exchange is one gather of
2*(1+nxterms) doubles per point at the end.  Points nearer the real axis need more iterations, so the
static partition interleaves points (round-robin after sorting by |Im omega|) to balance the load.
"""
import numpy as np


def partition(omegas, world):
    """Return a list of index arrays, one per rank (round-robin over points sorted by expected cost)."""
    om = np.asarray(omegas, dtype=complex)
    order = np.argsort(np.abs(om.imag), kind="stable")        # most expensive (small |Im|) first
    return [order[r::world] for r in range(world)]


def partition_tasks(n_operators, omegas, world):
    """(operator, omega point) tasks of a full contour run -> per rank a list of (operator index, point indices).

    The flattened operator-major task list is cut into `world` contiguous runs of equal length, so a rank holds whole
    operators plus at most two partial ones (large batches per solve), and the totals differ by at most one point.
    Inside an operator the points are listed cost-interleaved (stride permutation of the order by |Im omega|), so
    every run gets the same mix of cheap and expensive points."""
    om = np.asarray(omegas, dtype=complex)
    npts = len(om)
    order = np.argsort(np.abs(om.imag), kind="stable")
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
