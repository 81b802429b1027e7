#!/usr/bin/env python
"""bench.py -- pnFAM throughput on B200: FAM iterations/s and omega-points/s.

Workload (config.workload): 162Gd, SkO', 16 HO shells, 40x40 Gauss grid (N=1958, nghl=1600, nxy=103926),
Gamow-Teller K=0, synthetic contour sweep: CIRCLE contours (pynfam/strength/contour.py:212-283 restated) of 64
Gauss-Legendre nodes on [0, 10+0.25g] MeV, `--points` points PER GPU (default 128 = two contours; 1024 points at 8
GPUs, the span BASELINE.json configs[4] names), every point solved to convergence (eps=1e-7, M=50, max_iter=300).
Inputs: tests/golden/Gd162_SKOP_16sh/ (made with the reference's own hfbtho_main).

One "step" = one full batched contour solve on each rank (all FAM iterations of all its points).
  value   = FAM iterations/s, whole job, device-resident (CUDA events around the iteration loop)
  e2e     = same metric through the C ABI from HOST buffers: context creation (H2D of the model) +
            operator upload + solve + D2H of the strengths, wall clock, every step (after one untimed
            warm-up of that path)
  impl=reference : the reference's unmodified pnfam_main.x (oracle/_ref) on the host cores, a bounded
            sample of the same workload (one omega point, `--ref-iters` iterations), iterations/s from its
            own per-iteration timer.
"""
import argparse
import json
import os
import re
import shutil
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SHELLS = 16     # --shells 12|20|24 switch to tests/golden/Gd162_SKOP_<n>sh (the basis sizes of BASELINE.json configs[3]/[4])


def case_dir():
    return os.path.join(ROOT, "tests", "golden", "Gd162_SKOP_%dsh" % SHELLS)


def workload():
    return ("Gd162 SkO' %d shells 40x40 grid, GT- K=0, synthetic contour sweep (CIRCLE contours of %d Gauss-Legendre nodes "
            "on [0, 10+0.25g] MeV)" % (SHELLS, NODES_PER_CIRCLE))

FAM_NML = """&general
    fam_output_filename = 'GT-K0'
    print_stdout = .true.
    use_fam_storage = 0
    real_eqrpa = {re}
    imag_eqrpa = {im}
/
&ext_field
    beta_type = '-'
    operator_name = 'GT'
    operator_k = 0
    compute_crossterms = .true.
    two_body_current_mode = 0
/
&interaction
    interaction_name = 'SKOP'
    require_self_consistency = .true.
    require_gauge_invariance = .true.
    force_j2_terms = .false.
    vpair_t0 = -346.352
    vpair_t1 = ,
    override_cs0 = 128.279
    override_csr = 0.0
    override_cds = 0.0
/
&solver
    max_iter = {max_iter}
    convergence_epsilon = 1e-07
    broyden_history_size = 50
/
"""


def config_of(args):
    """The `config` object of the JSON line -- the same keys and values in both arms."""
    return {"workload": workload(), "points_per_gpu": args.points, "shells": SHELLS, "nghl": 1600, "eps": 1e-7,
            "broyden_history": 50, "max_iter": 300}


def parity_sample(all_strength, conv):
    """Outside the timed region: the strengths this run produced against the reference binary's own converged results
    for the same sweep points (tests/golden/Gd162_SKOP_<n>sh/prod_points.json or points.json, made by
    tests/golden/make_production.py).  Returns (max relative error on S and the cross-terms, points compared)."""
    import numpy as np
    for fn in ("prod_points.json", "sweep_points.json", "points.json"):
        path = os.path.join(case_dir(), fn)
        if not os.path.isfile(path):
            continue
        worst, n = 0.0, 0
        for pt in json.load(open(path))["points"].get("GT-K0", []):
            i = pt.get("sweep_index")
            if i is None or i >= len(all_strength) or not pt["conv"] or not conv[i]:
                continue
            g = complex(float(pt["rows"]["Strength"][0]), float(pt["rows"]["Strength"][1]))
            worst = max(worst, abs(all_strength[i, 0] - g) / abs(g))
            n += 1
        if n:
            return worst, n
    return None, 0


def circle_contour(npts, emin=0.0, emax=10.0):
    """CIRCLE contour of pynfam (strength/contour.py:212-283): omega_k = r0 + r exp(i theta_k), theta on
    Gauss-Legendre nodes over [pi, 3pi].  All npts nodes are returned (no symmetry shortcut)."""
    import numpy as np
    x, _ = np.polynomial.legendre.leggauss(npts)
    theta = np.pi + (x + 1.0) * np.pi
    r0, r = 0.5 * (emin + emax), 0.5 * (emax - emin)
    return r0 + r * np.exp(1j * theta)


NODES_PER_CIRCLE = 64   # pynfam's default contour has 60 nodes; a single circle with hundreds of nodes would put dozens
                        # of points within 1e-3 MeV of the real axis (on the poles of the response)


def sweep_contour(npts):
    """Synthetic contour sweep of `npts` points: CIRCLE contours of 64 Gauss-Legendre nodes on [0, 10 + 0.25 g] MeV,
    g = 0, 1, ... (the last one with the remaining nodes).  --points 32 is one 32-node circle on [0, 10]."""
    import numpy as np
    out, g = [], 0
    while len(out) < npts:
        n = min(NODES_PER_CIRCLE, npts - len(out))
        out.extend(circle_contour(n, 0.0, 10.0 + 0.25 * g))
        g += 1
    return np.array(out)


def stage(wd, omega, max_iter):
    os.makedirs(wd, exist_ok=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(os.path.join(case_dir(), f), wd)
    with open(os.path.join(wd, "GT-K0.in"), "w") as f:
        f.write(FAM_NML.format(re=repr(float(omega.real)), im=repr(float(omega.imag)), max_iter=max_iter))


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


ITER_RE = re.compile(r"^#\s+\d+[LB]\s+\S+\s+\S+\s+\S+\s+(\S+)\s*$", re.M)


def reference_farm(args, nproc):
    """The way pynfam itself uses the reference (README: MPI over serial executables): `nproc` single-threaded
    pnfam_main.x processes side by side, one omega point each, args.ref_iters iterations.  Returns (aggregate
    iterations/s from the processes' own per-iteration timers, iterations, wall seconds, mean setup seconds)."""
    from oracle import refrun
    oms = sweep_contour(max(args.points, nproc))
    res = [None] * nproc

    def work(k):
        wd = tempfile.mkdtemp()
        stage(wd, oms[(k * len(oms)) // nproc], args.ref_iters)
        dat, wall, out = refrun.run_pnfam(wd, "GT-K0.in", threads=1)
        res[k] = (wall, [float(m.group(1)) for m in ITER_RE.finditer(out)])
        shutil.rmtree(wd, ignore_errors=True)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(nproc)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    ok = [r for r in res if r and r[1]]
    if not ok:
        return None
    rate = sum(len(t) / sum(t) for _, t in ok)
    return rate, sum(len(t) for _, t in ok), wall, sum(w - sum(t) for w, t in ok) / len(ok)


def run_reference(args, rank):
    """Reference arm: the unmodified prebuilt pnfam_main.x on the host cores (kind 'reference')."""
    if rank != 0:
        return
    from oracle import refrun
    cores = os.cpu_count()
    if not refrun.ensure_built():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/pnfam_main.x or the OpenBLAS wheel is missing"}))
        return
    om = sweep_contour(args.points)[min(args.points, NODES_PER_CIRCLE) // 3]
    wd = tempfile.mkdtemp()
    stage(wd, om, args.ref_iters)
    per_iter = []
    setup = []
    # bounded: every repetition relaunches the binary (9 s of HFB reconstruction each at 16 shells), so at most one
    # warm-up and two timed launches of the threaded mode, one warm-up and four timed rounds of the farm mode
    n_warm, n_thr, n_farm = min(args.warmup, 1), min(args.steps, 2), min(args.steps, 4)
    for step in range(n_warm + n_thr):
        dat, wall, out = refrun.run_pnfam(wd, "GT-K0.in", threads=cores)
        times = [float(m.group(1)) for m in ITER_RE.finditer(out)]
        if not times:
            print(json.dumps({"impl": "reference", "unavailable": "reference run produced no iteration table"}))
            return
        if step >= n_warm:
            per_iter += times
            setup.append(wall - sum(times))
    ips_threaded = len(per_iter) / sum(per_iter)
    # second mode: one single-threaded process per core (how pynfam farms the executable); the better mode is reported
    farm = [reference_farm(args, cores) for _ in range(n_warm + n_farm)][n_warm:]
    farm = [f for f in farm if f]
    ips_farm = sum(f[0] for f in farm) / len(farm) if farm else 0.0
    ips = max(ips_threaded, ips_farm)
    mode = ("%d single-threaded processes side by side, 1 omega point x %d iterations each (setup %.1f s/process excluded)"
            % (cores, args.ref_iters, farm[0][3])) if ips_farm >= ips_threaded and farm else \
           ("1 process, OMP=OPENBLAS threads=%d, 1 omega point x %d iterations per step (setup %.1f s/launch excluded)"
            % (cores, args.ref_iters, sum(setup) / len(setup)))
    line = {
        "impl": "reference", "metric": "FAM iterations/s (omega-points/s in omega_points_per_s)", "value": ips,
        "unit": "iterations/s", "n_gpus": args.gpus,
        # one step of this arm = one timed round of the reported mode (the rounds are bounded: every one relaunches the
        # binary and repeats its HFB reconstruction); steps x ms_per_step is the wall time of those rounds
        "steps": (n_farm if ips_farm >= ips_threaded and farm else n_thr), "warmup": n_warm,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * (sum(f[2] for f in farm) / len(farm) if ips_farm >= ips_threaded and farm else sum(per_iter) / max(1, n_thr) + sum(setup) / max(1, len(setup))),
        "timed_rounds": {"threaded": n_thr, "farm": n_farm},
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args),
        "omega_points_per_s": ips / args.assumed_iters_per_point,
        "cpu_baseline": {"value": ips, "unit": "iterations/s", "cores": cores, "kind": "reference",
                         "sample": mode + "; per-iteration times from pnfam_main.x's own timer",
                         "threaded_iterations_per_s": ips_threaded, "farm_iterations_per_s": ips_farm},
        "e2e": {"value": ips, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=128,
                    help="omega points per GPU (weak scaling): 128 -> 1024 at 8 GPUs, the span BASELINE.json configs[4] names")
    ap.add_argument("--shells", type=int, default=16, choices=[12, 16, 20, 24], help="HO shells of the Gd162 basis (fixture)")
    ap.add_argument("--slots", type=int, default=0, help="omega points iterated side by side per GPU (0 = the library's default)")
    ap.add_argument("--ref-iters", type=int, default=4)
    ap.add_argument("--assumed-iters-per-point", type=float, default=25.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-contour", action="store_true",
                    help="BASELINE.json configs[3], the north-star target run: all 14 allowed + first-forbidden (operator, K) x "
                         "the computed half of pynfam's 60-node CIRCLE contour, cross-terms on, sharded over the GPUs; one "
                         "step = the whole nucleus from the two HFB files to OP.out / OP.out.ctr (use --shells 20)")
    args = ap.parse_args()
    global SHELLS
    SHELLS = args.shells
    if args.full_contour and args.impl == "b200":
        run_full_contour(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from pynfam_b200 import gpu, host, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the FAM iteration has no CPU fallback")
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there from C code, so file descriptor 1 points
        # to stderr until rank 0 prints the result
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- this rank's shard of the contour (independent omega points: no data-path collective) --------
    npts = args.points * world
    omegas = sweep_contour(npts)
    my_idx = shard.partition(omegas, world)[rank]
    mine = omegas[my_idx]
    wd = tempfile.mkdtemp()
    stage(wd, mine[0], 300)
    t0 = time.time()
    prob = host.Problem(wd, "GT-K0.in")
    setup_s = time.time() - t0
    ctx = gpu.Context(prob, device=local)
    nghl, nxy, dqp = prob.iscalar("nghl"), prob.iscalar("nxy"), prob.iscalar("dqp")
    db, r2c = prob.i32("db"), prob.i32("f_ir2c")
    t3 = sum(int(db[i]) ** 2 * int(db[j - 1]) + int(db[i]) * int(db[j - 1]) ** 2 for i, j in enumerate(r2c) if j > 0)
    f_iter = 88.0 * nghl * nxy + 32 * 2.0 * t3 + 650.0 * nghl * dqp

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    dmma_peak = gpu.dmma_peak_tflops(local)
    for _ in range(args.warmup):
        ctx.solve(prob, omegas=mine, slots=args.slots)
    barrier()
    sampler = ClockSampler(local)
    if not os.environ.get("PNFAM_BENCH_NO_SAMPLER"):
        sampler.start()
    dev_s, iters, launches = 0.0, 0, 0
    dens_s = proj_s = dens_fl = proj_fl = 0.0
    dens_n = proj_n = 0
    last = None
    for _ in range(args.steps):
        r = ctx.solve(prob, omegas=mine, slots=args.slots)
        st = r["stats"]
        dev_s += st["seconds_device"]; iters += st["iterations"]; launches += st["kernel_launches"]
        dens_s += st["seconds_density"]; proj_s += st["seconds_projection"]
        dens_fl += st["flops_density"]; proj_fl += st["flops_projection"]
        dens_n += st["launches_density"]; proj_n += st["launches_projection"]
        last = r
    barrier()
    # ---- e2e: host buffers -> context + operator upload -> solve -> strengths back, wall clock ------
    e2e_s, e2e_iters, h2d, d2h = 0.0, 0, 0, 0
    # one untimed warm-up of the e2e path: the first context created next to a live one carves its small arrays out of
    # the pooled blocks the previous solve returned, and the 10 GB Broyden history then needs fresh device memory once
    # (the untimed warm-up steps of the contract apply to this path as well: the memory pools of the library settle after
    # two or three create / solve / destroy cycles)
    for _ in range(max(1, min(args.warmup, 3))):
        cw = gpu.Context(prob, device=local)
        cw.solve(prob, omegas=mine, slots=args.slots)
        del cw
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        c2 = gpu.Context(prob, device=local)
        r2 = c2.solve(prob, omegas=mine, slots=args.slots)
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
        if rank == 0:
            print("e2e step %.1f ms (C ABI solve %.1f ms, device loop %.1f ms)" % (1e3 * (time.perf_counter() - t0),
                  1e3 * r2["stats"]["seconds_total"], 1e3 * r2["stats"]["seconds_device"]), file=sys.stderr)
        e2e_iters += r2["stats"]["iterations"]
        h2d += c2.h2d_bytes + r2["stats"]["h2d_bytes"]
        d2h += r2["stats"]["d2h_bytes"]
        del c2
    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join(timeout=2)

    # ---- aggregate over ranks: time = max over ranks, work = sum -------------------------------------
    vals = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    sums = torch.tensor([iters, e2e_iters, launches, len(mine) * args.steps, h2d, d2h], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    # the only exchange the path has: gather the strengths of all points (NCCL all_gather over NVLink)
    all_strength = shard.gather_strengths(my_idx, last["strength"], npts, dist=dist if world > 1 else None, device="cuda")
    assert np.isfinite(all_strength).all() and np.abs(all_strength[:, 0]).min() > 0
    conv_all = shard.gather_strengths(my_idx, last["conv"].astype(np.float64).reshape(-1, 1), npts,
                                      dist=dist if world > 1 else None, device="cuda")[:, 0].real > 0.5
    iters_conv = torch.tensor([float(last["iters"][last["conv"] > 0].sum()) * args.steps, float((last["conv"] > 0).sum()) * args.steps],
                              dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(iters_conv, op=dist.ReduceOp.SUM)
    iters_conv = iters_conv.cpu().numpy()
    vals, sums = vals.cpu().numpy(), sums.cpu().numpy()
    if rank == 0:
        value = sums[0] / vals[0]
        pts_per_s = iters_conv[1] / vals[0]          # converged points only
        par_rel, par_n = parity_sample(all_strength, conv_all)
        # roofline of the dominant kernels (tensor-bound, FP64 DMMA): algorithmic flops / CUDA-event time
        top = ("density", dens_fl, dens_s, dens_n) if dens_s >= proj_s else ("projection", proj_fl, proj_s, proj_n)
        ach = top[1] / top[2] / 1e12 if top[2] > 0 else 0.0
        cap = ncu_capture(top[0], SHELLS)
        line = {
            "metric": "FAM iterations/s (omega-points/s in omega_points_per_s)", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * vals[0] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args),
            "problem": {"basis_states": dqp, "nxy": nxy, "batch_slots": last["stats"]["batch_slots"],
                        "lock_steps_per_solve": last["stats"]["lock_steps"],
                        "l2": "per-step working set (Broyden history 2*50*4*nxy*8 B per slot = %.1f GB) >> 126 MB L2; "
                              "no flush needed" % (last["stats"]["batch_slots"] * 2 * 50 * 4 * nxy * 8 / 1e9)},
            "omega_points_per_s": pts_per_s,
            "converged_fraction": float(conv_all.mean()),
            "converged_iterations_fraction": float(iters_conv[0] / max(1.0, sums[0])),
            "parity_max_rel": par_rel, "parity_points": par_n,
            "parity_note": "strengths of this run vs the reference binary's converged results for the same sweep points "
                           "(tests/golden/.../prod_points.json), checked outside the timed region",
            "iterations_per_point": sums[0] / sums[3],
            "iterations_per_s_per_gpu": value / world,
            "host_setup_s": setup_s,
            "e2e": {"value": sums[1] / vals[1], "unit": "iterations/s", "h2d_bytes_per_step": int(sums[4] / args.steps / world),
                    "d2h_bytes_per_step": int(sums[5] / args.steps / world), "omega_points_per_s": sums[3] / vals[1]},
            "gpu_launches": int(sums[2]),
            "roofline": {"bound": "tensor", "kernel": top[0], "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": ach / dmma_peak if dmma_peak else None, "traffic": (cap or {}).get("traffic"),
                         "achieved_note": "ALGORITHMIC FP64 flops of the reference's formulation (SURVEY 8d: density 20, projection 24 "
                                          "x nghl x nxy per point and pass) / CUDA-event time of the kernel family; > peak because "
                                          "the factorised kernels execute far fewer operations (see executed)",
                         "executed": (cap or {}).get("executed"),
                         "peak_source": "FP64 DMMA (mma.sync m8n8k4) = FP64 FMA rate, measured live by pnfam_b200_dmma_peak; "
                                        "MEASURED_PEAKS.json carries no FP64 figure",
                         "whole_iteration": {"F_iter": f_iter, "achieved": f_iter * value / world / 1e12,
                                             "frac": f_iter * value / world / 1e12 / dmma_peak if dmma_peak else None,
                                             "note": "SURVEY 8(d): F_iter = 88*Ng*nxy + 32*2*T3 + 650*Ng*N algorithmic FP64 flop "
                                                     "per FAM iteration, x iterations/s per GPU, / measured DMMA peak"},
                         "density": {"tflops": dens_fl / dens_s / 1e12 if dens_s else None, "share_of_step": dens_s / dev_s,
                                     "ms_per_launch": 1e3 * dens_s / max(1, dens_n)},
                         "projection": {"tflops": proj_fl / proj_s / 1e12 if proj_s else None, "share_of_step": proj_s / dev_s,
                                        "ms_per_launch": 1e3 * proj_s / max(1, proj_n)}},
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


FULL_OPERATORS = [("F-", 0), ("GT-", 0), ("GT-", 1), ("RS0-", 0), ("PS0-", 0), ("R-", 0), ("P-", 0), ("RS1-", 0), ("R-", 1), ("P-", 1),
                  ("RS1-", 1), ("RS2-", 0), ("RS2-", 1), ("RS2-", 2)]   # operators sharing cross-term fields are neighbours


def run_full_contour(args):
    """The north-star target run through the public driver (pynfam_b200.strength.run_contours_sharded): one step = the
    full beta-decay contour of 162Gd from hfbtho_NAMELIST.dat + hfbtho_output.hel to the strength files, wall clock
    between barriers, max over ranks.  Timed steps start in a FRESH run directory (no set-up cache); `warm_cache_s` is one
    more run in a directory that already holds the cache.  The strengths are compared, outside the timed region, with the
    reference binary's converged results at the contour points of tests/golden/Gd162_SKOP_<n>sh/prod_points.json."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pynfam_b200.strength import famContour, run_contours_sharded
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # torchrun pins OMP_NUM_THREADS=1; the host set-up is OpenMP code: an even share of the cores per rank (rank 0 takes
    # all of them while it reconstructs the HFB solution for everybody)
    from pynfam_b200 import host
    host.set_threads(max(1, (os.cpu_count() or 1) // world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the FAM iteration has no CPU fallback")
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.all_reduce(torch.zeros(1, device="cuda"))           # NCCL communicator set-up is not part of the workload
    contour = famContour("CIRCLE", {"energy_min": 0.0, "energy_max": 10.0, "nr_points": 60})

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_run(wd):
        if rank == 0:
            stage(wd, 1.0 + 1.0j, 300)
        sync()
        dest = os.path.join(wd, "out_%d" % rank)
        os.makedirs(dest, exist_ok=True)
        t0 = time.perf_counter()
        fss = run_contours_sharded(wd, "GT-K0.in", FULL_OPERATORS, contour, dest=dest, dist=dist if world > 1 else None, device=local)
        sync()
        return time.perf_counter() - t0, fss, dict(run_contours_sharded.last_timing), dest

    def shared_dir(tag):
        # every rank works in the same run directory (rank 0 names it)
        name = [tempfile.mkdtemp(prefix="fullctr_%s_" % tag)] if rank == 0 else [None]
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        return name[0]

    wd = shared_dir("warm")
    for _ in range(max(1, min(args.warmup, 2))):
        one_run(wd)
    sampler = ClockSampler(local)
    sampler.start()
    walls, parts = [], []
    fss = dest = None
    for k in range(args.steps):
        w, fss, tm, dest = one_run(shared_dir("step%d" % k))
        walls.append(w)
        parts.append(tm)
    sampler.stop_flag = True
    warm_s, _, warm_tm, _ = one_run(wd)
    t = torch.tensor([sum(walls), warm_s] + [sum(p[k] for p in parts) for k in ("host_setup_s", "solve_s", "gather_and_write_s")],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().numpy()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    if rank == 0:
        iters = int(sum(int(np.sum(f.iters)) for f in fss))
        npts = len(FULL_OPERATORS) * contour.nr_compute
        # parity: the reference binary's converged strengths at one contour point of every operator
        worst, npar = None, 0
        path = os.path.join(case_dir(), "prod_points.json")
        if os.path.isfile(path):
            gold = json.load(open(path))["points"]
            byname = {f.opname: f for f in fss}
            worst = 0.0
            for name, lst in gold.items():
                for pt in lst:
                    f, i = byname.get(name), pt.get("contour_index")
                    if f is None or i is None or i >= contour.nr_compute:
                        continue
                    g = complex(float(pt["rows"]["Strength"][0]), float(pt["rows"]["Strength"][1]))
                    got = complex(f.str_df["Re(Strength)"].values[i], f.str_df["Im(Strength)"].values[i])
                    worst = max(worst, abs(got - g) / abs(g))
                    npar += 1
        sec = t[0] / args.steps
        line = {
            "metric": "FAM omega-points/s, full beta-decay contour of one nucleus (iterations/s in iterations_per_s)",
            "value": npts / sec, "unit": "omega-points/s", "n_gpus": world, "steps": args.steps, "warmup": max(1, min(args.warmup, 2)),
            "ms_per_step": 1e3 * sec, "seconds": sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Gd162 SkO' %d shells 40x40 grid, full beta-decay contour: %d (operator, K) x %d computed points of "
                                   "the 60-node CIRCLE contour on [0, 10] MeV, cross-terms on" % (SHELLS, len(FULL_OPERATORS), contour.nr_compute),
                       "shells": SHELLS, "nghl": 1600, "eps": 1e-7, "broyden_history": 50, "max_iter": 300},
            "omega_points": npts, "iterations": iters, "iterations_per_s": iters / sec,
            "all_converged": bool(all(f.meta["Conv"] == "Yes" for f in fss)),
            "max_over_ranks_s": {"host_setup": t[2] / args.steps, "solves": t[3] / args.steps, "gather_and_write": t[4] / args.steps},
            "warm_cache_s": float(t[1]),
            "e2e": {"value": npts / sec, "unit": "omega-points/s", "note": "the step IS end to end: HFB files -> set-up on the host -> "
                    "context + operator uploads -> solves -> strengths back -> one all_reduce -> OP.out / OP.out.ctr on disk"},
            "parity_max_rel": worst, "parity_points": npar,
            "host_threads": os.cpu_count(), "clocks": sampler.summary(),
            "files": sorted(os.listdir(dest))[:4] + ["..."],
        }
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


NCU_FAMILY = {"density": ["sf2_density_kernel<0, 2>", "sf2_density_kernel<0, 1>", "sf2_kappa_density4_kernel", "sf2_pack_kernel"],
              "projection": ["sf2_kappa_kernel<0>", "sf2_kappa_kernel<1>", "sf2_radial_kernel<0>", "sf2_radial_kernel<1>"]}


def ncu_capture(family, shells):
    """What the committed `ncu --set full` capture of this workload (profiles/ncu_kernels.json, written by
    scripts/ncu_kernels.py) says about the kernels of a family, per launch of its dominant kernel: DRAM bytes read + written
    (`traffic`), FP64 flops actually EXECUTED and the unit that limits it.  The capture is taken at 16 shells with 8 omega
    points per launch; nothing is reported for another basis size.  None if no capture is committed."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_kernels.json")))
        if shells != d.get("shells", 16):
            return None
        ks = {k["kernel"]: k for k in d["kernels"]}
        top = ks[[n for n in NCU_FAMILY[family] if n in ks][0]]       # the rho density runs as <0, 2> (two il per CTA) or <0, 1>
        launches = {"density": 1.0, "projection": 2.0}      # radial kernels launch once per pass, the others once for both
        fam_flops = sum(ks[n]["fp64_thread_flops_per_launch"] * (launches[family] if "radial" in n else 1.0) for n in NCU_FAMILY[family] if n in ks)
        units = {"l1": top["l1_throughput_pct"], "fp64_pipe": top["fp64_pipe_pct"], "l2": top["l2_throughput_pct"], "dram": top["dram_throughput_pct"],
                 "issue": top["issue_active_pct"]}
        lim = max(units, key=units.get)
        return {"traffic": {"bytes_per_launch": top["dram_bytes_per_launch"], "kernel": top["kernel"], "omega_points_per_launch": d.get("points", 8),
                            "source": d["source"]},
                "executed": {"kernel": top["kernel"], "fp64_flops_per_launch": top["fp64_thread_flops_per_launch"],
                             "fp64_flops_family_per_point_iteration": fam_flops / d.get("points", 8),
                             "tflops_under_ncu": top["fp64_thread_TFLOPs"], "limiter": lim, "limiter_pct_of_peak": units[lim],
                             "fp64_pipe_pct": top["fp64_pipe_pct"], "l1_throughput_pct": top["l1_throughput_pct"],
                             "note": "the fully factorised kernels execute ~14x fewer FP64 operations than the reference's GEMM "
                                     "formulation counts (roofline.achieved uses that count, SURVEY 8d); they are FP64-FMA gather "
                                     "kernels bound by L1/shared-memory throughput, not by the FP64 pipe"}}
    except Exception:
        return None


def cpu_baseline(args):
    """Rank 0, N=1: the reference binary (kind 'reference') or, if it did not travel, the numpy oracle port, on a
    bounded sample of the same workload."""
    from oracle import refrun
    cores = os.cpu_count()
    om = sweep_contour(args.points)[min(args.points, NODES_PER_CIRCLE) // 3]
    wd = tempfile.mkdtemp()
    stage(wd, om, args.ref_iters)
    if refrun.ensure_built():
        dat, wall, out = refrun.run_pnfam(wd, "GT-K0.in", threads=cores)
        times = [float(m.group(1)) for m in ITER_RE.finditer(out)]
        if times:
            threaded = len(times) / sum(times)
            farm = reference_farm(args, cores)
            if farm and farm[0] > threaded:
                return {"value": farm[0], "unit": "iterations/s", "cores": cores, "kind": "reference",
                        "sample": "oracle/_ref/pnfam_main.x, %d single-threaded processes side by side, 1 omega point x %d "
                                  "iterations each; wall %.1f s incl. %.1f s setup per process" % (cores, args.ref_iters, farm[2], farm[3]),
                        "threaded_iterations_per_s": threaded}
            return {"value": threaded, "unit": "iterations/s", "cores": cores, "kind": "reference",
                    "sample": "oracle/_ref/pnfam_main.x, 1 omega point x %d iterations, %d threads; wall %.1f s incl. %.1f s setup"
                              % (len(times), cores, wall, wall - sum(times)),
                    "farm_iterations_per_s": farm[0] if farm else None}
    from oracle import fam_oracle as fo
    from pynfam_b200 import host
    p = host.Problem(wd, "GT-K0.in")
    s = fo.solver_from_problem(p)
    t0 = time.time()
    s.solve(2, 1e-7)
    return {"value": 2 / (time.time() - t0), "unit": "iterations/s", "cores": 1, "kind": "port",
            "sample": "numpy oracle, 1 omega point x 2 iterations"}


if __name__ == "__main__":
    main()
