"""Development diagnostic (run on a GPU box): stage-by-stage comparison of the CUDA path with the oracle."""
import json, os, re, shutil, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from pynfam_b200 import host, gpu
from oracle import fam_oracle as fo

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def stage(case, op, idx, wd):
    gd = os.path.join(ROOT, "tests", "golden", case)
    d = json.load(open(gd + "/points.json"))["points"]
    os.makedirs(wd, exist_ok=True)
    for f in ("hfbtho_NAMELIST.dat", "hfbtho_output.hel"):
        shutil.copy(gd + "/" + f, wd)
    pt = d[op][idx]
    nml = re.sub(r"two_body_current_mode\s*=\s*114", "two_body_current_mode = 0", pt["namelist"])
    open(wd + "/x.in", "w").write(nml)
    return pt


def main():
    print("DMMA peak TFLOP/s:", gpu.dmma_peak_tflops())
    wd = tempfile.mkdtemp()
    for case, op, idx in (("S40_SKOP_6sh", "GT-K0", 10), ("S40_GT_All", "RS0-K0", 5), ("Gd162_GT_open_6sh", "GT-K1", 40)):
        pt = stage(case, op, idx, wd)
        p = host.Problem(wd, "x.in")
        model = fo.model_from_problem(p)
        ctx = gpu.Context(p)
        # --- calc_hamiltonian on the oracle's first-iteration densities
        s = fo.solver_from_problem(p, model)
        s.iterate(0)
        s.iterate(1)   # dRsp now holds non-trivial densities, dHsp the oracle's fields
        order = [(11, 're'), (11, 'im'), (12, 're'), (12, 'im'), (22, 're'), (22, 'im'), (21, 're'), (21, 'im')]
        ins = [(s.dRsp_re if c == 're' else s.dRsp_im).m[q].copy() for q, c in order]
        ref = [(s.dHsp_re if c == 're' else s.dHsp_im).m[q].copy() for q, c in order]
        outs = [r.copy() for r in ref]
        for o in outs:
            o.elem = np.zeros_like(o.elem)
        ctx.calc_hamiltonian(ins, outs)
        for (q, c), o, r in zip(order, outs, ref):
            sc = np.abs(r.elem).max() + 1e-300
            print(f"  {case} {op} calc_hamiltonian out m{q} {c}: max abs err {np.abs(o.elem - r.elem).max():.3e} (scale {sc:.3e})")
        # --- trajectories
        for mi in (1, 2, 3, 5):
            so = fo.solver_from_problem(p, model)
            it, si, st = so.solve(mi, 1e-7)
            r = ctx.solve(p, max_iter=mi)
            print(f"  max_iter={mi}: oracle S={st[0]:.15g} si={si:.6e} | gpu S={r['strength'][0,0]:.15g} si={r['si'][0]:.6e} rel {abs(r['strength'][0,0]-st[0])/abs(st[0]):.2e}")
        t0 = time.time()
        r = ctx.solve(p, want_trace=True)
        g = complex(float(pt["rows"]["Strength"][0]), float(pt["rows"]["Strength"][1]))
        print(f"  full: gpu iters {r['iters'][0]} (gold {pt['iters']}) S={r['strength'][0,0]:.17g} gold={g:.17g} rel {abs(r['strength'][0,0]-g)/abs(g):.2e}  wall {time.time()-t0:.3f}s")
        for k, l in enumerate(r["labels"][1:], 1):
            if l in pt["rows"]:
                gg = complex(float(pt["rows"][l][0]), float(pt["rows"][l][1]))
                print(f"     {l}: rel {abs(r['strength'][0,k]-gg)/abs(gg):.2e}")
        print("  stats", r["stats"])
        # batch of 8 omegas
        oms = [complex(p.scalar("real_eqrpa") + 0.3 * k, p.scalar("imag_eqrpa")) for k in range(8)]
        t0 = time.time()
        rb = ctx.solve(p, omegas=oms)
        print(f"  batch8: iters {rb['iters']} wall {time.time()-t0:.3f}s S0 rel vs single {abs(rb['strength'][0,0]-r['strength'][0,0])/abs(r['strength'][0,0]):.2e}")
        print("  stats", rb["stats"])


if __name__ == "__main__":
    main()
