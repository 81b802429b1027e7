#!/usr/bin/env python
"""Per-kernel table from the raw CSV of one `ncu --set full` capture (scripts/gpu_evidence.sh exports it on the GPU box):
time, launch shape, pipe / memory utilisation, DRAM bytes and EXECUTED FP64 flops per launch (2 x DFMA + DADD + DMUL thread
instructions; DMMA m8n8k4 = 512 flop per warp instruction).  Writes <out>.csv and <out>.json (bench.py reads the JSON for
roofline.traffic and the executed-flop figures).
Usage: python scripts/ncu_kernels.py <raw.csv> <out prefix> "<command>" [shells] [omega points per launch] """
import csv, json, sys

raw, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
shells = int(sys.argv[4]) if len(sys.argv) > 4 else 16
points = int(sys.argv[5]) if len(sys.argv) > 5 else 8
rows = list(csv.reader(open(raw)))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3}


def val(r, k, default=0.0):
    if k not in ix or r[ix[k]] in ("", "n/a"):
        return default
    return float(r[ix[k]].replace(",", "")) * scale.get(units[ix[k]], 1.0)


groups = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("pnfam::", "")
    groups.setdefault((name, r[ix["launch__grid_size"]]), []).append(r)
table = []
for (name, grid), rs in groups.items():
    n = len(rs)
    avg = lambda k: sum(val(r, k) for r in rs) / n
    cyc = avg("smsp__cycles_elapsed.max") or avg("sm__cycles_elapsed.max")
    nsmsp = 148 * 4
    thread_ops = {k: avg("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k) * cyc for k in ("dfma", "dadd", "dmul")}
    dmma_pct = avg("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active")
    rec = {
        "kernel": name, "grid": int(float(grid)), "launches_averaged": n, "us_per_launch": 1e6 * avg("gpu__time_duration.sum"),
        "registers": avg("launch__registers_per_thread"), "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"), "dmma_pipe_pct": dmma_pct,
        "fp64_pipe_pct": avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "l1_throughput_pct": avg("l1tex__throughput.avg.pct_of_peak_sustained_active"),
        "l2_throughput_pct": avg("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": avg("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_bytes_per_launch": avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
        "dram_read": avg("dram__bytes_read.sum"), "dram_write": avg("dram__bytes_write.sum"),
        "l2_hit_pct": avg("lts__t_sector_hit_rate.pct"),
        "fp64_thread_flops_per_launch": 2 * thread_ops["dfma"] + thread_ops["dadd"] + thread_ops["dmul"],
        "shared_bank_conflicts": avg("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    }
    t = rec["us_per_launch"] * 1e-6
    rec["dram_GBps"] = rec["dram_bytes_per_launch"] / t / 1e9 if t else 0.0
    rec["fp64_thread_TFLOPs"] = rec["fp64_thread_flops_per_launch"] / t / 1e12 if t else 0.0
    table.append(rec)
table.sort(key=lambda r: -r["us_per_launch"] * r["launches_averaged"])
keys = list(table[0].keys())
with open(out + ".csv", "w") as f:
    w = csv.writer(f)
    w.writerow(keys)
    for r in table:
        w.writerow([("%.4g" % r[k]) if isinstance(r[k], float) else r[k] for k in keys])
json.dump({"source": "ncu --set full --clock-control none, %s" % cmd, "shells": shells, "points": points, "kernels": table}, open(out + ".json", "w"), indent=1)
for r in table:
    print("%-28s grid %6d %8.1f us  warps %4.1f%% issue %4.1f%% dmma %4.1f%% fp64 %4.1f%% L1 %4.1f%% L2 %4.1f%% dram %4.1f%% (%6.0f GB/s, %6.1f MB)  fp64 thread flops %.3g (%.2f TF/s)" % (
        r["kernel"][:28], r["grid"], r["us_per_launch"], r["warps_active_pct"], r["issue_active_pct"], r["dmma_pipe_pct"], r["fp64_pipe_pct"],
        r["l1_throughput_pct"], r["l2_throughput_pct"], r["dram_throughput_pct"], r["dram_GBps"], r["dram_bytes_per_launch"] / 1e6,
        r["fp64_thread_flops_per_launch"], r["fp64_thread_TFLOPs"]))
