"""Summarise an ncu --set full report: per-kernel key metrics + instruction mix (dev tool)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'lts__t_sector_hit_rate.pct']
w = csv.writer(sys.stdout)
w.writerow([k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '') for k in keep])
w.writerow([rows[1][h.index(k)] for k in keep])
seen = set()
for r in rows[2:]:
    w.writerow([r[h.index(k)][:60] for k in keep])
if len(sys.argv) > 2:
    for pat in sys.argv[2:]:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        rs = list(csv.reader(src.splitlines()))
        hh = rs[1]
        ia, ie, isamp = hh.index('Source'), hh.index('Instructions Executed'), hh.index('# Samples')
        data = [(int(r[ie] or 0), int(r[isamp] or 0), r[ia]) for r in rs[2:] if len(r) > ie]
        tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
        op, ops = collections.Counter(), collections.Counter()
        for e, s, sc in data:
            o = sc.split()[0] if not sc.startswith('@') else sc.split()[1]
            op[o] += e; ops[o] += s
        print("# instruction mix", rs[0][1][:70], "total", tot)
        for k, v in op.most_common(14):
            print("#   %-14s %12d %5.1f%%  samples %5.1f%%" % (k, v, 100 * v / tot, 100 * ops[k] / ts))
