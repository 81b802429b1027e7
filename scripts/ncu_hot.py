"""Hot spots of one kernel in an ncu report: SASS instructions with the most stall samples (dev tool).
usage: ncu_hot.py report.ncu-rep kernel_regex [ntop]"""
import csv, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
# kernel_regex, or "#N" = the N-th profiled launch of the report (1-based)
sel = ["--kernel-id", ":::" + pat[1:]] if pat.startswith("#") else ["--kernel-name", "regex:" + pat, "--launch-count", "1"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rs = list(csv.reader(src.splitlines()))
print(rs[0][1][:100])
hh = rs[1]
ia, ie, isamp = hh.index('Source'), hh.index('Instructions Executed'), hh.index('# Samples')
data = [(i, int(r[isamp] or 0), int(r[ie] or 0), r[ia].strip()) for i, r in enumerate(rs[2:]) if len(r) > ie]
ts, te = sum(d[1] for d in data), sum(d[2] for d in data)
print("instructions", len(data), "samples", ts, "executed", te)
op = collections.Counter(); ops = collections.Counter()
for i, s, e, sc in data:
    o = sc.split()[0] if not sc.startswith('@') else sc.split()[1]
    op[o] += e; ops[o] += s
for k, v in ops.most_common(12):
    print("  %-16s samples %5.1f%%  executed %5.1f%%" % (k, 100 * v / ts, 100 * op[k] / te))
print("top instructions by samples:")
for i, s, e, sc in sorted(data, key=lambda d: -d[1])[:ntop]:
    print("  #%-5d %5.1f%%  exec %9d  %s" % (i, 100 * s / ts, e, sc[:90]))
