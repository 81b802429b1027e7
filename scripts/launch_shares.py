"""Per-kernel share of one bench step from an ncu launch list (gpu__time_duration.sum --csv) (dev tool).
usage: launch_shares.py launches.csv > profiles/<tag>_launch_shares.txt"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, start = r, i + 1
        break
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
t, n = collections.defaultdict(float), collections.Counter()
for r in rows[start:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[r[mu]]
    name = r[kn].split('(')[0]
    t[name] += v; n[name] += 1
tot = sum(t.values())
print("# per-kernel share of one bench step under ncu (gpu__time_duration.sum, --clock-control none; serialised, cold cache)")
print("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline")
print("# total %.1f ms over %d launches" % (tot, sum(n.values())))
print("%-62s %8s %10s %7s" % ("kernel", "launches", "ms", "share"))
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print("%-62s %8d %10.2f %6.1f%%" % (k, n[k], v, 100 * v / tot))
