#!/usr/bin/env python
"""Instruction census of the shipped CUDA library per kernel (cuobjdump -sass): total instructions and the counts of the
mnemonics that tell which hardware paths a kernel uses (DMMA = FP64 tensor cores; LDGSTS / UBLKCP = asynchronous copies to
shared memory, SYNCS = mbarrier; UTMALDG / UTC*MMA would be TMA tensor copies / tcgen05, which have no FP64 type)."""
import collections, re, subprocess, sys
so = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, tot, cnt = None, collections.Counter(), collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and fn:
        tot[fn] += 1
        cnt[fn][m.group(2)] += 1
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("# cuobjdump -sass %s : cubins for %s" % (so, ", ".join(arch)))
keys = ["DMMA", "DFMA", "DMUL", "DADD", "LDGSTS", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "UTMALDG", "UTCQMMA", "UTCHMMA"]
for f in sorted(tot):
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
    print("%-60s total %6d  " % (name[:60], tot[f]) + " ".join("%s=%d" % (k, cnt[f][k]) for k in keys if cnt[f][k]))
allc = collections.Counter()
for f in cnt:
    allc.update(cnt[f])
print("# library total: " + " ".join("%s=%d" % (k, allc[k]) for k in keys))
