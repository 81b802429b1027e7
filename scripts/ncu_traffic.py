"""Extract the DRAM traffic per launch of the two dominant kernels from an `ncu --set full` report of the bench
workload and write profiles/ncu_traffic.json (read by bench.py for roofline.traffic).
Usage: python scripts/ncu_traffic.py <report.ncu-rep> "<command the report was captured with>" """
import csv, io, json, os, subprocess, sys

rep, cmd = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for fam, key in (("projection", "projection_kernel<0"), ("density", "density_kernel<0")):
    sel = [r for r in rows[2:] if key in r[ik]]
    if not sel:
        continue
    rd = sum(float(r[ir]) * scale[units[ir]] for r in sel) / len(sel)
    wr = sum(float(r[iw]) * scale[units[iw]] for r in sel) / len(sel)
    res[fam] = {"kernel": sel[0][ik].split("(")[0], "bytes_per_launch": rd + wr, "read": rd, "write": wr, "launches_averaged": len(sel),
                "ms_per_launch_under_ncu": sum(float(r[it]) for r in sel) / len(sel)}
res["source"] = "ncu --set full --clock-control none, %s; %s" % (cmd, os.path.basename(rep))
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
