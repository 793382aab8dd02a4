"""Per-launch summary of the LAST training iteration in an ncu launch list.
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file L.csv <cmd>
  python tools/ncu_launch_list.py L.csv [launches per iteration = 21]
ncu times are cold-cache and serialised: read the SHARES, not the absolute values."""
import csv
import sys

path = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 21
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
h = rows[0]
iid, ik, im, iu, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
launches = {}
for r in rows[1:]:
    k = r[ik]
    if not ("b200" in k or "apply_" in k or "dgrad_stats" in k or "eval_kernel" in k):
        continue
    d = launches.setdefault(int(r[iid]), {"name": k})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    v *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "msecond": 1e3, "ms": 1e3, "usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6}.get(u, 1.0)
    d[r[im]] = v
ids = sorted(launches)[-per:]
tot = dict(t=0.0, r=0.0, w=0.0)
print("# last %d launches of: %s" % (per, path))
for i in ids:
    d = launches[i]
    name = d["name"].split("(")[0].split("::")[-1][:28]
    t, rd, wr = d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)
    tot["t"] += t; tot["r"] += rd; tot["w"] += wr
    print("%-28s %8.1f us  read %8.1f MB  write %8.1f MB  tensor %5.1f%%" % (name, t, rd, wr, tp))
print("sum: %.1f us, DRAM read %.2f GB, write %.2f GB" % (tot["t"], tot["r"] / 1e3, tot["w"] / 1e3))
