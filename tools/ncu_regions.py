"""Stall samples of one launch of an ncu report, summed per SASS function region (RET / EXIT delimited) and the top
instructions of each region.  python tools/ncu_regions.py report.ncu-rep [launch index] [top n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
k0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h, data = rows[k0], rows[k0 + 1:]
iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
print(rows[0][:2], "total samples", sum(int(r[iS]) for r in data))
start = 0
for b in [k for k, r in enumerate(data) if "RET." in r[iSrc] or r[iSrc].strip().startswith("EXIT")] + [len(data) - 1]:
    seg = range(start, b + 1)
    s = sum(int(data[k][iS]) for k in seg)
    if s >= 40:
        agg = sorted(((sum(int(data[k][i]) for k in seg), h[i]) for i in stall), reverse=True)[:5]
        print("region %d..%d: %d samples, %d warp instructions; %s" % (start, b, s, sum(int(data[k][iEx]) for k in seg),
                                                                       ", ".join("%s %d" % (n, v) for v, n in agg)))
        for k in sorted(seg, key=lambda k: -int(data[k][iS]))[:topn]:
            r = data[k]
            st = sorted(((int(r[i]), h[i]) for i in stall), reverse=True)[0]
            print("    %6d  %5s samples  %8s exec  %-70s %s" % (k, r[iS], r[iEx], r[iSrc].strip()[:70], st[1]))
    start = b + 1
