"""Instruction histogram of the built library (evidence that the hot loops are Blackwell-native):

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt

Per kernel of bilinear_b200/lib/libbilinear_b200.so: counts of the tcgen05 / TMA / TMEM mnemonics
(UTCHMMA = tcgen05.mma kind::f16, .2CTA = cta_group::2; UTMALDG = cp.async.bulk.tensor; LDTM = tcgen05.ld;
UTCBAR = tcgen05.commit; USETMAXREG = setmaxnreg), of legacy tensor instructions (HMMA: must be 0) and of
local-memory traffic (STL / LDL = register spills and stack frames).  Needs cuobjdump (CUDA toolkit), no GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bilinear_b200", "lib", "libbilinear_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "USETMAXREG", "HMMA", "STL", "LDL", "REDG", "ATOMG",
        "LDG", "STG", "FFMA", "FFMA2", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            base = op.split(".")[0]
            if base in ("UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "USETMAXREG", "HMMA", "STL", "LDL", "REDG", "ATOMG",
                        "LDG", "STG", "FFMA", "FFMA2"):
                cur[base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
    print("%-62s %s" % ("kernel", " ".join("%8s" % k[-8:] for k in KEYS)))
    for name, c in per.items():
        print("%-62s %s" % (name[-62:], " ".join("%8d" % c[k] for k in KEYS)))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("%-62s %s" % ("ALL", " ".join("%8d" % tot[k] for k in KEYS)))
    assert tot["HMMA"] == 0, "legacy mma.sync instructions found"


if __name__ == "__main__":
    sys.exit(main())
