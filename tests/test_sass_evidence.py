"""The built library must be Blackwell-native: tcgen05 MMAs (incl. cta_group::2), TMA loads and TMEM reads in the persistent
kernels, TMA loads in the evaluation kernel, packed fp32 pairs in the P-MPJPE math -- and no legacy mma.sync (HMMA).
Runs on CPU (cuobjdump disassembles the sm_100a cubins); skipped when the toolkit or the library is absent."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bilinear_b200", "lib", "libbilinear_b200.so")


@pytest.fixture(scope="module")
def histogram():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_histogram.py")], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    cols = lines[0].split()[1:]
    table = {}
    for l in lines[1:]:
        parts = l.rsplit(None, len(cols))
        if len(parts) == len(cols) + 1:
            table[parts[0].strip()] = dict(zip(cols, (int(x) for x in parts[1:])))
    return table


def _row(table, needle):
    rows = [v for k, v in table.items() if needle in k]
    assert rows, "no kernel matching %r in %s" % (needle, sorted(table))
    return rows[0]


def test_persistent_kernels_use_tcgen05_tma_tmem(histogram):
    solo = _row(histogram, "solo_variant::b200_mega_kernel")
    pair = _row(histogram, "pair_variant::b200_mega_kernel")
    for r in (solo, pair):
        assert r["UTCHMMA"] > 0 and r["UTMALDG"] > 0 and r["LDTM"] > 0 and r["UTCBAR"] > 0
        assert r["HMMA"] == 0
    assert pair["MMA.2CTA"] == pair["UTCHMMA"] > 0      # every MMA of the pair build is a cta_group::2 MMA
    assert solo["MMA.2CTA"] == 0


def test_evaluation_kernel_is_tma_fed_and_packed(histogram):
    r = _row(histogram, "eval_kernel_tma<false, true, true>")
    assert r["UTMALDG"] > 0 and r["HMMA"] == 0
    assert r["FFMA2"] > 100                              # packed two-wide fp32 (pose_metrics16)
    assert r["STL"] == 0 and r["LDL"] == 0 or r["STL"] <= 32   # 3 CTAs per SM: no register spills in the main path
