"""Optimizer phase alone (not a pytest module): python tests/gpu_adambench.py [fp32|bf16]"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bilinear_b200.engine import get_engine
from bilinear_b200.model import bilinear as mb


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net, opt, _, _ = mb.load(torch.device("cuda"))
    net.set_precision(prec)
    net.train()
    x = torch.randn(64, 32, device="cuda")
    y = torch.randn(64, 48, device="cuda")
    for _ in range(3):
        net.train_step(opt, x, y)
    eng = get_engine(net)
    adam = eng.adam_args(1e-3, 5, 1.0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for warm in (True, False):
        ts = []
        for k in range(12):
            if not warm:
                flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.adam_step(adam, 64)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print("%s optimizer launch (sumsq + adam, %s L2): %s us" % (prec, "warm" if warm else "flushed", " ".join("%.1f" % t for t in ts[2:])))
    for name, ph in (("adam_step", eng.debug_phase_ns(64)),):
        for i, (ns, kind, tiles, sub) in enumerate(ph):
            print("   %s phase %d kind %d: %.1f us" % (name, i, kind, ns / 1e3))


if __name__ == "__main__":
    main()
