"""Per-phase timing of the persistent kernel (not a pytest module): python tests/gpu_phases.py fp32:64 bf16:65536"""
import os
import sys
import warnings

# per-phase %globaltimer stamps describe ONE launch of the persistent kernel: keep the large-batch apply passes inside
# it for this script (by default they run as kernels of their own between launches, csrc/api.cu run_sched)
os.environ.setdefault("B200_SPLIT_EW", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bilinear_b200.engine import get_engine
from bilinear_b200._lib import PLAN_INFER, PLAN_TRAIN
from bilinear_b200.model import bilinear as mb

EPI = ["STORE", "PARTIAL", "FWD_LARGE", "FWD_EVAL", "DECODE", "DECODE_T", "DGRAD_LARGE", "WGRAD"]
EW = ["PREP", "DROPBITS", "BN_FOLD", "BN_FWD_APPLY", "BN_BWD_APPLY", "SPLIT_DP", "SUMSQ", "ADAM", "SPLIT_W",
      "BN_FWD_SMALL", "BN_BWD_SMALL", "DECODE_SMALL"]


def show(tag, ph):
    tot = sum(p[0] for p in ph)
    print("== %s: %d phases, %.1f us total" % (tag, len(ph), tot / 1e3))
    for i, (ns, kind, tiles, sub) in enumerate(ph):
        if kind == 0:
            name = "GEMM %-12s tiles=%d" % (EPI[tiles >> 24], tiles & 0xffffff)
        else:
            name = "EW   %-12s jobs=%d" % (EW[kind - 1], tiles)
        print("   %2d %-34s %9.1f us   [prod0 %s prodEnd %s mmaEnd %s acc0 %s epiEnd %s work %s]" %
              ((i, name, ns / 1e3) + tuple("%.1f" % (s / 1e3) if s >= 0 else "-" for s in sub[:6])))
        if i == 0 and len(sub) > 6:
            print("      (kernel entry -> first phase: %.1f us)" % (sub[6] / 1e3))
    sys.stdout.flush()


def main():
    cases = sys.argv[1:] or ["fp32:64", "bf16:64", "bf16:65536"]
    for c in cases:
        prec, b = c.split(":")
        b = int(b)
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net, opt, _, _ = mb.load(torch.device("cuda"))
        net.set_precision(prec)
        net.train()
        x = torch.randn(b, 32, device="cuda")
        y = torch.randn(b, 48, device="cuda")
        for _ in range(3):
            net.train_step(opt, x, y)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        evs[0].record()
        for k in range(20):
            net.train_step(opt, x, y)
            evs[k + 1].record()
        torch.cuda.synchronize()
        print("   back-to-back steps, event-timed: %s us" % " ".join("%.0f" % (evs[k].elapsed_time(evs[k + 1]) * 1e3) for k in range(20)))
        eng = get_engine(net)
        show("train_step %s B=%d" % (prec, b), eng.debug_phase_ns(b))
        net.eval()
        with torch.no_grad():
            for _ in range(2):
                net(x)
        show("forward_eval %s B=%d" % (prec, b), eng.debug_phase_ns(b))
        del net, opt, eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
