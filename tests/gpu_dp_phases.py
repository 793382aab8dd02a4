"""Per-phase timing of the fused data-parallel step (not a pytest module; launched by torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
        tests/gpu_dp_phases.py [fp32|bf16] [batch per GPU]
"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

from bilinear_b200 import parallel
from bilinear_b200.engine import get_engine
from bilinear_b200.model import bilinear as mb
from gpu_phases import EPI

EW = ["PREP", "DROPBITS", "BN_FOLD", "BN_FWD_APPLY", "BN_BWD_APPLY", "SPLIT_DP", "SUMSQ", "ADAM", "SPLIT_W",
      "BN_FWD_SMALL", "BN_BWD_SMALL", "DECODE_SMALL", "COMM_BARRIER", "ALLREDUCE"]


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rank, local_rank, world = parallel.init()
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net, opt, _, _ = mb.load(dev)
    net.set_precision(prec)
    net.train()
    net.enable_data_parallel()
    x = torch.randn(per, 32, device=dev)
    y = torch.randn(per, 48, device=dev)
    for _ in range(5):
        net.train_step_dp(opt, x, y, max_norm=1.0)
    torch.cuda.synchronize()
    dist.barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    evs[0].record()
    for k in range(20):
        net.train_step_dp(opt, x, y, max_norm=1.0)
        evs[k + 1].record()
    torch.cuda.synchronize()
    ph = get_engine(net).debug_phase_ns(per)
    dist.barrier()
    if rank == 0:
        print("world %d, %s, batch %d per GPU; back-to-back steps: %s us" %
              (world, prec, per, " ".join("%.0f" % (evs[k].elapsed_time(evs[k + 1]) * 1e3) for k in range(20))))
        tot = sum(p[0] for p in ph)
        print("== fused data-parallel step: %d phases, %.1f us total" % (len(ph), tot / 1e3))
        for i, (ns, kind, tiles, sub) in enumerate(ph):
            name = ("GEMM %-12s tiles=%d" % (EPI[tiles >> 24], tiles & 0xffffff)) if kind == 0 else "EW   %-12s" % EW[kind - 1]
            print("   %2d %-34s %9.1f us" % (i, name, ns / 1e3))
    parallel.shutdown() if hasattr(parallel, "shutdown") else None


if __name__ == "__main__":
    main()
