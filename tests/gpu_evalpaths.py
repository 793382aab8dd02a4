"""Timing of the evaluation entry points per code path (not a pytest module): two-kernel scheme (per-pose results +
per-action reduction kernel) against the single kernel with in-kernel per-action accumulation, CUDA events, L2
flushed, 1 M poses.   python tests/gpu_evalpaths.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bilinear_b200 import metrics

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(11)
gt = torch.randn(n, 48, device=dev, generator=g)
pred = gt + 0.2 * torch.randn(n, 48, device=dev, generator=g)
mean = 100.0 * torch.randn(48, device=dev, generator=g)
std = 50.0 + 200.0 * torch.rand(48, device=dev, generator=g)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in evs:
        flush.fill_(1.0)
        torch.cuda._sleep(600_000)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    return t[len(t) // 2], t[0]


for na in (15, 17):     # <= 16 buckets: two-kernel scheme; more: in-kernel accumulation
    act = torch.randint(0, na, (n,), device=dev, generator=g, dtype=torch.int32)
    acc = (torch.zeros(na, dtype=torch.float64, device=dev), torch.zeros(na, dtype=torch.float64, device=dev),
           torch.zeros(na, dtype=torch.int64, device=dev))
    for name, kw in (("mpjpe", dict(want_pmpjpe=False)), ("pmpjpe", dict(want_mpjpe=False)), ("fused", dict())):
        for keep in (True, False):
            fn = lambda: metrics.evaluate(pred, gt, mean, std, act, na, joints=16, accum=acc, keep_per_pose=keep, **kw)
            med, best = timed(fn)
            print("n_actions %2d %-6s keep_per_pose=%-5s median %6.1f us  best %6.1f us  -> %4.0f GB/s (384 B/pose)"
                  % (na, name, keep, med, best, 384.0 * n / med / 1e3))
