"""A handful of fused training iterations (ncu target; not a pytest module):
python tests/gpu_onestep.py [fp32|bf16] [batch] [iters]

Under ncu, large batches need B200_CG=1 in the environment: ncu cannot replay the cooperative cluster launch of
the CTA-pair build (LaunchFailed)."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bilinear_b200.model import bilinear as mb

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net, opt, _, _ = mb.load(torch.device("cuda"))
net.set_precision(prec)
net.train()
x = torch.randn(b, 32, device="cuda")
y = torch.randn(b, 48, device="cuda")
for _ in range(iters):
    loss, gn = net.train_step(opt, x, y)
torch.cuda.synchronize()
print("ok loss %.5f gnorm %.4f" % (float(loss), float(gn)))
if os.environ.get("B200_TIME"):      # B200_TIME=n: n more iterations, event-timed one by one (A/B runs of build switches)
    n = int(os.environ["B200_TIME"])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        net.train_step(opt, x, y)
        ev[i + 1].record()
    torch.cuda.synchronize()
    us = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n))
    print("step time: median %.1f us, min %.1f us, max %.1f us over %d iterations" % (us[n // 2], us[0], us[-1], n))
