"""A handful of eval-mode forwards (valid_bilinear.py:52) at one batch size (ncu target / A/B timing; not a pytest module):
python tests/gpu_inferstep.py [fp32|bf16] [batch] [iters]      B200_TIME=n: n more forwards, event-timed one by one."""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bilinear_b200.model import bilinear as mb

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net, _, _, _ = mb.load(torch.device("cuda"))
net.set_precision(prec)
net.eval()
x = torch.randn(b, 32, device="cuda")
with torch.no_grad():
    for _ in range(iters):
        p = net(x)
torch.cuda.synchronize()
print("ok |p| %.5f" % float(p.norm()))
if os.environ.get("B200_TIME"):
    n = int(os.environ["B200_TIME"])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    with torch.no_grad():
        ev[0].record()
        for i in range(n):
            net(x)
            ev[i + 1].record()
    torch.cuda.synchronize()
    us = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n))
    print("forward time: median %.1f us, min %.1f us, max %.1f us over %d iterations" % (us[n // 2], us[0], us[-1], n))
