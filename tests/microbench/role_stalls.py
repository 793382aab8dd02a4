import os, sys, warnings
sys.path.insert(0, "/root/repo")
import torch
from bilinear_b200.engine import get_engine
from bilinear_b200.model import bilinear as mb
sys.path.insert(0, "/root/repo/tests")
from gpu_phases import EPI, EW
b = 65536
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net, opt, _, _ = mb.load(torch.device("cuda"))
net.set_precision("bf16"); net.train()
x = torch.randn(b, 32, device="cuda"); y = torch.randn(b, 48, device="cuda")
for _ in range(5): net.train_step(opt, x, y)
torch.cuda.synchronize()
ph = get_engine(net).debug_phase_ns(b)
for i, (ns, kind, tiles, sub) in enumerate(ph):
    if kind != 0: continue
    prod_stall, prod_tot, mma_full, epi_stall, mma_issue, _, mma_acc = sub
    f = lambda v: 100.0 * v / max(prod_tot, 1)
    print("%2d %-12s %8.1f us  producer cycles %9d: wait-empty %5.1f%% | mma wait-full %5.1f%% wait-acc %5.1f%% issue %5.1f%% | epilogue wait-acc-full %5.1f%%"
          % (i, EPI[tiles >> 24], ns / 1e3, prod_tot, f(prod_stall), f(mma_full), f(mma_acc), f(mma_issue), f(epi_stall)))
