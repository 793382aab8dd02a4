"""Where does the host time of one e2e training iteration go?  (not a pytest module)"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
loop = bench.TrainLoop("fp32", 64, dev, 1, nbatches=8)
for _ in range(20):
    loop.one_e2e()
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    loop.one_e2e()
t1 = time.perf_counter()
print("e2e wall per step: %.1f us" % ((t1 - t0) / n * 1e6))
# device-only back to back (no sync per step)
t0 = time.perf_counter()
for _ in range(n):
    loop.one()
torch.cuda.synchronize()
t1 = time.perf_counter()
print("device-tensor steps, one sync at the end: %.1f us" % ((t1 - t0) / n * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    loop.one_e2e()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(18)
