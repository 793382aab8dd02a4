"""MPJPE / P-MPJPE kernel timing only (not a pytest module): python tests/gpu_evalbench.py [poses] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.cuda.set_device(0)
r = bench.bench_eval(n, steps, 5, torch.device("cuda", 0), 0, 1, bench.load_peaks())
print(json.dumps(r, indent=1))
