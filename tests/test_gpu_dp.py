"""Multi-rank parity of the data-parallel training step under pytest: one process per visible GPU, launched
with torch.distributed.run on 127.0.0.1, each rank checking itself against the oracle "reference per shard
(local BatchNorm), gradients averaged, clip + Adam on the average" (tests/gpu_dp_check.py; SURVEY.md section
8(e)(i); the step is train_bilinear.py:75-83).  Both exchange paths are covered: `fused` = reduce-scatter +
all-gather INSIDE the persistent kernel over NVLink peer memory (the default), `nccl` = NCCL all-reduce between
two kernels.  Skipped on a single-GPU box; run with `gpurun --gpus N -- python -m pytest tests/test_gpu_dp.py -m gpu`
(logs of the 2- / 4- / 8-GPU runs are kept under profiles/)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(900)
@pytest.mark.parametrize("variant", ["fused", "nccl"])
def test_data_parallel_step_matches_per_shard_reference(cuda_lib, variant):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus N)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "gpu_dp_check.py"), variant]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=840, cwd=ROOT)
    log_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(log_dir, exist_ok=True)
    with open(os.path.join(log_dir, "dp_check_%s_%dgpu.log" % (variant, n)), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # ranks share one pipe and their lines may run together: count the markers, not the lines
    assert r.stdout.count("-> OK") == n, r.stdout[-3000:]
    assert r.stdout.count("[%s] rank" % variant) == n  and "-> FAIL" not in r.stdout, r.stdout[-3000:]
    assert r.stdout.count("parameters identical across ranks: True") == n
