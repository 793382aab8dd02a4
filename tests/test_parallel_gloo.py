"""Multi-process plumbing (bilinear_b200/parallel.py) with world_size 2 on CPU / gloo: the gradient
all-reduce in per-layer buckets, the per-action sum combination of the sharded evaluation, max-over-ranks
timing, and the data-parallel algebra itself -- "shard the batch, local BatchNorm, average the gradients,
clip + Adam on the average" -- checked with the ORACLE standing in for the device step on every rank
(the CUDA step is checked against the same oracle in tests/test_gpu_network.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from bilinear_b200 import parallel
    from oracle import bilinear_oracle as bo
    r, lr_, w = parallel.init(backend="gloo")
    assert (r, w) == (rank, world) and parallel.env_world() == (rank, rank, world)

    # ---- gradient all-reduce: six buckets, decode first, must equal one flat all-reduce
    assert parallel.BUCKETS[0][0] == 0 and parallel.BUCKETS[-1][1] == 4291632
    assert all(a[1] == b[0] for a, b in zip(parallel.BUCKETS, parallel.BUCKETS[1:]))
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(4291632, generator=g)
    mine = flat.clone()
    parallel.allreduce_gradients(flat, bucketed=True)
    other = torch.randn(4291632, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    assert torch.equal(flat, mine + other) or torch.allclose(flat, mine + other, rtol=0, atol=1e-6)
    flat2 = mine.clone()
    parallel.allreduce_gradients(flat2, bucketed=False)
    assert torch.equal(flat, flat2)

    # ---- evaluation: shards by pose, per-action sums add up
    n = 1001
    lo, hi = parallel.shard_range(n, rank, world)
    vals = torch.arange(n, dtype=torch.float64)
    acts = (torch.arange(n) * 7) % 15
    sums = torch.zeros(15, dtype=torch.float64).index_add_(0, acts[lo:hi], vals[lo:hi])
    psums = 2 * sums
    cnt = torch.zeros(15, dtype=torch.int64).index_add_(0, acts[lo:hi], torch.ones(hi - lo, dtype=torch.int64))
    parallel.combine_action_sums(sums, psums, cnt)
    full = torch.zeros(15, dtype=torch.float64).index_add_(0, acts, vals)
    assert torch.equal(sums, full) and torch.equal(psums, 2 * full) and int(cnt.sum()) == n
    assert parallel.max_over_ranks(10.0 + rank, torch.device("cpu")) == 10.0 + world - 1

    # ---- replicas start identical: rank 0's parameters / buffers / dropout base seed reach every rank, and the
    # data permutation seed is rank 0's, even when the processes were seeded differently
    import torch.nn as nn

    class _Eng:
        seed = 1000 + rank
        bumped = 0

        def note_external_param_write(self):
            self.bumped += 1
    torch.manual_seed(50 + rank)                     # per-rank seeding: the replicas differ before the broadcast
    toy = nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3))
    toy[1].running_mean.fill_(float(rank))
    eng = _Eng()
    parallel.broadcast_module_state(toy, eng)
    torch.manual_seed(50)
    ref_toy = nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3))
    assert torch.equal(toy[0].weight, ref_toy[0].weight) and float(toy[1].running_mean[0]) == 0.0
    assert eng.seed == 1000 and eng.bumped == 1
    assert parallel.broadcast_int(777 + rank, torch.device("cpu")) == 777
    # ragged tail: train_bilinear.py drops a global batch that leaves some rank with fewer than two rows
    assert [n_tail < 2 * world for n_tail in (1, 3, 4, 5)] == [True, True, False, False]

    # ---- data-parallel step == "reference per shard, gradients averaged" (local BatchNorm), identical on all ranks
    torch.manual_seed(0)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import reference_init_state
    st = bo.State(reference_init_state(0), dtype=torch.float64)
    x, y, masks = bo.synthetic_batch(5, 16)
    lo, hi = parallel.shard_range(16, rank, world)
    xs, ys, ms = x[lo:hi], y[lo:hi], [m[lo:hi] for m in masks]
    p, cache = bo.forward(st, xs.double(), train=True, masks=ms)
    loss, dP = bo.mse_loss(p, ys.double())
    grads = bo.backward(st, cache, dP)
    names = bo.param_names()
    flat_g = torch.cat([grads[k].reshape(-1) for k in names]).float()
    parallel.allreduce_gradients(flat_g)
    off = 0
    avg = {}
    for k in names:
        nelem = grads[k].numel()
        avg[k] = flat_g[off:off + nelem].view(grads[k].shape).double()
        off += nelem
    norm = bo.clip_and_adam(st, avg, bo.lr_schedule(1), max_norm=1.0, grad_scale=1.0 / world)
    torch.save({"p": {k: v.clone() for k, v in st.p.items()}, "norm": float(norm), "loss": float(loss)},
               os.path.join(out_dir, "rank%d.pt" % rank))
    parallel.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_two_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"))
    assert a["norm"] == b["norm"]                       # every rank clips with the same global norm ...
    for k in a["p"]:
        assert torch.equal(a["p"][k], b["p"][k]), k     # ... and ends the step with identical parameters
    assert a["loss"] != b["loss"]                       # each rank saw its own shard
