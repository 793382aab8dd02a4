"""Benchmark of the pose-lifting hot path on B200 (contract: see README / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train_b64|train_b65536|infer_10m|eval_1m]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Headline workload (BASELINE.json configs[1]): one training iteration of train_bilinear.py:75-83 at batch
64 per GPU -- zero_grad, forward, MSELoss, backward, clip_grad_norm_(1), Adam -- as ONE persistent
sm_100a kernel, fp32-parity mode.  A "step" is one such iteration over one batch of synthetic H36M-shaped
data (x ~ N(0,1) [64,32], y ~ N(0,1) [64,48], random-init weights, in-kernel Philox dropout).  `value` is
whole-job poses/s with the inputs resident in HBM; `e2e` is the same through the public nn.Module API with
pinned HOST inputs copied in and the loss read back every step.  The other BASELINE configs (large-batch
bf16 training, 10M-pose inference, 1M-pose MPJPE/P-MPJPE) are measured in the same run and reported
under "also", each with its own roofline.

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_POSE_TRAIN = 25591808      # fwd 8 552 448 + dgrad 8 486 912 + wgrad 8 552 448
FLOP_PER_POSE_FWD = 8552448
PARAM_NUMEL = 4291632
ADAM_BYTES_PER_STEP = 6 * 4 * PARAM_NUMEL   # p, m, v read + written (SURVEY.md section 8(d), B=64 row)
SPACER_CYCLES = 600_000                     # ~0.3 ms at 1.9 GHz
EVAL_BYTES_PER_POSE = 384                   # 2 x 16 x 3 x 4 read (+ 4 written when per-pose output is kept)


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary under profiles/ (bytes), or None."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        tot = 0.0
        with open(path) as f:
            for line in f:
                k, unit, v = line.rstrip("\n").split(",")[:3]
                if k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return tot or None
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def make_model(precision, device, seed=0):
    import torch
    from bilinear_b200.model import bilinear as mb
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net, opt, step, _ = mb.load(device)          # model/bilinear.py:58-92: random init, Adam(lr=1e-3)
    net.set_precision(precision)
    net.train()
    return net, opt, step


def lr_for(step):
    return 1.0e-3 * 0.96 ** (step / 100000)        # util/config.py:22


class TrainLoop:
    """The hot loop of train_bilinear.py:60-90 on device-resident synthetic batches."""

    def __init__(self, precision, batch, device, world, nbatches):
        import torch
        from bilinear_b200 import parallel
        self.torch, self.parallel = torch, parallel
        self.net, self.opt, self.step = make_model(precision, device)
        self.batch, self.world, self.device = batch, world, device
        self.fused_dp = world > 1 and os.environ.get("BILINEAR_B200_DP", "fused") == "fused"
        if self.fused_dp:
            self.net.enable_data_parallel()
        g = torch.Generator(device="cpu").manual_seed(1234 + parallel.env_world()[0])
        self.xs = [torch.randn(batch, 32, generator=g).to(device) for _ in range(nbatches)]
        self.ys = [torch.randn(batch, 48, generator=g).to(device) for _ in range(nbatches)]
        # pinned host batches as a DataLoader(pin_memory=True) delivers them; x and y of one batch share a
        # pinned block [x | y] so that the library can move them with a single copy
        self.hx, self.hy = [], []
        for x, y in zip(self.xs[:8], self.ys[:8]):
            blk = torch.empty(batch * 80, dtype=torch.float32).pin_memory()
            hx, hy = blk[:batch * 32].view(batch, 32), blk[batch * 32:].view(batch, 48)
            hx.copy_(x.cpu()); hy.copy_(y.cpu())
            self.hx.append(hx); self.hy.append(hy)
        self.i = 0

    def _lr(self):
        if self.step % 100000 == 0 or self.step == 1:   # util/config.py:21, train_bilinear.py:66-70
            for grp in self.opt.param_groups:
                grp["lr"] = lr_for(self.step)

    def one(self, x=None, y=None):
        self._lr()
        if x is None:
            x, y = self.xs[self.i % len(self.xs)], self.ys[self.i % len(self.ys)]
        self.i += 1
        if self.world == 1:
            out = self.net.train_step(self.opt, x, y, max_norm=1.0)
        elif self.fused_dp:
            out = self.net.train_step_dp(self.opt, x, y, max_norm=1.0)
        else:
            from bilinear_b200.engine import get_engine
            out = self.net.train_step(self.opt, x, y, max_norm=1.0, update=False)
            self.parallel.allreduce_gradients(get_engine(self.net).flat_g)
            self.opt.finish_step(max_norm=1.0, grad_scale=1.0 / self.world)
        self.step += 1
        return out

    def one_e2e(self):
        """train_bilinear.py:62-88 through the public API with pinned HOST batches: H2D of x and y, the step,
        and the loss read back (train_bilinear.py:88) every iteration."""
        k = self.i % len(self.hx)
        if self.world == 1:
            self._lr()
            self.i += 1
            loss, _ = self.net.train_step_host(self.opt, self.hx[k], self.hy[k], max_norm=1.0)
            self.step += 1
            return loss
        x = self.hx[k].to(self.device, non_blocking=True)
        y = self.hy[k].to(self.device, non_blocking=True)
        loss, _ = self.one(x, y)
        return float(loss.item())


class L2Flusher:
    def __init__(self, device):
        import torch
        self.buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # 2 x L2 (126 MB)
        self.k = 0

    def flush(self):
        self.k += 1
        self.buf.fill_(float(self.k))


def timed_steps(torch, parallel, fn, steps, flusher, device):
    """Sum of per-step CUDA-event times (ms) on the launching stream, L2 flushed between steps (outside
    the event pairs); barrier + synchronize on both sides; max over ranks."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    parallel.barrier()
    torch.cuda.synchronize()
    for a, b in evs:
        if flusher is not None:
            flusher.flush()
        # ~0.3 ms device-side spacer: the host enqueues the start event, the step and the stop event while
        # the GPU is still busy, so the event pair measures the step on the device and not the Python /
        # ctypes launch path (that path is what `e2e` measures, on the wall clock)
        torch.cuda._sleep(SPACER_CYCLES)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    parallel.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    return parallel.max_over_ranks(ms, device)


def timed_wall(torch, parallel, fn, steps, device):
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    parallel.barrier()
    return parallel.max_over_ranks(ms, device)


# ------------------------------------------------------------------------------------------------ workloads
def bench_train(args, precision, batch, steps, warmup, device, world, peaks, flush=True, with_e2e=True):
    import torch
    from bilinear_b200 import _lib, parallel
    loop = TrainLoop(precision, batch, device, world, nbatches=(64 if batch <= 4096 else 3))
    for _ in range(max(warmup, 3)):
        loop.one()
    flusher = L2Flusher(device) if flush else None
    n0 = _lib.lib().b200_launch_count()
    ms = timed_steps(torch, parallel, loop.one, steps, flusher, device)
    launches = _lib.lib().b200_launch_count() - n0
    ms_step = ms / steps
    poses = batch * world
    res = dict(value=poses / (ms_step * 1e-3), ms_per_step=ms_step, launches=launches)
    if flush:   # steady state, parameters L2-resident (no flush) -- reported beside the flushed number
        ms_warm = timed_steps(torch, parallel, loop.one, steps, None, device)
        res["ms_per_step_warm_l2"] = ms_warm / steps
    if with_e2e:
        for _ in range(3):
            loop.one_e2e()
        ms_e = timed_wall(torch, parallel, loop.one_e2e, steps, device)
        res["e2e"] = dict(value=poses / (ms_e / steps * 1e-3), unit="poses/s", h2d_bytes_per_step=batch * (32 + 48) * 4,
                          d2h_bytes_per_step=4, ms_per_step=ms_e / steps)
    if batch <= 256:
        ach = ADAM_BYTES_PER_STEP / (ms_step * 1e-3) / 1e9
        res["roofline"] = dict(bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"],
                               traffic=ncu_traffic("r1p_train_step_b64_fp32_ncu_full.csv") if precision == "fp32" else None,
                               traffic_source="profiles/r1p_train_step_b64_fp32_ncu_full.csv (bytes per launch)",
                               peak_source=peaks["source"],
                               algorithmic="6 x 4 B x 4 291 632 parameters (Adam p,m,v read+write) per step")
    else:
        ach = FLOP_PER_POSE_TRAIN * batch / (ms_step * 1e-3) / 1e12
        pk = peaks["bf16_sustained"]
        res["roofline"] = dict(bound="tensor", achieved=ach, peak=pk, unit="TFLOP/s", frac=ach / pk,
                               traffic=ncu_traffic("r1p_train_step_b65536_bf16_ncu_full.csv") if batch == 65536 else None,
                               peak_source=peaks["source"] + " (sustained bf16 cuBLAS)",
                               algorithmic="25 591 808 FLOP/pose x %d poses per step" % batch)
    return res


def bench_infer(total, chunk, precision, steps, warmup, device, rank, world, peaks):
    """BASELINE config 4: eval-mode lifting of `total` poses sharded by pose index (no communication)."""
    import torch
    from bilinear_b200 import parallel
    net, _, _ = make_model(precision, device)
    net.eval()
    lo, hi = parallel.shard_range(total, rank, world)
    n = hi - lo
    g = torch.Generator(device=device).manual_seed(7 + rank)
    x = torch.randn(n, 32, device=device, generator=g)
    out = torch.empty(n, 48, device=device)

    def run():
        with torch.no_grad():
            for s in range(0, n, chunk):
                e = min(n, s + chunk)
                out[s:e] = net(x[s:e])
    for _ in range(max(1, min(warmup, 2))):
        run()
    ms = timed_steps(torch, parallel, run, steps, None, device) / steps
    ach = FLOP_PER_POSE_FWD * n / (ms * 1e-3) / 1e12
    pk = peaks["bf16_sustained"]
    return dict(workload="infer_%d_%s" % (total, precision), value=total / (ms * 1e-3), unit="poses/s", ms=ms,
                chunk=chunk, l2="inputs (%.0f MB per rank) exceed L2" % (n * 128 / 1e6),
                roofline=dict(bound="tensor", achieved=ach, peak=pk, unit="TFLOP/s", frac=ach / pk, traffic=None))


def bench_eval(total, steps, warmup, device, rank, world, peaks):
    """BASELINE config 5: MPJPE + P-MPJPE over `total` poses sharded by pose index."""
    import torch
    from bilinear_b200 import metrics, parallel
    lo, hi = parallel.shard_range(total, rank, world)
    n = hi - lo
    g = torch.Generator(device=device).manual_seed(11 + rank)
    gt = torch.randn(n, 48, device=device, generator=g)
    pred = gt + 0.2 * torch.randn(n, 48, device=device, generator=g)
    mean = 100.0 * torch.randn(48, device=device, generator=g)
    std = 50.0 + 200.0 * torch.rand(48, device=device, generator=g)
    act = torch.randint(0, 15, (n,), device=device, generator=g, dtype=torch.int32)
    acc = (torch.zeros(15, dtype=torch.float64, device=device), torch.zeros(15, dtype=torch.float64, device=device),
           torch.zeros(15, dtype=torch.int64, device=device))
    flusher = L2Flusher(device)
    out = {}
    for name, kw in (("mpjpe", dict(want_pmpjpe=False)), ("pmpjpe", dict(want_mpjpe=False)), ("fused", dict())):
        fn = lambda: metrics.evaluate(pred, gt, mean, std, act, 15, joints=16, accum=acc, **kw)
        for _ in range(max(warmup, 3)):
            fn()
        ms = timed_steps(torch, parallel, fn, steps, flusher, device) / steps
        ach = (EVAL_BYTES_PER_POSE + (8 if name == "fused" else 4) + 4) * n / (ms * 1e-3) / 1e9
        out[name] = dict(value=total / (ms * 1e-3), unit="poses/s", ms=ms,
                         roofline=dict(bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s",
                                       frac=ach / peaks["hbm"], traffic=None))
    r = metrics.evaluate(pred, gt, mean, std, act, 15, joints=16)
    parallel.combine_action_sums(r.sum, r.psum, r.count)
    out["mpjpe_mm"], out["pmpjpe_mm"] = r.mpjpe, r.pmpjpe
    out["workload"] = "eval_%d" % total
    out["l2"] = "flushed between iterations"
    return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_train_sample(batch, seconds_budget, min_steps=3):
    """The reference's CPU path for one training iteration (train_bilinear.py:75-83), restated by the
    oracle port in float32 on all host threads.  Returns (poses/s, steps timed, threads)."""
    import torch
    from oracle import bilinear_oracle as bo
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    from bilinear_b200.model import bilinear as mb
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net, _, _, _ = mb.load(torch.device("cpu"))
    st = bo.State(net.state_dict(), dtype=torch.float32)
    x, y, masks = bo.synthetic_batch(1, batch)
    for _ in range(2):
        bo.train_step(st, x, y, masks, lr_for(1))
    t0 = time.perf_counter()
    n = 0
    while n < min_steps or (time.perf_counter() - t0) < seconds_budget:
        bo.train_step(st, x, y, masks, lr_for(1))
        n += 1
        if n >= 2000:
            break
    dt = time.perf_counter() - t0
    return batch * n / dt, n, threads, dt


METRIC = "train poses/sec, bilinear MLP (full train_bilinear.py iteration: fwd+MSE+bwd+clip+Adam)"


def workload_config(workload, batch, world, fused_dp=True):
    """`config` of the JSON line -- shared by the B200 arm and the reference (CPU) arm."""
    if world > 1:
        par = ("dp%d (local BatchNorm, gradient all-reduce inside the persistent kernel over NVLink peer memory)"
               if fused_dp else "dp%d (local BatchNorm, NCCL all-reduce of the flat gradient)") % world
    else:
        par = "single GPU"
    return {"workload": "%s: batch %d per GPU, 16 joints, 1024 hidden, 2 residual blocks, random init, "
                        "Philox dropout, clip_grad_norm 1, Adam" % (workload, batch),
            "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": par,
            "l2": "flushed (256 MB fill) between timed iterations" if batch <= 4096 else "activations exceed L2"}


def run_reference(args, emit):
    """--impl reference: the CPU arm alone, same metric / config / unit as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import bilinear_oracle as bo
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    batch = 64 if args.workload == "train_b64" else 65536
    if batch > 4096:
        batch_sample = 4096       # bounded sample of the large-batch step
    else:
        batch_sample = batch
    from bilinear_b200.model import bilinear as mb
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net, _, _, _ = mb.load(torch.device("cpu"))
    st = bo.State(net.state_dict(), dtype=torch.float32)
    x, y, masks = bo.synthetic_batch(1, batch_sample)
    for _ in range(max(args.warmup, 1) if batch_sample <= 256 else 1):
        bo.train_step(st, x, y, masks, lr_for(1))
    steps = args.steps if batch_sample <= 256 else min(args.steps, 5)
    steps = min(steps, 2000)
    t0 = time.perf_counter()
    for _ in range(steps):
        bo.train_step(st, x, y, masks, lr_for(1))
    dt = time.perf_counter() - t0
    value = batch_sample * steps / dt
    cfg = workload_config(args.workload, batch, max(args.gpus, 1))
    cfg["reference_arm"] = ("oracle/bilinear_oracle.py (CPU restatement of the reference's PyTorch path; the reference "
                            "itself is a Python package and does not travel to the GPU box), batch %d per step on %d "
                            "host threads, rank 0 only" % (batch_sample, threads))
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "poses/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "poses/s", "cores": threads, "kind": "port",
                         "sample": "%d iterations of batch %d (oracle/bilinear_oracle.py, float32, torch CPU ops)"
                                   % (steps, batch_sample)},
        "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ main
def main():
    # the contract is ONE JSON line on stdout: anything libraries write to file descriptor 1 (e.g. NCCL's
    # version banner) is sent to stderr instead, and the line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train_b64", choices=["train_b64", "train_b65536"])
    ap.add_argument("--no-also", action="store_true", help="skip the secondary BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, emit)
        return

    import torch
    from bilinear_b200 import _lib, parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    _lib.lib()
    rank, local_rank, world = parallel.init()
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun with --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    peaks = load_peaks()
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if args.workload == "train_b64":
        main_res = bench_train(args, "fp32", 64, steps, warmup, device, world, peaks)
        dtype, batch = "f32 (3 x bf16 split operands on tcgen05, fp32 accumulate)", 64
    else:
        main_res = bench_train(args, "bf16", 65536, min(steps, 20), warmup, device, world, peaks, flush=False)
        dtype, batch = "bf16", 65536
        steps = min(steps, 20)
    clocks = sampler.stop() if sampler else None

    also = []
    if not args.no_also:
        s2 = 10
        try:
            if args.workload != "train_b65536":
                r = bench_train(args, "bf16", 65536, s2, 3, device, world, peaks, flush=False, with_e2e=True)
                r["workload"] = "train_b65536_bf16"; r["unit"] = "poses/s"
                r["l2"] = "activations (>1 GB per step) exceed L2"
                also.append(r)
            r = bench_train(args, "bf16", 64, steps, warmup, device, world, peaks, with_e2e=False)
            r["workload"] = "train_b64_bf16"; r["unit"] = "poses/s"; r["l2"] = "flushed between iterations"
            also.append(r)
            also.append(bench_infer(10_000_000, 65536, "bf16", 3, 1, device, rank, world, peaks))
            also.append(bench_infer(10_000_000, 65536, "fp32", 2, 1, device, rank, world, peaks))
            also.append(bench_eval(1_000_000, 50, 5, device, rank, world, peaks))
        except Exception as e:  # a secondary config must not take the headline down with it
            also.append({"error": repr(e)[:400]})

    cpu = None
    if rank == 0 and not args.no_cpu:
        v, n, threads, dt = cpu_train_sample(64, 12.0)
        cpu = {"value": v, "unit": "poses/s", "cores": threads, "kind": "port",
               "sample": "%d train iterations of batch 64 in %.1f s (oracle/bilinear_oracle.py, float32, torch CPU ops)"
                         % (n, dt)}
    parallel.barrier()

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": main_res["value"], "unit": "poses/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(args.workload, batch, world, os.environ.get("BILINEAR_B200_DP", "fused") == "fused"),
            "roofline": main_res["roofline"], "e2e": main_res.get("e2e"),
            "gpu_launches": main_res["launches"], "clocks": clocks, "cpu_baseline": cpu,
            "ms_per_step_warm_l2": main_res.get("ms_per_step_warm_l2"),
            "also": also,
        }
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
