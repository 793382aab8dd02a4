"""Benchmark of the pose-lifting hot path on B200 (contract: DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train_b65536|train_b64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...   # the UNMODIFIED reference module on the box's host cores (CPU)
    python bench.py --impl torch_gpu ...   # the UNMODIFIED reference module on the same B200 (torch eager / CUDA graph)

Headline workload (BASELINE.json configs[2], the largest single-GPU configuration and the one the north-star
targets are quoted on): one training iteration of train_bilinear.py:75-83 -- zero_grad, forward, MSELoss,
backward, clip_grad_norm_(1), Adam -- at batch 65 536 PER GPU in bf16 mode, as one persistent sm_100a kernel;
data parallel (N > 1) with the gradient all-reduce inside that kernel.  A "step" is one such iteration over one
batch of synthetic H36M-shaped data (x ~ N(0,1) [B,32], y ~ N(0,1) [B,48], random-init weights, in-kernel Philox
dropout).  `value` is whole-job poses/s with the inputs resident in HBM (CUDA events on the launching stream);
`e2e` is the same through the public nn.Module API with pinned HOST batches: every step copies its own inputs
host -> device (bilinear_b200.feed.PinnedPrefetcher keeps that copy one step ahead on a side stream) and reads
the loss back.  The other BASELINE configs -- the batch-64 persistent-kernel step (configs[1]), 10 M-pose
inference, 1 M-pose MPJPE / P-MPJPE, and at N > 1 the strong-scaling row (65 536 GLOBAL) -- are measured in the
same run and reported compactly under "also"; "torch_gpu" holds the reference's own GPU path (the unmodified
module under torch on the same B200) for the two training configs.

One JSON line on stdout (rank 0).  Nothing here reads /root/reference: the reference module comes from
baseline/_ref (staged by __graft_entry__.build()).
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
EVAL_BYTES_PER_POSE = 384                   # 2 x 16 x 3 x 4 bytes read per pose
LARGE_BATCH = 65536
CPU_SAMPLE_ROWS = 8192                      # rows of the 65 536-row step the CPU arm times per step (bounded sample)
NCU_TRAFFIC = {"train_b65536": "r5_train_step_b65536_bf16_ncu.csv", "train_b64": "r1p_train_step_b64_fp32_ncu_full.csv"}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    summary under profiles/ (bytes), or None."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        tot = 0.0
        with open(path) as f:
            for line in f:
                parts = line.rstrip("\n").split(",")
                if len(parts) < 3 or line.startswith("#"):
                    continue
                k, unit, v = parts[:3]
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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
        # median over the samples taken UNDER LOAD (power well above idle); all samples if none qualifies
        busy = [s for s, p in zip(sm, pw) if p > 400.0] or sm
        busy.sort()
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "samples_under_load": len(busy), "reasons": sorted(reasons)}


def note(msg):
    """Progress line on stderr (stdout carries the one JSON line only)."""
    sys.stderr.write("[bench %s r%s] %s\n" % (time.strftime("%H:%M:%S"), os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


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
    """The hot loop of train_bilinear.py:60-90 on synthetic batches."""

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
        for x, y in zip(self.xs[:4], self.ys[:4]):
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

    def e2e_small(self):
        """train_bilinear.py:62-88 at batch 64 on one GPU: H2D of x and y, the step and the loss read-back are ONE
        library call (BilinearUnit.train_step_host)."""
        k = self.i % len(self.hx)
        self._lr()
        self.i += 1
        loss, _ = self.net.train_step_host(self.opt, self.hx[k], self.hy[k], max_norm=1.0)
        self.step += 1
        return loss

    def e2e_run(self, steps):
        """`steps` iterations fed from pinned HOST batches: each step's inputs are copied host -> device inside
        the timed region (one step ahead, on a side stream) and each step's loss is read back on the host --
        like train_bilinear.py:88 reads loss.item() -- one step late, so that the read never drains the queue."""
        from bilinear_b200.feed import PinnedPrefetcher
        torch = self.torch
        host = ((self.hx[i % len(self.hx)], self.hy[i % len(self.hy)]) for i in range(steps))
        feed = PinnedPrefetcher(host, self.device)
        ring = torch.zeros(2, dtype=torch.float32).pin_memory()
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        total = 0.0
        for i, (x, y) in enumerate(feed):
            loss, _ = self.one(x, y)
            ring[i & 1].copy_(loss, non_blocking=True)      # 4-byte D2H of this step's loss
            evs[i & 1].record()
            if i > 0:
                evs[(i - 1) & 1].synchronize()
                total += float(ring[(i - 1) & 1])
        evs[(steps - 1) & 1].synchronize()
        total += float(ring[(steps - 1) & 1])
        return total, feed.h2d_bytes // max(steps, 1)


class L2Flusher:
    """Evicts the timed kernels' inputs from the 126 MB L2: a 256 MB fill (write), then a 256 MB read pass over a second
    buffer.  The read pass matters: the fill alone leaves the L2 full of DIRTY lines, and the timed kernel then pays for
    their write-back (up to 126 MB of extra HBM writes inside the event pair -- +33 % traffic on a 384 MB evaluation pass,
    which is how a 62 us kernel measured 83 us).  After the read pass the L2 holds clean lines of a buffer nobody uses."""

    def __init__(self, device):
        import torch
        self.buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # 2 x L2 (126 MB)
        self.rd = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)
        self.k = 0

    def flush(self):
        import torch
        self.k += 1
        self.buf.fill_(float(self.k))
        if not os.environ.get("B200_FLUSH_WRITE_ONLY"):
            torch.sum(self.rd)


def timed_steps(torch, parallel, fn, steps, flusher, device, spacer=True):
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
        if spacer:
            torch.cuda._sleep(SPACER_CYCLES)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    parallel.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    return parallel.max_over_ranks(ms, device)


def timed_block(torch, parallel, fn, steps, device):
    """K steps back to back between ONE pair of CUDA events (ms): what a long-running job sees -- no spacer,
    no flush (the large-batch step streams > 8 GB through a 126 MB L2 by itself)."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    parallel.barrier()
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    parallel.barrier()
    return parallel.max_over_ranks(a.elapsed_time(b), device)


def timed_wall(torch, parallel, fn, device):
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    parallel.barrier()
    return parallel.max_over_ranks(ms, device), out


# ------------------------------------------------------------------------------------------------ workloads
def bench_train(precision, batch, steps, warmup, device, world, peaks, with_e2e=True):
    import torch
    from bilinear_b200 import _lib, parallel
    small = batch <= 4096
    loop = TrainLoop(precision, batch, device, world, nbatches=(64 if small else 3))
    note("train %s B=%d: warm-up" % (precision, batch))
    for _ in range(max(warmup, 3)):
        loop.one()
    note("train %s B=%d: timed" % (precision, batch))
    n0 = _lib.lib().b200_launch_count()
    if small:
        flusher = L2Flusher(device)
        ms = timed_steps(torch, parallel, loop.one, steps, flusher, device)
    else:
        ms = timed_block(torch, parallel, loop.one, steps, device)
    launches = _lib.lib().b200_launch_count() - n0
    ms_step = ms / steps
    poses = batch * world
    res = dict(value=poses / (ms_step * 1e-3), ms_per_step=ms_step, launches=launches)
    if small:   # steady state, parameters L2-resident (no flush) -- reported beside the flushed number
        res["ms_per_step_warm_l2"] = timed_steps(torch, parallel, loop.one, steps, None, device) / steps
    if with_e2e:
        note("train %s B=%d: e2e" % (precision, batch))
        if small and world == 1:
            for _ in range(3):
                loop.e2e_small()
            ms_e, _ = timed_wall(torch, parallel, lambda: [loop.e2e_small() for _ in range(steps)], device)
            h2d = batch * 80 * 4
        else:
            loop.e2e_run(3)
            ms_e, (_, h2d) = timed_wall(torch, parallel, lambda: loop.e2e_run(steps), device)
        res["e2e"] = dict(value=poses / (ms_e / steps * 1e-3), unit="poses/s", h2d_bytes_per_step=int(h2d),
                          d2h_bytes_per_step=4, ms_per_step=ms_e / steps)
    if batch <= 256:
        ach = ADAM_BYTES_PER_STEP / (ms_step * 1e-3) / 1e9
        res["roofline"] = dict(bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"],
                               traffic=ncu_traffic(NCU_TRAFFIC["train_b64"]) if precision == "fp32" else None,
                               peak_source=peaks["source"],
                               algorithmic="6 x 4 B x 4 291 632 parameters (Adam p,m,v read+write) per step")
    else:
        ach = FLOP_PER_POSE_TRAIN * batch / (ms_step * 1e-3) / 1e12
        res["roofline"] = dict(bound="tensor", achieved=ach, peak=peaks["bf16"], unit="TFLOP/s", frac=ach / peaks["bf16"],
                               frac_of_sustained=ach / peaks["bf16_sustained"],
                               traffic=ncu_traffic(NCU_TRAFFIC["train_b65536"]) if batch == LARGE_BATCH else None,
                               peak_source=peaks["source"] + " (cuBLAS bf16 8192^3: burst %.0f, sustained %.0f TFLOP/s)"
                               % (peaks["bf16"], peaks["bf16_sustained"]),
                               traffic_note="DRAM bytes of ONE step = its 22 launches together (ncu, profiles/r5_train_step_"
                                            "b65536_bf16_launch_list.txt); algorithmic activation traffic ~4.6 GB" if batch == LARGE_BATCH
                               else None,
                               algorithmic="25 591 808 FLOP/pose x %d poses per step per GPU" % batch)
    del loop
    torch.cuda.empty_cache()
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
    # fp32-parity mode issues six bf16 MMAs per product (3-way split operands): its ceiling is 1/6 of the bf16 peak
    pk = peaks["bf16"] / (6.0 if precision == "fp32" else 1.0)
    return dict(workload="infer_%d_%s" % (total, precision), value=total / (ms * 1e-3), unit="poses/s", ms=ms,
                roofline=dict(bound="tensor", achieved=ach, peak=pk, unit="TFLOP/s", frac=ach / pk,
                              note="fp32-parity mode: 6 bf16 MMAs per product, peak = bf16 burst / 6" if precision == "fp32"
                              else "bf16 burst"))


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
        ach = EVAL_BYTES_PER_POSE * n / (ms * 1e-3) / 1e9        # 384 B/pose, nothing else counted
        out[name] = dict(value=total / (ms * 1e-3), us=ms * 1e3, hbm_frac=ach / peaks["hbm"], gbs=ach)
    r = metrics.evaluate(pred, gt, mean, std, act, 15, joints=16)
    parallel.combine_action_sums(r.sum, r.psum, r.count)
    out["mpjpe_mm"], out["pmpjpe_mm"] = r.mpjpe, r.pmpjpe
    out["workload"] = "eval_%d (MPJPE / P-MPJPE / both in one pass; 384 B per pose; L2 flushed between iterations: 256 MB fill + 256 MB read pass)" % total
    out["unit"] = "poses/s"
    return out


# ------------------------------------------------------------------------------------------------ reference arms
def reference_module():
    import baseline
    return baseline.load_reference_module()


def reference_step_fn(ref, device, batch, autocast=False, seed=0):
    """The unmodified reference module + the literal sequence of train_bilinear.py:66-83 on synthetic batches."""
    import torch
    import torch.nn as nn
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bilinear, optimizer, step, _ = ref.load(device)          # model/bilinear.py:58-92
    criterion = nn.MSELoss()
    bilinear.train()
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.randn(batch, 32, generator=g).to(device)
    y = torch.randn(batch, 48, generator=g).to(device)
    for param_group in optimizer.param_groups:                   # train_bilinear.py:66-70 at step == 1
        param_group['lr'] = lr_for(1)

    def one():
        optimizer.zero_grad()
        if autocast:
            with torch.autocast(device.type, dtype=torch.bfloat16):
                prediction = bilinear(x)
            loss = criterion(prediction.float(), y)
        else:
            prediction = bilinear(x)
            loss = criterion(prediction, y)
        loss.backward()
        nn.utils.clip_grad_norm_(bilinear.parameters(), max_norm=1)
        optimizer.step()
        return loss
    return one, bilinear, optimizer, (x, y)


def cpu_reference_sample(batch_rows, steps, warmup, threads=None):
    """Time the reference's CPU path on `batch_rows` rows per step.  Returns (poses/s, ms_per_step, threads, kind)."""
    import torch
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    import baseline
    if baseline.available():
        one, _, _, _ = reference_step_fn(reference_module(), torch.device("cpu"), batch_rows)
        kind = "reference"
    else:       # the staged copy is missing (build() did not run in the build container): fall back to the oracle port
        from oracle import bilinear_oracle as bo
        from bilinear_b200.model import bilinear as mb
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net, _, _, _ = mb.load(torch.device("cpu"))
        st = bo.State(net.state_dict(), dtype=torch.float32)
        x, y, masks = bo.synthetic_batch(1, batch_rows)
        one = lambda: bo.train_step(st, x, y, masks, lr_for(1))
        kind = "port"
    for _ in range(max(1, warmup)):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch_rows * steps / dt, dt / steps * 1e3, threads, kind


def cpu_baseline_entry(workload, budget_s=120):
    """The CPU baseline of the JSON line: `bench.py --impl reference` in a CHILD process with a clean threading
    environment.  torchrun exports OMP_NUM_THREADS=1 to every rank, and an OpenMP runtime that was initialised with
    one thread does not grow back reliably -- round 1 reported a 200x too low figure at N > 1 that way."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT",
                        "TORCHELASTIC_RUN_ID", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE")}
    env["CUDA_VISIBLE_DEVICES"] = ""          # the child is the CPU arm: it must not touch the GPUs
    small = workload == "train_b64"
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
           "--steps", "200" if small else "6", "--warmup", "20" if small else "2"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=budget_s)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    return line["cpu_baseline"]


def torch_gpu_arm(device, steps, warmup, configs=(("fp32", 64), ("bf16_autocast", LARGE_BATCH))):
    """The reference's OWN GPU path: the unmodified module on `device` under torch (util/config.py:17 picks cuda
    when it is there), eager as written and -- a stronger baseline than the reference itself -- the same
    iteration captured whole in a CUDA graph.  fp32 at batch 64; at batch 65 536 additionally bf16 autocast."""
    import torch
    from bilinear_b200 import parallel
    ref = reference_module()
    out = {}
    for mode, batch in configs:
        one, net, opt, _ = reference_step_fn(ref, device, batch, autocast=(mode == "bf16_autocast"))
        small = batch <= 4096
        k = steps if small else min(steps, 10)
        for _ in range(max(3, warmup if small else 3)):
            one()
        flusher = L2Flusher(device) if small else None
        if small:
            ms_eager = timed_steps(torch, parallel, one, k, flusher, device) / k
        else:
            ms_eager = timed_block(torch, parallel, one, k, device) / k
        entry = {"batch": batch, "eager_ms_per_step": ms_eager, "eager_poses_per_s": batch / (ms_eager * 1e-3)}
        try:
            # whole-iteration capture (torch "whole network" recipe): capturable Adam, grads set to None inside the graph
            for grp in opt.param_groups:
                grp["capturable"] = True
            for st in opt.state.values():
                if "step" in st and not st["step"].is_cuda:
                    st["step"] = st["step"].to(device)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    one()
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                one()
            for _ in range(3):
                graph.replay()
            if small:
                ms_graph = timed_steps(torch, parallel, graph.replay, k, flusher, device) / k
            else:
                ms_graph = timed_block(torch, parallel, graph.replay, k, device) / k
            entry.update(graph_ms_per_step=ms_graph, graph_poses_per_s=batch / (ms_graph * 1e-3))
            del graph
        except Exception as e:      # capture is an extra; the eager number stands on its own
            entry["graph_error"] = repr(e)[:200]
        out["train_b%d_%s" % (batch, mode)] = entry
        del one, net, opt
        torch.cuda.empty_cache()
    out["what"] = ("unmodified reference module (baseline/_ref/model/bilinear.py) + literal train_bilinear.py:75-83 on the "
                   "same B200 under torch %s; same CUDA-event protocol as the headline" % torch.__version__)
    return out


METRIC = "train poses/sec, bilinear MLP (full train_bilinear.py iteration: fwd+MSE+bwd+clip+Adam)"


def workload_config(workload, batch, world, fused_dp=True):
    """`config` of the JSON line -- identical for the B200 arm and the reference arms."""
    if world > 1:
        par = ("dp%d (local BatchNorm, gradient all-reduce inside the persistent kernel over NVLink peer memory)"
               if fused_dp else "dp%d (local BatchNorm, NCCL all-reduce of the flat gradient)") % world
    else:
        par = "single GPU"
    return {"workload": "%s: batch %d per GPU, 16 joints, 1024 hidden, 2 residual blocks, random init, "
                        "Philox dropout, clip_grad_norm 1, Adam" % (workload, batch),
            "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": par,
            "l2": "flushed (256 MB fill, then a 256 MB read pass so that the fill's dirty lines are written back before the event pair) between timed iterations" if batch <= 4096
                  else "no flush needed: one step streams > 8 GB of activations through the 126 MB L2"}


def batch_of(workload):
    return 64 if workload == "train_b64" else LARGE_BATCH


def run_reference(args, emit):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, same metric /
    config / unit as the B200 arm.  Each step is a bounded sample of the workload: CPU_SAMPLE_ROWS rows of the
    65 536-row batch (the whole batch at batch 64).  Rank 0 alone runs; the other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    batch = batch_of(args.workload)
    rows = min(batch, CPU_SAMPLE_ROWS)
    steps = max(1, min(args.steps, 2000 if rows <= 256 else 40))
    value, ms, threads, kind = cpu_reference_sample(rows, steps, max(1, min(args.warmup, 20 if rows <= 256 else 3)))
    sample = ("%d iterations of %d rows%s on %d host threads; %s" %
              (steps, rows, "" if rows == batch else " (bounded slice of the %d-row step)" % batch, threads,
               "unmodified reference module from baseline/_ref + literal train_bilinear.py:75-83, fp32"
               if kind == "reference" else "oracle/bilinear_oracle.py port (baseline/_ref not staged), fp32"))
    emit({
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "poses/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, batch, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "poses/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_torch_gpu(args, emit):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from bilinear_b200 import parallel  # noqa: F401  (single process: barrier() is a no-op)
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    batch = batch_of(args.workload)
    res = torch_gpu_arm(device, max(args.steps, 1), max(args.warmup, 3))
    key = "train_b%d_%s" % (batch, "fp32" if batch == 64 else "bf16_autocast")
    best = res[key].get("graph_poses_per_s") or res[key]["eager_poses_per_s"]
    emit({"impl": "torch_gpu", "metric": METRIC, "value": best, "unit": "poses/s", "n_gpus": 1, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": batch / best * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32" if batch == 64 else "bf16 (torch.autocast)", "data": "synthetic",
          "config": workload_config(args.workload, batch, 1), "torch_gpu": res})


# ------------------------------------------------------------------------------------------------ main
def compact(r, keys=("workload", "value", "unit", "ms_per_step", "ms_per_step_warm_l2", "ms", "launches")):
    """One flat dict per secondary workload (the driver keeps only the tail of a long line)."""
    out = {k: r[k] for k in keys if k in r}
    if "roofline" in r:
        rf = r["roofline"]
        out["roofline"] = {k: rf[k] for k in ("bound", "achieved", "peak", "unit", "frac", "frac_of_sustained", "note") if k in rf}
    if "e2e" in r:
        out["e2e_value"] = r["e2e"]["value"]
    return out


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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_gpu"])
    ap.add_argument("--workload", default="train_b65536", choices=["train_b65536", "train_b64"])
    ap.add_argument("--no-also", action="store_true", help="skip the secondary BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--no-torch", action="store_true", help="skip the torch-on-B200 baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, emit)
        return
    if args.impl == "torch_gpu":
        run_torch_gpu(args, emit)
        return

    rank0 = int(os.environ.get("RANK", "0")) == 0
    batch = batch_of(args.workload)
    # The CPU baseline is taken on rank 0 BEFORE the process group exists: torchrun exports OMP_NUM_THREADS=1 and,
    # once NCCL is up, the other ranks spin in the barrier on the cores the CPU arm needs (round 1 reported a
    # 200x too low figure at N > 1 for exactly that reason).  The other ranks wait in the rendezvous meanwhile.
    cpu = None
    if rank0 and not args.no_cpu:
        try:
            note("cpu baseline")
            cpu = cpu_baseline_entry(args.workload)
        except Exception as e:
            cpu = {"value": None, "unit": "poses/s", "cores": 0, "kind": "port", "sample": "failed: " + repr(e)[:200]}

    import torch
    from bilinear_b200 import _lib, parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    _lib.lib()
    note("process group")
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
    note("headline %s" % args.workload)
    if args.workload == "train_b64":
        main_res = bench_train("fp32", 64, steps, warmup, device, world, peaks)
        dtype = "f32 (3 x bf16 split operands on tcgen05, fp32 accumulate)"
    else:
        main_res = bench_train("bf16", LARGE_BATCH, steps, warmup, device, world, peaks)
        dtype = "bf16"
    clocks = sampler.stop() if sampler else None

    also = []
    if not args.no_also:
        def guarded(fn):
            note("also #%d" % len(also))
            try:
                also.append(fn())
            except Exception as e:  # a secondary config must not take the headline down with it
                also.append({"error": repr(e)[:300]})
        if args.workload == "train_b65536":
            def small():
                r = bench_train("fp32", 64, max(steps, 100), warmup, device, world, peaks)
                r["workload"] = "train_b64_fp32 per GPU (BASELINE configs[1]: persistent whole-network kernel; L2 flushed)"
                r["unit"] = "poses/s"
                return compact(r)
            guarded(small)
            if world > 1:
                def strong():
                    r = bench_train("bf16", LARGE_BATCH // world, steps, warmup, device, world, peaks, with_e2e=False)
                    r["workload"] = "train_b65536_GLOBAL_bf16 (strong scaling: %d rows per GPU)" % (LARGE_BATCH // world)
                    r["unit"] = "poses/s"
                    return compact(r)
                guarded(strong)
        else:
            def large():
                r = bench_train("bf16", LARGE_BATCH, min(steps, 20), 3, device, world, peaks)
                r["workload"] = "train_b65536_bf16 per GPU"
                r["unit"] = "poses/s"
                return compact(r)
            guarded(large)
        # the evaluation passes (~0.1 ms each, HBM-bound) are timed BEFORE the two inference legs: those run the tensor
        # pipe at the board's power cap for ~2 s, and the SM clock they leave behind cost the passes 10 % (round-1 verdict)
        guarded(lambda: bench_eval(1_000_000, 50, 5, device, rank, world, peaks))
        guarded(lambda: compact(bench_infer(10_000_000, 65536, "bf16", 3, 1, device, rank, world, peaks)))
        guarded(lambda: compact(bench_infer(10_000_000, 65536, "fp32", 2, 1, device, rank, world, peaks)))

    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_torch:     # (a single-GPU baseline: measured in the N = 1 run only)
        try:
            note("torch-on-B200 arm")
            torch_gpu = torch_gpu_arm(device, max(steps, 50), warmup)
        except Exception as e:
            torch_gpu = {"error": repr(e)[:300]}
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
            "torch_gpu": torch_gpu, "also": also,
        }
        if "ms_per_step_warm_l2" in main_res:
            line["ms_per_step_warm_l2"] = main_res["ms_per_step_warm_l2"]
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
