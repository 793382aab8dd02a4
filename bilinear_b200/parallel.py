"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU
box, gloo in the CPU tests).  The reference is single-process; this is the new
data-parallel layer of SURVEY.md section 8(e):

  * training shards the batch row-wise; every rank holds a full replica; the ONE exchange step per
    iteration is an all-reduce(sum) of the flat 4 291 632-float gradient buffer, issued in six
    per-layer buckets in backward order; the division by world size and the global-norm clipping happen
    inside the fused clip+Adam kernel on the averaged gradient, identically on every rank.
    BatchNorm uses each rank's own shard statistics ("local-BN", what DistributedDataParallel does);
    its oracle is "reference per shard, gradients averaged".
  * inference and evaluation shard by pose index with no data-path collective; the <= 15 per-action
    (sum, count) pairs are combined with one tiny all-reduce at the end.
"""
import os

import torch
import torch.distributed as dist

# offsets of the per-layer buckets in the flat parameter order (encode | 4 hidden | decode)
_ENCODE = 32 * 1024 + 3 * 1024
_HIDDEN = 1024 * 1024 + 3 * 1024
_DECODE = 48 * 1024 + 48
BUCKETS = []
_o = 0
for _n in (_ENCODE, _HIDDEN, _HIDDEN, _HIDDEN, _HIDDEN, _DECODE):
    BUCKETS.append((_o, _o + _n))
    _o += _n
assert _o == 4291632


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n independent units (poses) owned by `rank`; sizes differ by <= 1."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(flat_g, group=None, bucketed=True):
    """Sum the flat gradient buffer over ranks (in place), decode bucket first -- the order in which
    backward finishes them.  Returns the async work handles' completion (blocking on the caller's
    stream only, not the host)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat_g
    if not bucketed:
        dist.all_reduce(flat_g, op=dist.ReduceOp.SUM, group=group)
        return flat_g
    works = []
    for lo, hi in reversed(BUCKETS):
        works.append(dist.all_reduce(flat_g[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    return flat_g


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier: lets torch view memory the library allocated."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerComm:
    """Gradient exchange buffers for the IN-KERNEL all-reduce (csrc/mega.cu: ew_allreduce): every rank
    cudaMalloc's its flat gradient buffer (+ control block) through the library, the 64-byte CUDA IPC handles
    travel through one torch.distributed all_gather, and every rank maps every peer's buffer.  torch only
    carries the handles; the data path is NVLink loads / stores issued by the persistent kernel."""

    def __init__(self, device, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._L = C, _lib.lib()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 8:
            raise RuntimeError("the in-kernel gradient exchange supports up to 8 ranks (one NVSwitch domain)")
        self.device = torch.device(device)
        nbytes = self._L.b200_comm_bytes()
        own = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            _lib.check(self._L.b200_comm_alloc(nbytes, C.byref(own), handle))
        self.own = own.value
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        if self.world > 1:
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().view(self.world, 64)
        else:
            allh = mine.cpu().view(1, 64)
        self.ptrs, self._opened = [], []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(self.own)
                continue
            hb = (C.c_uint8 * 64)(*[int(v) for v in allh[r]])
            p = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self._L.b200_comm_open(hb, C.byref(p)))
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        self.grads = torch.as_tensor(_RawCuda(self.own, _lib.PARAM_NUMEL), device=self.device)
        self.epoch = 1
        barrier()

    def desc(self):
        from ._lib import CommDesc
        d = CommDesc()
        d.rank, d.world, d.epoch = self.rank, self.world, self.epoch
        for r, p in enumerate(self.ptrs):
            d.buffers[r] = p
        return d

    def close(self):
        barrier()
        for p in self._opened:
            self._L.b200_comm_close(self._C.c_void_p(p))
        self._opened = []
        if self.own:
            self.grads = None
            self._L.b200_comm_free(self._C.c_void_p(self.own))
            self.own = None


def broadcast_module_state(module, engine, group=None, src=0):
    """Rank `src`'s parameters, buffers and dropout base seed -> every rank (no-op for a world of one)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
    dev = next(module.parameters()).device
    seed = torch.tensor([engine.seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
    dist.broadcast(seed, src=src, group=group)
    engine.seed = int(seed.item())
    engine.note_external_param_write()


def broadcast_int(value, device, group=None, src=0):
    """One Python int from rank `src` to all ranks (e.g. the seed of the data permutation)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return int(value)
    t = torch.tensor([int(value) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
    dist.broadcast(t, src=src, group=group)
    return int(t.item())


def combine_action_sums(sums, psums, counts, group=None):
    """valid_bilinear.py:61-70 across ranks: per-action fp64 sums and int64 counts add up."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, group=group)
        dist.all_reduce(psums, group=group)
        dist.all_reduce(counts, group=group)
    return sums, psums, counts


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
