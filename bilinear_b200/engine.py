"""Host-side engine under the drop-in nn.Module: owns the flat parameter / gradient / optimizer
buffers the C ABI works on, caches one plan per (batch, precision, kind), and keeps the bf16 operand
planes in sync with the fp32 master parameters.

PyTorch is used for device memory and streams only; every arithmetic step of the hot path runs in
bilinear_b200/lib/libbilinear_b200.so (csrc/mega.cu).  There is no CPU or eager fallback: a CPU
tensor, a missing library or a failing call raises.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import PLAN_INFER, PLAN_TRAIN, PREC_BF16, PREC_FP32, AdamArgs, PlanDesc, TrainArgs, check, ptr

PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16}


def default_precision():
    """Reference-equal default is the fp32-parity mode; BILINEAR_B200_PRECISION=bf16 selects bf16."""
    return os.environ.get("BILINEAR_B200_PRECISION", "fp32")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Plan:
    def __init__(self, handle, workspace, batch, precision, kind):
        self.handle = handle
        self.workspace = workspace
        self.batch = batch
        self.precision = precision
        self.kind = kind
        self.weights_version = -1

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().b200_plan_destroy(self.handle)
        except Exception:
            pass


class Engine:
    """One per BilinearUnit instance (created lazily on the first CUDA call)."""

    def __init__(self, module):
        self.module = module
        self.device = None
        self.flat_p = self.flat_g = self.flat_m = self.flat_v = None
        self.plans = {}
        self.shared_ws = {}
        self.weights_version = 0      # bumped whenever the fp32 master parameters change
        self._seen_torch_version = None
        self.masks = None             # injected dropout keep masks [5][B][1024] uint8 (tests / parity)
        self.seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.fwd_counter = 0
        self.pending_backward = None  # batch size of the forward whose activations sit in the workspace
        self.scalars = None
        self.host_out = None          # pinned [loss, grad norm] of the *_host entry points
        self._bns = None
        self._plist = None
        self.comm = None              # parallel.PeerComm: in-kernel data-parallel gradient exchange

    # ------------------------------------------------------------------ flat buffers
    def _params(self):
        # the 22 Parameter OBJECTS never change (only their .data is re-pointed): walk the module tree once
        if self._plist is None:
            self._plist = list(self.module.parameters())
        return self._plist

    def _is_flat(self, thorough=False):
        if self.flat_p is None:
            return False
        base = self.flat_p.data_ptr()
        params = self._params()
        if not thorough:
            # hot path: .to() / load_state_dict / a foreign optimizer re-point all parameters or none
            last = params[-1]
            return (params[0].data_ptr() == base and
                    last.data_ptr() == base + 4 * (_lib.PARAM_NUMEL - last.numel()))
        off = 0
        for p in params:
            if p.data_ptr() != base + 4 * off or p.device != self.flat_p.device or p.dtype != torch.float32:
                return False
            off += p.numel()
        return True

    def ensure_flat(self):
        params = self._params()
        if not params[0].is_cuda:
            raise RuntimeError("bilinear_b200 runs on CUDA (sm_100a) only: move the module to a CUDA device "
                               "(no CPU fallback exists)")
        if self._is_flat():
            return
        dev = params[0].device
        n = sum(p.numel() for p in params)
        assert n == _lib.PARAM_NUMEL, n
        flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        if self.comm is not None:
            if self.comm.device != dev:
                raise RuntimeError("the gradient exchange buffers live on %s, the module on %s" % (self.comm.device, dev))
            flat_g = self.comm.grads          # gradients are written straight into the peer-mapped buffer
            flat_g.zero_()
        else:
            flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                flat_p[off:off + k].copy_(p.detach().reshape(-1).float())
                p.data = flat_p[off:off + k].view(p.shape)
                p.grad = None
                off += k
        if self.flat_m is None or self.flat_m.device != dev:
            self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
            self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_p, self.flat_g, self.device = flat_p, flat_g, dev
        self.plans.clear()
        self.shared_ws.clear()
        self.scalars = torch.zeros(8, dtype=torch.float32, device=dev)
        self.weights_version += 1
        self._seen_torch_version = self._torch_versions()

    def attach_comm(self, comm):
        """Data parallel: from now on the flat gradient buffer is `comm`'s peer-mapped buffer."""
        self.comm = comm
        self.flat_p = None            # re-flatten (and rebuild the plans) on the next call
        self.plans.clear()

    def grad_views(self):
        out, off = [], 0
        for p in self._params():
            k = p.numel()
            out.append(self.flat_g[off:off + k].view(p.shape))
            off += k
        return out

    def state_views(self):
        m, v, off = [], [], 0
        for p in self._params():
            k = p.numel()
            m.append(self.flat_m[off:off + k].view(p.shape))
            v.append(self.flat_v[off:off + k].view(p.shape))
            off += k
        return m, v

    def note_external_param_write(self):
        self.weights_version += 1

    def _torch_versions(self):
        return sum(p._version for p in self._params())

    def _sync_weight_version(self):
        # in-place torch writes to a parameter (load_state_dict, nn.init, a foreign optimizer) bump its
        # autograd version counter; our own kernels do not.  Writes through `.data` are invisible to
        # this check: call note_external_param_write() after those.
        v = self._torch_versions()
        if v != self._seen_torch_version:
            self._seen_torch_version = v
            self.weights_version += 1

    # ------------------------------------------------------------------ plans
    def _bn_layers(self):
        if self._bns is None:
            m = self.module
            self._bns = [m.encode[1], m.bilinear[0][0][1], m.bilinear[0][1][1], m.bilinear[1][0][1],
                         m.bilinear[1][1][1]]
        return self._bns

    def plan(self, batch, kind):
        self.ensure_flat()
        precision = PRECISIONS[self.module.precision]
        bns = self._bn_layers()
        key = (batch, precision, kind, tuple(b.running_mean.data_ptr() for b in bns))
        pl = self.plans.get(key)
        L = _lib.lib()
        if pl is None:
            for b in bns:
                if b.running_mean is None or not b.running_mean.is_cuda:
                    raise RuntimeError("BatchNorm buffers must live on the CUDA device")
            need = L.b200_workspace_bytes(batch, precision, kind)
            if need < 0:
                raise RuntimeError("bilinear_b200: bad plan parameters")
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            off = (-ws.data_ptr()) % 1024
            d = PlanDesc()
            d.batch, d.precision, d.kind = batch, precision, kind
            d.params = self.flat_p.data_ptr()
            d.grads = self.flat_g.data_ptr()
            d.exp_avg = self.flat_m.data_ptr()
            d.exp_avg_sq = self.flat_v.data_ptr()
            for i, b in enumerate(bns):
                d.running_mean[i] = b.running_mean.data_ptr()
                d.running_var[i] = b.running_var.data_ptr()
                d.num_batches_tracked[i] = b.num_batches_tracked.data_ptr()
            d.workspace = ws.data_ptr() + off
            d.workspace_bytes = need
            h = C.c_void_p()
            check(L.b200_plan_create(C.byref(d), C.byref(h)))
            pl = _Plan(h, ws, batch, precision, kind)
            if self.comm is not None and kind == PLAN_TRAIN:
                cd = self.comm.desc()
                check(L.b200_plan_set_comm(h, C.byref(cd)))
            if len(self.plans) > 8:      # ragged tail batches come and go; keep the cache bounded
                self.plans.pop(next(iter(self.plans)))
            self.plans[key] = pl
        self._sync_weight_version()
        if pl.weights_version != self.weights_version:
            check(L.b200_refresh_weights(pl.handle, _stream()))
            pl.weights_version = self.weights_version
        return pl

    def _momenta(self):
        return [(-1.0 if b.momentum is None else float(b.momentum)) for b in self._bn_layers()]

    # ------------------------------------------------------------------ calls
    @staticmethod
    def _check_input(x, cols):
        if not x.is_cuda:
            raise RuntimeError("bilinear_b200 runs on CUDA (sm_100a) only: got a CPU tensor (no CPU fallback exists)")
        if x.dim() != 2 or x.shape[1] != cols:
            raise RuntimeError("expected a [B, %d] tensor, got %s" % (cols, list(x.shape)))
        return x.detach().contiguous().float()

    def forward_eval(self, x):
        x = self._check_input(x, 32)
        B = x.shape[0]
        kind = PLAN_TRAIN if (B, PRECISIONS[self.module.precision], PLAN_TRAIN) in {k[:3] for k in self.plans} else PLAN_INFER
        pl = self.plan(B, kind)
        p = torch.empty(B, 48, dtype=torch.float32, device=x.device)
        check(_lib.lib().b200_forward_eval(pl.handle, ptr(x), ptr(p), _stream()))
        return p

    def _train_args(self, x, y, p_out, B):
        t = TrainArgs()
        t.x = x.data_ptr() if x is not None else None
        t.y = y.data_ptr() if y is not None else None
        t.p_out = p_out.data_ptr() if p_out is not None else None
        if self.masks is not None:
            m = self.masks
            if tuple(m.shape) != (5, B, 1024) or m.dtype != torch.uint8 or not m.is_cuda:
                raise RuntimeError("injected dropout masks must be a CUDA uint8 tensor of shape [5, B, 1024]")
            t.masks = m.contiguous().data_ptr()
        t.seed = self.seed
        t.step = self.fwd_counter
        for i, mom in enumerate(self._momenta()):
            t.momentum[i] = mom
        return t

    def forward_train(self, x):
        x = self._check_input(x, 32)
        B = x.shape[0]
        if B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                             % str(torch.Size([B, 1024])))
        pl = self.plan(B, PLAN_TRAIN)
        p = torch.empty(B, 48, dtype=torch.float32, device=x.device)
        t = self._train_args(x, None, p, B)
        check(_lib.lib().b200_forward_train(pl.handle, C.byref(t), _stream()))
        self.fwd_counter += 1
        self.pending_backward = B
        return p

    def backward(self, dP):
        B = self.pending_backward
        if B is None or dP.shape[0] != B:
            raise RuntimeError("backward() without a matching training-mode forward")
        pl = self.plan(B, PLAN_TRAIN)
        dP = dP.detach().contiguous().float()
        check(_lib.lib().b200_backward(pl.handle, ptr(dP), _stream()))
        self.pending_backward = None

    def adam_args(self, lr, step, max_norm, grad_scale=1.0, betas=(0.9, 0.999), eps=1e-8):
        a = AdamArgs()
        a.lr, a.beta1, a.beta2, a.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        a.step = int(step)
        a.max_norm = float(max_norm) if max_norm else 0.0
        a.grad_scale = float(grad_scale)
        return a

    def train_step(self, x, y, adam, update=True, p_out=None):
        """One iteration of train_bilinear.py:75-83 as one persistent kernel.  Returns device scalars
        (loss, pre-clip gradient norm) without synchronising."""
        x = self._check_input(x, 32)
        y = self._check_input(y, 48)
        B = x.shape[0]
        if B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                             % str(torch.Size([B, 1024])))
        pl = self.plan(B, PLAN_TRAIN)
        t = self._train_args(x, y, p_out, B)
        t.loss_out = self.scalars[0:1].data_ptr()
        t.gnorm_out = self.scalars[1:2].data_ptr()
        check(_lib.lib().b200_train_step(pl.handle, C.byref(t), C.byref(adam), 1 if update else 0, _stream()))
        self.fwd_counter += 1
        self.pending_backward = None
        if update:
            pl.weights_version = self.weights_version = self.weights_version + 1
        return self.scalars[0], self.scalars[1]

    def train_step_dp(self, x, y, adam):
        """Data-parallel iteration as ONE kernel per rank: forward, backward, all-reduce of the flat gradient
        over NVLink peer memory, clip + Adam on the average.  Every rank must call it for the same step."""
        if self.comm is None:
            raise RuntimeError("attach_comm(parallel.PeerComm(...)) first")
        x = self._check_input(x, 32)
        y = self._check_input(y, 48)
        B = x.shape[0]
        if B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                             % str(torch.Size([B, 1024])))
        pl = self.plan(B, PLAN_TRAIN)
        t = self._train_args(x, y, None, B)
        # every rank draws ITS OWN dropout bits: the Philox key carries the rank (the counter carries step,
        # layer, feature and local batch word), so row i of two shards never shares a keep mask
        t.seed = (self.seed + 0x9E3779B97F4A7C15 * (self.comm.rank + 1)) & 0xFFFFFFFFFFFFFFFF
        t.loss_out = self.scalars[0:1].data_ptr()
        t.gnorm_out = self.scalars[1:2].data_ptr()
        ep = C.c_uint32(self.comm.epoch)
        check(_lib.lib().b200_train_step_dp(pl.handle, C.byref(t), C.byref(adam), C.byref(ep), _stream()))
        self.comm.epoch = ep.value
        self.fwd_counter += 1
        self.pending_backward = None
        pl.weights_version = self.weights_version = self.weights_version + 1
        return self.scalars[0], self.scalars[1]

    def train_step_host(self, x_host, y_host, adam, update=True):
        """train_bilinear.py:62-88 with its transfers: x_host [B, 32] / y_host [B, 48] are CPU tensors
        (pinned for asynchronous copies); copies in, the one-kernel step, and the (loss, grad norm) written
        back to pinned host memory are enqueued by ONE library call; returns the two Python floats after
        synchronising the stream (the reference reads loss.item() every step as well)."""
        if x_host.is_cuda or y_host.is_cuda:
            raise RuntimeError("train_step_host() takes CPU tensors; use train_step() for CUDA tensors")
        B = x_host.shape[0]
        if x_host.dim() != 2 or x_host.shape[1] != 32 or y_host.dim() != 2 or y_host.shape[0] != B or y_host.shape[1] != 48:
            raise RuntimeError("expected x [B, 32] and y [B, 48]")
        if x_host.dtype != torch.float32 or y_host.dtype != torch.float32 or not x_host.is_contiguous() \
                or not y_host.is_contiguous():
            x_host, y_host = x_host.float().contiguous(), y_host.float().contiguous()
        if B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                             % str(torch.Size([B, 1024])))
        pl = self.plan(B, PLAN_TRAIN)
        if self.host_out is None:
            self.host_out = torch.zeros(2, dtype=torch.float32).pin_memory()
            self.host_out_np = self.host_out.numpy()
        t = self._train_args(None, None, None, B)
        stream = torch.cuda.current_stream()
        check(_lib.lib().b200_train_step_host(pl.handle, x_host.data_ptr(), y_host.data_ptr(), C.byref(t), C.byref(adam),
                                              1 if update else 0, self.host_out.data_ptr(), stream.cuda_stream))
        self.fwd_counter += 1
        self.pending_backward = None
        if update:
            pl.weights_version = self.weights_version = self.weights_version + 1
        stream.synchronize()
        return float(self.host_out_np[0]), float(self.host_out_np[1])

    def forward_eval_host(self, x_host, out_host=None):
        """valid_bilinear.py:39-52 with its transfers: CPU x [B, 32] -> CPU prediction [B, 48] (pinned)."""
        if x_host.is_cuda:
            raise RuntimeError("forward_eval_host() takes a CPU tensor")
        x_host = x_host.float().contiguous()
        B = x_host.shape[0]
        self.ensure_flat()
        kind = PLAN_TRAIN if (B, PRECISIONS[self.module.precision], PLAN_TRAIN) in {k[:3] for k in self.plans} else PLAN_INFER
        pl = self.plan(B, kind)
        if out_host is None:
            out_host = torch.empty(B, 48, dtype=torch.float32).pin_memory()
        stream = torch.cuda.current_stream()
        check(_lib.lib().b200_forward_eval_host(pl.handle, C.c_void_p(x_host.data_ptr()),
                                                C.c_void_p(out_host.data_ptr()), C.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        return out_host

    def adam_step(self, adam, batch_hint=None):
        """clip_grad_norm_ + Adam.step on the flat gradient buffer (train_bilinear.py:81-83)."""
        self.ensure_flat()
        pl = None
        for k, v in self.plans.items():
            if k[2] == PLAN_TRAIN and (batch_hint is None or k[0] == batch_hint):
                pl = v
        if pl is None:
            pl = self.plan(batch_hint or 64, PLAN_TRAIN)
        self._sync_weight_version()
        check(_lib.lib().b200_adam_step(pl.handle, C.byref(adam), self.scalars[1:2].data_ptr(), _stream()))
        pl.weights_version = self.weights_version = self.weights_version + 1
        return self.scalars[1]


    def debug_phase_ns(self, batch, kind=PLAN_TRAIN):
        """[(duration_ns, kind, tiles)] per phase of the last launch on the plan for `batch`."""
        pl = self.plan(batch, kind)
        n = 160
        ns = (C.c_int64 * (8 * n))()
        kinds = (C.c_int32 * n)()
        tiles = (C.c_int32 * n)()
        cnt = C.c_int32(0)
        check(_lib.lib().b200_debug_phase_ns(pl.handle, ns, kinds, tiles, n, C.byref(cnt)))
        return [(int(ns[8 * i]), int(kinds[i]), int(tiles[i]), [int(ns[8 * i + k]) for k in range(1, 8)])
                for i in range(cnt.value)]

    def debug_keep_bits(self, layer, batch):
        """[batch, 1024] bool: dropout keep AND (BatchNorm output > 0) of the last training forward."""
        pl = self.plan(batch, PLAN_TRAIN)
        ldw = (batch + 31) // 32
        words = torch.empty(1024, ldw, dtype=torch.int32, device=self.device)
        check(_lib.lib().b200_debug_keep_bits(pl.handle, layer, ptr(words), _stream()))
        bits = (words.unsqueeze(-1) >> torch.arange(32, device=self.device, dtype=torch.int32)) & 1
        return bits.reshape(1024, ldw * 32)[:, :batch].t().contiguous().bool()


def get_engine(module):
    eng = module.__dict__.get("_b200_engine")
    if eng is None:
        eng = Engine(module)
        module.__dict__["_b200_engine"] = eng
    return eng
