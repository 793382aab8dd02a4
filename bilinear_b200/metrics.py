"""H36M evaluation math of the reference's valid_bilinear.py on the GPU:

    valid_bilinear.py:53-60   un-normalise with the train-set mean/std, per-pose sum over joints of the
                              Euclidean distance (mm)
    valid_bilinear.py:61-70   per-action accumulation in float64 -- the reference does one `.cpu()`
                              round trip per pose; here it is a device-side segmented sum, one D2H at the end
    valid_bilinear.py:76-83   MPJPE_a = total_dist[a] / (count_a * J); average over all poses

plus the Procrustes-aligned variant (P-MPJPE; not in the reference).  All
arithmetic runs in libbilinear_b200.so (csrc/eval.cu); there is no torch / CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import check, ptr


@dataclass
class EvalResult:
    dist: torch.Tensor               # [N] fp32 per-pose sum over joints (valid_bilinear.py:60)
    pdist: torch.Tensor              # [N] fp32 same after Procrustes alignment (None if not computed)
    sum: torch.Tensor                # [n_actions] fp64
    psum: torch.Tensor               # [n_actions] fp64
    count: torch.Tensor              # [n_actions] int64
    joints: int

    @property
    def per_action_mpjpe(self):
        return self.sum / (self.count.double() * self.joints)

    @property
    def per_action_pmpjpe(self):
        return self.psum / (self.count.double() * self.joints)

    @property
    def mpjpe(self):
        c = int(self.count.sum())
        return float(self.sum.sum()) / (c * self.joints) if c else float("nan")

    @property
    def pmpjpe(self):
        c = int(self.count.sum())
        return float(self.psum.sum()) / (c * self.joints) if c else float("nan")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(t, name, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("bilinear_b200.metrics runs on CUDA (sm_100a) only: %s is a CPU tensor "
                           "(no CPU fallback exists)" % name)
    return t.detach().to(dtype).contiguous()


def evaluate(pred, gt, mean=None, stddev=None, action=None, n_actions=1, joints=16, fused=True, accum=None,
             want_pmpjpe=True, want_mpjpe=True, keep_per_pose=True):
    """pred, gt: [N, 3*joints] (z-scored if mean/stddev [3*joints] are given, else mm).  action: [N] int32
    bucket ids or None.  `accum` = (sum, psum, count) device tensors to keep accumulating into across
    batches (what the loop of valid_bilinear.py:38-70 does); otherwise fresh zeroed accumulators."""
    pred = _prep(pred, "pred")
    gt = _prep(gt, "gt")
    if pred.dim() != 2 or pred.shape != gt.shape or pred.shape[1] != 3 * joints or joints not in (16, 17):
        raise RuntimeError("expected two [N, %d] tensors with joints in {16, 17}; got %s / %s"
                           % (3 * joints, list(pred.shape), list(gt.shape)))
    mean = _prep(mean, "mean")
    stddev = _prep(stddev, "stddev")
    if (mean is None) != (stddev is None):
        raise RuntimeError("mean and stddev must be given together")
    if mean is not None and (mean.numel() != 3 * joints or stddev.numel() != 3 * joints):
        raise RuntimeError("mean / stddev must have %d elements (the train-set statistics)" % (3 * joints))
    action = _prep(action, "action", torch.int32)
    n = pred.shape[0]
    dev = pred.device
    if accum is None:
        acc = torch.zeros(n_actions, dtype=torch.float64, device=dev)
        pacc = torch.zeros(n_actions, dtype=torch.float64, device=dev)
        cnt = torch.zeros(n_actions, dtype=torch.int64, device=dev)
    else:
        acc, pacc, cnt = accum
    dist = torch.empty(n, dtype=torch.float32, device=dev) if (keep_per_pose and want_mpjpe) else None
    pdist = torch.empty(n, dtype=torch.float32, device=dev) if (keep_per_pose and want_pmpjpe) else None
    L = _lib.lib()
    if n > 0:
        if want_mpjpe and want_pmpjpe and fused:
            check(L.b200_mpjpe_pmpjpe(ptr(pred), ptr(gt), ptr(mean), ptr(stddev), n, joints, ptr(action), n_actions,
                                      ptr(dist), ptr(pdist), ptr(acc), ptr(pacc), ptr(cnt), _stream()))
        else:
            if want_mpjpe:
                check(L.b200_mpjpe(ptr(pred), ptr(gt), ptr(mean), ptr(stddev), n, joints, ptr(action), n_actions,
                                   ptr(dist), ptr(acc), ptr(cnt), _stream()))
            if want_pmpjpe:
                check(L.b200_pmpjpe(ptr(pred), ptr(gt), ptr(mean), ptr(stddev), n, joints, ptr(action), n_actions,
                                    ptr(pdist), ptr(pacc), None if want_mpjpe else ptr(cnt), _stream()))
    return EvalResult(dist, pdist, acc, pacc, cnt, joints)
