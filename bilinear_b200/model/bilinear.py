"""Drop-in for the reference's model/bilinear.py: same names, same no-argument constructor, same
37-key state_dict, same factory signature and checkpoint format -- with the arithmetic executed by
the sm_100a library instead of torch ops.

    reference                         here
    model/bilinear.py:7-13            heavy_linear(in_features, out_features, bias=True)
    model/bilinear.py:16-41           BilinearUnit()  .forward(in_tensor)
    model/bilinear.py:43-55           BilinearUnit.reset_statistics()
    model/bilinear.py:58-92           load(device, parameter_dir=None, learning_rate=1.0e-3)
                                        -> (module, Adam-compatible optimizer, step, epoch_to_load)

The sub-modules stay real nn.Linear / nn.BatchNorm1d / nn.ReLU / nn.Dropout instances at the same
attribute paths (load() relies on isinstance(m, nn.Linear); reset_statistics on layer[int] and
reset_running_stats()), but they only HOLD the parameters and buffers; forward never calls them.
"""
import os

import torch
import torch.nn as nn

from ..engine import default_precision, get_engine
from ..optim import FusedAdam


def heavy_linear(in_features, out_features, bias=True):
    # model/bilinear.py:7-13 (construction order matters: it fixes the RNG stream of the default init)
    return nn.Sequential(
        nn.Linear(in_features, out_features, bias=bias),
        nn.BatchNorm1d(out_features),
        nn.ReLU(),
        nn.Dropout(p=0.5),
    )


class _LiftFunction(torch.autograd.Function):
    """Training-mode forward/backward of the whole network as two library calls.  Gradients are
    written straight into the engine's flat gradient buffer and exposed as param.grad views."""

    @staticmethod
    def forward(ctx, x, module, *params):
        eng = get_engine(module)
        ctx.module = module
        return eng.forward_train(x)

    @staticmethod
    def backward(ctx, dP):
        module = ctx.module
        eng = get_engine(module)
        params = list(module.parameters())
        views = eng.grad_views()
        # accumulate semantics if a gradient is already present (the reference always zero_grad()s first)
        stash = None
        if any(p.grad is not None for p in params):
            stash = [None if p.grad is None else p.grad.detach().clone() for p in params]
        eng.backward(dP)
        for i, (p, v) in enumerate(zip(params, views)):
            if stash is not None and stash[i] is not None:
                v.add_(stash[i])
            p.grad = v
        return (None, None) + (None,) * len(params)


class BilinearUnit(nn.Module):
    def __init__(self):
        super().__init__()

        num_joint = 17 - 1

        self.encode = heavy_linear(in_features=2 * num_joint, out_features=1024)
        self.bilinear = nn.ModuleList([
            nn.Sequential(
                heavy_linear(in_features=1024, out_features=1024),
                heavy_linear(in_features=1024, out_features=1024),
            ) for _ in range(2)
        ])
        self.decode = nn.Linear(in_features=1024, out_features=3 * num_joint, bias=True)
        # additional knob, reference-equal default: "fp32" (3 x bf16 split, 1e-5 parity) | "bf16"
        self.__dict__["precision"] = default_precision()

    def set_precision(self, precision):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.__dict__["precision"] = precision
        return self

    def inject_dropout_masks(self, masks):
        """Parity testing: use these keep masks ([5, B, 1024] uint8 on the device, 1 = keep) instead of
        the in-kernel Philox stream for subsequent training-mode forwards; None restores Philox."""
        get_engine(self).masks = masks

    def forward(self, in_tensor):
        eng = get_engine(self)
        if not self.training:
            # valid_bilinear.py:31,52 -- BN running statistics, dropout identity
            return eng.forward_eval(in_tensor)
        if in_tensor.requires_grad:
            raise RuntimeError("gradients with respect to the 2D input are not part of the hot path")
        if torch.is_grad_enabled():
            return _LiftFunction.apply(in_tensor, self, *self.parameters())
        return eng.forward_train(in_tensor)

    def reset_statistics(self):
        """model/bilinear.py:43-55: every BatchNorm layer forgets its running statistics and switches to
        the cumulative moving average (momentum=None => factor 1/num_batches_tracked)."""
        for layer in self.modules():
            if isinstance(layer, nn.BatchNorm1d):
                layer.reset_running_stats()
                layer.momentum = None

    # ---- fused fast path (what train_bilinear.py of this package drives)
    def train_step(self, optimizer, in_image_space, in_camera_space, max_norm=1.0, update=True, grad_scale=1.0):
        """zero_grad + forward + nn.MSELoss + backward + clip_grad_norm_(max_norm) + optimizer.step()
        (train_bilinear.py:75-83) as ONE persistent kernel.  Returns (loss, grad_norm) device scalars."""
        if not self.training:
            raise RuntimeError("train_step() needs the module in training mode")
        return optimizer.fused_step(self, in_image_space, in_camera_space, max_norm, update, grad_scale)


def _train_step_dp(self, optimizer, in_image_space, in_camera_space, max_norm=1.0):
    """Data-parallel variant of train_step(): this rank's rows of the global batch in, gradients averaged over
    all ranks INSIDE the kernel (NVLink peer memory), identical parameters on every rank out.  Needs
    `bilinear_b200.parallel.PeerComm` attached (see enable_data_parallel)."""
    if not self.training:
        raise RuntimeError("train_step_dp() needs the module in training mode")
    eng = get_engine(self)
    return optimizer.fused_step_dp(self, in_image_space, in_camera_space, max_norm, eng.comm.world)


def _enable_data_parallel(self, group=None):
    """Allocate and exchange the peer-mapped gradient buffers (once, after .to(device))."""
    from .. import parallel
    eng = get_engine(self)
    dev = next(self.parameters()).device
    if not dev.type == "cuda":
        raise RuntimeError("bilinear_b200 runs on CUDA (sm_100a) only: move the module to a CUDA device first")
    eng.attach_comm(parallel.PeerComm(dev, group))
    # Replicas must start identical and stay identical (what DistributedDataParallel does in its constructor):
    # rank 0's parameters, BatchNorm buffers and Philox base seed go to every rank, so a per-rank torch seed
    # cannot silently fork the model.  The dropout stream itself is keyed per rank (Engine.train_step_dp).
    parallel.broadcast_module_state(self, eng, group)
    return self


BilinearUnit.train_step_dp = _train_step_dp
BilinearUnit.enable_data_parallel = _enable_data_parallel


def _train_step_host(self, optimizer, in_image_space, in_camera_space, max_norm=1.0, update=True, grad_scale=1.0):
    """The body of the loop train_bilinear.py:62-88 INCLUDING its transfers, for CPU (pinned) batches as a
    DataLoader(pin_memory=True) delivers them: host->device copies, the one-kernel training iteration and
    the loss read-back are one library call.  Returns (loss, grad_norm) as Python floats."""
    if not self.training:
        raise RuntimeError("train_step_host() needs the module in training mode")
    if not next(self.parameters()).is_cuda:
        raise RuntimeError("bilinear_b200 runs on CUDA (sm_100a) only: move the module to a CUDA device")
    return optimizer.fused_step_host(self, in_image_space, in_camera_space, max_norm, update, grad_scale)


BilinearUnit.train_step_host = _train_step_host


def _latest_epoch(parameter_dir):
    """Largest N among the '{N}.save' files below parameter_dir (0: none).  Like the reference
    (model/bilinear.py:63-72) any file whose name is not '<int>.<ext>' raises ValueError."""
    latest = 0
    if parameter_dir is None:
        return latest
    for _, _, files in os.walk(parameter_dir):
        for file in files:
            stem, _ext = file.split('.')
            latest = max(latest, int(stem))
    return latest


def load(device, parameter_dir=None, learning_rate=1.0e-3):
    """model/bilinear.py:58-92: build module + optimizer, resume from the newest checkpoint if there
    is one, otherwise Kaiming-normal initialise every nn.Linear weight (biases and BatchNorm keep
    their defaults).  Returns (module, optimizer, step, epoch_to_load)."""
    bilinear = BilinearUnit().to(device)
    optimizer = FusedAdam(bilinear, lr=learning_rate)
    step = 1
    epoch_to_load = _latest_epoch(parameter_dir)
    if epoch_to_load != 0:
        parameter = torch.load('{0}/{1}.save'.format(parameter_dir, epoch_to_load), map_location=device)
        bilinear.load_state_dict(parameter['state'])
        optimizer.load_state_dict(parameter['optimizer'])
        step = parameter['step']
    else:
        bilinear.apply(_kaiming_on_linear)
    return bilinear, optimizer, step, epoch_to_load


def _kaiming_on_linear(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight)
