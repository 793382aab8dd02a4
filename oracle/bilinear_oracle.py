"""ORACLE (test infrastructure, never on the product path): CPU restatement of the lifting-MLP hot
path of nulledge/bilinear, written as explicit tensor algebra so that every intermediate the CUDA
kernels fuse away can be inspected.

The reference composes this algorithm from PyTorch modules (PyTorch is a third-party dependency the
reference does not vendor or pin; its notebook metadata suggests torch 0.4.1, Sept 2018 -- the
semantics below are those of torch 2.11.0, the version installed here, which is what the golden
vectors under tests/golden/ were generated with by tests/golden/make_golden.py running the
UNMODIFIED /root/reference/model/bilinear.py):

    model/bilinear.py:7-13    heavy_linear = Linear -> BatchNorm1d -> ReLU -> Dropout(0.5)
    model/bilinear.py:16-29   BilinearUnit: encode 32->1024, 2 x [heavy, heavy] + skip, decode 1024->48
    model/bilinear.py:31-41   forward (residual add after the second dropout, no activation after it)
    train_bilinear.py:49,78   nn.MSELoss()           mean over B*48 elements
    train_bilinear.py:79      loss.backward()        (formulas: SURVEY.md Appendix A)
    train_bilinear.py:81      clip_grad_norm_(params, max_norm=1)   torch/nn/utils/clip_grad.py
    train_bilinear.py:83      Adam.step()            torch/optim/adam.py (_single_tensor_adam)
    util/config.py:21-22      lr(step) = 1e-3 * 0.96 ** (step / 100000), refreshed when
                              step % 100000 == 0 or step == 1 (train_bilinear.py:66-70)

Pinned: tests/test_oracle_golden.py checks this file against every golden vector (forward, loss, all
22 gradients, gradient norm, running statistics, Adam state and parameters after one step, eval-mode
forward) for seeds {0,1,2} x batch {2,64,257}.

dtype is a parameter: float64 arbitrates tolerances, float32 (with all host threads) is what bench.py
times as the CPU baseline ("port").

bf16 mode (a NEW capability, the reference has no reduced-precision mode): `Rounding.bf16()` restates
what "bf16 mode" means in this build -- every GEMM operand (activations, weights, upstream gradients)
is rounded to bfloat16 (round-to-nearest-even), products are accumulated in full precision, everything
else (bias, BatchNorm, ReLU, dropout, loss, optimizer, master weights) stays in `dtype`; the pre-BN
activation kept for backward is the bf16-rounded one.  Against this restatement the CUDA bf16 path
agrees to fp32 round-off; against the unrounded oracle it agrees to the 2e-2 of the north star in the
forward quantities, while per-tensor gradients differ by ~sqrt(fraction of ReLU decisions flipped by
the 4e-3 operand rounding) ~ 1e-1 -- a property of ANY bf16 evaluation of this network: torch.autocast
(bfloat16) on the UNMODIFIED reference module against its own fp32 run shows 9.7e-2 / 1.0e-1 worst per-tensor
gradient error at cosine > 0.99 (tests/golden/make_autocast_golden.py -> tests/golden/autocast_bf16_study.json,
pinned by tests/test_oracle_golden.py::test_bf16_gradient_error_is_a_property_of_the_network).

ReLU is discontinuous in its derivative: an element whose BatchNorm output y is within round-off of 0
may legitimately be treated as active or inactive by a different evaluation order (the fp32 reference
against this fp64 restatement included).  `ambiguous_relu` lists such elements and `backward` accepts
explicit overrides so that tests can admit either decision for them.
"""
import math

import torch

HIDDEN = 1024
DIN = 32
DOUT = 48
NUM_HEAVY = 5
BN_EPS = 1e-5



class Rounding:
    """Where values are rounded to bfloat16 (None = nowhere: the reference's fp32 semantics)."""

    def __init__(self, operands=False, saved_z=False, saved_dy=False, fwd_z=False):
        self.operands, self.saved_z, self.saved_dy, self.fwd_z = operands, saved_z, saved_dy, fwd_z

    @staticmethod
    def bf16(large_batch=False):
        # large_batch (B > 256): BatchNorm needs the whole column before any element can be normalised,
        # so z (forward) and dy (backward) make one round trip through HBM between the two passes -- in
        # bf16; the normalisation then starts from the rounded z.  The gradient kept for a residual edge
        # makes the same round trip (it doubles as the stored dy of its layer)
        return Rounding(operands=True, saved_z=True, saved_dy=large_batch, fwd_z=large_batch)

    @staticmethod
    def _r(t):
        return t.to(torch.bfloat16).to(t.dtype)

    def op(self, t):
        return self._r(t) if self.operands else t

    def z(self, t):
        return self._r(t) if self.saved_z else t

    def dy(self, t):
        return self._r(t) if self.saved_dy else t


EXACT = Rounding()

# state_dict keys of the reference module, in order (37 entries)
HEAVY_PREFIXES = ["encode", "bilinear.0.0", "bilinear.0.1", "bilinear.1.0", "bilinear.1.1"]


def param_names():
    """Order of BilinearUnit.parameters() == order of torch.optim.Adam's param indices."""
    names = []
    for pre in HEAVY_PREFIXES:
        names += [pre + ".0.weight", pre + ".0.bias", pre + ".1.weight", pre + ".1.bias"]
    names += ["decode.weight", "decode.bias"]
    return names


def param_shapes():
    shapes = {}
    for i, pre in enumerate(HEAVY_PREFIXES):
        k = DIN if i == 0 else HIDDEN
        shapes[pre + ".0.weight"] = (HIDDEN, k)
        shapes[pre + ".0.bias"] = (HIDDEN,)
        shapes[pre + ".1.weight"] = (HIDDEN,)
        shapes[pre + ".1.bias"] = (HIDDEN,)
    shapes["decode.weight"] = (DOUT, HIDDEN)
    shapes["decode.bias"] = (DOUT,)
    return shapes


def lr_schedule(step):
    """util/config.py:22"""
    return 1.0e-3 * 0.96 ** (step / 100000)


def lr_refresh_due(step):
    """util/config.py:21"""
    return step % 100000 == 0 or step == 1


class State:
    """Parameters, BatchNorm buffers and Adam state as plain tensors of one dtype."""

    def __init__(self, state_dict, dtype=torch.float64):
        self.dtype = dtype
        self.p = {k: state_dict[k].detach().clone().to(dtype) for k in param_names()}
        self.rm = [state_dict[pre + ".1.running_mean"].detach().clone().to(dtype) for pre in HEAVY_PREFIXES]
        self.rv = [state_dict[pre + ".1.running_var"].detach().clone().to(dtype) for pre in HEAVY_PREFIXES]
        self.nbt = [int(state_dict[pre + ".1.num_batches_tracked"]) for pre in HEAVY_PREFIXES]
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.adam_step = 0

    def state_dict(self):
        out = {}
        for i, pre in enumerate(HEAVY_PREFIXES):
            out[pre + ".0.weight"] = self.p[pre + ".0.weight"]
            out[pre + ".0.bias"] = self.p[pre + ".0.bias"]
            out[pre + ".1.weight"] = self.p[pre + ".1.weight"]
            out[pre + ".1.bias"] = self.p[pre + ".1.bias"]
            out[pre + ".1.running_mean"] = self.rm[i]
            out[pre + ".1.running_var"] = self.rv[i]
            out[pre + ".1.num_batches_tracked"] = torch.tensor(self.nbt[i], dtype=torch.int64)
        out["decode.weight"] = self.p["decode.weight"]
        out["decode.bias"] = self.p["decode.bias"]
        return out


def _heavy_forward(st, i, h, train, mask, momentum, rnd=EXACT):
    """model/bilinear.py:7-13.  Returns (output, cache)."""
    pre = HEAVY_PREFIXES[i]
    W, b = st.p[pre + ".0.weight"], st.p[pre + ".0.bias"]
    g, be = st.p[pre + ".1.weight"], st.p[pre + ".1.bias"]
    z = rnd.op(h) @ rnd.op(W).t() + b                   # nn.Linear
    B = z.shape[0]
    if train:
        if B < 2:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                             % str(torch.Size(z.shape)))
        mean = z.mean(dim=0)
        var = ((z - mean) ** 2).mean(dim=0)             # biased: used for normalisation
        st.nbt[i] += 1                                  # incremented before the factor is derived
        fac = (1.0 / st.nbt[i]) if momentum is None else momentum
        st.rm[i] = (1 - fac) * st.rm[i] + fac * mean
        st.rv[i] = (1 - fac) * st.rv[i] + fac * var * (B / (B - 1))   # unbiased into running_var
    else:
        mean, var = st.rm[i], st.rv[i]
    rstd = 1.0 / torch.sqrt(var + BN_EPS)
    xhat = ((rnd.z(z) if (train and rnd.fwd_z) else z) - mean) * rstd
    y = g * xhat + be                                   # nn.BatchNorm1d (affine)
    a = torch.clamp(y, min=0)                           # nn.ReLU
    if train:
        m = mask.to(z.dtype) if mask is not None else torch.ones_like(z)
        d = a * m * 2.0                                 # nn.Dropout(p=0.5): keep * 1/(1-p)
    else:
        m = None
        d = a
    xhat_saved = (rnd.z(z) - mean) * rstd               # what backward recomputes from the saved z
    return d, dict(h_in=h, xhat=xhat_saved, rstd=rstd, y=y, pos=(y > 0), m=m)


def forward(st, x, train=False, masks=None, momentum=(0.1,) * NUM_HEAVY, rnd=EXACT):
    """model/bilinear.py:31-41.  masks: list of 5 keep-masks [B,1024] (train only; None = keep all).
    momentum: per-layer BatchNorm momentum, None entries mean cumulative average."""
    x = x.to(st.dtype)
    caches = []
    mk = (lambda i: None if masks is None else masks[i])
    h, c = _heavy_forward(st, 0, x, train, mk(0), momentum[0], rnd)
    caches.append(c)
    for blk in range(2):
        skip = rnd.op(h)                                # the residual edge reads the stored operand
        h, c = _heavy_forward(st, 1 + 2 * blk, h, train, mk(1 + 2 * blk), momentum[1 + 2 * blk], rnd)
        caches.append(c)
        h, c = _heavy_forward(st, 2 + 2 * blk, h, train, mk(2 + 2 * blk), momentum[2 + 2 * blk], rnd)
        caches.append(c)
        h = h + skip
    p = rnd.op(h) @ rnd.op(st.p["decode.weight"]).t() + st.p["decode.bias"]
    return p, dict(layers=caches, h_last=h)


def ambiguous_relu(cache, tol=2e-5):
    """[(layer, row, feature)] of elements whose BatchNorm output is within `tol` of the ReLU kink."""
    out = []
    for i, c in enumerate(cache["layers"]):
        idx = torch.nonzero(c["y"].abs() < tol)
        out += [(i, int(r), int(f)) for r, f in idx]
    return out


def mse_loss(p, y):
    """nn.MSELoss(): mean over all B*48 elements; returns (loss, dL/dp)."""
    y = y.to(p.dtype)
    diff = p - y
    n = diff.numel()
    return (diff ** 2).sum() / n, 2.0 * diff / n


def backward(st, cache, dP, rnd=EXACT, relu_pos=None):
    """Gradients of all 22 parameters for upstream gradient dP.
    relu_pos: optional list of five [B,1024] bool tensors that REPLACE the oracle's own ReLU decisions
    (y > 0); tests use it to adopt another evaluation's decisions after checking that they differ only
    where |y| is within round-off of the kink (see ambiguous_relu)."""
    grads = {}
    h_last = cache["h_last"]
    dPr = rnd.op(dP)
    grads["decode.weight"] = dPr.t() @ rnd.op(h_last)
    grads["decode.bias"] = dP.sum(dim=0)
    dh = dPr @ rnd.op(st.p["decode.weight"])            # gradient at h4 (block-1 output incl. skip)

    def heavy_backward(i, dd):
        pre = HEAVY_PREFIXES[i]
        c = cache["layers"][i]
        W = st.p[pre + ".0.weight"]
        g = st.p[pre + ".1.weight"]
        B = dd.shape[0]
        pos = c["pos"] if relu_pos is None else relu_pos[i]
        da = dd * c["m"] * 2.0
        dy = da * pos.to(dd.dtype)
        dbeta = dy.sum(dim=0)
        dgamma = (dy * c["xhat"]).sum(dim=0)
        dz = g * c["rstd"] * (rnd.dy(dy) - dbeta / B - c["xhat"] * dgamma / B)
        dzr = rnd.op(dz)
        grads[pre + ".1.weight"] = dgamma
        grads[pre + ".1.bias"] = dbeta
        grads[pre + ".0.weight"] = dzr.t() @ rnd.op(c["h_in"])
        grads[pre + ".0.bias"] = dz.sum(dim=0)
        return dzr @ rnd.op(W)

    for blk in (1, 0):
        skip_grad = rnd.dy(dh)                          # out = block(h) + h
        dx = heavy_backward(2 + 2 * blk, dh)
        dx = heavy_backward(1 + 2 * blk, dx)
        dh = dx + skip_grad
    heavy_backward(0, dh)                               # the network input needs no gradient
    return grads


def total_grad_norm(grads):
    """clip_grad_norm_: 2-norm of the per-tensor 2-norms."""
    return torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)


def clip_and_adam(st, grads, lr, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    """train_bilinear.py:81-83.  Returns the pre-clip total norm."""
    if grad_scale != 1.0:
        grads = {k: g * grad_scale for k, g in grads.items()}
    norm = total_grad_norm(grads)
    coef = 1.0
    if max_norm is not None and max_norm > 0:
        coef = min(float(max_norm) / (float(norm) + 1e-6), 1.0)
    st.adam_step += 1
    t = st.adam_step
    b1, b2 = betas
    bc1 = 1 - b1 ** t
    bc2_sqrt = math.sqrt(1 - b2 ** t)
    step_size = lr / bc1
    for k in param_names():
        g = grads[k] * coef
        st.m[k] = st.m[k] + (1 - b1) * (g - st.m[k])                 # exp_avg.lerp_(grad, 1 - beta1)
        st.v[k] = st.v[k] * b2 + (1 - b2) * g * g                    # mul_(beta2).addcmul_(g, g, 1 - beta2)
        denom = st.v[k].sqrt() / bc2_sqrt + eps
        st.p[k] = st.p[k] - step_size * (st.m[k] / denom)            # addcdiv_
    return norm


def train_step(st, x, y, masks, lr, momentum=(0.1,) * NUM_HEAVY, max_norm=1.0, grad_scale=1.0, rnd=EXACT,
               relu_pos=None):
    """One iteration of train_bilinear.py:75-83.  Returns (p, loss, grads (unclipped), total_norm)."""
    p, cache = forward(st, x, train=True, masks=masks, momentum=momentum, rnd=rnd)
    loss, dP = mse_loss(p, y)
    rp = relu_pos(cache) if callable(relu_pos) else relu_pos
    grads = backward(st, cache, dP, rnd=rnd, relu_pos=rp)
    norm = clip_and_adam(st, grads, lr, max_norm=max_norm, grad_scale=grad_scale)
    return p, loss, grads, norm


def synthetic_batch(seed, batch, with_masks=True):
    """x ~ N(0,1) [B,32], y ~ N(0,1) [B,48], five Bernoulli(0.5) keep masks."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, DIN, generator=g)
    y = torch.randn(batch, DOUT, generator=g)
    masks = None
    if with_masks:
        masks = [(torch.rand(batch, HIDDEN, generator=g) < 0.5).to(torch.uint8) for _ in range(NUM_HEAVY)]
    return x, y, masks
