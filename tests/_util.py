"""Shared helpers of the test-suite: golden-vector comparison and error metrics."""
import numpy as np
import torch

NSAMPLE = 1024


def positions_for(name, numel):
    # must match tests/golden/make_golden.py
    seed = sum(ord(c) * (i + 1) for i, c in enumerate(name)) % (2 ** 31)
    return np.random.default_rng(seed).integers(0, numel, size=NSAMPLE)


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 in float64 (the 'relative' of the north-star tolerances)."""
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1) if not torch.is_tensor(a) else a.detach().double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1) if not torch.is_tensor(b) else b.detach().double().cpu().reshape(-1)
    den = float(b.norm())
    num = float((a - b).norm())
    return num / den if den > 0 else num


def rel_max(a, b):
    """max|a - b| / max|b|"""
    a = a.detach().double().cpu().reshape(-1) if torch.is_tensor(a) else torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = b.detach().double().cpu().reshape(-1) if torch.is_tensor(b) else torch.as_tensor(np.asarray(b)).double().reshape(-1)
    den = float(b.abs().max())
    num = float((a - b).abs().max())
    return num / den if den > 0 else num


def golden_error(gold, prefix, name, t, scale=None):
    """Error of tensor `t` against the packed golden entry prefix/name: relative l2 when the whole tensor
    was stored, else max over (sampled-values relative l2, |l2 digest| relative, |sum digest| / (l2*sqrt n)).
    `scale` overrides the denominator (tensors that are analytically zero)."""
    a = t.detach().cpu().double().numpy().reshape(-1)
    key = prefix + "/" + name
    if key + "#full" in gold.files:
        ref = gold[key + "#full"].astype(np.float64).reshape(-1)
        den = scale if scale is not None else np.linalg.norm(ref)
        return float(np.linalg.norm(a - ref) / max(den, 1e-300))
    l2 = float(gold[key + "#l2"])
    sm = float(gold[key + "#sum"])
    den = scale if scale is not None else l2
    errs = [abs(np.linalg.norm(a) - l2) / max(den, 1e-300),
            abs(a.sum() - sm) / max(den * np.sqrt(a.size), 1e-300)]
    if key + "#samp" in gold.files:
        pos = positions_for(name, a.size)
        ref = gold[key + "#samp"].astype(np.float64)
        # sampled positions: compare against the per-element scale of the tensor
        errs.append(np.linalg.norm(a[pos] - ref) / max(den * np.sqrt(NSAMPLE / a.size), 1e-300))
    return float(max(errs))


def reference_init_state(seed):
    """State dict of model.bilinear.load(cpu) after torch.manual_seed(seed) -- reproduced with this
    repo's drop-in module, whose constructor consumes the RNG stream exactly like the reference's
    (model/bilinear.py:16-29, :85-90); pinned against the golden `init` digests in test_oracle_golden."""
    import warnings
    from bilinear_b200.model import bilinear as mb
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net, _, step, epoch = mb.load(torch.device("cpu"))
    assert step == 1 and epoch == 0
    return {k: v.clone() for k, v in net.state_dict().items()}
