"""ORACLE (test infrastructure, never on the product path): CPU restatement of the H36M evaluation
math of the reference, in numpy float64.

MPJPE follows the reference literally:
    valid_bilinear.py:53-54   prediction = stddev * prediction + mean ; gt = stddev * gt + mean
    valid_bilinear.py:56-57   view(-1, 16, 3)
    valid_bilinear.py:60      dist = sum_j sqrt(sum_xyz (pred - gt)^2)
    valid_bilinear.py:61-70   per-action accumulation in np.double
    valid_bilinear.py:76-83   MPJPE_a = total_dist[a] / (count_a * 16) ; avg = sum dist / sum(count * 16)

P-MPJPE: PARITY UNPINNED.  The reference contains no Procrustes code (H36M/protocol.py:1-4 is an
enum of 2D-input sources) and no golden vectors, so there is nothing of the reference's to pin this
against.  The restatement follows SURVEY.md Appendix B (standard similarity Procrustes as used by the
Martinez et al. evaluation) and is implemented twice, independently -- per-pose numpy SVD
(`pmpjpe_per_pose`) and batched torch.linalg.svd (`pmpjpe_per_pose_torch`) -- and cross-checked in
tests/test_oracle_metrics.py together with known-answer cases.
"""
import numpy as np


def unnormalise(a, mean, std):
    """valid_bilinear.py:53-54 (mean/std are the TRAIN-set statistics, broadcast over the batch)."""
    a = np.asarray(a, dtype=np.float64)
    if mean is None:
        return a
    return np.asarray(std, np.float64) * a + np.asarray(mean, np.float64)


def mpjpe_per_pose(pred, gt, mean=None, std=None):
    """Per-pose SUM over joints of the Euclidean distance, valid_bilinear.py:56-60.  pred, gt: [N, 3J]
    (normalised units when mean/std are given, else mm)."""
    p = unnormalise(pred, mean, std)
    g = unnormalise(gt, mean, std)
    n = p.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.float64)
    p = p.reshape(n, -1, 3)
    g = g.reshape(n, -1, 3)
    return np.sqrt(((p - g) ** 2).sum(axis=2)).sum(axis=1)


def per_action_report(dist, action, n_actions, joints):
    """valid_bilinear.py:61-83: float64 accumulation per action, then per-action and overall means."""
    total = np.zeros(n_actions, dtype=np.float64)
    count = np.zeros(n_actions, dtype=np.int64)
    np.add.at(total, action, np.asarray(dist, np.float64))
    np.add.at(count, action, 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        per_action = total / (count * joints)
    avg = total.sum() / (count.sum() * joints) if count.sum() else float("nan")
    return total, count, per_action, avg


DEGENERATE_REL = 1e-10


def procrustes_align(P, G):
    """SURVEY.md Appendix B for one pose.  P, G: [J, 3] float64.  Returns the aligned prediction."""
    muG = G.mean(axis=0)
    muP = P.mean(axis=0)
    G0 = G - muG
    P0 = P - muP
    nG = np.sqrt((G0 ** 2).sum())
    nP = np.sqrt((P0 ** 2).sum())
    # convention (stated, since nothing in the reference defines it): a pose whose extent about its
    # centroid is below 1e-5 of its raw magnitude (extent^2 <= DEGENERATE_REL * sum |x|^2) has no
    # orientation -> alignment undefined -> the UNALIGNED prediction is measured
    if nG ** 2 <= DEGENERATE_REL * (G ** 2).sum() or nP ** 2 <= DEGENERATE_REL * (P ** 2).sum():
        return P.copy()
    G0n = G0 / nG
    P0n = P0 / nP
    A = G0n.T @ P0n
    U, S, Vt = np.linalg.svd(A)
    V = Vt.T
    R = V @ U.T
    if np.linalg.det(R) < 0:  # forbid reflection
        V[:, 2] *= -1
        S = S.copy()
        S[2] *= -1
        R = V @ U.T
    tr = S.sum()
    scale = tr * nG / nP
    return scale * (P0 @ R) + muG


def pmpjpe_per_pose(pred, gt, mean=None, std=None):
    """Per-pose SUM over joints of the distance after Procrustes alignment (numpy SVD per pose)."""
    p = unnormalise(pred, mean, std)
    g = unnormalise(gt, mean, std)
    n = p.shape[0]
    p = p.reshape(n, -1, 3)
    g = g.reshape(n, -1, 3)
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        a = procrustes_align(p[i], g[i])
        out[i] = np.sqrt(((a - g[i]) ** 2).sum(axis=1)).sum()
    return out


def pmpjpe_per_pose_torch(pred, gt, mean=None, std=None, dtype=None):
    """Independent second implementation: batched torch.linalg.svd (float64 by default; float32 is
    what bench.py times as the CPU baseline for this metric)."""
    import torch
    dtype = dtype or torch.float64
    p = torch.as_tensor(unnormalise(pred, mean, std)).to(dtype)
    g = torch.as_tensor(unnormalise(gt, mean, std)).to(dtype)
    n = p.shape[0]
    p = p.reshape(n, -1, 3)
    g = g.reshape(n, -1, 3)
    muG = g.mean(dim=1, keepdim=True)
    muP = p.mean(dim=1, keepdim=True)
    G0 = g - muG
    P0 = p - muP
    nG = G0.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
    nP = P0.pow(2).sum(dim=(1, 2), keepdim=True).sqrt()
    ok = ((nG ** 2 > DEGENERATE_REL * g.pow(2).sum(dim=(1, 2), keepdim=True))
          & (nP ** 2 > DEGENERATE_REL * p.pow(2).sum(dim=(1, 2), keepdim=True))).flatten()
    nGs = torch.where(nG > 0, nG, torch.ones_like(nG))
    nPs = torch.where(nP > 0, nP, torch.ones_like(nP))
    A = (G0 / nGs).transpose(1, 2) @ (P0 / nPs)
    U, S, Vt = torch.linalg.svd(A)
    V = Vt.transpose(1, 2)
    R = V @ U.transpose(1, 2)
    sign = torch.sign(torch.linalg.det(R))
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    V = V.clone()
    V[:, :, 2] *= sign[:, None]
    S = S.clone()
    S[:, 2] *= sign
    R = V @ U.transpose(1, 2)
    tr = S.sum(dim=1).reshape(n, 1, 1)
    aligned = tr * nG / nPs * (P0 @ R) + muG
    aligned = torch.where(ok.reshape(n, 1, 1), aligned, p)
    return (aligned - g).pow(2).sum(dim=2).sqrt().sum(dim=1).double().numpy()


def synthetic_eval_poses(n, joints=16, seed=0):
    """G = 150 N(0,1) mm root-relative; P = s R G + t + 30 N(0,1);
    mu ~ 100 N(0,1), sigma ~ U(50, 250); action ids U{0..14}.  Returns normalised float32 arrays
    (what valid_bilinear.py receives) plus the statistics."""
    rng = np.random.default_rng(seed)
    G = 150.0 * rng.standard_normal((n, joints, 3))
    G = G - G[:, :1, :]
    Q, _ = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    det = np.linalg.det(Q)
    Q[:, :, 0] *= np.sign(det)[:, None]
    s = rng.uniform(0.8, 1.2, size=(n, 1, 1))
    t = 50.0 * rng.standard_normal((n, 1, 3))
    P = s * (G @ Q) + t + 30.0 * rng.standard_normal((n, joints, 3))
    mean = (100.0 * rng.standard_normal(3 * joints)).astype(np.float32)
    std = rng.uniform(50.0, 250.0, size=3 * joints).astype(np.float32)
    pred_norm = ((P.reshape(n, -1) - mean) / std).astype(np.float32)
    gt_norm = ((G.reshape(n, -1) - mean) / std).astype(np.float32)
    action = rng.integers(0, 15, size=n).astype(np.int32)
    return dict(pred_norm=pred_norm, gt_norm=gt_norm, mean=mean, std=std, action=action)
