"""GPU parity for the evaluation math: MPJPE (valid_bilinear.py:53-83) and P-MPJPE (parity unpinned by
the reference; oracle = SURVEY.md Appendix B in float64) within 0.05 mm, through the C-ABI."""
import numpy as np
import pytest
import torch

from bilinear_b200 import metrics
from oracle import pose_metrics as pm

pytestmark = pytest.mark.gpu
TOL_MM = 0.05


def _run(d, J, fused=True):
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(a).to(dev)
    return metrics.evaluate(t(d["pred_norm"]), t(d["gt_norm"]), t(d["mean"]), t(d["std"]), t(d["action"]), 15,
                            joints=J, fused=fused)


@pytest.mark.parametrize("n,J", [(1, 16), (127, 16), (128, 16), (1000, 16), (901, 17), (200000, 16), (65536, 17)])
def test_metrics_match_oracle(cuda_lib, n, J):
    d = pm.synthetic_eval_poses(n, J, seed=n % 7)
    r = _run(d, J)
    ref_m = pm.mpjpe_per_pose(d["pred_norm"], d["gt_norm"], d["mean"], d["std"])
    ref_p = pm.pmpjpe_per_pose_torch(d["pred_norm"], d["gt_norm"], d["mean"], d["std"])
    assert np.abs(r.dist.double().cpu().numpy() - ref_m).max() / J < TOL_MM
    assert np.abs(r.pdist.double().cpu().numpy() - ref_p).max() / J < TOL_MM
    tot, cnt, per_action, avg = pm.per_action_report(ref_m, d["action"], 15, J)
    ptot, _, pper, pavg = pm.per_action_report(ref_p, d["action"], 15, J)
    assert (r.count.cpu().numpy() == cnt).all()
    assert abs(r.mpjpe - avg) < TOL_MM and abs(r.pmpjpe - pavg) < TOL_MM
    got = r.per_action_mpjpe.cpu().numpy()
    ok = cnt > 0
    assert np.abs(got[ok] - per_action[ok]).max() < TOL_MM
    assert np.abs(r.per_action_pmpjpe.cpu().numpy()[ok] - pper[ok]).max() < TOL_MM


def test_separate_kernels_equal_fused(cuda_lib):
    d = pm.synthetic_eval_poses(5000, 16, seed=9)
    a = _run(d, 16, fused=True)
    b = _run(d, 16, fused=False)
    assert torch.equal(a.dist, b.dist) and torch.equal(a.pdist, b.pdist)


def test_known_answers(cuda_lib):
    rng = np.random.default_rng(0)
    n = 300
    G = 150.0 * rng.standard_normal((n, 16, 3))
    Q, _ = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    Q[:, :, 0] *= np.sign(np.linalg.det(Q))[:, None]
    P = 1.2 * (G @ Q) + 40.0 * rng.standard_normal((n, 1, 3))
    dev = torch.device("cuda")
    g = torch.from_numpy(G.reshape(n, -1).astype(np.float32)).to(dev)
    p = torch.from_numpy(P.reshape(n, -1).astype(np.float32)).to(dev)
    r = metrics.evaluate(p, g, None, None, None, 1, joints=16)
    assert float(r.pdist.max()) / 16 < 5e-3          # similarity-transformed copy aligns to ~0
    assert float(r.dist.min()) / 16 > 1.0
    r0 = metrics.evaluate(g, g, None, None, None, 1, joints=16)
    assert float(r0.dist.abs().max()) == 0.0 and float(r0.pdist.abs().max()) / 16 < 5e-3
    # reflection must not be allowed
    Pm = G.copy()
    Pm[:, :, 0] *= -1
    pmir = torch.from_numpy(Pm.reshape(n, -1).astype(np.float32)).to(dev)
    rm = metrics.evaluate(pmir, g, None, None, None, 1, joints=16)
    ref = pm.pmpjpe_per_pose_torch(Pm.reshape(n, -1).astype(np.float32), G.reshape(n, -1).astype(np.float32))
    assert float(rm.pdist.min()) > 1.0
    assert np.abs(rm.pdist.double().cpu().numpy() - ref).max() / 16 < TOL_MM


def test_large_rotations_both_adjugate_paths(cuda_lib):
    """The kernel takes the optimal quaternion from the FIRST column of the adjugate when q0^2 > 0.1 (turns below ~143
    degrees) and from the full pivoted adjugate otherwise (eval.cu, horn_top): sweep the turn angle across that boundary up
    to 180 degrees, about random axes, with noise, and compare with the float64 oracle."""
    rng = np.random.default_rng(5)
    n = 4096
    G = 200.0 * rng.standard_normal((n, 16, 3))
    axis = rng.standard_normal((n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = np.linspace(np.deg2rad(100.0), np.pi, n)
    K = np.zeros((n, 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    R = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1.0 - np.cos(ang))[:, None, None] * (K @ K)   # Rodrigues
    P = 0.8 * (G @ R) + 25.0 * rng.standard_normal((n, 1, 3)) + 10.0 * rng.standard_normal((n, 16, 3))
    g32 = G.reshape(n, -1).astype(np.float32)
    p32 = P.reshape(n, -1).astype(np.float32)
    dev = torch.device("cuda")
    r = metrics.evaluate(torch.from_numpy(p32).to(dev), torch.from_numpy(g32).to(dev), None, None, None, 1, joints=16)
    ref = pm.pmpjpe_per_pose_torch(p32, g32)
    assert np.isfinite(r.pdist.cpu().numpy()).all()
    assert np.abs(r.pdist.double().cpu().numpy() - ref).max() / 16 < TOL_MM


def test_degenerate_poses_finite(cuda_lib):
    rng = np.random.default_rng(1)
    G = (100.0 * rng.standard_normal((8, 16, 3))).astype(np.float32)
    P = (100.0 * rng.standard_normal((8, 16, 3))).astype(np.float32)
    G[0, :, 2] = 0.0
    P[1, :, 1:] = 0.0
    P[2] = P[2, :1]
    G[3] = G[3, :1]
    P[4] = 0.0
    dev = torch.device("cuda")
    r = metrics.evaluate(torch.from_numpy(P.reshape(8, -1)).to(dev), torch.from_numpy(G.reshape(8, -1)).to(dev),
                         None, None, None, 1, joints=16)
    assert torch.isfinite(r.pdist).all() and torch.isfinite(r.dist).all()
    ref = pm.pmpjpe_per_pose_torch(P.reshape(8, -1), G.reshape(8, -1))
    got = r.pdist.double().cpu().numpy()
    for i in (0, 2, 3, 4, 5, 6, 7):        # the collinear case (1) has a non-unique optimal rotation
        assert abs(got[i] - ref[i]) / 16 < TOL_MM, (i, got[i], ref[i])
    assert got[1] <= r.dist.double().cpu().numpy()[1] + 1e-2


def test_empty_and_errors(cuda_lib):
    dev = torch.device("cuda")
    z = torch.zeros(0, 48, device=dev)
    r = metrics.evaluate(z, z, None, None, None, 15, joints=16)
    assert r.dist.numel() == 0 and int(r.count.sum()) == 0 and np.isnan(r.mpjpe)
    with pytest.raises(RuntimeError):
        metrics.evaluate(torch.zeros(4, 45, device=dev), torch.zeros(4, 45, device=dev), None, None, None, 15, joints=15)
    with pytest.raises(RuntimeError):
        metrics.evaluate(torch.zeros(4, 48), torch.zeros(4, 48), None, None, None, 15, joints=16)


def test_full_size_properties(cuda_lib):
    """BASELINE config 5 size (1M poses): size-independent properties instead of the per-pose oracle --
    invariance of P-MPJPE under a similarity transform of the prediction, count conservation, and the
    per-action sums adding up to the per-pose sums."""
    n = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    G = 150.0 * torch.randn(n, 16, 3, device="cuda", generator=g)
    P = G + 30.0 * torch.randn(n, 16, 3, device="cuda", generator=g)
    act = torch.randint(0, 15, (n,), device="cuda", generator=g, dtype=torch.int32)
    r = metrics.evaluate(P.reshape(n, 48), G.reshape(n, 48), None, None, act, 15, joints=16)
    assert int(r.count.sum()) == n
    assert abs(float(r.sum.sum()) - float(r.dist.double().sum())) < 1e-6 * float(r.dist.double().sum())
    assert abs(float(r.psum.sum()) - float(r.pdist.double().sum())) < 1e-6 * float(r.pdist.double().sum())
    c, s = np.cos(0.7), np.sin(0.7)
    R = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], device="cuda", dtype=torch.float32)
    P2 = 0.9 * (P @ R) + torch.tensor([10.0, -20.0, 30.0], device="cuda")
    r2 = metrics.evaluate(P2.reshape(n, 48), G.reshape(n, 48), None, None, act, 15, joints=16)
    assert float((r2.pdist - r.pdist).abs().max()) / 16 < TOL_MM
    assert abs(r2.pmpjpe - r.pmpjpe) < 1e-3
    # (Procrustes minimises the sum of SQUARED distances, so pdist <= dist holds on average, not per pose)
    assert r.pmpjpe <= r.mpjpe
