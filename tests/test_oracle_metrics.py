"""oracle/pose_metrics.py: MPJPE follows valid_bilinear.py:53-83 literally and is checked against the
literal torch expression of that file; P-MPJPE is PARITY UNPINNED (no reference code or vectors exist,
SURVEY.md section 0) -- two independent implementations are cross-checked and known-answer cases hold."""
import numpy as np
import pytest
import torch

from oracle import pose_metrics as pm


def _literal_valid_bilinear(pred, gt, mean, std):
    # valid_bilinear.py:52-60 verbatim semantics (mean/stddev arrive replicated per sample)
    prediction = torch.from_numpy(pred)
    in_camera_space = torch.from_numpy(gt)
    m = torch.from_numpy(mean).expand_as(prediction)
    s = torch.from_numpy(std).expand_as(prediction)
    prediction = s * prediction + m
    in_camera_space = s * in_camera_space + m
    prediction = prediction.view(-1, 16, 3)
    in_camera_space = in_camera_space.view(-1, 16, 3)
    return torch.sum(torch.sqrt(torch.sum((prediction - in_camera_space) ** 2, dim=2)), dim=1)


def test_mpjpe_matches_reference_expression():
    d = pm.synthetic_eval_poses(4097, 16, seed=3)
    ref = _literal_valid_bilinear(d["pred_norm"], d["gt_norm"], d["mean"], d["std"]).double().numpy()
    got = pm.mpjpe_per_pose(d["pred_norm"], d["gt_norm"], d["mean"], d["std"])
    assert np.abs(got - ref).max() / 16 < 1e-3          # fp32 reference vs fp64 restatement, mm per joint
    total, count, per_action, avg = pm.per_action_report(got, d["action"], 15, 16)
    # valid_bilinear.py:61-83 with a python loop
    tot = {}
    cnt = {}
    for i in range(len(got)):
        a = int(d["action"][i])
        tot[a] = tot.get(a, 0.0) + np.double(ref[i])
        cnt[a] = cnt.get(a, 0) + 1
    for a in tot:
        assert count[a] == cnt[a]
        assert abs(per_action[a] - tot[a] / (cnt[a] * 16)) < 1e-3
    assert abs(avg - sum(tot.values()) / sum(c * 16 for c in cnt.values())) < 1e-3


@pytest.mark.parametrize("J", [16, 17])
def test_pmpjpe_two_implementations_agree(J):
    d = pm.synthetic_eval_poses(2000, J, seed=4)
    a = pm.pmpjpe_per_pose(d["pred_norm"], d["gt_norm"], d["mean"], d["std"])
    b = pm.pmpjpe_per_pose_torch(d["pred_norm"], d["gt_norm"], d["mean"], d["std"])
    assert np.abs(a - b).max() / J < 1e-8
    m = pm.mpjpe_per_pose(d["pred_norm"], d["gt_norm"], d["mean"], d["std"])
    assert (a <= m + 1e-6).all()                         # alignment can only reduce the error


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    return q


def test_pmpjpe_known_answers():
    rng = np.random.default_rng(0)
    G = 150.0 * rng.standard_normal((50, 16, 3))
    # identical poses -> 0
    assert np.abs(pm.pmpjpe_per_pose(G.reshape(50, -1), G.reshape(50, -1))).max() < 1e-9
    # pred = s R gt + t -> 0
    P = np.stack([1.3 * (g @ _rot(rng)) + rng.standard_normal(3) * 40 for g in G])
    assert np.abs(pm.pmpjpe_per_pose(P.reshape(50, -1), G.reshape(50, -1))).max() < 1e-8
    assert np.abs(pm.pmpjpe_per_pose_torch(P.reshape(50, -1), G.reshape(50, -1))).max() < 1e-8
    # a reflected pose must NOT align to zero
    Pm = G.copy()
    Pm[:, :, 0] *= -1
    r = pm.pmpjpe_per_pose(Pm.reshape(50, -1), G.reshape(50, -1))
    assert (r > 1.0).all()
    assert np.abs(r - pm.pmpjpe_per_pose_torch(Pm.reshape(50, -1), G.reshape(50, -1))).max() < 1e-6


def test_pmpjpe_degenerate_is_finite():
    rng = np.random.default_rng(1)
    G = 100.0 * rng.standard_normal((8, 16, 3))
    P = 100.0 * rng.standard_normal((8, 16, 3))
    G[0, :, 2] = 0.0                     # planar ground truth
    P[1, :, 1:] = 0.0                    # collinear prediction
    P[2] = P[2, :1]                      # zero-extent prediction -> unaligned distance by convention
    G[3] = G[3, :1]
    for f in (pm.pmpjpe_per_pose, pm.pmpjpe_per_pose_torch):
        r = f(P.reshape(8, -1), G.reshape(8, -1))
        assert np.isfinite(r).all()
    r = pm.pmpjpe_per_pose(P.reshape(8, -1), G.reshape(8, -1))
    m = pm.mpjpe_per_pose(P.reshape(8, -1), G.reshape(8, -1))
    assert abs(r[2] - m[2]) < 1e-9 and abs(r[3] - m[3]) < 1e-9


def test_empty_input():
    z = np.zeros((0, 48), np.float32)
    assert pm.mpjpe_per_pose(z, z).shape == (0,)
    total, count, per_action, avg = pm.per_action_report(np.zeros(0), np.zeros(0, np.int64), 15, 16)
    assert count.sum() == 0 and np.isnan(avg)
