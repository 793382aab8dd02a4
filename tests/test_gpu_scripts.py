"""Entry points on a synthetic Human3.6M (tests/_h36m_synth.py): train_bilinear.py writes reference-format
checkpoints that plain torch reads back, resumes from them, and valid_bilinear.py reports the MPJPE the
oracle computes for the saved state (valid_bilinear.py:76-83)."""
import logging
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import _h36m_synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _cfg(data_dir, comment):
    from bilinear_b200 import config
    c = config._Map(dict(config.bilinear))
    c.lr_decay = config.bilinear.lr_decay
    c.data_dir, c.comment, c.epochs, c.device = data_dir, comment, 2, torch.device("cuda")
    return c


def _quiet_root_logger():
    root = logging.getLogger()
    for h in list(root.handlers):
        root.removeHandler(h)


def test_train_then_valid_scripts(cuda_lib, tmp_path):
    import train_bilinear
    import valid_bilinear
    from bilinear_b200 import h36m
    from oracle import bilinear_oracle as bo
    from oracle import pose_metrics as pm
    data_dir = _h36m_synth.write_dataset(str(tmp_path), n_train=256, n_valid=96)
    save_root = str(tmp_path / "save")
    cfg = _cfg(data_dir, "unit test")
    torch.manual_seed(0)
    _quiet_root_logger()
    loss, step = train_bilinear.main(cfg, save_root=save_root)
    assert np.isfinite(loss) and step == 1 + 2 * 4            # 256 poses / batch 64 = 4 iterations per epoch
    pdir = os.path.join(save_root, "unit test", "parameter")
    assert sorted(os.listdir(pdir)) == ["1.save", "2.save"]   # '{epoch}.save' (train_bilinear.py:92-104)

    # the checkpoint is what the reference writes: plain torch reads it into an equally shaped module + Adam
    ck = torch.load(os.path.join(pdir, "2.save"), map_location="cpu")
    assert set(ck) == {"epoch", "step", "state", "optimizer"} and ck["epoch"] == 2 and ck["step"] == step
    from bilinear_b200.model.bilinear import BilinearUnit
    plain = BilinearUnit()                                     # same nn.Module tree as model/bilinear.py:16-29
    assert list(ck["state"].keys()) == list(plain.state_dict().keys()) and len(ck["state"]) == 37
    plain.load_state_dict(ck["state"])
    adam = torch.optim.Adam(plain.parameters(), lr=1e-3)
    adam.load_state_dict(ck["optimizer"])
    assert len(adam.state_dict()["state"]) == 22
    assert all(float(s["step"]) == 8.0 for s in adam.state_dict()["state"].values())

    # resume: two more epochs start from epoch 3 and continue the step counter (model/bilinear.py:63-83)
    _quiet_root_logger()
    loss2, step2 = train_bilinear.main(cfg, save_root=save_root)
    assert step2 == step + 8 and sorted(os.listdir(pdir)) == ["1.save", "2.save", "3.save", "4.save"]

    # validation: per-action and average MPJPE of the newest checkpoint == oracle on the same state
    _quiet_root_logger()
    res = valid_bilinear.main(cfg, save_root=save_root)
    ck = torch.load(os.path.join(pdir, "4.save"), map_location="cpu")
    st = bo.State({k: v.clone() for k, v in ck["state"].items()}, dtype=torch.float64)
    ds = h36m.PoseDataset(data_dir, "valid")
    p, _ = bo.forward(st, ds.part.double(), train=False)
    dist = pm.mpjpe_per_pose(p.float().numpy(), ds.S.numpy(), ds.mean_S.numpy(), ds.stddev_S.numpy())
    rep = pm.per_action_report(dist, ds.action_id.numpy(), len(ds.action_names), 16)
    ref_avg = float(dist.sum() / (len(ds) * 16))
    assert abs(res["avg"] - ref_avg) < 0.05, (res["avg"], ref_avg)
    pd = pm.pmpjpe_per_pose_torch(p.float().numpy(), ds.S.numpy(), ds.mean_S.numpy(), ds.stddev_S.numpy())
    assert abs(res["avg_pmpjpe"] - float(pd.sum() / (len(ds) * 16))) < 0.05
    for i, name in enumerate(ds.action_names):
        if name in res:
            a = (ds.action_id.numpy() == i)
            assert abs(res[name] - float(dist[a].sum() / (a.sum() * 16))) < 0.05, name
    _quiet_root_logger()
