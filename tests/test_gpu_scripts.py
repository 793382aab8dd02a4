"""Entry points on a synthetic Human3.6M (tests/_h36m_synth.py): train_bilinear.py writes reference-format
checkpoints that plain torch reads back, resumes from them, and valid_bilinear.py reports the MPJPE the
oracle computes for the saved state (valid_bilinear.py:76-83)."""
import logging
import warnings
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


def test_detector_protocol_runs_through_the_scripts(cuda_lib, tmp_path):
    """`protocol = Protocol.SH` (H36M/protocol.py:3, util/config.py:24) makes both scripts
    read {train,valid}_SH.bin (H36M/data.py:33) -- same model, noisier 2D inputs.  The MPJPE valid_bilinear.py
    reports must be the oracle's on the SH validation inputs (and differ from the GT inputs')."""
    import train_bilinear
    import valid_bilinear
    from bilinear_b200 import h36m
    from oracle import bilinear_oracle as bo
    from oracle import pose_metrics as pm
    data_dir = _h36m_synth.write_dataset(str(tmp_path), n_train=192, n_valid=80, protocols=("GT", "SH"))
    save_root = str(tmp_path / "save")
    cfg = _cfg(data_dir, "Bilinear SH")
    cfg.protocol = h36m.Protocol.SH
    cfg.epochs = 1
    torch.manual_seed(0)
    _quiet_root_logger()
    loss, step = train_bilinear.main(cfg, save_root=save_root)
    assert np.isfinite(loss) and step == 1 + 3
    _quiet_root_logger()
    res = valid_bilinear.main(cfg, save_root=save_root)
    ck = torch.load(os.path.join(save_root, "Bilinear SH", "parameter", "1.save"), map_location="cpu")
    st = bo.State({k: v.clone() for k, v in ck["state"].items()}, dtype=torch.float64)
    avgs = {}
    for protocol in (h36m.Protocol.SH, h36m.Protocol.GT):
        ds = h36m.PoseDataset(data_dir, "valid", protocol)
        p, _ = bo.forward(st, ds.part.double(), train=False)
        dist = pm.mpjpe_per_pose(p.float().numpy(), ds.S.numpy(), ds.mean_S.numpy(), ds.stddev_S.numpy())
        avgs[protocol] = float(dist.sum() / (len(ds) * 16))
    assert abs(res["avg"] - avgs[h36m.Protocol.SH]) < 0.05, (res["avg"], avgs)
    assert abs(avgs[h36m.Protocol.SH] - avgs[h36m.Protocol.GT]) > 0.05      # the protocol really changed the inputs
    _quiet_root_logger()


@pytest.mark.parametrize("move_after_load", [False, True])
def test_reference_written_checkpoint_resumes_on_the_device(cuda_lib, tmp_path, move_after_load):
    """SURVEY.md section 8(f)-2 on the GPU: a checkpoint written by the UNMODIFIED reference (baseline/_ref/model/
    bilinear.py + the literal step / save of train_bilinear.py:75-104, three CPU iterations) is resumed by
    `bilinear_b200.model.bilinear.load`, and the next fused iteration -- which needs the reference's parameters,
    BatchNorm buffers AND Adam moments / step count -- matches the oracle continued from the same checkpoint.
    `move_after_load`: load(device=cpu, ...) followed by .to('cuda') -- the moments arrive as CPU tensors and must
    be copied into the flat device buffers before the kernel reads them."""
    import baseline
    if not baseline.available():
        pytest.skip("baseline/_ref/model/bilinear.py was not staged (run __graft_entry__.build() in the build container)")
    from bilinear_b200.model import bilinear as mb
    from oracle import bilinear_oracle as bo
    ref = baseline.load_reference_module()
    pdir = str(tmp_path / "parameter")
    torch.manual_seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_net, r_opt, r_step, _ = ref.load(torch.device("cpu"), pdir)
    r_net.train()
    for i in range(3):
        x, y, _ = bo.synthetic_batch(700 + i, 64, with_masks=False)
        r_opt.zero_grad()
        nn.MSELoss()(r_net(x), y).backward()
        nn.utils.clip_grad_norm_(r_net.parameters(), max_norm=1)
        r_opt.step()
        r_step += 1
    os.makedirs(pdir)
    torch.save({'epoch': 1, 'step': r_step, 'state': r_net.state_dict(), 'optimizer': r_opt.state_dict()},
               '{0}/{1}.save'.format(pdir, 1))

    if move_after_load:
        net, opt, step, epoch = mb.load(torch.device("cpu"), pdir)
        net.to("cuda")
    else:
        net, opt, step, epoch = mb.load(torch.device("cuda"), pdir)
    assert (step, epoch) == (4, 1)
    # oracle continued from the same file
    ck = torch.load(os.path.join(pdir, "1.save"), map_location="cpu")
    st = bo.State({k: v.clone() for k, v in ck["state"].items()}, dtype=torch.float64)
    names = bo.param_names()
    for i, n in enumerate(names):
        st.m[n] = ck["optimizer"]["state"][i]["exp_avg"].double().clone()
        st.v[n] = ck["optimizer"]["state"][i]["exp_avg_sq"].double().clone()
    st.adam_step = int(float(ck["optimizer"]["state"][0]["step"]))
    net.train()
    x, y, masks = bo.synthetic_batch(41, 64)
    net.inject_dropout_masks(torch.stack(masks).to(torch.uint8).cuda())
    lr = ck["optimizer"]["param_groups"][0]["lr"]
    before = {k: v.detach().double().cpu().clone() for k, v in net.named_parameters()}
    loss, gnorm = net.train_step(opt, x.cuda(), y.cuda(), max_norm=1.0)
    _, loss_ref, _, norm_ref = bo.train_step(st, x, y, masks, lr)
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * float(loss_ref)
    assert abs(float(gnorm) - float(norm_ref)) < 1e-5 * float(norm_ref)
    for name, p in net.named_parameters():
        if name.endswith(".0.bias") and not name.startswith("decode"):
            continue
        delta = p.detach().double().cpu() - before[name]
        delta_ref = st.p[name] - before[name]
        assert float((delta - delta_ref).norm()) < 1e-3 * float(delta_ref.norm()) + 1e-7, name
    osd = opt.state_dict()
    assert all(float(s["step"]) == 4.0 for s in osd["state"].values())
    i = names.index("bilinear.0.0.0.weight")
    assert float((osd["state"][i]["exp_avg"].double().cpu() - st.m[names[i]]).norm()) < 1e-4 * float(st.m[names[i]].norm())
    # ... and the reference reads what this package writes after that step
    torch.save({'epoch': 2, 'step': step + 1, 'state': net.state_dict(), 'optimizer': osd}, '{0}/{1}.save'.format(pdir, 2))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n2, o2, s2, e2 = ref.load(torch.device("cpu"), pdir)
    assert (s2, e2) == (step + 1, 2) and len(o2.state_dict()["state"]) == 22
