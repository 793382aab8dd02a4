"""the reference's `bilinear.sh` alternates `python train_bilinear.py` /
`python valid_bilinear.py` forty times and carries ALL state between the processes through
`save/{comment}/parameter/{epoch}.save` (train_bilinear.py:92-104 writes, model/bilinear.py:63-83 reads).  For
the two implementations to alternate inside that loop a checkpoint written by either must load into the other.

Direction 1  the UNMODIFIED reference (model/bilinear.py + the literal step of train_bilinear.py:75-83 and the
             literal save of :92-104) writes 1.save  ->  bilinear_b200.model.bilinear.load() resumes from it:
             same 37 state entries, same Adam moments / step counts, same global step, and an optimizer
             state_dict() that round-trips unchanged.
Direction 2  this package writes 2.save in the same layout  ->  the unmodified reference's load() picks it as
             the newest epoch, restores module + torch.optim.Adam from it and trains on.

Runs on the CPU (loading, saving and state bookkeeping involve no kernel); the reference tree exists in the
build container only, so the module is skipped elsewhere.  The GPU half -- a reference-written checkpoint
resumed on the device and stepped against the oracle -- is tests/test_gpu_scripts.py.
"""
import importlib.util
import os
import sys
import warnings

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF = "/root/reference/model/bilinear.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree exists in the build container only")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_bilinear_ckpt", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_trains_and_saves(ref, parameter_dir, seed, iterations, epoch=1):
    """train_bilinear.py:45-104, literally, on synthetic batches (CPU)."""
    from oracle import bilinear_oracle as bo
    torch.manual_seed(seed)
    bilinear, optimizer, step, train_epoch = ref.load(torch.device("cpu"), parameter_dir)
    criterion = nn.MSELoss()
    bilinear.train()
    for i in range(iterations):
        x, y, _ = bo.synthetic_batch(300 + step, 64, with_masks=False)
        if step % 100000 == 0 or step == 1:
            for param_group in optimizer.param_groups:
                param_group['lr'] = 1.0e-3 * 0.96 ** (step / 100000)
        optimizer.zero_grad()
        loss = criterion(bilinear(x), y)
        loss.backward()
        nn.utils.clip_grad_norm_(bilinear.parameters(), max_norm=1)
        optimizer.step()
        step = step + 1
    os.makedirs(parameter_dir, exist_ok=True)
    torch.save({'epoch': epoch, 'step': step, 'state': bilinear.state_dict(), 'optimizer': optimizer.state_dict()},
               '{0}/{1}.save'.format(parameter_dir, epoch))
    return bilinear, optimizer, step


def test_reference_checkpoint_resumes_here_and_back(tmp_path):
    warnings.filterwarnings("ignore")
    from bilinear_b200.model import bilinear as mb
    ref = load_reference()
    pdir = str(tmp_path / "save" / "Bilinear GT" / "parameter")
    ref_net, ref_opt, ref_step = reference_trains_and_saves(ref, pdir, seed=3, iterations=3)
    assert sorted(os.listdir(pdir)) == ["1.save"]

    # ---- direction 1: reference -> this package
    net, opt, step, epoch = mb.load(torch.device("cpu"), pdir)
    assert (step, epoch) == (ref_step, 1) == (4, 1)
    sd, ref_sd = net.state_dict(), ref_net.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys()) and len(sd) == 37
    for k in sd:
        assert sd[k].dtype == ref_sd[k].dtype and torch.equal(sd[k], ref_sd[k]), k
    osd, ref_osd = opt.state_dict(), ref_opt.state_dict()
    assert osd["param_groups"][0]["params"] == ref_osd["param_groups"][0]["params"] == list(range(22))
    assert osd["param_groups"][0]["lr"] == ref_osd["param_groups"][0]["lr"]
    assert set(osd["state"].keys()) == set(ref_osd["state"].keys()) == set(range(22))
    for i in range(22):
        assert float(osd["state"][i]["step"]) == float(ref_osd["state"][i]["step"]) == 3.0
        assert torch.equal(osd["state"][i]["exp_avg"], ref_osd["state"][i]["exp_avg"])
        assert torch.equal(osd["state"][i]["exp_avg_sq"], ref_osd["state"][i]["exp_avg_sq"])
    assert opt._t == 3      # the fused kernel continues Adam's bias correction from the reference's count

    # ---- direction 2: this package -> reference (epoch 2 written in the layout of train_bilinear.py:92-104)
    torch.save({'epoch': 2, 'step': step + 5, 'state': net.state_dict(), 'optimizer': opt.state_dict()},
               '{0}/{1}.save'.format(pdir, 2))
    net2, opt2, step2, epoch2 = ref.load(torch.device("cpu"), pdir)
    assert (step2, epoch2) == (step + 5, 2)
    for k, v in net2.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    assert all(float(s["step"]) == 3.0 for s in opt2.state_dict()["state"].values())
    # ... and the reference trains on from it (the 40 x (train, valid) loop of bilinear.sh)
    _, _, step3 = reference_trains_and_saves(ref, pdir, seed=4, iterations=2, epoch=3)
    assert step3 == step2 + 2 and sorted(os.listdir(pdir)) == ["1.save", "2.save", "3.save"]
    net3, opt3, step4, epoch3 = mb.load(torch.device("cpu"), pdir)
    assert (step4, epoch3) == (step3, 3) and opt3._t == 5


def test_loader_file_name_contract(tmp_path):
    """model/bilinear.py:63-72: every file below parameter_dir must be '<int>.<ext>', the largest int wins; anything
    else raises ValueError -- in both implementations."""
    from bilinear_b200.model import bilinear as mb
    ref = load_reference()
    pdir = str(tmp_path / "parameter")
    os.makedirs(pdir)
    open(os.path.join(pdir, "notes.txt"), "w").close()
    with pytest.raises(ValueError):
        ref.load(torch.device("cpu"), pdir)
    with pytest.raises(ValueError):
        mb.load(torch.device("cpu"), pdir)
