"""Generates tests/golden/bilinear_golden.npz by running the UNMODIFIED reference module
(/root/reference/model/bilinear.py) under the installed torch (CPU, fp32) -- the reference has no
tests or golden vectors of its own, so these are "outputs of the reference itself
run here".  /root/reference exists only in the build container; the .npz is what travels.

    python tests/golden/make_golden.py

Per case (seed s in {0,1,2}, batch B in {2,64,257}):
  * weights  = model.bilinear.load(cpu) after torch.manual_seed(s)          (model/bilinear.py:58-92)
  * x, y, masks = oracle.bilinear_oracle.synthetic_batch(1000 + s, B)       
  * the five nn.Dropout children are swapped for a module multiplying by the injected mask * 2
    (state_dict unchanged: Dropout has no parameters)
  * literal sequence of train_bilinear.py:66-83 at step == 1:
        lr = config lr_decay.function(1); zero_grad; forward; MSELoss; backward;
        clip_grad_norm_(max_norm=1); optimizer.step()
  * then eval-mode forward of the updated model on the same x          (valid_bilinear.py:31,52)
Tensors above 4096 elements are stored as (l2 norm, sum, values at 1024 fixed random positions), smaller
ones whole; Adam moments as (l2 norm, sum) digests only; the initial state once per seed.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bilinear_oracle as bo  # noqa: E402

REF = "/root/reference/model/bilinear.py"
SEEDS = (0, 1, 2)
BATCHES = (2, 64, 257)
NSAMPLE = 1024
WHOLE_LIMIT = 4096


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_bilinear", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class InjectedDropout(nn.Module):
    """Dropout(p=0.5) with a caller-supplied keep mask (train) / identity (eval)."""

    def __init__(self):
        super().__init__()
        self.mask = None

    def forward(self, x):
        if not self.training:
            return x
        return x * self.mask.to(x.dtype) * 2.0


def inject(net):
    holders = []
    for seq in [net.encode, net.bilinear[0][0], net.bilinear[0][1], net.bilinear[1][0], net.bilinear[1][1]]:
        assert isinstance(seq[3], nn.Dropout) and seq[3].p == 0.5
        seq[3] = InjectedDropout()
        holders.append(seq[3])
    return holders


def positions_for(name, numel):
    # deterministic, independent of PYTHONHASHSEED
    seed = sum(ord(c) * (i + 1) for i, c in enumerate(name)) % (2 ** 31)
    return np.random.default_rng(seed).integers(0, numel, size=NSAMPLE)


def pack(out, prefix, name, t, digest_only=False):
    a = t.detach().cpu().numpy().astype(np.float32).reshape(-1)
    key = prefix + "/" + name
    if digest_only:
        out[key + "#l2"] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))
        out[key + "#sum"] = np.float64(a.astype(np.float64).sum())
    elif a.size <= WHOLE_LIMIT:
        out[key + "#full"] = a.reshape(tuple(t.shape))
    else:
        out[key + "#l2"] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))
        out[key + "#sum"] = np.float64(a.astype(np.float64).sum())
        out[key + "#samp"] = a[positions_for(name, a.size)]


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(8)
    ref = load_reference()
    out = {}
    for s in SEEDS:
        for B in BATCHES:
            prefix = "s%d_b%d" % (s, B)
            torch.manual_seed(s)
            net, opt, step, epoch = ref.load(torch.device("cpu"))
            assert step == 1 and epoch == 0
            holders = inject(net)
            if B == BATCHES[0]:
                for k, v in net.state_dict().items():
                    pack(out, "s%d/init" % s, k, v.float())
            x, y, masks = bo.synthetic_batch(1000 + s, B)
            for h, m in zip(holders, masks):
                h.mask = m
            net.train()
            lr = bo.lr_schedule(step)                      # train_bilinear.py:66-70 (step == 1 triggers)
            for grp in opt.param_groups:
                grp["lr"] = lr
            criterion = nn.MSELoss()
            opt.zero_grad()
            p = net(x)
            loss = criterion(p, y)
            loss.backward()
            pack(out, prefix, "p_train", p)
            out[prefix + "/loss"] = np.float64(loss.item())
            for (name, prm) in net.named_parameters():
                pack(out, prefix + "/grad", name, prm.grad)
            norm = nn.utils.clip_grad_norm_(net.parameters(), max_norm=1)
            out[prefix + "/total_norm"] = np.float64(float(norm))
            opt.step()
            for k, v in net.state_dict().items():
                pack(out, prefix + "/after", k, v.float())
            osd = opt.state_dict()
            names = [n for n, _ in net.named_parameters()]
            for i, name in enumerate(names):
                pack(out, prefix + "/exp_avg", name, osd["state"][i]["exp_avg"], digest_only=True)
                pack(out, prefix + "/exp_avg_sq", name, osd["state"][i]["exp_avg_sq"], digest_only=True)
                assert float(osd["state"][i]["step"]) == 1.0
            net.eval()
            with torch.no_grad():
                pack(out, prefix, "p_eval_after", net(x))
            print(prefix, "loss", float(loss), "norm", float(norm))
    # B = 1 in training mode: the reference raises (BatchNorm1d) -- pin the message
    torch.manual_seed(0)
    net, _, _, _ = ref.load(torch.device("cpu"))
    net.train()
    try:
        net(torch.zeros(1, 32))
        msg = ""
    except ValueError as e:
        msg = str(e)
    out["b1_train_error"] = np.array(msg)
    np.savez_compressed(os.path.join(HERE, "bilinear_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "bilinear_golden.npz"), os.path.getsize(os.path.join(HERE, "bilinear_golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
