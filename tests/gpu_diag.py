"""Diagnostic table (not a pytest module): per-tensor errors of the CUDA path against the fp64 oracle."""
import json
import sys
import os
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn as nn
from _util import rel_l2
from oracle import bilinear_oracle as bo
from test_gpu_network import make_net, warmed_state, as_float_state, masks_tensor, ZERO_GRAD
from bilinear_b200.engine import get_engine


def grads_table(gv, grads):
    gn = float(bo.total_grad_norm(grads))
    out = {}
    for name in bo.param_names():
        if name in ZERO_GRAD:
            out[name] = float((gv[name].double().cpu() - grads[name]).norm()) / gn
        else:
            out[name] = rel_l2(gv[name], grads[name])
    return out


def main():
    cases = [("fp32", 64), ("fp32", 300), ("bf16", 64), ("bf16", 1000)]
    if len(sys.argv) > 1:
        cases = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
    for precision, batch in cases:
        for mode in ("autograd", "fused"):
            wseed = int(os.environ.get("DIAG_WSEED", "0"))
            st = warmed_state(wseed)
            net, opt = make_net(wseed, precision, as_float_state(st))
            net.train()
            x, y, masks = bo.synthetic_batch(int(os.environ.get("DIAG_BSEED", "17")), batch)
            net.inject_dropout_masks(masks_tensor(masks))
            lr = bo.lr_schedule(1)
            for g in opt.param_groups:
                g["lr"] = lr
            eng = get_engine(net)
            names = [n for n, _ in net.named_parameters()]
            p_ref, cache = bo.forward(bo.State(as_float_state(st), dtype=torch.float64), x, train=True, masks=masks)
            if mode == "autograd":
                opt.zero_grad()
                p = net(x.cuda())
                loss = nn.MSELoss()(p, y.cuda())
                loss.backward()
                gnorm = torch.sqrt(sum((q.grad.double() ** 2).sum() for q in net.parameters()))
                gv = {n: q.grad for n, q in net.named_parameters()}
                perr = rel_l2(p, p_ref)
            else:
                loss, gnorm = net.train_step(opt, x.cuda(), y.cuda(), max_norm=1.0)
                gv = dict(zip(names, eng.grad_views()))
                perr = None
            _, loss_ref, grads, norm_ref = bo.train_step(st, x, y, masks, lr)
            tab = grads_table(gv, grads)
            print("== %s B=%d %s: p_err %s loss %.8g ref %.8g gnorm %.8g ref %.8g" %
                  (precision, batch, mode, perr, float(loss), float(loss_ref), float(gnorm), float(norm_ref)))
            for k, v in tab.items():
                print("   %-28s %.3e" % (k, v))
            sd = net.state_dict()
            for i, pre in enumerate(bo.HEAVY_PREFIXES):
                print("   rm%d %.2e rv%d %.2e nbt %d/%d" % (i, rel_l2(sd[pre + ".1.running_mean"], st.rm[i]), i,
                                                       rel_l2(sd[pre + ".1.running_var"], st.rv[i]),
                                                       int(sd[pre + ".1.num_batches_tracked"]), st.nbt[i]))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
