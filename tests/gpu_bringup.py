"""GPU bring-up harness (not a pytest module): exercises the tcgen05 GEMM path and the evaluation
kernels case by case, each case in its own subprocess so a trapped kernel cannot poison the rest.

    python tests/gpu_bringup.py            # run all groups, write gpurun_out/bringup.json
    python tests/gpu_bringup.py <group>    # run one group in-process (used by the driver above)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _gemm_case(torch, L, m, n, k, planes, a_mn, b_mn, bn, seed=0):
    from bilinear_b200._lib import check, ptr
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(m, k, generator=g)
    B = torch.randn(n, k, generator=g)
    dev = torch.device("cuda")
    A_d, B_d = A.to(dev), B.to(dev)
    Ap = torch.empty(planes, m, k, dtype=torch.bfloat16, device=dev)
    Bp = torch.empty(planes, n, k, dtype=torch.bfloat16, device=dev)
    check(L.b200_debug_split(ptr(A_d), ptr(Ap), A.numel(), planes, None))
    check(L.b200_debug_split(ptr(B_d), ptr(Bp), B.numel(), planes, None))
    torch.cuda.synchronize()
    # what the planes actually hold
    A_eff = Ap.double().sum(0).cpu()
    B_eff = Bp.double().sum(0).cpu()
    split_err = float((A_eff - A.double()).abs().max() / A.abs().max())
    A_in = Ap.transpose(1, 2).contiguous() if a_mn else Ap
    B_in = Bp.transpose(1, 2).contiguous() if b_mn else Bp
    D = torch.full((m, n), float("nan"), device=dev)
    scratch = torch.zeros(1 << 17, dtype=torch.uint8, device=dev)
    off = (-scratch.data_ptr()) % 1024
    import ctypes as C
    check(L.b200_debug_gemm(ptr(A_in), ptr(B_in), ptr(D), m, n, k, planes, int(a_mn), int(b_mn), bn,
                            C.c_void_p(scratch.data_ptr() + off), 65536, None))
    torch.cuda.synchronize()
    ref = A_eff @ B_eff.t()
    got = D.double().cpu()
    err = (got - ref).abs()
    nan = int(torch.isnan(got).sum())
    rel = float(err.nan_to_num(1e30).max() / ref.abs().max())
    info = dict(m=m, n=n, k=k, planes=planes, a_mn=int(a_mn), b_mn=int(b_mn), bn=bn, rel_err=rel, nan=nan,
                split_err=split_err)
    if rel > 1e-3:
        bad = (err > 1e-3 * ref.abs().max()) | torch.isnan(got)
        info["bad_frac"] = float(bad.double().mean())
        info["bad_rows"] = [int(i) for i in torch.nonzero(bad.any(1)).flatten()[:8]]
        info["bad_cols"] = [int(i) for i in torch.nonzero(bad.any(0)).flatten()[:8]]
        info["got_sample"] = [float(v) for v in got.flatten()[:4]]
        info["ref_sample"] = [float(v) for v in ref.flatten()[:4]]
    return info


def group_gemm_basic(torch, L):
    out = []
    for (m, n, k, bn) in [(128, 64, 64, 64), (128, 64, 256, 64), (256, 256, 128, 256), (1024, 64, 1024, 64),
                          (128, 32, 64, 32), (1024, 512, 1024, 256)]:
        out.append(_gemm_case(torch, L, m, n, k, 1, 0, 0, bn))
    return out


def group_gemm_ragged(torch, L):
    out = []
    for (m, n, k, bn) in [(48, 64, 1024, 64), (1024, 2, 32, 16), (1024, 257, 1024, 256), (1024, 70, 48, 80),
                          (1024, 32, 200, 32), (100, 300, 72, 128)]:
        out.append(_gemm_case(torch, L, m, n, k, 1, 0, 0, bn))
    return out


def group_gemm_planes3(torch, L):
    out = []
    for (m, n, k, bn) in [(128, 64, 64, 64), (1024, 64, 1024, 64), (1024, 257, 1024, 128), (48, 64, 1024, 64)]:
        out.append(_gemm_case(torch, L, m, n, k, 3, 0, 0, bn))
    return out


def group_gemm_mn_a(torch, L):
    out = []
    for (m, n, k, bn) in [(128, 64, 64, 64), (256, 128, 256, 128), (1024, 64, 1024, 64)]:
        out.append(_gemm_case(torch, L, m, n, k, 1, 1, 0, bn))
    return out


def group_gemm_mn_b(torch, L):
    out = []
    for (m, n, k, bn) in [(128, 64, 64, 64), (256, 128, 256, 128), (1024, 256, 1024, 256)]:
        out.append(_gemm_case(torch, L, m, n, k, 1, 0, 1, bn))
    out.append(_gemm_case(torch, L, 256, 128, 256, 1, 1, 1, 128))
    return out


def group_eval(torch, L):
    import ctypes as C
    from bilinear_b200._lib import check, ptr
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import pose_metrics as pm
    out = []
    dev = torch.device("cuda")
    for (n, J) in [(1, 16), (1000, 16), (128 * 7 + 5, 17), (200000, 16)]:
        d = pm.synthetic_eval_poses(n, J, seed=1)
        pred, gt, mean, std, act = d["pred_norm"], d["gt_norm"], d["mean"], d["std"], d["action"]
        ref_m = pm.mpjpe_per_pose(pred, gt, mean, std)
        ref_p = pm.pmpjpe_per_pose(pred, gt, mean, std)
        t = lambda a: torch.from_numpy(a).to(dev)
        pr, g, me, sd, ac = t(pred), t(gt), t(mean), t(std), t(act)
        dist = torch.empty(n, device=dev)
        pdist = torch.empty(n, device=dev)
        acc = torch.zeros(15, dtype=torch.float64, device=dev)
        pacc = torch.zeros(15, dtype=torch.float64, device=dev)
        cnt = torch.zeros(15, dtype=torch.int64, device=dev)
        check(L.b200_mpjpe_pmpjpe(ptr(pr), ptr(g), ptr(me), ptr(sd), n, J, ptr(ac), 15, ptr(dist), ptr(pdist),
                                  ptr(acc), ptr(pacc), ptr(cnt), None))
        torch.cuda.synchronize()
        dm = (dist.double().cpu().numpy() - ref_m)
        dp = (pdist.double().cpu().numpy() - ref_p)
        import numpy as np
        out.append(dict(n=n, J=J, mpjpe_max_abs_mm=float(np.abs(dm).max() / J), pmpjpe_max_abs_mm=float(np.abs(dp).max() / J),
                        mpjpe_agg_diff_mm=float(abs(acc.sum().item() - ref_m.sum()) / (n * J)),
                        pmpjpe_agg_diff_mm=float(abs(pacc.sum().item() - ref_p.sum()) / (n * J)),
                        count_ok=bool(int(cnt.sum().item()) == n)))
    return out


GROUPS = {
    "gemm_basic": group_gemm_basic,
    "gemm_ragged": group_gemm_ragged,
    "gemm_planes3": group_gemm_planes3,
    "gemm_mn_a": group_gemm_mn_a,
    "gemm_mn_b": group_gemm_mn_b,
    "eval": group_eval,
}


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] in GROUPS:
        import torch
        from bilinear_b200._lib import lib
        res = GROUPS[sys.argv[1]](torch, lib())
        print("RESULT " + json.dumps(res))
        return
    names = sys.argv[1:] or list(GROUPS)
    summary = {}
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=300)
            res = None
            for line in p.stdout.splitlines():
                if line.startswith("RESULT "):
                    res = json.loads(line[7:])
            summary[name] = dict(rc=p.returncode, secs=round(time.time() - t0, 1), result=res,
                                 tail=(p.stdout[-1500:] + "\n--stderr--\n" + p.stderr[-2500:]) if res is None else "")
        except subprocess.TimeoutExpired as e:
            summary[name] = dict(rc="timeout", secs=round(time.time() - t0, 1), result=None,
                                 tail=str(e.stdout)[-1000:] + str(e.stderr)[-1000:])
        print(name, json.dumps(summary[name])[:3000], flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "bringup.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
