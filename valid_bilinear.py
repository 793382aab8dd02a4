"""Drop-in for the reference's valid_bilinear.py: same flow and the same log lines
('{action}: {MPJPE}' per action, then 'avg: {MPJPE}'), followed by the Procrustes-aligned numbers.

    python valid_bilinear.py                          # one process, one B200
    torchrun --nproc-per-node N valid_bilinear.py     # the validation set is sharded by pose index; the only
                                                      # communication is one all-reduce of <= 15 per-action sums

The reference un-normalises, takes per-pose joint distances on the device and then does ONE `.cpu()` round
trip PER POSE to accumulate per action in float64 (valid_bilinear.py:61-70).  Here the per-action float64
sums are kept on the device (csrc/eval.cu) and read back once at the end.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bilinear_b200 import config as _config
from bilinear_b200 import h36m as H36M
from bilinear_b200 import metrics, parallel
from bilinear_b200.log import get_logger
from bilinear_b200.model import bilinear as model_bilinear

try:
    from tqdm import tqdm
except Exception:                       # pragma: no cover
    tqdm = None


def main(cfg=None, save_root='save'):
    cfg = cfg or _config.bilinear
    assert cfg.comment is not None
    rank, local_rank, world = parallel.init()
    device = torch.device('cuda', local_rank) if (world > 1 or str(cfg.device).startswith('cuda')) else cfg.device
    logger, log_dir, comment = get_logger(comment=cfg.comment, root=save_root)

    data = H36M.PoseDataset(data_dir=cfg.data_dir, task=H36M.Task.Valid, protocol=cfg.protocol)
    lo, hi = parallel.shard_range(len(data), rank, world)
    part = data.part[lo:hi].to(device)
    S = data.S[lo:hi].to(device)
    action = data.action_id[lo:hi].to(device)
    mean, stddev = data.mean_S.to(device), data.stddev_S.to(device)
    n_actions = len(data.action_names)

    bilinear, optimizer, step, train_epoch = model_bilinear.load(
        device=device,
        parameter_dir='{log_dir}/parameter'.format(log_dir=log_dir),
    )
    bilinear.set_precision(cfg.get('precision', 'fp32'))
    bilinear.eval()

    acc = (torch.zeros(n_actions, dtype=torch.float64, device=device),
           torch.zeros(n_actions, dtype=torch.float64, device=device),
           torch.zeros(n_actions, dtype=torch.int64, device=device))
    n = hi - lo
    nb = (n + cfg.batch_size - 1) // cfg.batch_size
    progress = tqdm(total=nb, desc='%d epoch' % train_epoch) if (tqdm and rank == 0) else None
    with torch.set_grad_enabled(False):
        for b0 in range(0, n, cfg.batch_size):
            in_image_space = part[b0:b0 + cfg.batch_size]
            in_camera_space = S[b0:b0 + cfg.batch_size]
            prediction = bilinear(in_image_space)
            # restore mm units, per-pose joint distances, per-action float64 accumulation -- on the device
            metrics.evaluate(prediction, in_camera_space, mean, stddev, action[b0:b0 + cfg.batch_size], n_actions,
                             joints=16, accum=acc)
            if progress:
                progress.update(1)
    if progress:
        progress.close()
    parallel.combine_action_sums(*acc)
    total_dist, total_pdist, total = acc[0].cpu(), acc[1].cpu(), acc[2].cpu()

    result = {}
    if rank == 0:
        logger.info('===========================================================')
        dist = 0.0
        pdist = 0.0
        cnt = 0
        for i, key in enumerate(data.action_names):
            if int(total[i]) == 0:
                continue
            value = float(total_dist[i])
            logger.info('{action}: {MPJPE}'.format(action=key, MPJPE=value / (int(total[i]) * 16)))
            result[key] = value / (int(total[i]) * 16)
            dist = dist + value
            pdist = pdist + float(total_pdist[i])
            cnt = cnt + int(total[i]) * 16
        logger.info('avg: {MPJPE}'.format(MPJPE=dist / cnt))
        logger.info('===========================================================')
        for i, key in enumerate(data.action_names):
            if int(total[i]):
                logger.info('{action} (Procrustes-aligned): {v}'.format(action=key, v=float(total_pdist[i]) / (int(total[i]) * 16)))
        logger.info('avg (Procrustes-aligned): {v}'.format(v=pdist / cnt))
        logger.info('===========================================================')
        result['avg'] = dist / cnt
        result['avg_pmpjpe'] = pdist / cnt
    parallel.barrier()
    return result


if __name__ == '__main__':
    main()
