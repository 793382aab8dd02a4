"""Drop-in for the reference's train_bilinear.py: same flow, same log lines, same checkpoint files -- the
hot loop (train_bilinear.py:60-90) runs as one persistent sm_100a kernel per iteration.

    python train_bilinear.py                                 # one process, one B200 (reference behaviour)
    torchrun --nproc-per-node N train_bilinear.py            # data parallel: batch_size per GPU, local BN,
                                                             # NCCL all-reduce of the flat gradient

Configuration lives in bilinear_b200/config.py (util/config.py of the reference, values verbatim).

Differences that are deliberate (hard part 5):
  * the dataset is resident on the device and shuffled there -- no DataLoader workers (H36M/data.py:74-112);
  * the per-step `writer.add_scalar` / `loss.item()` of train_bilinear.py:86-88 force two device syncs per
    ~0.25 ms step; the same scalars are kept in a device ring and drained once per epoch, so TensorBoard
    receives the same (step, loss) pairs and the epoch log line shows the same last loss.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bilinear_b200 import config as _config
from bilinear_b200 import h36m as H36M
from bilinear_b200 import parallel
from bilinear_b200.log import get_logger
from bilinear_b200.model import bilinear as model_bilinear

try:                                    # tensorboardX is optional here (not installed in every image)
    from tensorboardX import SummaryWriter
except Exception:                       # pragma: no cover
    SummaryWriter = None
try:
    from tqdm import tqdm
except Exception:                       # pragma: no cover
    tqdm = None


def main(cfg=None, save_root='save'):
    cfg = cfg or _config.bilinear
    rank, local_rank, world = parallel.init()
    device = torch.device('cuda', local_rank) if (world > 1 or str(cfg.device).startswith('cuda')) else cfg.device
    logger, log_dir, comment = get_logger(comment=cfg.comment, root=save_root)

    if cfg.comment is None or not os.path.exists('{log_dir}/parameter'.format(log_dir=log_dir)):
        logger.info('                                                           ')
        logger.info('                                                           ')
        logger.info('===========================================================')
        logger.info('Comment        : ' + comment + '                           ')
        logger.info('===========================================================')
        logger.info('Architecture   : ' + 'Bilinear' + '                        ')
        logger.info('   -protocol   : ' + cfg.protocol + '          ')
        logger.info('   -task       : ' + H36M.Task.Train + '                   ')
        logger.info('   -device     : ' + str(device) + '       ')
        logger.info('===========================================================')
        logger.info('Data           : ' + 'Human3.6M' + '                       ')
        logger.info('   -directory  : ' + cfg.data_dir + '          ')
        logger.info('   -mini batch : ' + str(cfg.batch_size) + '   ')
        logger.info('   -shuffle    : ' + 'True' + '                            ')
        logger.info('   -worker     : ' + str(cfg.num_workers) + '  ')
        logger.info('===========================================================')

    data = H36M.PoseDataset(data_dir=cfg.data_dir, task=H36M.Task.Train, protocol=cfg.protocol).to(device)

    bilinear, optimizer, step, train_epoch = model_bilinear.load(
        device=device,
        parameter_dir='{log_dir}/parameter'.format(log_dir=log_dir) if cfg.comment is not None else None,
    )
    bilinear.set_precision(cfg.get('precision', 'fp32'))
    writer = None
    if SummaryWriter is not None and rank == 0:
        writer = SummaryWriter(log_dir='{log_dir}/visualize'.format(log_dir=log_dir))

    bilinear.train()
    if world > 1:
        # one process per GPU: peer-mapped gradient buffers for the in-kernel all-reduce
        bilinear.enable_data_parallel()
    # the SAME permutation on every rank: rank 0's seed is broadcast (per-rank seeding must not fork the shards)
    gen = torch.Generator().manual_seed(parallel.broadcast_int(int(torch.initial_seed()) & 0x7fffffff, device))
    global_batch = cfg.batch_size * world
    steps_per_epoch = (len(data) + global_batch - 1) // global_batch
    ring = torch.zeros(steps_per_epoch, dtype=torch.float32, device=device)
    last_loss = float('nan')

    for epoch in range(train_epoch + 1, train_epoch + cfg.get('epochs', 10) + 1):
        progress = tqdm(total=steps_per_epoch, desc='%d epoch' % epoch) if (tqdm and rank == 0) else None
        first_step = step
        k = 0
        for part, S, _ in data.batches(global_batch, shuffle=True, generator=gen):
            if world > 1:       # this rank's rows of the global batch
                if part.shape[0] < 2 * world:
                    # ragged tail with fewer than two rows for some rank: BatchNorm cannot train on one row
                    # (the single-GPU reference fails the same way at N % 64 == 1).  Every rank sees the same
                    # shape, so every rank drops the batch -- none is left waiting in the gradient exchange.
                    if rank == 0:
                        logger.info('tail batch of {n} poses dropped ({w} ranks need 2 rows each)'.format(
                            n=part.shape[0], w=world))
                    continue
                lo, hi = parallel.shard_range(part.shape[0], rank, world)
                part, S = part[lo:hi], S[lo:hi]

            # Learning rate decay
            if cfg.lr_decay.activate and cfg.lr_decay.condition(step):
                lr = cfg.lr_decay.function(step)
                logger.info('Learning rate decay to {lr} (step: {step})'.format(lr=lr, step=step))
                for param_group in optimizer.param_groups:
                    param_group['lr'] = lr

            # zero_grad, forward, MSELoss, backward, clip_grad_norm_(max_norm=1), optimizer.step
            if world == 1:
                loss, _ = bilinear.train_step(optimizer, part, S, max_norm=1)
            else:   # gradients averaged over the ranks inside the same kernel (NVLink peer memory)
                loss, _ = bilinear.train_step_dp(optimizer, part, S, max_norm=1)
            ring[k].copy_(loss)
            k += 1
            if progress:
                progress.update(1)
            step = step + 1

        losses = ring[:k].cpu()                      # the epoch's one device -> host transfer
        if k:
            last_loss = float(losses[-1])
        if writer is not None:
            for i in range(k):
                writer.add_scalar('BI/loss', float(losses[i]), first_step + i)
        if progress:
            progress.set_postfix(loss=last_loss)
            progress.close()

        if rank == 0:
            save_dir = '{log_dir}/parameter'.format(log_dir=log_dir)
            if not os.path.exists(save_dir):
                os.makedirs(save_dir)
            save_to = '{save_dir}/{epoch}.save'.format(save_dir=save_dir, epoch=epoch, )
            torch.save(
                {
                    'epoch': epoch,
                    'step': step,
                    'state': bilinear.state_dict(),
                    'optimizer': optimizer.state_dict(),
                },
                save_to,
            )
            logger.info('Epoch {epoch} saved (loss: {loss})'.format(epoch=epoch, loss=last_loss))
        parallel.barrier()

    if writer is not None:
        writer.close()
    return last_loss, step


if __name__ == '__main__':
    main()
