"""util/config.py:13-25 of the reference, values verbatim, without the dotmap dependency.

Additional knobs of this implementation (reference-equal defaults):
    precision   'fp32' (1e-5 parity, 3 x bf16 split operands) | 'bf16'
    epochs      epochs per invocation of train_bilinear.py (the reference hard-codes 10, train_bilinear.py:56)
"""
import torch

from .h36m import Protocol


class _Map(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


bilinear = _Map({
    'comment': 'Bilinear GT',
    'batch_size': 64,
    'num_workers': 8,
    'device': torch.device('cuda' if torch.cuda.is_available() else 'cpu'),
    'data_dir': 'data/Human3.6M',
    'lr_decay': _Map({
        'activate': True,
        'condition': (lambda step: step % 100000 == 0 or step == 1),
        'function': (lambda step: 1.0e-3 * 0.96 ** (step / 100000)),
    }),
    'protocol': Protocol.GT,
    # ---- additions
    'precision': 'fp32',
    'epochs': 10,
})
