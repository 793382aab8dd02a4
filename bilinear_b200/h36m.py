"""Human3.6M constants and the pose-only data feed of the reference.

    H36M/protocol.py:1-4      Protocol    which 2D detections feed the lifter (file {task}_{protocol}.bin)
    H36M/annotation.py:1-9    Annotation  keys of the per-sample dict
    H36M/task.py:1-9          Task, tasks
    H36M/util.py:13-22        decode_image_name
    H36M/data.py:17-59,71-112 Dataset with position_only=True: pickle -> drop 2D joint 9 -> root-centre the
                              3D pose and drop the pelvis -> flatten -> z-score with the TRAIN statistics

The reference serves one pose per __getitem__ through a DataLoader with 8 workers and re-sends the (48,)
mean / std with every sample.  Here the whole z-scored set is built once as two arrays -- (N, 32) and
(N, 48) float32 -- which `PoseDataset.to(device)` keeps resident in HBM; batches are index_select's of a
device-side permutation.  Values are bit-identical to the reference's per-sample arithmetic (same numpy
float32 operations in the same order; tests/test_h36m.py).
"""
import os
import pickle

import numpy as np
import torch


class Protocol:
    GT = 'GT'  # Ground-truth
    SH = 'SH'  # Stacked hourglass
    SH_FT = 'SH+FT'  # Fine tuned stacked hourglass


class Annotation:
    S = 'S'  # 3D position
    Center = 'center'  # 2D position in image space
    Part = 'part'  # 2D position
    Scale = 'scale'  # Bounding box scale
    Image = 'image'  # Image name
    Root_Of = 'root of '  # Root position
    Mean_Of = 'mean of '  # Mean of position
    Stddev_Of = 'stddev of '  # Stddev of position


class Task:
    Train = 'train'
    Valid = 'valid'


tasks = [Task.Train, Task.Valid]

# the 15 Human3.6M actions (image names look like S1_Walking_1.54138969_000001.jpg)
ACTIONS = ['Directions', 'Discussion', 'Eating', 'Greeting', 'Phoning', 'Photo', 'Posing', 'Purchases', 'Sitting',
           'SittingDown', 'Smoking', 'Waiting', 'WalkDog', 'Walking', 'WalkTogether']


def decode_image_name(image_name):
    """H36M/util.py:13-22."""
    subject_action, camera_frame, _ = image_name.split('.')
    split = subject_action.split('_')
    subject = split[0]
    action = split[1]
    if len(split) >= 3:
        action = action + '_' + split[2]
    camera, frame = camera_frame.split('_')
    return subject, action, camera, frame


def action_key(action):
    """valid_bilinear.py:64: sub-action 1 / 2 are not told apart."""
    return action.split('_')[0]


class _Opaque:
    """Stand-in for pickled objects of classes that are not importable here (the real pickles hold
    vectormath.Vector2 centres, which the pose-only path never touches)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return _Opaque


def _prepare(raw):
    """H36M/data.py:36-59 for one task: returns (part (N,32), S (N,48)) float32 before z-scoring and their
    per-column mean / std."""
    out = {}
    for dim, anno in zip([2, 3], [Annotation.Part, Annotation.S]):
        a = np.asarray(raw[anno], dtype=np.float32)
        if anno == Annotation.Part:
            a = np.delete(a, 9, axis=1)                       # 16 joints, the nose is dropped
        else:
            root = np.expand_dims(a[:, 0, :], 1)
            a = (a - root)[:, 1:, :]                          # root-centred, pelvis removed
        a = np.reshape(a, (-1, dim * 16))
        out[anno] = a
        out[Annotation.Mean_Of + anno] = np.mean(a, axis=0)
        out[Annotation.Stddev_Of + anno] = np.std(a, axis=0)
    return out


class PoseDataset:
    """Pose-only view of H36M.Dataset(data_dir, task, position_only=True, protocol)."""

    def __init__(self, data_dir, task, protocol=Protocol.GT):
        assert task in tasks
        assert protocol in [Protocol.GT, Protocol.SH, Protocol.SH_FT]
        assert os.path.exists(data_dir) and 'Human3.6M' in data_dir      # H36M/data.py:21-23
        self.data_dir, self.task, self.protocol = data_dir, task, protocol
        data = {}
        for t in tasks:
            path = "{data_dir}/{task}_{protocol}.bin".format(data_dir=data_dir, task=t, protocol=protocol)
            with open(path, 'rb') as f:
                raw = _TolerantUnpickler(f).load()
            data[t] = _prepare(raw)
            data[t][Annotation.Image] = list(raw[Annotation.Image])
        tr = data[Task.Train]
        self.mean = {a: tr[Annotation.Mean_Of + a] for a in (Annotation.Part, Annotation.S)}
        self.stddev = {a: tr[Annotation.Stddev_Of + a] for a in (Annotation.Part, Annotation.S)}
        mine = data[task]
        # H36M/data.py:108-110: z-score with the TRAIN-set statistics (also for the validation set)
        self.part = torch.from_numpy((mine[Annotation.Part] - self.mean[Annotation.Part]) / self.stddev[Annotation.Part])
        self.S = torch.from_numpy((mine[Annotation.S] - self.mean[Annotation.S]) / self.stddev[Annotation.S])
        self.images = mine[Annotation.Image]
        self.actions = [decode_image_name(n)[1] for n in self.images]
        keys = sorted({action_key(a) for a in self.actions})
        self.action_names = keys
        lut = {k: i for i, k in enumerate(keys)}
        self.action_id = torch.tensor([lut[action_key(a)] for a in self.actions], dtype=torch.int32)
        self.mean_S = torch.from_numpy(np.asarray(self.mean[Annotation.S], dtype=np.float32))
        self.stddev_S = torch.from_numpy(np.asarray(self.stddev[Annotation.S], dtype=np.float32))

    def __len__(self):
        return len(self.images)

    def __getitem__(self, index):
        """Same 4-tuple as the reference's position-only __getitem__ (H36M/data.py:74-112)."""
        sample = {
            Annotation.Image: self.images[index],
            Annotation.S: self.S[index].numpy(),
            Annotation.Part: self.part[index].numpy(),
            Annotation.Mean_Of + Annotation.S: self.mean_S.numpy(),
            Annotation.Stddev_Of + Annotation.S: self.stddev_S.numpy(),
        }
        return sample, -1, -1, self.actions[index]

    def to(self, device, pin=False):
        """Keep the whole set on `device` (HBM-resident feed) or pinned on the host."""
        for name in ("part", "S", "action_id", "mean_S", "stddev_S"):
            t = getattr(self, name)
            t = t.pin_memory() if (pin and not t.is_cuda and device == "cpu") else t.to(device)
            setattr(self, name, t)
        return self

    def batches(self, batch_size, shuffle=True, generator=None, drop_last=False):
        """Yield (part [b,32], S [b,48], action_id [b]) like DataLoader(shuffle=...) would order them; the
        permutation lives on the data's device, so no per-sample host work is left."""
        n = len(self)
        dev = self.part.device
        if shuffle:
            perm = torch.randperm(n, generator=generator).to(dev)
        else:
            perm = torch.arange(n, device=dev)
        for lo in range(0, n, batch_size):
            idx = perm[lo:lo + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            yield self.part.index_select(0, idx), self.S.index_select(0, idx), self.action_id.index_select(0, idx)


def shard(dataset_len, rank, world):
    """Contiguous slice of independent poses owned by `rank` (evaluation shards by pose, no collective)."""
    base, rem = divmod(dataset_len, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
