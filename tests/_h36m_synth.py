"""Deterministic synthetic Human3.6M pickles ({train,valid}_GT.bin) in the layout H36M/data.py:33 reads:
keys 'S' (N,17,3), 'part' (N,17,2), 'center' (objects with .x/.y), 'scale', 'image'."""
import os
import pickle

import numpy as np

ACTS = ["Walking", "Walking_1", "Eating", "Eating_2", "Sitting_1", "Photo", "WalkDog_1", "Directions"]


class Center:
    def __init__(self, x, y):
        self.x, self.y = x, y


def make_raw(n, seed):
    rng = np.random.RandomState(seed)
    S = (rng.randn(n, 17, 3) * 300.0 + rng.randn(1, 17, 3) * 100.0).astype(np.float64)
    part = (rng.rand(n, 17, 2) * 1000.0).astype(np.float64)
    images = []
    for i in range(n):
        images.append("S%d_%s.%d_%06d.jpg" % (1 + i % 7, ACTS[(i * 5 + seed) % len(ACTS)], 54138969 + i % 4, i + 1))
    return {"S": [s for s in S], "part": [p for p in part], "center": [Center(float(i), float(2 * i)) for i in range(n)],
            "scale": [1.0 + 0.01 * i for i in range(n)], "image": images}


def make_raw_detected(n, seed, protocol):
    """{task}_SH.bin / {task}_SH+FT.bin (H36M/data.py:33; written by `SH preprocess.ipynb` cells 7-13): the SAME
    poses and image names as the GT file, with the 2D joints replaced by stacked-hourglass detections -- here
    the GT joints plus detector noise (SH+FT: the fine-tuned detector, less noise)."""
    raw = make_raw(n, seed)
    rng = np.random.RandomState(1000 + seed + (0 if protocol == "SH" else 1))
    sigma = 12.0 if protocol == "SH" else 6.0
    raw["part"] = [p + rng.randn(17, 2) * sigma for p in raw["part"]]
    return raw


def write_dataset(root, n_train=40, n_valid=24, protocols=("GT",)):
    d = os.path.join(root, "data", "Human3.6M")
    os.makedirs(d, exist_ok=True)
    for protocol in protocols:
        for task, n, seed in (("train", n_train, 1), ("valid", n_valid, 2)):
            raw = make_raw(n, seed) if protocol == "GT" else make_raw_detected(n, seed, protocol)
            with open(os.path.join(d, "%s_%s.bin" % (task, protocol)), "wb") as f:
                pickle.dump(raw, f)
    return d
