"""Host-side data path: PoseDataset against golden outputs of the UNMODIFIED reference H36M.Dataset
(tests/golden/make_h36m_golden.py), constants, config literals, checkpoint discovery, sharding."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _h36m_synth  # noqa: E402

from bilinear_b200 import config, h36m, parallel  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "h36m_golden.npz"))


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    return _h36m_synth.write_dataset(str(tmp_path_factory.mktemp("h36m")))


@pytest.mark.parametrize("task", ["train", "valid"])
def test_pose_dataset_is_bit_identical_to_reference(gold, data_dir, task):
    ds = h36m.PoseDataset(data_dir, task, h36m.Protocol.GT)
    assert len(ds) == gold[task + "/part"].shape[0]
    assert np.array_equal(ds.part.numpy(), gold[task + "/part"])
    assert np.array_equal(ds.S.numpy(), gold[task + "/S"])
    # the statistics are the TRAIN set's for both tasks (H36M/data.py:87-94,108-110)
    assert np.array_equal(ds.mean_S.numpy(), gold["train/mean_S"])
    assert np.array_equal(ds.stddev_S.numpy(), gold["train/stddev_S"])
    assert np.array_equal(gold[task + "/mean_S"], gold["train/mean_S"])
    assert list(ds.actions) == list(gold[task + "/action"])
    sample, image, heatmap, action = ds[3]
    assert image == -1 and heatmap == -1 and action == gold[task + "/action"][3]
    assert np.array_equal(sample[h36m.Annotation.Part], gold[task + "/part"][3])
    assert ds.part.dtype == torch.float32 and ds.S.shape[1] == 48


def test_action_buckets_follow_valid_bilinear(data_dir):
    ds = h36m.PoseDataset(data_dir, "valid")
    for a, i in zip(ds.actions, ds.action_id.tolist()):
        assert ds.action_names[i] == a.split('_')[0]            # valid_bilinear.py:64
    assert h36m.decode_image_name("S1_Walking_1.54138969_000001.jpg") == ("S1", "Walking_1", "54138969", "000001")
    assert h36m.decode_image_name("S11_Photo.60457274_000123.jpg") == ("S11", "Photo", "60457274", "000123")


def test_batches_cover_every_pose_once(data_dir):
    ds = h36m.PoseDataset(data_dir, "train")
    g = torch.Generator().manual_seed(5)
    seen = []
    sizes = []
    for part, S, act in ds.batches(16, shuffle=True, generator=g):
        sizes.append(part.shape[0])
        for row in part:
            seen.append(int((ds.part == row).all(dim=1).nonzero()[0]))
    assert sizes == [16, 16, 8] and sorted(seen) == list(range(40)) and seen != list(range(40))
    assert [p.shape[0] for p, _, _ in ds.batches(16, shuffle=False, drop_last=True)] == [16, 16]


def test_missing_or_wrong_directory_asserts(tmp_path):
    with pytest.raises(AssertionError):
        h36m.PoseDataset(str(tmp_path / "nope" / "Human3.6M"), "train")
    (tmp_path / "elsewhere").mkdir()
    with pytest.raises(AssertionError):
        h36m.PoseDataset(str(tmp_path / "elsewhere"), "train")     # path must contain 'Human3.6M' (H36M/data.py:23)
    with pytest.raises(AssertionError):
        h36m.PoseDataset(str(tmp_path), "test")


def test_constants_and_config_are_verbatim():
    assert (h36m.Protocol.GT, h36m.Protocol.SH, h36m.Protocol.SH_FT) == ("GT", "SH", "SH+FT")
    assert h36m.Annotation.Mean_Of + h36m.Annotation.S == "mean of S"
    assert h36m.Annotation.Stddev_Of + h36m.Annotation.S == "stddev of S"
    assert h36m.tasks == ["train", "valid"]
    c = config.bilinear
    assert (c.comment, c.batch_size, c.num_workers, c.data_dir, c.protocol) == ("Bilinear GT", 64, 8, "data/Human3.6M", "GT")
    assert c.lr_decay.activate and c.lr_decay.condition(1) and c.lr_decay.condition(200000)
    assert not c.lr_decay.condition(2)
    assert c.lr_decay.function(1) == 1.0e-3 * 0.96 ** (1 / 100000) and c.lr_decay.function(1) < 1.0e-3


def test_checkpoint_discovery_like_reference(tmp_path):
    from bilinear_b200.model.bilinear import _latest_epoch
    assert _latest_epoch(None) == 0 and _latest_epoch(str(tmp_path)) == 0
    for e in (1, 2, 10):
        (tmp_path / ("%d.save" % e)).write_bytes(b"")
    assert _latest_epoch(str(tmp_path)) == 10
    (tmp_path / "notes.txt.bak").write_bytes(b"")
    with pytest.raises(ValueError):                                 # model/bilinear.py:68: file.split('.') must give two parts
        _latest_epoch(str(tmp_path))


def test_shards_partition_the_poses():
    for n, w in ((109867, 8), (10, 3), (7, 8)):
        r = [parallel.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
        assert [h36m.shard(n, k, w) for k in range(w)] == r


# ------------------------------------------------------------------------------------------ detector protocols
@pytest.fixture(scope="module")
def data_dir_all(tmp_path_factory):
    return _h36m_synth.write_dataset(str(tmp_path_factory.mktemp("h36m_all")), protocols=("GT", "SH", "SH+FT"))


@pytest.mark.parametrize("protocol", [h36m.Protocol.SH, h36m.Protocol.SH_FT])
@pytest.mark.parametrize("task", ["train", "valid"])
def test_detector_protocols_are_bit_identical_to_reference(gold, data_dir_all, protocol, task):
    """H36M/protocol.py:3-4, H36M/data.py:33: `Protocol.SH` / `Protocol.SH_FT` read {task}_SH.bin / {task}_SH+FT.bin --
    the same poses with stacked-hourglass 2D detections in place of the ground-truth joints.  Golden: the UNMODIFIED
    H36M.Dataset on the same synthetic pickles (tests/golden/make_h36m_golden.py)."""
    ds = h36m.PoseDataset(data_dir_all, task, protocol)
    pre = protocol + "/" + task
    assert np.array_equal(ds.part.numpy(), gold[pre + "/part"])
    assert np.array_equal(ds.S.numpy(), gold[pre + "/S"])
    assert np.array_equal(ds.mean_S.numpy(), gold[protocol + "/train/mean_S"])
    assert np.array_equal(ds.stddev_S.numpy(), gold[protocol + "/train/stddev_S"])
    assert list(ds.actions) == list(gold[pre + "/action"])
    # the 3D side is the GT file's; only the 2D inputs (and therefore their z-score statistics) differ
    gt = h36m.PoseDataset(data_dir_all, task, h36m.Protocol.GT)
    assert np.array_equal(ds.S.numpy(), gt.S.numpy()) and not np.array_equal(ds.part.numpy(), gt.part.numpy())


def test_protocol_selects_the_file(tmp_path):
    d = _h36m_synth.write_dataset(str(tmp_path))            # GT files only
    with pytest.raises(FileNotFoundError):
        h36m.PoseDataset(d, "train", h36m.Protocol.SH)      # data/Human3.6M/train_SH.bin (H36M/data.py:33)
    with pytest.raises(AssertionError):
        h36m.PoseDataset(d, "train", "SH_FT")               # not a Protocol value (H36M/data.py:22)
