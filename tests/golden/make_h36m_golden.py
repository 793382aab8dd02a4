"""Generates tests/golden/h36m_golden.npz by running the UNMODIFIED reference dataset
(/root/reference/H36M/data.py, position_only=True) on the synthetic pickles of tests/_h36m_synth.py.
`vectormath` (imported at module level by H36M/data.py:8 and H36M/util.py:6, unused on the pose-only
path) is not installed here and is stubbed.  /root/reference exists only in the build container.

    python tests/golden/make_h36m_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
import _h36m_synth  # noqa: E402

vm = types.ModuleType("vectormath")
vm.Vector2 = type("Vector2", (), {})
sys.modules["vectormath"] = vm
sys.modules["__main__"].Center = _h36m_synth.Center
sys.path.insert(0, "/root/reference")
import H36M  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as tmp:
    d = _h36m_synth.write_dataset(tmp)
    for task in (H36M.Task.Train, H36M.Task.Valid):
        ds = H36M.Dataset(data_dir=d, task=task, protocol=H36M.Protocol.GT)
        parts, Ss, acts = [], [], []
        for i in range(len(ds)):
            sample, image, heatmap, action = ds[i]
            assert image == -1 and heatmap == -1
            parts.append(sample[H36M.Annotation.Part])
            Ss.append(sample[H36M.Annotation.S])
            acts.append(action)
            mean = sample[H36M.Annotation.Mean_Of + H36M.Annotation.S]
            std = sample[H36M.Annotation.Stddev_Of + H36M.Annotation.S]
        out[task + "/part"] = np.stack(parts)
        out[task + "/S"] = np.stack(Ss)
        out[task + "/action"] = np.array(acts)
        out[task + "/mean_S"] = mean
        out[task + "/stddev_S"] = std
np.savez_compressed(os.path.join(HERE, "h36m_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
