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
    d = _h36m_synth.write_dataset(tmp, protocols=("GT", "SH", "SH+FT"))
    # GT entries keep their historic keys ("train/part"); the detector protocols (H36M/protocol.py:3-4) are
    # stored under "SH/train/part", "SH+FT/train/part", ...
    for protocol, task in [(p, t) for p in (H36M.Protocol.GT, H36M.Protocol.SH, H36M.Protocol.SH_FT)
                           for t in (H36M.Task.Train, H36M.Task.Valid)]:
        ds = H36M.Dataset(data_dir=d, task=task, protocol=protocol)
        prefix = task if protocol == H36M.Protocol.GT else protocol + "/" + task
        parts, Ss, acts = [], [], []
        for i in range(len(ds)):
            sample, image, heatmap, action = ds[i]
            assert image == -1 and heatmap == -1
            parts.append(sample[H36M.Annotation.Part])
            Ss.append(sample[H36M.Annotation.S])
            acts.append(action)
            mean = sample[H36M.Annotation.Mean_Of + H36M.Annotation.S]
            std = sample[H36M.Annotation.Stddev_Of + H36M.Annotation.S]
        out[prefix + "/part"] = np.stack(parts)
        out[prefix + "/S"] = np.stack(Ss)
        out[prefix + "/action"] = np.array(acts)
        out[prefix + "/mean_S"] = mean
        out[prefix + "/stddev_S"] = std
np.savez_compressed(os.path.join(HERE, "h36m_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
