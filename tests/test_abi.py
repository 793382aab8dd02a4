"""The C-ABI library loads without a GPU and exports every symbol include/bilinear_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

from bilinear_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "bilinear_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 15
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), "missing export: " + n


def test_binding_covers_header():
    assert sorted(_lib.EXPORTS) == _declared()


def test_constants_and_argument_checks_without_gpu():
    L = _lib.lib()
    assert L.b200_version() >= 100
    assert L.b200_param_numel() == _lib.PARAM_NUMEL == 4291632
    assert L.b200_workspace_bytes(64, _lib.PREC_FP32, _lib.PLAN_TRAIN) > 0
    assert L.b200_workspace_bytes(64, _lib.PREC_BF16, _lib.PLAN_INFER) > 0
    assert L.b200_workspace_bytes(0, _lib.PREC_BF16, _lib.PLAN_INFER) == -1
    assert L.b200_workspace_bytes(64, 2, _lib.PLAN_INFER) == -1
    # argument errors are reported through the return code + b200_last_error, never a crash
    assert L.b200_plan_create(None, None) != 0
    assert b"null" in L.b200_last_error()
    assert L.b200_forward_eval(None, None, None, None) != 0
    assert L.b200_mpjpe(None, None, None, None, 0, 16, None, 15, None, None, None, None) != 0


def test_struct_sizes_match_header():
    # layout agreed with the C side (x86-64 SysV): see include/bilinear_b200.h
    assert ctypes.sizeof(_lib.AdamArgs) == 32
    assert ctypes.sizeof(_lib.TrainArgs) == 4 * 8 + 2 * 8 + 5 * 4 + 4 + 2 * 8
    assert ctypes.sizeof(_lib.PlanDesc) == 16 + 4 * 8 + 3 * 5 * 8 + 16
