"""The C-ABI library loads and exports every symbol include/sdrb200.h declares; without a GPU
the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import plan_path
from sdrreceiver_b200 import binding as B


def declared_functions(root):
    src = open(os.path.join(root, "include", "sdrb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdrb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(root):
    names = declared_functions(root)
    assert len(names) >= 30
    raw = C.CDLL(B.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    # and the binding covers the whole header
    L = B.lib()
    assert all(getattr(L, n).argtypes is not None for n in names)


def test_version_and_error_string():
    L = B.lib()
    assert b"sm_100a" in L.sdrb_version()
    h = C.c_void_p()
    assert L.sdrb_plan_from_ini(b"/nonexistent.ini", C.byref(h)) == -2
    assert b"cannot read" in L.sdrb_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    plan = B.Plan(plan_path("54W_288K"))
    with pytest.raises(B.SdrbError, match="no CUDA device"):
        B.Bank(plan, 1, 1)


def test_product_does_not_import_oracle(root):
    pkg = os.path.join(root, "sdrreceiver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("reference oracle", ""), (dirpath, f)
