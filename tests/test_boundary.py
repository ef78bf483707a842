"""CPU: the C-ABI library loads and exports every symbol include/goofer_b200.h declares; the product never
touches the oracle; compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "goofer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(goofer_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/goofer_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from goofer_b200 import capi
    assert sorted(capi.EXPORTS) == header_functions()


def test_version_and_struct_sizes(lib):
    from goofer_b200 import capi
    assert lib.goofer_version() == 5
    for which, rec in enumerate((capi.GooferSource, capi.GooferNote, capi.GooferNotePlanInfo, capi.GooferBatch, capi.GooferStats)):
        assert int(lib.goofer_struct_size(which)) == C.sizeof(rec)
    assert int(lib.goofer_struct_size(99)) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "goofer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in txt and "liboracle" not in txt and "goofer_oracle" not in txt, f"{f} references the oracle"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import numpy as np
    from goofer_b200 import capi
    x = np.zeros(1024, np.float32)
    S = np.zeros((513, 5, 2), np.float32)
    rc = lib.goofer_stft_batch(x.ctypes.data, 1, 1024, S.ctypes.data, None)
    assert rc == capi.ERR_CUDA
    assert b"CUDA" in lib.goofer_last_error() or b"device" in lib.goofer_last_error()


def test_invalid_descriptors_are_rejected(lib):
    from goofer_b200 import capi
    assert lib.goofer_plan_batch(None, None) == capi.ERR_INVALID
    b = capi.GooferBatch()
    b.n_notes = -1
    info = (capi.GooferNotePlanInfo * 1)()
    assert lib.goofer_plan_batch(C.byref(b), info) == capi.ERR_INVALID
    assert lib.goofer_render_batch(C.byref(b), None, 0, None) == capi.ERR_INVALID
    assert int(lib.goofer_workspace_bytes(C.byref(b), 0)) == 0
