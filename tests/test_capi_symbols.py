"""The C-ABI library loads on a machine without a GPU, exports every symbol that
include/graphlily_b200.h declares, the ctypes table covers them all, and compute entry
points fail loudly (no CPU fallback) when no CUDA device exists."""
import ctypes
import os
import re

import numpy as np
import pytest

from graphlily_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "graphlily_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(glb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 35
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature"
    assert set(capi.SIGNATURES) <= set(names)
    assert capi.lib.glb_version() == 100


def test_library_is_self_contained():
    # static cudart, NCCL via dlopen: nothing CUDA-ish in DT_NEEDED
    import subprocess
    out = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not [n for n in needed if "cuda" in n or "nccl" in n or "torch" in n], needed


def test_sm100a_cubin_embedded():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(capi.device_count() > 0, reason="this check is for machines without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(capi.GlbError):
        capi.Context(0)
    assert b"" != capi.lib.glb_last_error()
    # argument validation happens before any device work and reports through the same channel
    h = ctypes.c_void_p()
    assert capi.lib.glb_ctx_create(0, None, None) == 1   # GLB_EINVAL


def test_idx_val_layout():
    assert capi.IDX_VAL.itemsize == 8
    a = capi.sparse_to_numpy([5, 9], [0.5, 1.5])
    assert a["index"].tolist() == [2, 5, 9] and a["val"].tolist()[1:] == [0.5, 1.5]
    assert ctypes.sizeof(capi.Epilogue) == 32 or ctypes.sizeof(capi.Epilogue) == 24
