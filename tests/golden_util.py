import os

import numpy as np

from graphlily_b200.io import CSRMatrix

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
_cache = {}


def golden():
    if "z" not in _cache:
        _cache["z"] = dict(np.load(PATH))
    return _cache["z"]


def golden_csr(prefix):
    z = golden()
    shape = z[prefix + "_shape"]
    return CSRMatrix(int(shape[0]), int(shape[1]), z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"])
