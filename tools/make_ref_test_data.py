#!/usr/bin/env python
"""Writes the datasets the reference's own tests name (tests/test_module_spmv_spmspv.cpp:144-265,
tests/test_app.cpp:57-125: dense_32, dense_1K, uniform_10K_10, gplus_108K_13M, all *_csr_float32.npz) into
a directory, as seeded synthetic matrices of those shapes (the real files are not shipped); point
GLB_DATASET_DIR at it.  Usage: python tools/make_ref_test_data.py <dir> [gplus_scale]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlily_b200 import datasets, io  # noqa: E402
from graphlily_b200.io import CSRMatrix  # noqa: E402


def dense(n):
    return CSRMatrix(n, n, np.ones(n * n, np.float32), np.tile(np.arange(n, dtype=np.uint32), n),
                     (np.arange(n + 1, dtype=np.uint64) * n).astype(np.uint32))


if __name__ == "__main__":
    out = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    os.makedirs(out, exist_ok=True)
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    mats = {"dense_32": dense(32), "dense_1K": dense(1024),
            "uniform_10K_10": datasets.uniform_csr(10000, 10000, 10, seed=1, value=1.0),
            "gplus_108K_13M": datasets.c3_gplus(scale, device=dev)}
    for name, m in mats.items():
        io.save_csr_matrix_to_npz(os.path.join(out, f"{name}_csr_float32.npz"), m, compressed=False)
        print(f"{name}: {m.num_rows} x {m.num_cols}, nnz {m.nnz}")
