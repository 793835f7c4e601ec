"""Generates tests/golden/golden_v1.npz: seeded inputs plus the outputs of the REFERENCE'S OWN
CPU code (oracle/_ref/libgraphlily_ref.so, compiled from /root/reference by oracle/Makefile).

Run in the container that has /root/reference:   python tests/golden/make_golden.py
The fixture travels with the repo, so machines without the reference tree (the GPU box)
still check the C restatement and the CUDA kernels against reference-produced answers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from graphlily_b200 import datasets  # noqa: E402
from graphlily_b200.io import CSRMatrix  # noqa: E402
from util import MASKS, SEMIRINGS  # noqa: E402


def main():
    ref = oracle.ref
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    rng = np.random.default_rng(20240521)
    out = {}

    def put_csr(prefix, m):
        out[prefix + "_shape"] = np.array([m.num_rows, m.num_cols], np.int64)
        out[prefix + "_indptr"], out[prefix + "_indices"], out[prefix + "_data"] = m.indptr, m.indices, m.data

    # --- SpMV, the shape of tests/test_module_spmv_spmspv.cpp:104-112,174 (scaled down) -------
    m = datasets.uniform_csr(2048, 2048, 10, seed=1)          # A[:] = 1/N
    put_csr("spmv", m)
    x = rng.integers(0, 2, m.num_cols).astype(np.float32)      # rand() % 2
    mask = rng.integers(0, 2, m.num_rows).astype(np.float32)
    out["spmv_x"], out["spmv_mask"] = x, mask
    for op, zero in SEMIRINGS:
        for mt in MASKS:
            out[f"spmv_y_op{op}_m{mt}"] = ref.spmv(m, op, zero, mt, x, mask)
    # a power-law matrix with long rows, empty rows and general values
    p = datasets.powerlaw_csr(1500, 1500, 40000, seed=2, max_degree=6000)
    p.data = rng.random(p.nnz).astype(np.float32)
    put_csr("pl", p)
    xp = (rng.integers(0, 4, p.num_cols) * rng.random(p.num_cols)).astype(np.float32)
    out["pl_x"] = xp
    for op, zero in SEMIRINGS:
        out[f"pl_y_op{op}"] = ref.spmv(p, op, zero, 0, xp)

    # --- SpMSpV, tests/test_module_spmv_spmspv.cpp:197-214,286-313 --------------------------
    ip, ix, d = ref.csr2csc(m)
    csc = CSRMatrix(m.num_rows, m.num_cols, np.ones_like(d), ix, ip)
    put_csr("spmspv_csc", csc)
    idx = np.arange(0, m.num_cols, 4, dtype=np.uint32)            # stride N / nnz
    val = (rng.integers(0, 10, len(idx)) / 10).astype(np.float32)  # (rand() % 10) / 10
    out["spmspv_x_idx"], out["spmspv_x_val"] = idx, val
    for op, zero in SEMIRINGS:
        smask = np.where(rng.random(m.num_rows) < 0.5, np.float32(zero), np.float32(1)).astype(np.float32)
        out[f"spmspv_mask_op{op}"] = smask
        for mt in MASKS:
            out[f"spmspv_y_op{op}_m{mt}"] = ref.spmspv(csc, op, zero, mt, idx, val, smask)

    # --- apply operators, tests/test_module_apply.cpp:54-206 ------------------------------
    v = rng.random(8192).astype(np.float32)
    am = rng.integers(0, 2, 8192).astype(np.float32)
    out["apply_in"], out["apply_mask"] = v, am
    out["apply_ewise_add"] = ref.ewise_add(v, 0.25)
    out["apply_assign_dense_m1"] = ref.assign_dense(am, v, 23.0, 1)[1]
    out["apply_assign_dense_m2"] = ref.assign_dense(am, v, 23.0, 2)[1]
    sidx = rng.choice(8192, 2000, replace=False).astype(np.uint32)
    sval = rng.random(2000).astype(np.float32)
    out["apply_sparse_idx"], out["apply_sparse_val"] = sidx, sval
    out["apply_assign_sparse"] = ref.assign_sparse(sidx, v, 7.0)
    a, fi, fv = ref.assign_sparse_relax(sidx, sval, v)
    out["apply_relax_inout"], out["apply_relax_idx"], out["apply_relax_val"] = a, fi, fv

    # --- apps, tests/test_app.cpp:51-135 (source 0, 10 iterations) --------------------------
    g = datasets.powerlaw_graph(2048, 30000, seed=3, diagonal=False)   # already a multiple of 128
    put_csr("app", g)
    out["app_bfs"] = ref.bfs(g, 0, 10)
    gp = CSRMatrix(g.num_rows, g.num_cols, ref.normalize_outdegree(g) * np.float32(0.9), g.indices, g.indptr)
    out["app_pagerank"] = ref.pagerank(gp, 0.9, 10)
    sip, six, sd = ref.sssp_preprocess(g)
    out["app_sssp_indptr"], out["app_sssp_indices"], out["app_sssp_data"] = sip, six, sd
    out["app_sssp"] = ref.sssp(CSRMatrix(g.num_rows, g.num_cols, sd, six, sip), 0, 10)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
