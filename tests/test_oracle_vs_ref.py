"""Pins the C restatement (oracle/oracle.c) bit-for-bit against the reference's own CPU code
compiled from /root/reference (oracle/_ref, see oracle/ref_driver.cpp), on seeded inputs --
and checks that the fast member-install path of the driver equals the reference's full
load_and_format_matrix path.  Skipped where oracle/_ref is absent (it travels to the GPU box
prebuilt; the committed fixtures under tests/golden cover that case)."""
import numpy as np
import pytest

from graphlily_b200 import datasets
from graphlily_b200.io import CSRMatrix
from util import MASKS, SEMIRINGS, random_csr


@pytest.fixture(scope="module")
def both(oracle):
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built")
    return oracle.port, oracle.ref


def same(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mask_type", MASKS)
def test_spmv(both, op, zero, mask_type):
    port, ref = both
    rng = np.random.default_rng(100 + op * 10 + mask_type)
    for shape in [(1, 1), (37, 53), (300, 200), (128, 128)]:
        m = random_csr(rng, shape[0], shape[1], 0.1, values="rand" if op == 0 else "small")
        x = (rng.integers(0, 3, shape[1]) * rng.random(shape[1])).astype(np.float32)
        mask = rng.integers(0, 2, shape[0]).astype(np.float32)
        same(port.spmv(m, op, zero, mask_type, x, mask), ref.spmv(m, op, zero, mask_type, x, mask))


def test_spmv_fast_equals_full_reference_path(both):
    # fast=0 runs the reference's own load_and_format_matrix (CPSR formatter) first
    port, ref = both
    rng = np.random.default_rng(7)
    m = datasets.uniform_csr(1024, 1024, 10, seed=3)
    x = rng.integers(0, 2, 1024).astype(np.float32)
    for op, zero in SEMIRINGS:
        same(ref.spmv(m, op, zero, 0, x, fast=True), ref.spmv(m, op, zero, 0, x, fast=False))
        same(ref.spmv(m, op, zero, 0, x, fast=True), port.spmv(m, op, zero, 0, x))


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mask_type", MASKS)
def test_spmspv(both, op, zero, mask_type):
    port, ref = both
    rng = np.random.default_rng(200 + op * 10 + mask_type)
    for n in [1, 64, 500]:
        m = random_csr(rng, n, n, 0.08, values="rand" if op == 0 else "small")   # used as CSC
        k = int(rng.integers(0, n + 1))
        idx = rng.choice(n, size=k, replace=False).astype(np.uint32)
        val = (rng.integers(0, 10, k) / 10).astype(np.float32)
        mask = np.where(rng.random(n) < 0.5, np.float32(zero), np.float32(1)).astype(np.float32)
        same(port.spmspv(m, op, zero, mask_type, idx, val, mask), ref.spmspv(m, op, zero, mask_type, idx, val, mask))


def test_spmspv_fast_equals_full_reference_path(both):
    port, ref = both
    rng = np.random.default_rng(9)
    csr = datasets.uniform_csr(1024, 1024, 10, seed=5, value=1.0)
    ip, ix, d = port.csr2csc(csr)
    csc = CSRMatrix(1024, 1024, d, ix, ip)
    idx = np.arange(0, 1024, 7, dtype=np.uint32)
    val = np.ones(len(idx), np.float32)
    mask = np.zeros(1024, np.float32)
    a = ref.spmspv(csc, 1, 0.0, 1, idx, val, mask, fast=True)
    same(a, ref.spmspv(csc, 1, 0.0, 1, idx, val, mask, fast=False))
    same(a, port.spmspv(csc, 1, 0.0, 1, idx, val, mask))


def test_apply_ops(both):
    port, ref = both
    rng = np.random.default_rng(11)
    v = rng.random(1000).astype(np.float32)
    same(port.ewise_add(v, 0.3), ref.ewise_add(v, 0.3))
    mask = rng.integers(0, 2, 1000).astype(np.float32)
    for mt in (1, 2):
        rc_p, a = port.assign_dense(mask, v, 23.0, mt)
        rc_r, b = ref.assign_dense(mask, v, 23.0, mt)
        assert rc_p == rc_r == 0
        same(a, b)
    assert port.assign_dense(mask, v, 1.0, 0)[0] != 0 and ref.assign_dense(mask, v, 1.0, 0)[0] != 0
    idx = rng.choice(1000, 300, replace=False).astype(np.uint32)
    same(port.assign_sparse(idx, v, 5.0), ref.assign_sparse(idx, v, 5.0))
    val = rng.random(300).astype(np.float32)
    pa, pi, pv = port.assign_sparse_relax(idx, val, v)
    ra, ri, rv = ref.assign_sparse_relax(idx, val, v)
    same(pa, ra), same(pi, ri), same(pv, rv)
    # duplicates in the list: sequential semantics must match too
    idx2 = np.array([3, 3, 5, 3], np.uint32)
    val2 = np.array([0.5, 0.4, 0.1, 0.45], np.float32)
    for a, b in zip(port.assign_sparse_relax(idx2, val2, np.ones(8, np.float32)),
                    ref.assign_sparse_relax(idx2, val2, np.ones(8, np.float32))):
        same(a, b)


def test_io_helpers(both):
    port, ref = both
    rng = np.random.default_rng(13)
    for shape in [(1, 1), (50, 70), (200, 200)]:
        m = random_csr(rng, shape[0], shape[1], 0.05)
        for a, b in zip(port.csr2csc(m), ref.csr2csc(m)):
            same(a, b)
        same(port.normalize_outdegree(m), ref.normalize_outdegree(m))
        pr, rr = port.round_dim(m.num_rows, m.num_cols, m.indptr, 128, 128), ref.round_dim(m.num_rows, m.num_cols, m.indptr, 128, 128)
        assert pr[:2] == rr[:2]
        same(pr[2], rr[2])


def test_sssp_preprocess_including_missing_diagonals(both):
    # graphs WITHOUT a full diagonal exercise the reference's stale-row-end behaviour (sssp.h:31-32,58)
    port, ref = both
    rng = np.random.default_rng(17)
    for n, dens, empty in [(1, 0.5, 0.0), (12, 0.3, 0.3), (40, 0.2, 0.1), (150, 0.05, 0.2), (64, 0.5, 0.0)]:
        m = random_csr(rng, n, n, dens, empty_frac=empty)
        for a, b in zip(port.sssp_preprocess(m), ref.sssp_preprocess(m)):
            same(a, b)
    g = datasets.powerlaw_graph(256, 3000, seed=2, diagonal=True)
    for a, b in zip(port.sssp_preprocess(g), ref.sssp_preprocess(g)):
        same(a, b)


def test_apps(both):
    port, ref = both
    g = datasets.powerlaw_graph(1024, 12000, seed=4, diagonal=False)
    for iters in (1, 5):
        same(port.bfs(g, 0, iters), ref.bfs(g, 0, iters))
    gp = CSRMatrix(g.num_rows, g.num_cols, port.normalize_outdegree(g) * np.float32(0.9), g.indices, g.indptr)
    same(port.pagerank(gp, 0.9, 10), ref.pagerank(gp, 0.9, 10))
    gd = datasets.powerlaw_graph(1024, 12000, seed=4, diagonal=True)
    ip, ix, d = port.sssp_preprocess(gd)
    gs = CSRMatrix(gd.num_rows, gd.num_cols, d, ix, ip)
    same(port.sssp(gs, 0, 6), ref.sssp(gs, 0, 6))


def test_apps_full_reference_pipeline_from_npz(both, tmp_path):
    """The reference apps' own load_and_format_matrix(path) + compute_reference_results against the
    restated pre-processing + restated loops (uniform_10K_10-shaped input of tests/test_app.cpp:51-135)."""
    port, ref = both
    from graphlily_b200 import io
    m = datasets.uniform_csr(1000, 1000, 10, seed=8, value=1.0)
    path = str(tmp_path / "u.npz")
    io.save_csr_matrix_to_npz(path, m)
    # BFS: round to 128, all values 1
    nr, nc, ip = port.round_dim(m.num_rows, m.num_cols, m.indptr, 128, 128)
    mb = CSRMatrix(nr, nc, np.ones(m.nnz, np.float32), m.indices, ip)
    same(port.bfs(mb, 0, 10), ref.app_bfs_npz(path, 1000, 0, 10))
    # PageRank: round, normalise, * damping
    mp = CSRMatrix(nr, nc, m.data, m.indices, ip)
    mp.data = port.normalize_outdegree(mp) * np.float32(0.9)
    same(port.pagerank(mp, 0.9, 10), ref.app_pagerank_npz(path, 1000, 0.9, 10))
    # SSSP: preprocess THEN round
    sip, six, sd = port.sssp_preprocess(m)
    nr, nc, ip2 = port.round_dim(m.num_rows, m.num_cols, sip, 128, 128)
    same(port.sssp(CSRMatrix(nr, nc, sd, six, ip2), 0, 10), ref.app_sssp_npz(path, 1000, 0, 10))
