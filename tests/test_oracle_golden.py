"""The oracle (C restatement, and the compiled reference when present) against the golden
vectors the reference's own tests hold for this path: /root/reference/tests/test_io.cpp
(:83-140) and the closed-form answers of its two shipped fixtures (eye_10, line_8)."""
import numpy as np
import pytest

from graphlily_b200 import datasets
from util import csr_matrix_1

BACKENDS = ["port", "ref"]


def backend(oracle, name):
    b = getattr(oracle, name)
    if b is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return b


@pytest.mark.parametrize("name", BACKENDS)
def test_csr2csc_golden(oracle, name):
    # test_io.cpp:110-118
    ip, ix, d = backend(oracle, name).csr2csc(csr_matrix_1())
    assert d.tolist() == [1, 5, 2, 7, 3, 6, 4, 8]
    assert ix.tolist() == [0, 1, 0, 2, 0, 1, 0, 3]
    assert ip.tolist() == [0, 2, 4, 6, 8]


@pytest.mark.parametrize("name", BACKENDS)
def test_round_dim_golden(oracle, name):
    # test_io.cpp:121-130: 4x4 -> 6x5 with divisors 3, 5
    m = csr_matrix_1()
    nr, nc, ip = backend(oracle, name).round_dim(m.num_rows, m.num_cols, m.indptr, 3, 5)
    assert (nr, nc) == (6, 5)
    assert ip.tolist() == [0, 4, 6, 7, 8, 8, 8]


@pytest.mark.parametrize("name", BACKENDS)
def test_normalize_outdegree_golden(oracle, name):
    # test_io.cpp:133-140: the first four entries become 0.5
    d = backend(oracle, name).normalize_outdegree(csr_matrix_1())
    assert d[:4].tolist() == [0.5, 0.5, 0.5, 0.5]
    assert d.tolist() == [0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]


@pytest.mark.parametrize("name", BACKENDS)
def test_spmv_row_sums(oracle, name):
    # SURVEY 8c probe: 4x4 golden matrix times ones = row sums 10 11 7 8
    y = backend(oracle, name).spmv(csr_matrix_1(), 0, 0.0, 0, np.ones(4, np.float32))
    assert y.tolist() == [10, 11, 7, 8]


@pytest.mark.parametrize("name", BACKENDS)
def test_line8_closed_form(oracle, name):
    # tests/test_data/line_8_csr_float32.npz: A[i, i-1] = 1.  BFS from 0: distance[k] = k + 1
    # (bfs.h:108-110,123); SSSP with the zero-weight diagonal: dist[k] = k.
    b = backend(oracle, name)
    m = datasets.line_graph(8)
    dist = b.bfs(m, 0, 7)
    assert dist.tolist() == [1, 2, 3, 4, 5, 6, 7, 8]
    dist3 = b.bfs(m, 0, 3)
    assert dist3.tolist() == [1, 2, 3, 4, 0, 0, 0, 0]
    ip, ix, d = b.sssp_preprocess(m)
    from graphlily_b200.io import CSRMatrix
    ms = CSRMatrix(8, 8, d, ix, ip)
    assert b.sssp(ms, 0, 7).tolist() == [0, 1, 2, 3, 4, 5, 6, 7]
    assert b.sssp(ms, 0, 2).tolist() == [0, 1, 2, 255, 255, 255, 255, 255]


@pytest.mark.parametrize("name", BACKENDS)
def test_eye10_pagerank(oracle, name):
    # identity graph: every column has one entry, rank stays d*r + (1-d)/N = 1/N
    b = backend(oracle, name)
    m = datasets.eye(10)
    m.data = b.normalize_outdegree(m) * np.float32(0.9)
    r = b.pagerank(m, 0.9, 10)
    assert np.allclose(r, 0.1, rtol=1e-6)


def test_reference_npz_fixtures(oracle):
    # test_io.cpp:83-93 through the reference's own loader over our npz reader
    import os
    if oracle.ref is None or not os.path.exists("/root/reference/tests/test_data/eye_10_csr_float32.npz"):
        pytest.skip("reference tree not present")
    nr, nc, ip, ix, d = oracle.ref.load_npz("/root/reference/tests/test_data/eye_10_csr_float32.npz")
    assert (nr, nc) == (10, 10)
    assert d.tolist() == [1] * 10 and ix.tolist() == list(range(10)) and ip.tolist() == list(range(11))
    nr, nc, ip, ix, d = oracle.ref.load_npz("/root/reference/tests/test_data/line_8_csr_float32.npz")
    g = datasets.line_graph(8)
    assert (nr, nc) == (8, 8) and ip.tolist() == g.indptr.tolist() and ix.tolist() == g.indices.tolist()


def test_reference_constants(oracle):
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built")
    c = oracle.ref.constants()
    # zeros 0 / 0 / 255 (UFIXED_INF), ones 1 / 1 / 0, FLOAT_INF, 16 channels * pack 8 = 128
    assert c.tolist() == [0, 0, 255, 1, 1, 0, np.float32(999999999), 128]
