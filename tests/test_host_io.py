"""graphlily_b200.io (Python host mirror of graphlily::io) against the golden vectors of
/root/reference/tests/test_io.cpp and against the oracle on seeded inputs."""
import numpy as np

from graphlily_b200 import datasets, io
from graphlily_b200.io import CSRMatrix
from util import csr_matrix_1, random_csr


def test_create_csr_matrix():
    # test_io.cpp:68-80
    m = io.create_csr_matrix(5, 5, [1, 2, 3, 4, 5, 6, 7, 8, 9], [0, 1, 2, 3, 0, 2, 1, 3, 2], [0, 4, 6, 7, 8, 9])
    assert (m.num_rows, m.num_cols, m.nnz) == (5, 5, 9)
    assert m.data.dtype == np.float32 and m.indices.dtype == np.uint32 and m.indptr.dtype == np.uint32


def test_csr2csc_golden():
    c = io.csr2csc(csr_matrix_1())   # test_io.cpp:110-118
    assert c.data.tolist() == [1, 5, 2, 7, 3, 6, 4, 8]
    assert c.indices.tolist() == [0, 1, 0, 2, 0, 1, 0, 3]
    assert c.indptr.tolist() == [0, 2, 4, 6, 8]


def test_round_dim_golden():
    m = io.util_round_csr_matrix_dim(csr_matrix_1(), 3, 5)   # test_io.cpp:121-130
    assert (m.num_rows, m.num_cols) == (6, 5) and m.indptr.tolist() == [0, 4, 6, 7, 8, 8, 8]


def test_normalize_golden():
    m = io.util_normalize_csr_matrix_by_outdegree(csr_matrix_1())   # test_io.cpp:133-140
    assert m.data.tolist() == [0.5] * 8


def test_npz_roundtrip_and_reference_fixture_shape(tmp_path):
    m = datasets.eye(10)
    p = str(tmp_path / "eye.npz")
    io.save_csr_matrix_to_npz(p, m)
    r = io.load_csr_matrix_from_float_npz(p)     # test_io.cpp:83-93
    assert (r.num_rows, r.num_cols) == (10, 10)
    assert r.data.tolist() == [1] * 10 and r.indices.tolist() == list(range(10)) and r.indptr.tolist() == list(range(11))
    import scipy.sparse
    s = scipy.sparse.load_npz(p)                 # the writer is scipy-compatible
    assert s.shape == (10, 10) and s.nnz == 10


def test_against_oracle(oracle):
    rng = np.random.default_rng(5)
    for shape in [(1, 1), (33, 70), (257, 129)]:
        m = random_csr(rng, *shape, 0.07)
        ip, ix, d = oracle.port.csr2csc(m)
        c = io.csr2csc(m)
        assert c.indptr.tolist() == ip.tolist() and c.indices.tolist() == ix.tolist() and c.data.tobytes() == d.tobytes()
        assert io.util_normalize_csr_matrix_by_outdegree(
            CSRMatrix(m.num_rows, m.num_cols, m.data.copy(), m.indices, m.indptr)).data.tobytes() == \
            oracle.port.normalize_outdegree(m).tobytes()
        nr, nc, rip = oracle.port.round_dim(m.num_rows, m.num_cols, m.indptr, 128, 128)
        r = io.util_round_csr_matrix_dim(CSRMatrix(m.num_rows, m.num_cols, m.data, m.indices, m.indptr.copy()), 128, 128)
        assert (r.num_rows, r.num_cols) == (nr, nc) and r.indptr.tolist() == rip.tolist()


def test_sssp_preprocess_against_oracle(oracle):
    rng = np.random.default_rng(6)
    cases = [random_csr(rng, n, n, dens, empty_frac=e) for n, dens, e in [(1, 0.5, 0), (30, 0.2, 0.2), (90, 0.1, 0.1)]]
    cases.append(datasets.powerlaw_graph(512, 6000, seed=1, diagonal=True))   # vectorised full-diagonal path
    for m in cases:
        oip, oix, od = oracle.port.sssp_preprocess(m)
        r = io.sssp_preprocess(CSRMatrix(m.num_rows, m.num_cols, m.data.copy(), m.indices.copy(), m.indptr.copy()))
        assert r.indptr.tolist() == oip.tolist() and r.indices.tolist() == oix.tolist() and r.data.tobytes() == od.tobytes()


def test_generators_are_sorted_distinct_and_seeded():
    a = datasets.powerlaw_csr(4096, 4096, 1 << 16, seed=9)
    b = datasets.powerlaw_csr(4096, 4096, 1 << 16, seed=9)
    assert a.nnz == 1 << 16 and a.indices.tobytes() == b.indices.tobytes()
    rows = np.repeat(np.arange(a.num_rows, dtype=np.int64), np.diff(a.indptr.astype(np.int64)))
    key = rows * a.num_cols + a.indices
    assert (np.diff(key) > 0).all()
    deg = np.diff(a.indptr.astype(np.int64))
    assert deg.max() > 20 * deg.mean()            # heavy tail
    g = datasets.powerlaw_graph(1024, 20000, seed=2, diagonal=True)
    dense = np.zeros((1024, 1024), bool)
    grow = np.repeat(np.arange(1024), np.diff(g.indptr.astype(np.int64)))
    dense[grow, g.indices] = True
    assert (dense == dense.T).all() and dense.diagonal().all()
    u = datasets.uniform_csr(1000, 1000, 10, seed=0)
    assert (np.diff(u.indptr.astype(np.int64)) == 10).all() and u.data[0] == np.float32(1e-3)


def test_row_cuts_of_the_sharded_apps():
    """ModuleCollection._cuts: equal slots under the NCCL allgather, nnz-balanced 32-row-aligned cuts over an
    exchange (slices need not be equal there) -- computed identically on every rank, tiling all rows."""
    from graphlily_b200 import datasets
    from graphlily_b200.app import ModuleCollection
    m = datasets.powerlaw_csr(4096, 4096, 150000, seed=3, max_degree=3000)
    ip = m.indptr.astype(np.int64)
    for world in (1, 2, 4, 8):
        for exchange in (None, object()):
            ranges = []
            for rank in range(world):
                mc = ModuleCollection()
                mc.csr_matrix_ = m
                mc.set_sharding(rank, world, exchange)
                ranges.append(mc._row_range(m.num_rows))
            assert ranges[0][0] == 0 and ranges[-1][1] == m.num_rows
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))      # tiles, in rank order
            if exchange is None or world == 1:
                assert len({re - rb for rb, re in ranges}) == 1                          # equal slots
            else:
                assert all(rb % 32 == 0 for rb, _ in ranges)
                nnz = [int(ip[re] - ip[rb]) for rb, re in ranges]
                assert max(nnz) - min(nnz) <= 0.1 * m.nnz / world + 2 * 3000, nnz        # balanced up to one giant row
