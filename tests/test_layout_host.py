"""The lane-segment formatter (glb_csr_format_host, the counterpart of the reference's csr2cpsr,
data_formatter.h:457-534) checked on the CPU: a sequential model of the kernel schedule
(tests/layout_model.py) run over the produced arrays must reproduce the oracle's SpMV."""
import numpy as np
import pytest

from graphlily_b200 import capi, datasets, io
from layout_model import FLAG, run_model
from util import SEMIRINGS, random_csr


def check(oracle, m, rb=0, re=None, seed=0, tile_k=0):
    rng = np.random.default_rng(seed)
    re = m.num_rows if re is None else re
    L = capi.format_host(m, rb, re, tile_k)
    nnz = int(m.indptr[re]) - int(m.indptr[rb])
    assert L["nnz"] == nnz
    # invariants: chunks tile the group stream, hold 1..8 groups and at most row_cap flags;
    # one flag per non-empty row that does not open a chunk, plus one per padded chunk
    n_groups = np.diff(L["chunk_goff"].astype(np.int64))
    assert L["n_chunks"] == 0 or (n_groups.min() >= 1 and n_groups.max() <= L["max_groups"])
    assert int(L["chunk_goff"][-1]) == L["n_groups"] and L["n_groups"] * L["group"] - nnz < L["group"] * max(L["n_chunks"], 1)
    n_flags = sum(bin(int(w)).count("1") for w in L["flags"].ravel())
    n_fresh = int((L["chunk_first"] & FLAG != 0).sum())
    assert len(L["nz_rows"]) + len(L["empty_rows"]) == re - rb
    assert len(L["nz_rows"]) <= n_flags + n_fresh <= len(L["nz_rows"]) + L["n_chunks"]
    if L["n_chunks"]:
        assert max(sum(bin(int(w)).count("1") for w in row) for row in L["flags"]) <= L["row_cap"]
    if len(L["hot_cols"]):
        assert L["tile_k"] == len(L["hot_cols"]) == len(set(L["hot_cols"].tolist()))
    for op, zero in SEMIRINGS:
        x = rng.integers(0, 2, m.num_cols).astype(np.float32) if op != 2 else rng.integers(0, 5, m.num_cols).astype(np.float32)
        y, written = run_model(L, x, op, zero, m.num_rows)
        ref = oracle.port.spmv(m, op, zero, 0, x)
        assert (written[rb:re] == 1).all() and written[:rb].sum() == 0 and written[re:].sum() == 0
        if op == 0:
            assert np.allclose(y[rb:re], ref[rb:re], rtol=1e-5, atol=1e-7)
        else:
            assert np.array_equal(y[rb:re], ref[rb:re])
    return L


def test_powerlaw_and_shards(oracle):
    rng = np.random.default_rng(1)
    m = datasets.powerlaw_csr(3000, 3000, 60000, seed=5, max_degree=5000)
    m.data = rng.random(m.nnz).astype(np.float32)
    L = check(oracle, m)
    assert L["n_chunks"] >= 59 and len(L["fixups"]) <= 2 * L["n_chunks"]
    L = check(oracle, m, tile_k=64)          # hot columns renumbered
    assert len(L["hot_cols"]) == 64
    counts = np.bincount(m.indices, minlength=m.num_cols)
    assert counts[L["hot_cols"]].min() >= np.sort(counts)[-64]
    check(oracle, m, tile_k=1 << 20)         # every column hot: identity numbering
    check(oracle, m, 1000, 2000, tile_k=100)
    check(oracle, m, 1000, 2000)
    check(oracle, m, 0, 1)
    check(oracle, m, 2999, 3000)
    check(oracle, m, 1500, 1500)   # empty shard


@pytest.mark.parametrize("groups", ["1", "2", "4", None])
def test_chunk_sizes(oracle, monkeypatch, groups):
    # small shards get smaller chunks (more warps per launch); every size the formatter can pick, and its
    # own choice, through the same sequential model
    if groups is None:
        monkeypatch.delenv("GLB_SPMV_MAX_GROUPS", raising=False)
    else:
        monkeypatch.setenv("GLB_SPMV_MAX_GROUPS", groups)
    rng = np.random.default_rng(2)
    m = datasets.powerlaw_csr(3000, 3000, 60000, seed=7, max_degree=5000)
    m.data = rng.random(m.nnz).astype(np.float32)
    L = check(oracle, m, tile_k=64)
    assert L["max_groups"] == (int(groups) if groups else 1)      # 60 000 non-zeros: the smallest chunks
    assert L["n_chunks"] >= m.nnz // (128 * L["max_groups"])
    check(oracle, m, 700, 2100, tile_k=0)


def test_boundary_cases(oracle):
    # rows that are exactly one chunk, span several, end on a boundary, or are empty
    rng = np.random.default_rng(2)
    ip = [0]
    for d in [1024, 0, 0, 2048, 1, 1023, 5000, 0, 1024, 3, 0]:
        ip.append(ip[-1] + d)
    nnz = ip[-1]
    m = io.CSRMatrix(len(ip) - 1, 7000, rng.random(nnz).astype(np.float32),
                     rng.integers(0, 7000, nnz).astype(np.uint32), np.array(ip, np.uint32))
    L = check(oracle, m)
    assert L["empty_rows"].tolist() == [1, 2, 7, 10]
    for a, b in [(0, 3), (3, 4), (1, 2), (4, 11), (6, 7)]:
        check(oracle, m, a, b)
    one_long = io.CSRMatrix(1, 50, np.ones(40000, np.float32), rng.integers(0, 50, 40000).astype(np.uint32),
                            np.array([0, 40000], np.uint32))
    L = check(oracle, one_long)
    assert len(L["fixups"]) == 1 and L["fixups"][0].tolist()[:2] == [0, 0] and (L["fixups"][0][2] & 0x7fffffff) == 39


def test_row_cap_cuts_chunks(oracle):
    # many very short rows: more row ends per 1024 non-zeros than a chunk may hold
    rng = np.random.default_rng(7)
    deg = rng.integers(0, 4, 6000)
    deg[100:110] = 700
    ip = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    nnz = int(ip[-1])
    m = io.CSRMatrix(len(deg), 900, rng.random(nnz).astype(np.float32), rng.integers(0, 900, nnz).astype(np.uint32), ip)
    L = check(oracle, m, tile_k=50)
    assert (np.diff(L["chunk_goff"].astype(np.int64)) < L["max_groups"]).any()   # some chunk was cut early
    check(oracle, m, 17, 5000, tile_k=0)
    ones = io.CSRMatrix(5000, 5000, np.ones(5000, np.float32), np.arange(5000, dtype=np.uint32),
                        np.arange(5001, dtype=np.uint32))
    check(oracle, ones, tile_k=128)


def test_empty_and_tiny(oracle):
    check(oracle, io.CSRMatrix(5, 5, np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(6, np.uint32)))
    check(oracle, datasets.eye(10))
    check(oracle, datasets.line_graph(8))
    rng = np.random.default_rng(3)
    for n in (1, 2, 31, 33, 1025):
        check(oracle, random_csr(rng, n, n + 3, 0.3), seed=n)


def test_invalid_inputs_rejected():
    m = datasets.eye(4)
    bad = io.CSRMatrix(4, 4, m.data, np.array([0, 1, 2, 9], np.uint32), m.indptr)
    with pytest.raises(capi.GlbError):
        capi.format_host(bad)
    bad2 = io.CSRMatrix(4, 4, m.data, m.indices, np.array([0, 2, 1, 3, 4], np.uint32))
    with pytest.raises(capi.GlbError):
        capi.format_host(bad2)


def test_formatter_fuzz_row_lengths_shards_and_hot_sets(oracle):
    """Randomised: row lengths drawn around the lane / group / chunk sizes (31..33, 127..129,
    1023..1025, 3000), random shards and hot-set sizes from 0 to 'every column'.  (The same loop ran
    clean under AddressSanitizer with the library built -fsanitize=address.)"""
    rng = np.random.default_rng(123)
    for trial in range(25):
        rows, cols = int(rng.integers(1, 1500)), int(rng.integers(1, 4000))
        degs = np.minimum(rng.choice([0, 0, 1, 2, 5, 31, 32, 33, 127, 128, 129, 1023, 1024, 1025, 3000], rows), cols)
        ip = np.concatenate([[0], np.cumsum(degs)]).astype(np.uint32)
        ix = np.concatenate([np.sort(rng.choice(cols, d, replace=False)) for d in degs] + [np.zeros(0, np.int64)])
        m = io.CSRMatrix(rows, cols, rng.integers(1, 3, len(ix)).astype(np.float32), ix.astype(np.uint32), ip)
        rb = int(rng.integers(0, rows))
        re = int(rng.integers(rb, rows + 1))
        check(oracle, m, rb, re, seed=trial, tile_k=int(rng.choice([0, 1, 7, 64, 4096, 10**6])))
