"""Parity of the CUDA path against the oracle, through the C ABI (ctypes -> libgraphlily_b200.so).

Bars (north_star): or-and and min-plus results bit-exact; fp32 plus-times within 1e-5 relative
(the kernel reduces a row in a different order than the sequential reference loop).
Shapes follow tests/test_module_spmv_spmspv.cpp / test_module_apply.cpp of the reference plus the
edge cases of a warp-segment layout: empty rows, rows longer than a chunk, rows ending on a chunk
boundary, empty matrices, row shards."""
import numpy as np
import pytest

from golden_util import golden, golden_csr
from graphlily_b200 import capi, datasets, io
from graphlily_b200.capi import Epilogue
from graphlily_b200.io import CSRMatrix
from util import MASKS, SEMIRINGS, assert_close_rel, densify, random_csr

pytestmark = pytest.mark.gpu


def check_vec(got, ref, op):
    if op == capi.OP_MUL_ADD:
        assert_close_rel(got, ref, 1e-5)
    else:
        assert got.tobytes() == ref.tobytes(), f"mismatch at {np.nonzero(got != ref)[0][:8]}"


def gpu_spmv(ctx, m, op, zero, mt, x, mask, rb=0, re=None, epilogue=None, y_init=np.nan):
    A = capi.CsrMatrix(ctx, m, rb, re)
    dx, dy = ctx.to_device(np.asarray(x, np.float32)), ctx.to_device(np.full(m.num_rows, y_init, np.float32))
    dm = ctx.to_device(np.asarray(mask, np.float32)) if mask is not None else None
    A.spmv(op, zero, mt, dx, dm, dy, epilogue)
    y = dy.read(np.float32, m.num_rows)
    A.close()
    return y


# ------------------------------------------------------------------------------ SpMV
@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mt", MASKS)
def test_spmv_golden_fixture(ctx, op, zero, mt):
    z, m = golden(), golden_csr("spmv")
    y = gpu_spmv(ctx, m, op, zero, mt, z["spmv_x"], z["spmv_mask"])
    check_vec(y, z[f"spmv_y_op{op}_m{mt}"], op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmv_golden_powerlaw(ctx, op, zero):
    z, m = golden(), golden_csr("pl")
    check_vec(gpu_spmv(ctx, m, op, zero, 0, z["pl_x"], None), z[f"pl_y_op{op}"], op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mt", MASKS)
def test_spmv_c1_uniform_10k(ctx, oracle, op, zero, mt):
    # BASELINE configs[0]: 10k x 10k, 10 nnz/row, A = 1/N, x = rand % 2, mask = rand % 2
    rng = np.random.default_rng(1)
    m = datasets.uniform_csr(10000, 10000, 10, seed=0)
    x, mask = rng.integers(0, 2, 10000).astype(np.float32), rng.integers(0, 2, 10000).astype(np.float32)
    check_vec(gpu_spmv(ctx, m, op, zero, mt, x, mask), oracle.port.spmv(m, op, zero, mt, x, mask), op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmv_random_shapes_with_empty_rows(ctx, oracle, op, zero):
    rng = np.random.default_rng(10 + op)
    for shape in [(1, 1), (5, 9), (33, 70), (257, 129), (2000, 1500)]:
        m = random_csr(rng, *shape, 0.1, values="rand" if op == 0 else "small", empty_frac=0.3)
        x = (rng.integers(0, 3, shape[1]) * rng.random(shape[1])).astype(np.float32)
        mask = rng.integers(0, 2, shape[0]).astype(np.float32)
        for mt in MASKS:
            check_vec(gpu_spmv(ctx, m, op, zero, mt, x, mask), oracle.port.spmv(m, op, zero, mt, x, mask), op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmv_chunk_boundaries_and_long_rows(ctx, oracle, op, zero):
    rng = np.random.default_rng(20 + op)
    ip = [0]
    for d in [1024, 0, 0, 2048, 1, 1023, 5000, 0, 1024, 3, 0, 70000, 7, 0]:
        ip.append(ip[-1] + d)
    nnz = ip[-1]
    m = CSRMatrix(len(ip) - 1, 9000, (rng.random(nnz) if op == 0 else rng.integers(0, 3, nnz)).astype(np.float32),
                  rng.integers(0, 9000, nnz).astype(np.uint32), np.array(ip, np.uint32))
    x = rng.integers(0, 4, 9000).astype(np.float32)
    check_vec(gpu_spmv(ctx, m, op, zero, 0, x, None), oracle.port.spmv(m, op, zero, 0, x), op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmv_powerlaw_1m_nnz(ctx, oracle, op, zero):
    rng = np.random.default_rng(30 + op)
    m = datasets.powerlaw_csr(1 << 15, 1 << 15, 1 << 20, seed=11, max_degree=1 << 14)
    if op == 2:
        m.data = rng.integers(0, 3, m.nnz).astype(np.float32)
    x = rng.integers(0, 2, m.num_cols).astype(np.float32) if op != 2 else rng.integers(0, 200, m.num_cols).astype(np.float32)
    check_vec(gpu_spmv(ctx, m, op, zero, 0, x, None), oracle.port.spmv(m, op, zero, 0, x), op)


@pytest.mark.parametrize("stored_zeros", [False, True])
def test_spmv_or_and_bitmap_hot_and_cold_columns(ctx, oracle, stored_zeros):
    # or-and packs x to one bit per stored column word (pack_bits_kernel): more columns than tile_k so
    # both the hot-rank and the cold-column bit ranges are read; with and without stored 0.0 values
    # (pattern-only kernel vs value-testing kernel); x holds negative, tiny, -0.0 and NaN entries.
    rng = np.random.default_rng(77)
    n = 100_003
    m = datasets.powerlaw_csr(6000, n, 400_000, seed=13, max_degree=50_000, value=3.0)
    if stored_zeros:
        m.data = rng.integers(0, 2, m.nnz).astype(np.float32) * np.float32(-2.5)
    x = np.zeros(n, np.float32)
    pick = rng.random(n)
    x[pick < 0.02] = 1.0
    x[(pick >= 0.02) & (pick < 0.03)] = -1e-42   # denormal: still true
    x[(pick >= 0.03) & (pick < 0.04)] = -0.0
    x[(pick >= 0.04) & (pick < 0.045)] = np.nan
    mask = rng.integers(0, 2, m.num_rows).astype(np.float32)
    for mt in MASKS:
        for zero in (0.0, 1.0):
            check_vec(gpu_spmv(ctx, m, 1, zero, mt, x, mask), oracle.port.spmv(m, 1, zero, mt, x, mask), 1)
    ref = oracle.port.spmv(m, 1, 0.0, 0, x)
    y = gpu_spmv(ctx, m, 1, 0.0, 0, x, None, 1000, 4097, y_init=-7.0)
    assert y[1000:4097].tobytes() == ref[1000:4097].tobytes()
    assert (y[:1000] == -7.0).all() and (y[4097:] == -7.0).all()


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("density", [0.0, 0.003, 0.5, 1.0])
def test_spmv_mask_driven_skipping(ctx, oracle, op, zero, density):
    # lanes / chunks whose rows are all masked out skip their gathers: masks from "everything
    # discarded" to "nothing discarded", over rows inside a lane, across lanes and across chunks
    rng = np.random.default_rng(40 + op)
    ip = [0]
    for d in rng.choice([0, 1, 2, 3, 8, 31, 32, 33, 100, 1024, 1500, 5000, 40000], 400):
        ip.append(ip[-1] + int(d))
    nnz, ncols = ip[-1], 70_000
    m = CSRMatrix(len(ip) - 1, ncols, (rng.random(nnz) + 0.5 if op == 0 else rng.integers(1, 3, nnz)).astype(np.float32),
                  rng.integers(0, ncols, nnz).astype(np.uint32), np.array(ip, np.uint32))
    x = (rng.integers(0, 2, ncols) * (1 + rng.random(ncols))).astype(np.float32)
    keep = (rng.random(m.num_rows) < density).astype(np.float32)
    for mt, mask in ((1, 1.0 - keep), (2, keep)):     # WriteToZero discards mask != 0, WriteToOne mask == 0
        check_vec(gpu_spmv(ctx, m, op, zero, mt, x, mask), oracle.port.spmv(m, op, zero, mt, x, mask), op)


def test_spmv_empty_matrix_and_zero_value(ctx, oracle):
    m = CSRMatrix(6, 4, np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(7, np.uint32))
    x = np.ones(4, np.float32)
    for op, zero in SEMIRINGS + [(0, 2.5), (2, 7.0)]:   # `zero` is a runtime value
        check_vec(gpu_spmv(ctx, m, op, zero, 0, x, None), oracle.port.spmv(m, op, zero, 0, x), op)
    rng = np.random.default_rng(3)
    m = random_csr(rng, 300, 300, 0.05, values="small")
    x = rng.integers(0, 6, 300).astype(np.float32)
    for op, zero in [(0, 2.5), (2, 3.0), (1, 1.0)]:
        check_vec(gpu_spmv(ctx, m, op, zero, 0, x, None), oracle.port.spmv(m, op, zero, 0, x), op)


def test_spmv_row_shards_cover_disjoint_slices(ctx, oracle):
    rng = np.random.default_rng(4)
    m = datasets.powerlaw_csr(4096, 4096, 1 << 17, seed=12, max_degree=3000)
    x = rng.random(4096).astype(np.float32)
    ref = oracle.port.spmv(m, 0, 0.0, 0, x)
    bounds = [0, 1024, 1031, 3000, 4096]
    for rb, re in zip(bounds[:-1], bounds[1:]):
        y = gpu_spmv(ctx, m, 0, 0.0, 0, x, None, rb, re, y_init=-7.0)
        assert_close_rel(y[rb:re], ref[rb:re])
        assert (y[:rb] == -7.0).all() and (y[re:] == -7.0).all()   # rows outside the shard untouched


def test_spmv_fused_epilogues(ctx, oracle):
    rng = np.random.default_rng(5)
    g = datasets.powerlaw_graph(2048, 30000, seed=6)
    # PageRank step: y = A x + c
    gp = CSRMatrix(g.num_rows, g.num_cols, oracle.port.normalize_outdegree(g) * np.float32(0.9), g.indices, g.indptr)
    x = rng.random(2048).astype(np.float32)
    c = float(np.float32(0.1) / np.float32(2048))
    y = gpu_spmv(ctx, gp, 0, 0.0, 0, x, None, epilogue=Epilogue(1, c, None, 0.0, 0))
    assert_close_rel(y, oracle.port.ewise_add(oracle.port.spmv(gp, 0, 0.0, 0, x), c))
    # BFS step: y = mask-to-zero(A or.and x); distance[y != 0] = level
    dist = np.zeros(2048, np.float32)
    frontier = np.zeros(2048, np.float32)
    seen = rng.choice(2048, 200, replace=False)
    dist[seen] = 1
    frontier[seen[:50]] = 1
    A = capi.CsrMatrix(ctx, g)
    dx, dd, dy = ctx.to_device(frontier), ctx.to_device(dist), ctx.zeros_f32(2048)
    A.spmv(1, 0.0, 1, dx, dd, dy, Epilogue(0, 0.0, dd.ptr, 5.0, 2))
    ref_y = oracle.port.spmv(g, 1, 0.0, 1, frontier, dist)
    _, ref_d = oracle.port.assign_dense(ref_y, dist, 5.0, 2)
    assert dy.read(np.float32, 2048).tobytes() == ref_y.tobytes()
    assert dd.read(np.float32, 2048).tobytes() == ref_d.tobytes()


def test_spmv_host_entry_point(ctx, oracle):
    rng = np.random.default_rng(6)
    m = datasets.uniform_csr(5000, 5000, 10, seed=2)
    x, mask = rng.integers(0, 2, 5000).astype(np.float32), rng.integers(0, 2, 5000).astype(np.float32)
    A = capi.CsrMatrix(ctx, m)
    y = np.zeros(5000, np.float32)
    A.spmv_host(0, 0.0, 2, x, mask, y)
    assert_close_rel(y, oracle.port.spmv(m, 0, 0.0, 2, x, mask))


def test_spmv_argument_errors(ctx):
    m = datasets.eye(8)
    A = capi.CsrMatrix(ctx, m)
    d = ctx.zeros_f32(8)
    e = ctx.zeros_f32(8)
    with pytest.raises(capi.GlbError):
        A.spmv(0, 0.0, 1, d, None, e)       # mask required
    with pytest.raises(capi.GlbError):
        A.spmv(0, 0.0, 0, d, None, d)       # y aliases x
    with pytest.raises(capi.GlbError):
        A.spmv(7, 0.0, 0, d, None, e)       # bad semiring
    with pytest.raises(capi.GlbError):
        capi.assign_dense(ctx, d, e, 8, 1.0, capi.MASK_NONE)   # assign_vector_dense_module.h:88-95


# ---------------------------------------------------------------------------- SpMSpV
def gpu_spmspv(ctx, csc, op, zero, mt, idx, val, mask, runs=1):
    A = capi.CscMatrix(ctx, csc)
    dx = ctx.to_device(capi.sparse_to_numpy(idx, val, csc.num_cols + 1))
    dy = ctx.to_device(np.zeros(csc.num_rows + 1, capi.IDX_VAL))
    dm = ctx.to_device(np.asarray(mask, np.float32)) if mask is not None else None
    for _ in range(runs):
        A.spmspv(op, zero, mt, dx, dm, dy)
    head = dy.read(capi.IDX_VAL, 1)
    oi, ov = dy.read_sparse()
    assert len(np.unique(oi)) == len(oi)
    assert head["val"][0] == np.float32(zero)           # head = {nnz, Zero}, kernel_spmspv_impl.h:551-555
    assert capi.sparse_count(ctx, dy) == len(oi)
    assert not (ov == np.float32(zero)).any()           # only entries != zero are listed
    A.close()
    return densify(oi, ov, csc.num_rows, zero)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mt", MASKS)
def test_spmspv_golden_fixture(ctx, op, zero, mt):
    z, m = golden(), golden_csr("spmspv_csc")
    y = gpu_spmspv(ctx, m, op, zero, mt, z["spmspv_x_idx"], z["spmspv_x_val"], z[f"spmspv_mask_op{op}"])
    check_vec(y, z[f"spmspv_y_op{op}_m{mt}"], op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmspv_sparsity_sweep_heavy_columns_and_reuse(ctx, oracle, op, zero):
    # vector sparsity 0 / 0.5 / 0.99 as tests/test_module_spmv_spmspv.cpp:286-313; a power-law matrix
    # has columns above the heavy-column threshold; runs=2 checks the accumulator is left clean
    rng = np.random.default_rng(40 + op)
    g = datasets.powerlaw_csr(1 << 14, 1 << 14, 1 << 19, seed=13, max_degree=1 << 13)
    ip, ix, d = oracle.port.csr2csc(g)
    csc = CSRMatrix(g.num_rows, g.num_cols, np.ones_like(d) if op else rng.random(len(d)).astype(np.float32), ix, ip)
    assert np.diff(ip.astype(np.int64)).max() > 2048
    n = g.num_cols
    mask = np.where(rng.random(n) < 0.5, np.float32(zero), np.float32(1)).astype(np.float32)
    for sparsity in (0.0, 0.5, 0.99):
        k = max(1, int(n * (1 - sparsity)))
        idx = (np.arange(k) * (n // k)).astype(np.uint32)
        val = ((rng.integers(0, 10, k)) / 10).astype(np.float32)
        for mt in MASKS:
            y = gpu_spmspv(ctx, csc, op, zero, mt, idx, val, mask, runs=2)
            check_vec(y, oracle.port.spmspv(csc, op, zero, mt, idx, val, mask), op)


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmspv_segment_boundaries_and_repeated_columns(ctx, oracle, op, zero):
    # columns of exactly 511 / 512 / 513 / 1024 / 1025 / 5000 non-zeros (the scatter works in segments of
    # 512), and a frontier that lists the long columns many times over (more queued segments than the
    # queue holds: those warps walk their column themselves)
    rng = np.random.default_rng(60 + op)
    lens = [0, 1, 511, 512, 513, 1024, 1025, 5000, 31, 33, 2048, 6000]
    n_rows = 7000
    ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    ix = np.concatenate([np.sort(rng.choice(n_rows, ln, replace=False)) for ln in lens]).astype(np.uint32)
    d = (np.ones(len(ix)) if op else rng.random(len(ix))).astype(np.float32)
    csc = CSRMatrix(n_rows, len(lens), d, ix, ip)
    mask = np.where(rng.random(n_rows) < 0.5, np.float32(zero), np.float32(1)).astype(np.float32)
    idx = np.arange(len(lens), dtype=np.uint32)
    val = (1 + rng.integers(0, 5, len(lens))).astype(np.float32)
    for mt in MASKS:
        check_vec(gpu_spmspv(ctx, csc, op, zero, mt, idx, val, mask, runs=2),
                  oracle.port.spmspv(csc, op, zero, mt, idx, val, mask), op)
    if op != 0:   # repeats change a plus-times result, or-and / min-plus are idempotent: same answer as the oracle's loop
        idx = np.tile(np.array([7, 11, 10, 5], np.uint32), 40)
        val = np.tile(np.array([1.0, 2.0, 3.0, 1.0], np.float32), 40)
        check_vec(gpu_spmspv(ctx, csc, op, zero, 0, idx, val, mask, runs=2),
                  oracle.port.spmspv(csc, op, zero, 0, idx, val, mask), op)


def test_spmspv_empty_frontier_and_rectangular(ctx, oracle):
    rng = np.random.default_rng(50)
    m = random_csr(rng, 70, 40, 0.2, values="small")           # as CSC: 40 rows, 70 columns
    csc = CSRMatrix(40, 70, m.data, m.indices, m.indptr)
    y = gpu_spmspv(ctx, csc, 0, 0.0, 0, [], [], None)
    assert (y == 0).all()
    idx = np.array([0, 69, 13], np.uint32)
    val = np.array([1.0, 2.0, 0.5], np.float32)
    check_vec(gpu_spmspv(ctx, csc, 2, 255.0, 0, idx, val, None), oracle.port.spmspv(csc, 2, 255.0, 0, idx, val, None), 2)


def test_chunk_sizes_small_shards(ctx, oracle, monkeypatch):
    # chunks of 1 / 2 / 4 / 8 groups and the automatic choice (small shards get small chunks so that the
    # launch has enough warps): three semirings, masked and not, against the oracle
    rng = np.random.default_rng(70)
    m = datasets.powerlaw_csr(6000, 6000, 150000, seed=23, max_degree=3000)
    x = rng.integers(0, 3, m.num_cols).astype(np.float32)
    mask = rng.integers(0, 2, m.num_rows).astype(np.float32)
    for groups in ("1", "2", "4", "8", None):
        if groups is None:
            monkeypatch.delenv("GLB_SPMV_MAX_GROUPS", raising=False)
        else:
            monkeypatch.setenv("GLB_SPMV_MAX_GROUPS", groups)
        A = capi.CsrMatrix(ctx, m)
        if groups is not None:
            assert A.info()["chunks"] >= m.nnz // (128 * int(groups))
        A.close()
        for op, zero in SEMIRINGS:
            for mt in (0, 1, 2):
                got = gpu_spmv(ctx, m, op, zero, mt, x, mask if mt else None)
                check_vec(got, oracle.port.spmv(m, op, zero, mt, x, mask if mt else None), op)


def test_spmspv_fused_push_levels(ctx, oracle):
    # glb_spmspv_fused: the sparse assign of a push level inside the SpMSpV launch, against the
    # reference's two-module sequence (SpMSpV.run(); SparseAssign.run(), bfs.h:147-151 / sssp.h:178-190)
    # evaluated with the oracle; repeated launches on one matrix (the accumulator, bitmap and counters
    # must come back to rest), frontiers from empty to dense
    rng = np.random.default_rng(71)
    g = datasets.social_graph(3000, 60000, 400, seed=12, diagonal=True)
    csc = io.csr2csc(g)
    n = g.num_rows
    A = capi.CscMatrix(ctx, csc)
    dy = ctx.to_device(np.zeros(n + 1, capi.IDX_VAL))
    dnf = ctx.to_device(np.zeros(n + 1, capi.IDX_VAL))
    for density in (0.0, 0.001, 0.05, 0.6, 0.0005):
        idx = np.nonzero(rng.random(n) < density)[0].astype(np.uint32)
        dx = ctx.to_device(capi.sparse_to_numpy(idx, np.ones(len(idx), np.float32), n + 1))
        # BFS level: or-and, mask = distance (write where 0), distance[row] = 7 on the listed rows
        dist = np.where(rng.random(n) < 0.5, 0.0, rng.integers(1, 5, n)).astype(np.float32)
        ddist = ctx.to_device(dist)
        ep = capi.SpmspvEpilogue(capi.SPMSPV_EP_ASSIGN, ddist.ptr, 7.0, None)
        A.spmspv(1, 0.0, capi.MASK_WRITE_TO_ZERO, dx, ddist, dy, ep)
        y_ref = oracle.port.spmspv(csc, 1, 0.0, capi.MASK_WRITE_TO_ZERO, idx, np.ones(len(idx), np.float32), dist)
        oi, ov = dy.read_sparse()
        assert densify(oi, ov, n, 0.0).tobytes() == y_ref.tobytes()
        assert ddist.read(np.float32, n).tobytes() == oracle.port.assign_sparse(np.nonzero(y_ref)[0], dist, 7.0).tobytes()
        # SSSP level: min-plus, no mask, relax of the distance vector + new frontier
        val = rng.integers(0, 6, len(idx)).astype(np.float32)
        dx = ctx.to_device(capi.sparse_to_numpy(idx, val, n + 1))
        dist = rng.integers(0, 9, n).astype(np.float32)
        ddist = ctx.to_device(dist)
        ep = capi.SpmspvEpilogue(capi.SPMSPV_EP_RELAX, ddist.ptr, 0.0, dnf.ptr)
        A.spmspv(2, 255.0, capi.MASK_NONE, dx, None, dy, ep)
        y_ref = oracle.port.spmspv(csc, 2, 255.0, capi.MASK_NONE, idx, val, None)
        oi, ov = dy.read_sparse()
        assert densify(oi, ov, n, 255.0).tobytes() == y_ref.tobytes()
        listed = np.nonzero(y_ref != 255.0)[0].astype(np.uint32)
        ref_inout, ref_fi, ref_fv = oracle.port.assign_sparse_relax(listed, y_ref[listed], dist)
        assert ddist.read(np.float32, n).tobytes() == ref_inout.tobytes()
        fi, fv = dnf.read_sparse()
        o, ro = np.argsort(fi), np.argsort(ref_fi)
        assert fi[o].tolist() == ref_fi[ro].tolist() and fv[o].tobytes() == ref_fv[ro].tobytes()
        assert dnf.read(capi.IDX_VAL, 1)["val"][0] == 0.0
        # plus-times through the same kernel (bitmap path), twice in a row
        for _ in range(2):
            A.spmspv(0, 0.0, capi.MASK_NONE, dx, None, dy)
            oi, ov = dy.read_sparse()
            check_vec(densify(oi, ov, n, 0.0), oracle.port.spmspv(csc, 0, 0.0, capi.MASK_NONE, idx, val, None), 0)
    A.close()


def test_spmspv_direction_decision_on_the_device(ctx, oracle):
    # the `next` block of glb_spmspv_fused: keep pushing iff !force_stop && float(count) / n < threshold
    # (bfs.h:186-190), read back with glb_spmspv_push_state; when pushing stops the launch builds the dense
    # input of the first pull level (scatter of the result list, or a copy of the distance vector)
    rng = np.random.default_rng(72)
    g = datasets.social_graph(2048, 30000, 300, seed=13)
    csc = io.csr2csc(g)
    n = g.num_rows
    A = capi.CscMatrix(ctx, csc)
    dy = ctx.to_device(np.zeros(n + 1, capi.IDX_VAL))
    idx = np.array([5], np.uint32)
    dx = ctx.to_device(capi.sparse_to_numpy(idx, [1.0], n + 1))
    y_ref = oracle.port.spmspv(csc, 1, 0.0, 0, idx, np.ones(1, np.float32), None)
    count = int((y_ref != 0).sum())
    assert count > 0
    A.reset_levels()
    cases = [(0, (count + 0.5) / n, True), (0, (count - 0.5) / n, False), (1, 1.0, False),
             (0, float(np.float32(count) / np.float32(n)), False)]   # equality is not "less than"
    for k, (force_stop, thr, keep_ref) in enumerate(cases):
        dense = ctx.zeros_f32(n, 0.0)
        src = ctx.to_device(rng.random(n).astype(np.float32))
        for mode in (capi.SPMSPV_DENSE_SCATTER, capi.SPMSPV_DENSE_COPY, capi.SPMSPV_DENSE_NONE):
            capi.check(capi.lib.glb_buffer_fill_f32(ctx.handle, dense.ptr, 0.0, n))
            nxt = capi.SpmspvNext(force_stop, thr, n, 0, mode, dense.ptr if mode else None,
                                  src.ptr if mode == capi.SPMSPV_DENSE_COPY else None, n)
            A.spmspv(1, 0.0, 0, dx, None, dy, None, nxt)
            keep, _levels = A.push_state()
            assert keep == keep_ref, (k, mode)
            got = dense.read(np.float32, n)
            if keep or mode == capi.SPMSPV_DENSE_NONE:
                assert not got.any()
            elif mode == capi.SPMSPV_DENSE_SCATTER:
                assert got.tobytes() == y_ref.tobytes()
            else:
                assert got.tobytes() == src.read(np.float32, n).tobytes()
    assert A.push_state()[1] == 3 * len(cases)
    A.close()


# ----------------------------------------------------------------------------- apply
def test_apply_golden_fixture(ctx):
    z = golden()
    n = len(z["apply_in"])
    din, dout = ctx.to_device(z["apply_in"]), ctx.zeros_f32(n)
    capi.ewise_add(ctx, din, dout, n, 0.25)
    assert dout.read(np.float32, n).tobytes() == z["apply_ewise_add"].tobytes()
    for mt in (1, 2):
        dm, dio = ctx.to_device(z["apply_mask"]), ctx.to_device(z["apply_in"])
        capi.assign_dense(ctx, dm, dio, n, 23.0, mt)
        assert dio.read(np.float32, n).tobytes() == z[f"apply_assign_dense_m{mt}"].tobytes()
    dl = ctx.to_device(capi.sparse_to_numpy(z["apply_sparse_idx"], z["apply_sparse_val"]))
    dio = ctx.to_device(z["apply_in"])
    capi.assign_sparse(ctx, dl, dio, 7.0)
    assert dio.read(np.float32, n).tobytes() == z["apply_assign_sparse"].tobytes()
    dio = ctx.to_device(z["apply_in"])
    dnf = ctx.to_device(np.zeros(len(z["apply_sparse_idx"]) + 1, capi.IDX_VAL))
    capi.assign_sparse_relax(ctx, dl, dio, dnf)
    assert dio.read(np.float32, n).tobytes() == z["apply_relax_inout"].tobytes()
    fi, fv = dnf.read_sparse()
    order, ref_order = np.argsort(fi), np.argsort(z["apply_relax_idx"])
    assert fi[order].tolist() == z["apply_relax_idx"][ref_order].tolist()      # same set, order unspecified
    assert fv[order].tobytes() == z["apply_relax_val"][ref_order].tobytes()
    assert dnf.read(capi.IDX_VAL, 1)["val"][0] == 0.0                           # head {count, 0}


def test_apply_shapes_of_reference_tests(ctx, oracle):
    # tests/test_module_apply.cpp: eWiseAdd len 128 (:54-75), dense assign val 23 (:78-103),
    # unaligned / odd lengths, in-place eWiseAdd, device copy + aliasing (:209-261)
    rng = np.random.default_rng(60)
    for n in (1, 3, 128, 1000, 8192, 100003):
        v = rng.random(n).astype(np.float32)
        d = ctx.to_device(v)
        capi.ewise_add(ctx, d, d, n, 1.5)
        assert d.read(np.float32, n).tobytes() == oracle.port.ewise_add(v, 1.5).tobytes()
    v = rng.random(4096).astype(np.float32)
    a, b = ctx.to_device(v), ctx.zeros_f32(4096)
    capi.d2d(ctx, b, a, 4096 * 4)
    assert b.read(np.float32, 4096).tobytes() == v.tobytes()
    lst = ctx.to_device(capi.sparse_to_numpy([3, 100, 4095], [1.0, 2.0, 3.0]))
    dense = ctx.zeros_f32(4096)
    capi.sparse_to_dense(ctx, lst, dense, 4096, 255.0)
    ref = np.full(4096, 255.0, np.float32)
    ref[[3, 100, 4095]] = [1, 2, 3]
    assert dense.read(np.float32, 4096).tobytes() == ref.tobytes()
    empty = ctx.to_device(capi.sparse_to_numpy([], []))
    capi.assign_sparse(ctx, empty, dense, 9.0)
    nf = ctx.to_device(np.zeros(4, capi.IDX_VAL))
    capi.assign_sparse_relax(ctx, empty, dense, nf)
    assert capi.sparse_count(ctx, nf) == 0 and dense.read(np.float32, 4096).tobytes() == ref.tobytes()


def test_sparse_relax_repeated_indices(ctx, oracle):
    # assign_vector_sparse_module.h:318-335 walks the list in order; the device applies the entries
    # concurrently with compare-and-swap, so with REPEATED indices: inout must still equal the
    # reference's (the minimum wins whatever the interleaving), every improved index appears in the
    # new frontier with its final minimum, and every listed entry is an input entry that lowered the
    # value it found (val < the value before the call).  Distinct-index lists are covered bit-for-bit
    # by test_apply_golden_fixture.
    rng = np.random.default_rng(61)
    n = 5000
    for trial in range(4):
        base = rng.integers(0, 40, n).astype(np.float32)
        cnt = 60000
        idx = rng.integers(0, n // (1 + trial * 3), cnt).astype(np.uint32)    # heavy repetition
        val = rng.integers(0, 40, cnt).astype(np.float32)
        ref_inout, ref_idx, ref_val = oracle.port.assign_sparse_relax(idx, val, base)
        dl, dio = ctx.to_device(capi.sparse_to_numpy(idx, val)), ctx.to_device(base)
        dnf = ctx.to_device(np.zeros(cnt + 1, capi.IDX_VAL))
        capi.assign_sparse_relax(ctx, dl, dio, dnf)
        got = dio.read(np.float32, n)
        assert got.tobytes() == ref_inout.tobytes()
        fi, fv = dnf.read_sparse()
        assert (fv < base[fi]).all()                                   # each entry lowered what it found ...
        pairs = set(zip(idx.tolist(), val.tolist()))
        assert all((int(i), float(v)) in pairs for i, v in zip(fi, fv))   # ... and is an input entry
        improved = np.nonzero(ref_inout < base)[0]
        best = {}
        for i, v in zip(fi.tolist(), fv.tolist()):
            best[i] = min(best.get(i, np.inf), v)
        assert sorted(best) == improved.tolist()
        assert all(best[i] == ref_inout[i] for i in improved.tolist())
        assert set(ref_idx.tolist()) == set(improved.tolist())         # the reference improves the same indices


# ------------------------------------------------------------- pieces of the row-sharded runs, on one GPU
@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmspv_row_shards_tile_the_result(ctx, oracle, op, zero):
    # glb_csc_create_rows: each shard lists only its own rows and together they give the full result;
    # the frontier round trip of the sharded push (list -> dense rows -> list) loses nothing
    rng = np.random.default_rng(70 + op)
    g = datasets.powerlaw_csr(6000, 5000, 150_000, seed=17, max_degree=4000)
    ip, ix, d = oracle.port.csr2csc(g)
    csc = CSRMatrix(g.num_rows, g.num_cols, np.ones_like(d) if op else rng.random(len(d)).astype(np.float32), ix, ip)
    n = csc.num_rows
    k = 300
    idx = np.sort(rng.choice(csc.num_cols, k, replace=False)).astype(np.uint32)
    val = (1 + rng.integers(0, 9, k)).astype(np.float32)
    mask = np.where(rng.random(n) < 0.5, np.float32(zero), np.float32(1)).astype(np.float32)
    ref = oracle.port.spmspv(csc, op, zero, 1, idx, val, mask)
    dx = ctx.to_device(capi.sparse_to_numpy(idx, val, csc.num_cols + 1))
    dm = ctx.to_device(mask)
    dense = ctx.to_device(np.full(n, -9.0, np.float32))
    bounds = [0, 1000, 1001, 4096, n]
    for rb, re in zip(bounds[:-1], bounds[1:]):
        S = capi.CscMatrix(ctx, csc, rb, re)
        dy = ctx.to_device(np.zeros(n + 1, capi.IDX_VAL))
        for _ in range(2):   # second run: the accumulator was left clean
            S.spmspv(op, zero, 1, dx, dm, dy)
        oi, ov = dy.read_sparse()
        assert ((oi >= rb) & (oi < re)).all() and len(np.unique(oi)) == len(oi)
        got = densify(oi, ov, n, zero)
        if op == 0:
            assert_close_rel(got[rb:re], ref[rb:re], 1e-5)
        else:
            assert got[rb:re].tobytes() == ref[rb:re].tobytes()
        capi.sparse_to_dense_rows(ctx, dy, dense, rb, re, zero)
        S.close()
    full = dense.read(np.float32, n)
    check_vec(full, ref, op)
    relisted = ctx.to_device(np.zeros(n + 1, capi.IDX_VAL))
    capi.dense_to_sparse(ctx, dense, n, zero, relisted)
    oi, ov = relisted.read_sparse()
    assert len(np.unique(oi)) == len(oi) and not (ov == np.float32(zero)).any()
    assert densify(oi, ov, n, zero).tobytes() == full.tobytes()
    assert relisted.read(capi.IDX_VAL, 1)["val"][0] == np.float32(zero)


def test_exchange_api_with_a_single_rank(ctx, oracle):
    # the peer-mapped exchange degenerates cleanly to one rank: same results as glb_spmv, and every
    # entry point of the family (create / export / connect / vector / spmv / allgather / barrier /
    # host batch / status / destroy) runs on a one-GPU box
    rng = np.random.default_rng(80)
    m = datasets.powerlaw_csr(4096, 4096, 1 << 16, seed=19, max_degree=2000)
    n = m.num_rows
    A = capi.CsrMatrix(ctx, m)
    xc = capi.Exchange(ctx, n, 0, 1, lambda b: [b], n_vectors=3)
    assert not xc.has_multicast()
    x = rng.random(n).astype(np.float32)
    xc.barrier()
    xc.buffer(0).write(x)
    ref = x
    for it in range(3):
        xc.spmv(A, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, it % 2, (it + 1) % 2)
        ref = oracle.port.spmv(m, 0, 0.0, 0, ref)
    assert_close_rel(xc.buffer(1).read(np.float32, n), ref, 1e-4)
    xc.allgather(1, 0, n)
    xs, ys = [capi.PinnedArray(n) for _ in range(3)], [capi.PinnedArray(n) for _ in range(3)]
    for k in range(3):
        xs[k].array[:] = np.random.default_rng(90 + k).random(n).astype(np.float32)
    xc.spmv_host_batch(A, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, [a.ptr for a in xs], None, [a.ptr for a in ys])
    for k in range(3):
        assert_close_rel(ys[k].array, oracle.port.spmv(m, 0, 0.0, 0, xs[k].array), 1e-5)
    assert not xc.timed_out()
    xc.close()
    A.close()


def test_host_batch_matches_single_calls(ctx, oracle):
    # glb_spmv_host_batch (three-stream pipeline) against glb_spmv_host and the oracle, masked and not,
    # odd batch sizes, pageable and page-locked buffers
    rng = np.random.default_rng(85)
    m = datasets.powerlaw_csr(5000, 3000, 1 << 16, seed=23, max_degree=2500)
    A = capi.CsrMatrix(ctx, m)
    for count, pinned in ((1, True), (2, True), (5, True), (3, False)):
        mk = (lambda k: capi.PinnedArray(k).array) if pinned else (lambda k: np.empty(k, np.float32))
        hold = [capi.PinnedArray(max(m.num_rows, m.num_cols)) for _ in range(3 * count)] if pinned else None
        xs, ms, ys = [], [], []
        for k in range(count):
            xa = hold[3 * k].array[:m.num_cols] if pinned else mk(m.num_cols)
            ma = hold[3 * k + 1].array[:m.num_rows] if pinned else mk(m.num_rows)
            ya = hold[3 * k + 2].array[:m.num_rows] if pinned else mk(m.num_rows)
            xa[:] = rng.random(m.num_cols).astype(np.float32)
            ma[:] = rng.integers(0, 2, m.num_rows).astype(np.float32)
            ya[:] = np.nan
            xs.append(xa), ms.append(ma), ys.append(ya)
        for mt in (capi.MASK_NONE, capi.MASK_WRITE_TO_ONE):
            A.spmv_host_batch(capi.OP_MUL_ADD, 0.0, mt, xs, ms if mt else None, ys)
            single = np.empty(m.num_rows, np.float32)
            for k in range(count):
                assert_close_rel(ys[k], oracle.port.spmv(m, 0, 0.0, mt, xs[k], ms[k] if mt else None), 1e-5)
                A.spmv_host(capi.OP_MUL_ADD, 0.0, mt, xs[k], ms[k] if mt else None, single)
                assert single.tobytes() == ys[k].tobytes()      # same kernels, same order: bit-identical
    A.close()


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmv_tma_tile_variant(ctx, oracle, op, zero, monkeypatch):
    # the optional persistent kernel that keeps the hot vector in shared memory (TMA bulk load +
    # mbarrier, GLB_SPMV_TILE_THREADS): same results as the default path, masked and not
    monkeypatch.setenv("GLB_SPMV_TILE_THREADS", "512")
    monkeypatch.setenv("GLB_SPMV_TILE_K", "8192")
    rng = np.random.default_rng(95 + op)
    m = datasets.powerlaw_csr(20_000, 60_000, 600_000, seed=29, max_degree=30_000)
    if op == 2:
        m.data = rng.integers(0, 3, m.nnz).astype(np.float32)
    x = (rng.integers(0, 3, m.num_cols) * (0.5 + rng.random(m.num_cols))).astype(np.float32)
    mask = (rng.random(m.num_rows) < 0.3).astype(np.float32)
    for mt in MASKS:
        check_vec(gpu_spmv(ctx, m, op, zero, mt, x, mask), oracle.port.spmv(m, op, zero, mt, x, mask), op)


def test_assign_dense_vector_path_and_tails(ctx, oracle):
    # the dense assign reads the mask in 128-bit words and stores a fully assigned group of four as one
    # 128-bit word: odd lengths, all-hit / no-hit / mixed groups, unaligned views, mask aliasing inout
    rng = np.random.default_rng(97)
    for n in (1, 2, 5, 127, 128, 1001, 65536 + 3):
        for density in (0.0, 0.5, 1.0):
            mask = (rng.random(n) < density).astype(np.float32) * np.float32(-2.0)
            inout = rng.random(n).astype(np.float32)
            for mt in (capi.MASK_WRITE_TO_ZERO, capi.MASK_WRITE_TO_ONE):
                dm, dio = ctx.to_device(mask), ctx.to_device(inout)
                capi.assign_dense(ctx, dm, dio, n, 23.0, mt)
                ref = inout.copy()
                ref[(mask == 0) if mt == capi.MASK_WRITE_TO_ZERO else (mask != 0)] = 23.0
                assert dio.read(np.float32, n).tobytes() == ref.tobytes(), (n, density, mt)
    n = 4099
    mask, inout = (rng.random(n + 1) < 0.5).astype(np.float32), rng.random(n + 1).astype(np.float32)
    dm, dio = ctx.to_device(mask), ctx.to_device(inout)
    capi.assign_dense(ctx, dm.ptr + 4, dio.ptr + 4, n, 7.0, capi.MASK_WRITE_TO_ONE)     # 4-byte-aligned views
    ref = inout.copy()
    ref[1:][mask[1:] != 0] = 7.0
    assert dio.read(np.float32, n + 1).tobytes() == ref.tobytes()
    v = (rng.random(1000) < 0.5).astype(np.float32)
    d = ctx.to_device(v)
    capi.assign_dense(ctx, d, d, 1000, 9.0, capi.MASK_WRITE_TO_ZERO)                   # mask is inout
    assert d.read(np.float32, 1000).tobytes() == np.where(v == 0, np.float32(9.0), v).tobytes()


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmspv_conflict_matrix_and_dense_1k(ctx, oracle, op, zero):
    # the reference's "bank conflict" case (tests/test_module_spmv_spmspv.cpp:268-284: 1024 columns of 128
    # non-zeros, column i holding rows j * 8 + i % 8 -- every eighth column hits the same rows, i.e.
    # maximal contention on the accumulator) at sparsity 0, and dense_1K at sparsity 0 / 0.5 / 0.99
    n = 1024
    ip = (np.arange(n + 1) * (n // 8)).astype(np.uint32)
    ix = np.concatenate([np.arange(n // 8) * 8 + i % 8 for i in range(n)]).astype(np.uint32)
    conflict = CSRMatrix(n, n, np.full(len(ix), 1.0 / n, np.float32), ix, ip)
    dense = CSRMatrix(n, n, np.full(n * n, 1.0 / n, np.float32), np.tile(np.arange(n, dtype=np.uint32), n),
                      (np.arange(n + 1) * n).astype(np.uint32))
    rng = np.random.default_rng(99 + op)
    mask = np.where(rng.random(n) < 0.5, np.float32(zero), np.float32(1)).astype(np.float32)
    for csc, sparsities in ((conflict, (0.0,)), (dense, (0.0, 0.5, 0.99))):
        for sparsity in sparsities:
            k = max(1, int(n * (1 - sparsity)))
            idx = (np.arange(k) * (n // k)).astype(np.uint32)
            val = (rng.integers(0, 10, k) / 10).astype(np.float32)
            for mt in MASKS:
                y = gpu_spmspv(ctx, csc, op, zero, mt, idx, val, mask, runs=2)
                check_vec(y, oracle.port.spmspv(csc, op, zero, mt, idx, val, mask), op)
