"""The C2 configuration at FULL size (4 194 304 rows, 134 217 728 non-zeros -- what bench.py times),
checked through properties that do not need the sequential oracle to walk 134 M non-zeros per case:

* or-and and min-plus have closed forms numpy evaluates exactly with ``reduceat`` (any / min over the
  row's columns), so those two semirings are compared BIT-EXACT at full size;
* plus-times: the first 2^18 rows against the oracle (1e-5 relative), every row against a float64
  ``reduceat`` evaluation (1e-5 relative), and the checksum identity sum(y) = sum_c colsum[c] * x[c];
* the masked variants discard exactly the masked rows; row shards tile the result.
"""
import numpy as np
import pytest

from graphlily_b200 import capi, datasets

pytestmark = pytest.mark.gpu

ROWS, NNZ = 4_194_304, 134_217_728


@pytest.fixture(scope="module")
def c2(ctx):
    import torch
    m = datasets.powerlaw_csr(ROWS, ROWS, NNZ, seed=42, device=torch.device("cuda", 0))
    assert m.nnz == NNZ
    starts = m.indptr[:-1].astype(np.int64)
    assert (np.diff(m.indptr.astype(np.int64)) > 0).all()      # no empty rows: reduceat segments are rows
    A = capi.CsrMatrix(ctx, m)
    yield m, starts, A
    A.close()


def run(ctx, A, m, op, zero, mt, x, mask=None):
    dx, dy = ctx.to_device(x), ctx.to_device(np.full(m.num_rows, np.nan, np.float32))
    dm = ctx.to_device(mask) if mask is not None else None
    A.spmv(op, zero, mt, dx, dm, dy)
    return dy.read(np.float32, m.num_rows)


def test_c2_or_and_bit_exact(ctx, c2):
    m, starts, A = c2
    rng = np.random.default_rng(1)
    x = (rng.random(m.num_cols) < 0.02).astype(np.float32) * np.float32(3.5)
    ref = np.maximum.reduceat((x[m.indices] != 0).astype(np.uint8), starts).astype(np.float32)
    got = run(ctx, A, m, capi.OP_LOGICAL_AND_OR, 0.0, capi.MASK_NONE, x)
    assert got.tobytes() == ref.tobytes()
    mask = rng.integers(0, 2, m.num_rows).astype(np.float32)
    got = run(ctx, A, m, capi.OP_LOGICAL_AND_OR, 0.0, capi.MASK_WRITE_TO_ZERO, x, mask)
    assert got.tobytes() == np.where(mask != 0, np.float32(0), ref).tobytes()


def test_c2_min_plus_bit_exact(ctx, c2):
    m, starts, A = c2
    rng = np.random.default_rng(2)
    x = rng.integers(0, 200, m.num_cols).astype(np.float32)
    a = np.float32(m.data[0])                                   # C2 stores one value, 1 / N
    assert (m.data == a).all()
    ref = np.minimum(np.float32(255.0), np.minimum.reduceat(a + x[m.indices], starts)).astype(np.float32)
    got = run(ctx, A, m, capi.OP_ADD_MIN, 255.0, capi.MASK_NONE, x)
    assert got.tobytes() == ref.tobytes()


def test_c2_plus_times_against_oracle_fp64_and_checksum(ctx, c2, oracle):
    m, starts, A = c2
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2, m.num_cols).astype(np.float32)
    got = run(ctx, A, m, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, x)
    # every row against float64
    prod64 = m.data.astype(np.float64) * x[m.indices].astype(np.float64)
    ref64 = np.add.reduceat(prod64, starts)
    err = np.abs(got - ref64) / np.maximum(np.abs(ref64), 1e-30)
    assert ((err <= 1e-5) | (np.abs(got - ref64) < 1e-12)).all(), float(err.max())
    # checksum of checksums: sum(y) = sum_c colsum[c] * x[c]
    colsum = np.bincount(m.indices, weights=m.data.astype(np.float64), minlength=m.num_cols)
    assert abs(got.astype(np.float64).sum() - float(colsum @ x.astype(np.float64))) <= 1e-6 * float(colsum @ x)
    # the head of the matrix against the sequential oracle itself
    from graphlily_b200.io import CSRMatrix
    k = 1 << 18
    e = int(m.indptr[k])
    head = CSRMatrix(k, m.num_cols, m.data[:e], m.indices[:e], m.indptr[:k + 1])
    ref = oracle.port.spmv(head, 0, 0.0, 0, x)
    err = np.abs(got[:k] - ref) / np.maximum(np.abs(ref), 1e-30)
    assert ((err <= 1e-5) | (np.abs(got[:k] - ref) < 1e-12)).all(), float(err.max())
    # linearity: A (x + x2) = A x + A x2 within the same tolerance
    x2 = rng.integers(0, 3, m.num_cols).astype(np.float32)
    y2 = run(ctx, A, m, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, x2)
    y12 = run(ctx, A, m, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, x + x2)
    s = got.astype(np.float64) + y2.astype(np.float64)
    assert (np.abs(y12 - s) <= 2e-5 * np.maximum(np.abs(s), 1e-30) + 1e-12).all()


def test_c2_row_shards_tile_the_result(ctx, c2):
    m, starts, A = c2
    rng = np.random.default_rng(4)
    x = rng.integers(0, 200, m.num_cols).astype(np.float32)
    full = run(ctx, A, m, capi.OP_ADD_MIN, 255.0, capi.MASK_NONE, x)
    bounds = [0, 1_000_000, 1_000_032, 3_000_000, m.num_rows]
    dx, dy = ctx.to_device(x), ctx.to_device(np.full(m.num_rows, -7.0, np.float32))
    for rb, re in zip(bounds[:-1], bounds[1:]):
        S = capi.CsrMatrix(ctx, m, rb, re)
        S.spmv(capi.OP_ADD_MIN, 255.0, capi.MASK_NONE, dx, None, dy)
        S.close()
    assert dy.read(np.float32, m.num_rows).tobytes() == full.tobytes()
