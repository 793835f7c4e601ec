"""Shared helpers for the tests (CPU side)."""
import numpy as np

from graphlily_b200.io import CSRMatrix

# the 4x4 matrix of /root/reference/tests/test_io.cpp:40-51 (csr_matrix_1)
#   [[1,2,3,4],[5,0,6,0],[0,7,0,0],[0,0,0,8]]
def csr_matrix_1():
    return CSRMatrix(4, 4, np.array([1, 2, 3, 4, 5, 6, 7, 8], np.float32),
                     np.array([0, 1, 2, 3, 0, 2, 1, 3], np.uint32), np.array([0, 4, 6, 7, 8], np.uint32))


def random_csr(rng, n_rows, n_cols, density, values="rand", empty_frac=0.2):
    """Random CSR with sorted distinct columns, a share of empty rows, float32 data."""
    indptr, indices = [0], []
    for _ in range(n_rows):
        if rng.random() < empty_frac:
            k = 0
        else:
            k = int(rng.binomial(n_cols, density))
        cols = np.sort(rng.choice(n_cols, size=min(k, n_cols), replace=False)) if k else np.zeros(0, np.int64)
        indices.append(cols)
        indptr.append(indptr[-1] + len(cols))
    indices = np.concatenate(indices).astype(np.uint32) if indices else np.zeros(0, np.uint32)
    nnz = len(indices)
    if values == "rand":
        data = rng.random(nnz).astype(np.float32)
    elif values == "ones":
        data = np.ones(nnz, np.float32)
    else:
        data = rng.integers(0, 4, nnz).astype(np.float32)
    return CSRMatrix(n_rows, n_cols, data, indices, np.asarray(indptr, np.uint32))


def densify(idx, val, n, zero):
    out = np.full(n, zero, np.float32)
    out[np.asarray(idx, np.int64)] = val
    return out


SEMIRINGS = [(0, 0.0), (1, 0.0), (2, 255.0)]   # (op, zero): Arithmetic, Logical, Tropical (global.h:96-99)
MASKS = [0, 1, 2]                                # kNoMask, kMaskWriteToZero, kMaskWriteToOne


def assert_close_rel(got, ref, rel=1e-5):
    """|g - r| <= rel * max(|r|, tiny): the fp32 tolerance north_star states (1e-5 relative)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    tiny = 1e-30
    err = np.abs(got - ref) / np.maximum(np.abs(ref), tiny)
    bad = np.where((err > rel) & (np.abs(got - ref) > 1e-12))[0]
    assert bad.size == 0, f"{bad.size} mismatches, first at {bad[:5]}: got {got[bad[:5]]} ref {ref[bad[:5]]}"
