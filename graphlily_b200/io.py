"""Host-side containers and matrix pre-processing (Python mirror of ``graphlily::io``).

Mirrors ``/root/reference/graphlily/io/data_loader.h`` and ``data_formatter.h:19-51``:
``CSRMatrix`` / ``CSCMatrix``, ``load_csr_matrix_from_float_npz``, ``csr2csc``,
``util_round_csr_matrix_dim``, ``util_normalize_csr_matrix_by_outdegree`` and the SSSP
``_preprocess`` (``app/sssp.h:16-62``).  The C++ host mirror lives in
``include/graphlily/io``; this module exists so the Python tests and ``bench.py`` can
prepare inputs the same way the apps do.  Pure host code: no device work, no oracle.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class CSRMatrix:
    """data_loader.h:19-31 (for a CSC matrix ``indptr`` runs over columns, data_loader.h:92-104)."""
    num_rows: int
    num_cols: int
    data: np.ndarray      # float32 [nnz]
    indices: np.ndarray   # uint32  [nnz]
    indptr: np.ndarray    # uint32  [num_rows + 1]

    @property
    def nnz(self):
        return int(self.indptr[-1])


CSCMatrix = CSRMatrix


def create_csr_matrix(num_rows, num_cols, data, indices, indptr):
    """data_loader.h:35-47"""
    return CSRMatrix(int(num_rows), int(num_cols), np.asarray(data, np.float32).copy(),
                     np.asarray(indices, np.uint32).copy(), np.asarray(indptr, np.uint32).copy())


def load_csr_matrix_from_float_npz(path):
    """data_loader.h:51-70: scipy.sparse.save_npz file with float32 data."""
    z = np.load(path)
    shape = z["shape"]
    return create_csr_matrix(int(shape[0]), int(shape[1]), z["data"], z["indices"], z["indptr"])


def save_csr_matrix_to_npz(path, m, compressed=True):
    """Write the layout scipy.sparse.save_npz produces (int32 indices, int64 shape, float32 data)."""
    fn = np.savez_compressed if compressed else np.savez
    fn(path, indices=m.indices.astype(np.int32), indptr=m.indptr.astype(np.int32), format=np.bytes_(b"csr"),
       shape=np.array([m.num_rows, m.num_cols], np.int64), data=m.data.astype(np.float32))


def csr2csc(m):
    """data_loader.h:108-144: counting-sort transpose; row order is preserved inside a column."""
    order = np.argsort(m.indices, kind="stable")
    counts = np.bincount(m.indices, minlength=m.num_cols).astype(np.uint64)
    indptr = np.zeros(m.num_cols + 1, np.uint32)
    indptr[1:] = np.cumsum(counts).astype(np.uint32)
    rows = np.repeat(np.arange(m.num_rows, dtype=np.uint32), np.diff(m.indptr.astype(np.int64)))
    return CSCMatrix(m.num_rows, m.num_cols, m.data[order].astype(np.float32), rows[order], indptr)


def util_round_csr_matrix_dim(m, row_divisor, col_divisor):
    """data_formatter.h:19-33 (in place): pad rows by repeating the last indptr, pad cols."""
    if m.num_rows % row_divisor:
        pad = row_divisor - m.num_rows % row_divisor
        m.indptr = np.concatenate([m.indptr, np.full(pad, m.indptr[m.num_rows], np.uint32)])
        m.num_rows += pad
    if m.num_cols % col_divisor:
        m.num_cols += col_divisor - m.num_cols % col_divisor
    return m


def util_normalize_csr_matrix_by_outdegree(m):
    """data_formatter.h:37-51 (in place): data[i] = float(1.0 / nnz_in_column(i)), double divide."""
    counts = np.bincount(m.indices, minlength=m.num_cols)
    m.data = (1.0 / counts[m.indices].astype(np.float64)).astype(np.float32)
    return m


def sssp_preprocess(m):
    """app/sssp.h:16-62 (in place): weights <- 1, weight-0 diagonal on every row.

    When every row already holds its diagonal (how the benches generate graphs) the
    function only rewrites weights and is vectorised.  Otherwise the reference's sequential
    insertion is followed literally, including its use of a not-yet-updated row end after
    earlier insertions (sssp.h:31-32,58)."""
    n = len(m.indptr) - 1
    ip = m.indptr.astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ip))
    is_diag = m.indices.astype(np.int64) == rows
    if n and np.array_equal(np.bincount(rows[is_diag], minlength=n), np.ones(n, np.int64)):
        m.data = np.where(is_diag, np.float32(0), np.float32(1)).astype(np.float32)
        return m
    idx, out_i, out_d, out_p, k = m.indices, [], [], [0], 0
    for r in range(n):
        s, e = int(ip[r]), int(ip[r + 1])
        length = e - s
        ins, zero_at = -1, -1
        if length == k:
            ins = s
        elif length > k:
            scan_end = e - k
            for i in range(s, scan_end):
                c = int(idx[i])
                if c == r:
                    zero_at = i
                    break
                if c > r or i == scan_end - 1:
                    ins = i
                    break
        if length == k:
            out_i.append(r)
            out_d.append(0.0)
        for i in range(s, e):
            if length != k and ins == i:
                out_i.append(r)
                out_d.append(0.0)
            out_i.append(int(idx[i]))
            out_d.append(0.0 if zero_at == i else 1.0)
        if ins >= 0:
            k += 1
        out_p.append(len(out_i))
    m.indices = np.asarray(out_i, np.uint32)
    m.data = np.asarray(out_d, np.float32)
    m.indptr = np.asarray(out_p, np.uint32)
    return m
