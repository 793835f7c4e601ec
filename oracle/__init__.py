"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the parity oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product
(``graphlily_b200/``, ``include/``) never does.

Two back ends with the same Python surface:

* ``port``  -- ``oracle/liboracle.so``: the plain-C restatement in ``oracle.c``.
* ``ref``   -- ``oracle/_ref/libgraphlily_ref.so``: the reference's own
  ``compute_reference_results`` code compiled from ``/root/reference`` by
  ``oracle/Makefile`` (``None`` when it has not been built).

Matrices are passed as any object with ``num_rows, num_cols, indptr, indices,
data`` attributes (uint32 / uint32 / float32 numpy arrays).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "liboracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libgraphlily_ref.so")
REF_UFIXED_PATH = os.path.join(_HERE, "_ref", "libgraphlily_ref_ufixed.so")
VALMODEL_PATH = os.path.join(_HERE, "libvaltype_model.so")

OP_MUL_ADD, OP_LOGICAL_AND_OR, OP_ADD_MIN = 0, 1, 2
MASK_NONE, MASK_WRITE_TO_ZERO, MASK_WRITE_TO_ONE = 0, 1, 2

_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)


def build():
    """Compile liboracle.so (and _ref when /root/reference exists)."""
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _pu(a):
    return a.ctypes.data_as(_u32p)


def _pf(a):
    return None if a is None else a.ctypes.data_as(_f32p)


class _Backend:
    """Common numpy-facing surface over either shared library."""

    def __init__(self, path, prefix, is_ref):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.is_ref = is_ref
        self.path = path
        for name in ("spmv_timed",):
            getattr(self.lib, prefix + name).restype = C.c_double
        getattr(self.lib, prefix + "sssp_preprocess").restype = C.c_int64

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _mat(self, m):
        ip, ix, d = _u32(m.indptr), _u32(m.indices), _f32(m.data)
        return (C.c_uint32(int(m.num_rows)), C.c_uint32(int(m.num_cols)), _pu(ip), _pu(ix), _pf(d)), (ip, ix, d)

    # ---- module level -------------------------------------------------------------
    def spmv(self, m, op, zero, mask_type, x, mask=None, fast=True):
        args, keep = self._mat(m)
        x = _f32(x)
        mask = None if mask is None else _f32(mask)
        if mask_type != MASK_NONE and mask is None:
            raise ValueError("mask required")
        y = np.empty(int(m.num_rows), dtype=np.float32)
        extra = (C.c_int(int(fast)),) if self.is_ref else ()
        rc = self._fn("spmv")(*args, C.c_int(op), C.c_float(zero), C.c_int(mask_type), _pf(x), _pf(mask), _pf(y), *extra)
        assert rc == 0
        return y

    def spmv_timed(self, m, op, zero, x, reps=1):
        args, keep = self._mat(m)
        x = _f32(x)
        y = np.empty(int(m.num_rows), dtype=np.float32)
        sec = self._fn("spmv_timed")(*args, C.c_int(op), C.c_float(zero), _pf(x), _pf(y), C.c_int(reps))
        return float(sec), y

    def spmspv(self, m, op, zero, mask_type, x_idx, x_val, mask=None, fast=True):
        """``m`` is the CSC matrix (indptr over columns, indices = row ids). Returns dense y."""
        args, keep = self._mat(m)
        x_idx, x_val = _u32(x_idx), _f32(x_val)
        mask = None if mask is None else _f32(mask)
        y = np.empty(int(m.num_rows), dtype=np.float32)
        extra = (C.c_int(int(fast)),) if self.is_ref else ()
        rc = self._fn("spmspv")(*args, C.c_int(op), C.c_float(zero), C.c_int(mask_type), _pu(x_idx), _pf(x_val),
                                C.c_uint32(len(x_idx)), _pf(mask), _pf(y), *extra)
        assert rc == 0
        return y

    def ewise_add(self, vec, val):
        vec = _f32(vec)
        out = np.empty_like(vec)
        self._fn("ewise_add")(_pf(vec), _pf(out), C.c_uint32(len(vec)), C.c_float(val))
        return out

    def assign_dense(self, mask, inout, val, mask_type):
        """Returns (rc, new_inout); rc != 0 mirrors the reference's print-and-exit on kNoMask."""
        mask, out = _f32(mask), _f32(inout).copy()
        rc = self._fn("assign_dense")(_pf(mask), _pf(out), C.c_uint32(len(out)), C.c_float(val), C.c_int(mask_type))
        return rc, out

    def assign_sparse(self, m_idx, inout, val):
        m_idx, out = _u32(m_idx), _f32(inout).copy()
        if self.is_ref:
            self._fn("assign_sparse")(_pu(m_idx), None, C.c_uint32(len(m_idx)), _pf(out), C.c_uint32(len(out)), C.c_float(val))
        else:
            self._fn("assign_sparse")(_pu(m_idx), C.c_uint32(len(m_idx)), _pf(out), C.c_float(val))
        return out

    def assign_sparse_relax(self, m_idx, m_val, inout):
        """Returns (new_inout, frontier_idx, frontier_val) with the frontier in list order."""
        m_idx, m_val, out = _u32(m_idx), _f32(m_val), _f32(inout).copy()
        n = len(m_idx)
        nf_i, nf_v = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.float32)
        if self.is_ref:
            cnt = self._fn("assign_sparse_relax")(_pu(m_idx), _pf(m_val), C.c_uint32(n), _pf(out), C.c_uint32(len(out)), _pu(nf_i), _pf(nf_v))
        else:
            cnt = self._fn("assign_sparse_relax")(_pu(m_idx), _pf(m_val), C.c_uint32(n), _pf(out), _pu(nf_i), _pf(nf_v))
        return out, nf_i[:cnt].copy(), nf_v[:cnt].copy()

    # ---- io level -----------------------------------------------------------------
    def csr2csc(self, m):
        args, keep = self._mat(m)
        nnz = int(keep[0][int(m.num_rows)])
        oip = np.empty(int(m.num_cols) + 1, np.uint32)
        oix, od = np.empty(max(nnz, 1), np.uint32), np.empty(max(nnz, 1), np.float32)
        self._fn("csr2csc")(*args, _pu(oip), _pu(oix), _pf(od))
        return oip, oix[:nnz], od[:nnz]

    def round_dim(self, num_rows, num_cols, indptr, row_div, col_div):
        indptr = _u32(indptr)
        nr = -(-num_rows // row_div) * row_div
        oip = np.empty(nr + 1, np.uint32)
        dims = np.zeros(2, np.uint32)
        self._fn("round_dim")(C.c_uint32(num_rows), C.c_uint32(num_cols), _pu(indptr), C.c_uint32(row_div), C.c_uint32(col_div), _pu(oip), _pu(dims))
        return int(dims[0]), int(dims[1]), oip

    def normalize_outdegree(self, m):
        args, keep = self._mat(m)
        d = keep[2].copy()
        self._fn("normalize_outdegree")(args[0], args[1], args[2], args[3], _pf(d))
        return d

    def sssp_preprocess(self, m):
        args, keep = self._mat(m)
        nnz = int(keep[0][int(m.num_rows)])
        cap = nnz + int(m.num_rows) + 1
        oip = np.empty(int(m.num_rows) + 1, np.uint32)
        oix, od = np.empty(cap, np.uint32), np.empty(cap, np.float32)
        n = self._fn("sssp_preprocess")(*args, _pu(oip), _pu(oix), _pf(od))
        return oip, oix[:n].copy(), od[:n].copy()

    # ---- app level (matrix already preprocessed by the app's load_and_format_matrix) ------
    def bfs(self, m, source, iters):
        args, keep = self._mat(m)
        out = np.empty(int(m.num_rows), np.float32)
        if self.is_ref:
            self.lib.ref_app_bfs_csr(args[0], args[2], args[3], args[4], C.c_uint32(source), C.c_uint32(iters), _pf(out))
        else:
            self.lib.oracle_bfs(args[0], args[2], args[3], args[4], C.c_uint32(source), C.c_uint32(iters), _pf(out))
        return out

    def pagerank(self, m, damping, iters):
        args, keep = self._mat(m)
        out = np.empty(int(m.num_rows), np.float32)
        if self.is_ref:
            self.lib.ref_app_pagerank_csr(args[0], args[2], args[3], args[4], C.c_float(damping), C.c_uint32(iters), _pf(out))
        else:
            self.lib.oracle_pagerank(args[0], args[2], args[3], args[4], C.c_float(damping), C.c_uint32(iters), _pf(out))
        return out

    def sssp(self, m, source, iters, zero=255.0):
        args, keep = self._mat(m)
        out = np.empty(int(m.num_rows), np.float32)
        if self.is_ref:
            # the reference's SSSP hard-wires TropicalSemiring (zero = 255, global.h:99)
            assert zero == 255.0
            self.lib.ref_app_sssp_csr(args[0], args[2], args[3], args[4], C.c_uint32(source), C.c_uint32(iters), _pf(out))
        else:
            self.lib.oracle_sssp(args[0], args[2], args[3], args[4], C.c_uint32(source), C.c_uint32(iters), C.c_float(zero), _pf(out))
        return out


class _Ref(_Backend):
    """Extras only the compiled reference offers (npz loading, full app pipelines)."""

    def __init__(self, path):
        super().__init__(path, "ref_", True)
        for n in ("ref_app_bfs_npz", "ref_app_pagerank_npz", "ref_app_sssp_npz"):
            getattr(self.lib, n).restype = C.c_int64

    def constants(self):
        out = np.zeros(8, np.float32)
        self.lib.ref_constants(_pf(out))
        return out

    def load_npz(self, path):
        dims = np.zeros(3, np.uint32)
        self.lib.ref_load_npz(path.encode(), _pu(dims), None, None, None)
        nr, nc, nnz = (int(v) for v in dims)
        ip, ix, d = np.empty(nr + 1, np.uint32), np.empty(max(nnz, 1), np.uint32), np.empty(max(nnz, 1), np.float32)
        self.lib.ref_load_npz(path.encode(), _pu(dims), _pu(ip), _pu(ix), _pf(d))
        return nr, nc, ip, ix[:nnz], d[:nnz]

    def _app_npz(self, fn, path, rows, *scalars):
        out = np.empty(-(-rows // 128) * 128 + 128, np.float32)
        n = fn(path.encode(), *scalars, _pf(out))
        return out[:n].copy()

    def app_bfs_npz(self, path, rows, source, iters):
        return self._app_npz(self.lib.ref_app_bfs_npz, path, rows, C.c_uint32(source), C.c_uint32(iters))

    def app_pagerank_npz(self, path, rows, damping, iters):
        return self._app_npz(self.lib.ref_app_pagerank_npz, path, rows, C.c_float(damping), C.c_uint32(iters))

    def app_sssp_npz(self, path, rows, source, iters):
        return self._app_npz(self.lib.ref_app_sssp_npz, path, rows, C.c_uint32(source), C.c_uint32(iters))


def _load_port():
    if not os.path.exists(PORT_PATH):
        build()
    return _Backend(PORT_PATH, "oracle_", False)


def _load_ref(path=REF_PATH):
    if not os.path.exists(path):
        return None
    try:
        return _Ref(path)
    except OSError:
        return None


class _ValModel:
    """Sequential model of the unsigned / Q8.24 value-type variants (oracle/valtype_model.cpp); words are uint32."""
    U32, UFIXED = 1, 2

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.vt_ufixed_from_double.restype = C.c_uint32
        self.lib.vt_ufixed_from_double.argtypes = [C.c_double]
        self.lib.vt_ufixed_to_double.restype = C.c_double
        self.lib.vt_ufixed_to_double.argtypes = [C.c_uint32]
        for n in ("vt_ufixed_mul", "vt_ufixed_add"):
            getattr(self.lib, n).restype = C.c_uint32
            getattr(self.lib, n).argtypes = [C.c_uint32, C.c_uint32]

    def spmv(self, vt, m, op, zero, mask_type, x, mask=None):
        ip, ix, d = _u32(m.indptr), _u32(m.indices), _u32(m.data)
        x = _u32(x)
        mask = None if mask is None else _u32(mask)
        y = np.empty(int(m.num_rows), np.uint32)
        rc = self.lib.vt_spmv(C.c_int(vt), C.c_uint32(int(m.num_rows)), C.c_uint32(int(m.num_cols)), _pu(ip), _pu(ix), _pu(d),
                              C.c_int(op), C.c_uint32(int(zero)), C.c_int(mask_type), _pu(x), None if mask is None else _pu(mask), _pu(y))
        assert rc == 0
        return y

    def spmspv(self, vt, m, op, zero, mask_type, x_idx, x_val, mask=None):
        ip, ix, d = _u32(m.indptr), _u32(m.indices), _u32(m.data)
        x_idx, x_val = _u32(x_idx), _u32(x_val)
        mask = None if mask is None else _u32(mask)
        y = np.empty(int(m.num_rows), np.uint32)
        rc = self.lib.vt_spmspv(C.c_int(vt), C.c_uint32(int(m.num_rows)), C.c_uint32(int(m.num_cols)), _pu(ip), _pu(ix), _pu(d),
                                C.c_int(op), C.c_uint32(int(zero)), C.c_int(mask_type), _pu(x_idx), _pu(x_val),
                                C.c_uint32(len(x_idx)), None if mask is None else _pu(mask), _pu(y))
        assert rc == 0
        return y

    def ewise_add(self, vt, vec, val):
        vec = _u32(vec)
        out = np.empty_like(vec)
        self.lib.vt_ewise_add(C.c_int(vt), _pu(vec), _pu(out), C.c_uint32(len(vec)), C.c_uint32(int(val)))
        return out

    def assign_sparse_relax(self, vt, m_idx, m_val, inout):
        m_idx, m_val, out = _u32(m_idx), _u32(m_val), _u32(inout).copy()
        n = len(m_idx)
        nf_i, nf_v = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.uint32)
        cnt = self.lib.vt_assign_sparse_relax(C.c_int(vt), _pu(m_idx), _pu(m_val), C.c_uint32(n), _pu(out), _pu(nf_i), _pu(nf_v))
        return out, nf_i[:cnt].copy(), nf_v[:cnt].copy()

    def to_ufixed(self, x):
        return np.array([self.lib.vt_ufixed_from_double(float(v)) for v in np.atleast_1d(x)], np.uint32)

    def from_ufixed(self, w):
        return np.array([self.lib.vt_ufixed_to_double(int(v)) for v in np.atleast_1d(w)], np.float64)


def _load_valmodel():
    if not os.path.exists(VALMODEL_PATH):
        build()
    return _ValModel(VALMODEL_PATH)


port = _load_port()
ref = _load_ref()                      # the reference compiled with val_t = float: the fp32 parity target
ref_ufixed = _load_ref(REF_UFIXED_PATH)   # ... with val_t = the software ap_ufixed<32, 8, AP_RND, AP_SAT>
valmodel = _load_valmodel()
