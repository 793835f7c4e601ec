"""the reference's other two `val_t` choices (global.h:60-64) -- `unsigned` and
ap_ufixed<32, 8, AP_RND, AP_SAT>, the type of the shipped bitstream.

CPU part (no GPU): the software ap_ufixed of oracle/shim/ap_fixed.h against known answers of the documented
semantics; the sequential value-type model (oracle/valtype_model.cpp) against an independent big-integer
restatement in Python; and the reference compiled BOTH ways (val_t = float / val_t = that ap_ufixed), which
agree wherever matrix values are representable in Q8.24 and differ exactly where spmspv_module.h:471 reads
the matrix through val_t.

GPU part: the glb_*_vt entry points against the value-type model, bit for bit.

PARITY UNPINNED: nothing the reference ships (no bitstream, no emulator, no golden vector) pins its device
numerics; the model follows the Vitis-HLS documentation of ap_ufixed and the ALUs of ufixed_pe_fwd.h:23-65."""
import numpy as np
import pytest

from graphlily_b200.io import CSRMatrix
from util import MASKS, random_csr

U32, UFX = 1, 2
MAXW = 0xFFFFFFFF


# ---- independent restatement with Python integers ---------------------------------------------------
def py_times(vt, a, b):
    a, b = int(a), int(b)
    if vt == U32:
        return (a * b) & MAXW
    return min(((a * b) + (1 << 23)) >> 24, MAXW)       # exact product, round half up, saturate


def py_plus(vt, a, b):
    s = int(a) + int(b)
    return s & MAXW if vt == U32 else min(s, MAXW)


def py_one(vt):
    return 1 if vt == U32 else 1 << 24


def py_semi(vt, op, acc, a, b):
    if op == 0:
        return py_plus(vt, acc, py_times(vt, a, b))
    if op == 1:
        return py_one(vt) if (acc != 0 or (a != 0 and b != 0)) else 0
    return min(acc, py_plus(vt, a, b))


def py_spmv(vt, m, op, zero, x):
    y = np.empty(m.num_rows, np.uint32)
    for r in range(m.num_rows):
        acc = int(zero)
        for i in range(int(m.indptr[r]), int(m.indptr[r + 1])):
            acc = py_semi(vt, op, acc, m.data[i], x[m.indices[i]])
        y[r] = acc
    return y


def words(rng, vt, n, kind):
    """Random words of the value type: small integers, fractions, and values that saturate / wrap."""
    if kind == "small":
        v = rng.integers(0, 6, n).astype(np.uint64)
        return (v if vt == U32 else v << 24).astype(np.uint32)
    if kind == "frac":
        return rng.integers(0, 1 << 26, n).astype(np.uint32)          # Q8.24: [0, 4)
    return rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)   # anything: products overflow


def word_matrix(rng, vt, n, m, density, kind):
    base = random_csr(rng, n, m, density, values="small")
    return CSRMatrix(n, m, words(rng, vt, base.nnz, kind).view(np.float32), base.indices, base.indptr)


def as_words(m):
    return CSRMatrix(m.num_rows, m.num_cols, np.asarray(m.data).view(np.uint32), m.indices, m.indptr)


# ---- CPU: the software ap_ufixed --------------------------------------------------------------------
def test_ap_ufixed_known_answers(oracle):
    vm = oracle.valmodel
    f = lambda x: int(vm.to_ufixed(x)[0])
    assert f(0.0) == 0 and f(1.0) == 1 << 24 and f(0.5) == 1 << 23 and f(255.0) == 255 << 24
    assert f(2.0 ** -24) == 1 and f(2.0 ** -25) == 1 and f(2.0 ** -25 * 0.999) == 0    # AP_RND: half an ulp rounds up
    assert f(1.0 + 3 * 2.0 ** -25) == (1 << 24) + 2                                     # 1.5 ulp -> 2
    assert f(256.0) == MAXW and f(999999999.0) == MAXW and f(-3.0) == 0                 # AP_SAT both ways
    assert abs(vm.from_ufixed(MAXW)[0] - (256 - 2.0 ** -24)) == 0
    lib = vm.lib
    assert lib.vt_ufixed_mul(1 << 24, 1 << 24) == 1 << 24                               # 1 * 1
    assert lib.vt_ufixed_mul(1 << 23, 1 << 23) == 1 << 22                               # 0.5 * 0.5
    assert lib.vt_ufixed_mul(1, 1 << 23) == 1 and lib.vt_ufixed_mul(1, (1 << 23) - 1) == 0   # rounding of the product
    assert lib.vt_ufixed_mul(200 << 24, 2 << 24) == MAXW                                # 400 saturates
    assert lib.vt_ufixed_add(200 << 24, 100 << 24) == MAXW and lib.vt_ufixed_add(MAXW, 0) == MAXW
    assert lib.vt_ufixed_add(255 << 24, 1 << 24) == MAXW                                # inf + 1 = inf: min-plus keeps UFIXED_INF


@pytest.mark.parametrize("vt", [U32, UFX])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_valtype_model_against_python_integers(oracle, vt, op):
    rng = np.random.default_rng(300 + 10 * vt + op)
    for kind in ("small", "frac", "any"):
        m = word_matrix(rng, vt, 40, 50, 0.2, kind)
        mw = as_words(m)
        x = words(rng, vt, 50, kind)
        zero = {0: 0, 1: 0, 2: MAXW if vt == U32 else 255 << 24}[op]
        ref = py_spmv(vt, mw, op, zero, x)
        assert oracle.valmodel.spmv(vt, mw, op, zero, 0, x).tolist() == ref.tolist()
        mask = rng.integers(0, 2, 40).astype(np.uint32)
        got = oracle.valmodel.spmv(vt, mw, op, zero, 1, x, mask)
        assert got.tolist() == np.where(mask != 0, 0, ref).tolist()
        # SpMSpV: the same arrays read as a CSC matrix (40 columns, 50 rows), a partial frontier
        idx = np.nonzero(rng.random(40) < 0.5)[0].astype(np.uint32)
        xv = words(rng, vt, len(idx), kind)
        exp = [int(zero)] * 50
        for k, c in enumerate(idx):
            for i in range(int(mw.indptr[c]), int(mw.indptr[c + 1])):
                r = int(mw.indices[i])
                exp[r] = py_semi(vt, op, exp[r], mw.data[i], xv[k])
        csc = CSRMatrix(50, 40, mw.data, mw.indices, mw.indptr)
        assert oracle.valmodel.spmspv(vt, csc, op, zero, 0, idx, xv).tolist() == exp
        m2 = rng.integers(0, 2, 50).astype(np.uint32) * np.uint32(zero if zero else 3)
        got = oracle.valmodel.spmspv(vt, csc, op, zero, 2, idx, xv, m2)       # kMaskWriteToOne: off where mask == zero
        assert got.tolist() == [int(zero) if m2[r] == zero else exp[r] for r in range(50)]


def test_reference_compiled_both_ways(oracle):
    """val_t = float and val_t = the software ap_ufixed give the same CPU results wherever the matrix values
    are Q8.24 numbers (all app matrices: weights 0 / 1); for other values the SpMSpV path differs -- it reads
    the matrix through val_t (spmspv_module.h:471) -- and equals the float build run on the quantised matrix."""
    if oracle.ref is None or oracle.ref_ufixed is None:
        pytest.skip("oracle/_ref not built")
    a, b, vm = oracle.ref, oracle.ref_ufixed, oracle.valmodel
    rng = np.random.default_rng(5)
    for op, zero in ((0, 0.0), (1, 0.0), (2, 255.0)):
        m = random_csr(rng, 120, 120, 0.1, values="rand" if op == 0 else "small")
        x = rng.random(120).astype(np.float32)
        mask = rng.integers(0, 2, 120).astype(np.float32)
        for mt in MASKS:
            assert a.spmv(m, op, zero, mt, x, mask).tobytes() == b.spmv(m, op, zero, mt, x, mask).tobytes()   # SpMV: float matrix
        idx = rng.choice(120, 40, replace=False).astype(np.uint32)
        val = (rng.integers(0, 10, 40) / 8).astype(np.float32)
        dyadic = CSRMatrix(120, 120, (rng.integers(0, 64, m.nnz) / 16).astype(np.float32), m.indices, m.indptr)
        assert a.spmspv(dyadic, op, zero, 0, idx, val, mask).tobytes() == b.spmspv(dyadic, op, zero, 0, idx, val, mask).tobytes()
    m = random_csr(rng, 120, 120, 0.1, values="rand")
    quantised = CSRMatrix(120, 120, vm.from_ufixed(vm.to_ufixed(m.data)).astype(np.float32), m.indices, m.indptr)
    idx = rng.choice(120, 40, replace=False).astype(np.uint32)
    val = rng.random(40).astype(np.float32)
    yb = b.spmspv(m, 0, 0.0, 0, idx, val, None)
    assert yb.tobytes() == a.spmspv(quantised, 0, 0.0, 0, idx, val, None).tobytes()
    assert yb.tobytes() != a.spmspv(m, 0, 0.0, 0, idx, val, None).tobytes()
    g = random_csr(rng, 256, 256, 0.05, values="ones")
    assert a.bfs(g, 0, 6).tobytes() == b.bfs(g, 0, 6).tobytes()


# ---- GPU: the _vt entry points ----------------------------------------------------------------------
def zero_of(vt, op):
    return {0: 0, 1: 0, 2: MAXW if vt == U32 else 255 << 24}[op]


@pytest.mark.gpu
@pytest.mark.parametrize("vt", [U32, UFX])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_gpu_spmv_value_types(ctx, oracle, vt, op):
    from graphlily_b200 import capi, datasets
    rng = np.random.default_rng(400 + 10 * vt + op)
    cases = [(random_csr(rng, 700, 500, 0.05, values="small"), "any"), (random_csr(rng, 300, 300, 0.2, values="small"), "frac"),
             (datasets.powerlaw_csr(3000, 3000, 60000, seed=31, max_degree=2500), "small")]
    for base, kind in cases:
        m = CSRMatrix(base.num_rows, base.num_cols, words(rng, vt, base.nnz, kind).view(np.float32), base.indices, base.indptr)
        mw = as_words(m)
        x = words(rng, vt, m.num_cols, kind)
        mask = (rng.integers(0, 2, m.num_rows) * rng.integers(1, 1 << 30, m.num_rows)).astype(np.uint32)
        zero = zero_of(vt, op)
        A = capi.CsrMatrix(ctx, m)
        dx, dm = ctx.to_device(x), ctx.to_device(mask)
        for mt in MASKS:
            dy = ctx.to_device(np.full(m.num_rows, 0xDEADBEEF, np.uint32))
            capi.check(capi.lib.glb_spmv_vt(ctx.handle, A.handle, vt, op, zero, mt, dx.ptr, dm.ptr if mt else None, dy.ptr, None))
            got = dy.read(np.uint32, m.num_rows)
            ref = oracle.valmodel.spmv(vt, mw, op, zero, mt, x, mask if mt else None)
            assert got.tobytes() == ref.tobytes(), (kind, mt, np.nonzero(got != ref)[0][:5])
        # fused epilogue: eWiseAdd of a word, then the dense assign on the result
        add = int(words(rng, vt, 1, kind)[0])
        inout = words(rng, vt, m.num_rows, "small")
        dio, dy = ctx.to_device(inout), ctx.to_device(np.zeros(m.num_rows, np.uint32))
        ep = capi.Epilogue(1, np.uint32(add).view(np.float32), dio.ptr, np.uint32(77).view(np.float32), capi.MASK_WRITE_TO_ONE)
        capi.check(capi.lib.glb_spmv_vt(ctx.handle, A.handle, vt, op, zero, 0, dx.ptr, None, dy.ptr, ep))
        ref = oracle.valmodel.ewise_add(vt, oracle.valmodel.spmv(vt, mw, op, zero, 0, x), add)
        assert dy.read(np.uint32, m.num_rows).tobytes() == ref.tobytes()
        assert dio.read(np.uint32, m.num_rows).tobytes() == np.where(ref != 0, 77, inout).astype(np.uint32).tobytes()
        A.close()


@pytest.mark.gpu
@pytest.mark.parametrize("vt", [U32, UFX])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_gpu_spmspv_and_apply_value_types(ctx, oracle, vt, op):
    from graphlily_b200 import capi, datasets, io
    rng = np.random.default_rng(500 + 10 * vt + op)
    g = datasets.social_graph(3000, 60000, 700, seed=14, diagonal=True)
    csc = io.csr2csc(g)
    n = g.num_rows
    zero = zero_of(vt, op)
    for kind in ("small", "frac", "any"):
        m = CSRMatrix(n, n, words(rng, vt, csc.nnz, kind).view(np.float32), csc.indices, csc.indptr)
        mw = as_words(m)
        A = capi.CscMatrix(ctx, m)
        dy = ctx.to_device(np.zeros(n + 1, capi.IDX_VAL))
        for density in (0.002, 0.3):
            idx = np.nonzero(rng.random(n) < density)[0].astype(np.uint32)
            xv = words(rng, vt, len(idx), kind)
            mask = np.where(rng.random(n) < 0.5, np.uint32(zero), np.uint32(5)).astype(np.uint32)
            dx = ctx.to_device(capi.sparse_to_numpy(idx, xv.view(np.float32), n + 1))
            dm = ctx.to_device(mask)
            for mt in MASKS:
                capi.check(capi.lib.glb_spmspv_vt(ctx.handle, A.handle, vt, op, zero, mt, dx.ptr, dm.ptr if mt else None, dy.ptr))
                cnt = capi.sparse_count(ctx, dy)
                body = dy.read(capi.IDX_VAL, cnt + 1)[1:]
                got = np.full(n, zero, np.uint32)
                got[body["index"]] = body["val"].view(np.uint32)
                ref = oracle.valmodel.spmspv(vt, mw, op, zero, mt, idx, xv, mask if mt else None)
                assert len(np.unique(body["index"])) == cnt and not (body["val"].view(np.uint32) == zero).any()
                assert got.tobytes() == ref.tobytes(), (kind, density, mt)
        A.close()
    # apply operators on words
    v = words(rng, vt, 10007, "any")
    val = int(words(rng, vt, 1, "any")[0])
    din, dout = ctx.to_device(v), ctx.to_device(np.zeros(10007, np.uint32))
    capi.check(capi.lib.glb_ewise_add_vt(ctx.handle, vt, din.ptr, dout.ptr, 10007, val))
    assert dout.read(np.uint32, 10007).tobytes() == oracle.valmodel.ewise_add(vt, v, val).tobytes()
    maskw = (rng.integers(0, 2, 10007) * rng.integers(1, 1 << 31, 10007)).astype(np.uint32)
    for mt in (1, 2):
        dio, dm = ctx.to_device(v), ctx.to_device(maskw)
        capi.check(capi.lib.glb_assign_dense_vt(ctx.handle, vt, dm.ptr, dio.ptr, 10007, 1234567, mt))
        hit = (maskw == 0) if mt == 1 else (maskw != 0)
        assert dio.read(np.uint32, 10007).tobytes() == np.where(hit, 1234567, v).astype(np.uint32).tobytes()
    li = rng.choice(10007, 3000, replace=False).astype(np.uint32)
    lv = words(rng, vt, 3000, "any")
    dl = ctx.to_device(capi.sparse_to_numpy(li, lv.view(np.float32)))
    dio, dnf = ctx.to_device(v), ctx.to_device(np.zeros(3001, capi.IDX_VAL))
    capi.check(capi.lib.glb_assign_sparse_relax_vt(ctx.handle, vt, dl.ptr, dio.ptr, dnf.ptr))
    ref_io, ref_i, ref_v = oracle.valmodel.assign_sparse_relax(vt, li, lv, v)
    assert dio.read(np.uint32, 10007).tobytes() == ref_io.tobytes()
    cnt = capi.sparse_count(ctx, dnf)
    body = dnf.read(capi.IDX_VAL, cnt + 1)[1:]
    o, ro = np.argsort(body["index"]), np.argsort(ref_i)
    assert body["index"][o].tolist() == ref_i[ro].tolist() and body["val"].view(np.uint32)[o].tolist() == ref_v[ro].tolist()
    dio = ctx.to_device(v)
    capi.check(capi.lib.glb_assign_sparse_vt(ctx.handle, vt, dl.ptr, dio.ptr, 4242))
    exp = v.copy()
    exp[li] = 4242
    assert dio.read(np.uint32, 10007).tobytes() == exp.tobytes()
