"""The C restatement against the committed reference-produced fixtures (tests/golden/golden_v1.npz,
made by tests/golden/make_golden.py from oracle/_ref).  Runs everywhere, including machines
without /root/reference."""
import numpy as np
import pytest

from golden_util import golden, golden_csr
from graphlily_b200.io import CSRMatrix
from util import MASKS, SEMIRINGS


def same(a, b):
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mt", MASKS)
def test_spmv(oracle, op, zero, mt):
    z, m = golden(), golden_csr("spmv")
    same(oracle.port.spmv(m, op, zero, mt, z["spmv_x"], z["spmv_mask"]), z[f"spmv_y_op{op}_m{mt}"])


@pytest.mark.parametrize("op,zero", SEMIRINGS)
def test_spmv_powerlaw(oracle, op, zero):
    z, m = golden(), golden_csr("pl")
    same(oracle.port.spmv(m, op, zero, 0, z["pl_x"]), z[f"pl_y_op{op}"])


@pytest.mark.parametrize("op,zero", SEMIRINGS)
@pytest.mark.parametrize("mt", MASKS)
def test_spmspv(oracle, op, zero, mt):
    z, m = golden(), golden_csr("spmspv_csc")
    y = oracle.port.spmspv(m, op, zero, mt, z["spmspv_x_idx"], z["spmspv_x_val"], z[f"spmspv_mask_op{op}"])
    same(y, z[f"spmspv_y_op{op}_m{mt}"])


def test_apply(oracle):
    z, p = golden(), oracle.port
    same(p.ewise_add(z["apply_in"], 0.25), z["apply_ewise_add"])
    same(p.assign_dense(z["apply_mask"], z["apply_in"], 23.0, 1)[1], z["apply_assign_dense_m1"])
    same(p.assign_dense(z["apply_mask"], z["apply_in"], 23.0, 2)[1], z["apply_assign_dense_m2"])
    same(p.assign_sparse(z["apply_sparse_idx"], z["apply_in"], 7.0), z["apply_assign_sparse"])
    a, fi, fv = p.assign_sparse_relax(z["apply_sparse_idx"], z["apply_sparse_val"], z["apply_in"])
    same(a, z["apply_relax_inout"]), same(fi, z["apply_relax_idx"]), same(fv, z["apply_relax_val"])


def test_apps(oracle):
    z, p, g = golden(), oracle.port, golden_csr("app")
    same(p.bfs(g, 0, 10), z["app_bfs"])
    gp = CSRMatrix(g.num_rows, g.num_cols, p.normalize_outdegree(g) * np.float32(0.9), g.indices, g.indptr)
    same(p.pagerank(gp, 0.9, 10), z["app_pagerank"])
    sip, six, sd = p.sssp_preprocess(g)
    same(sip, z["app_sssp_indptr"]), same(six, z["app_sssp_indices"]), same(sd, z["app_sssp_data"])
    same(p.sssp(CSRMatrix(g.num_rows, g.num_cols, sd, six, sip), 0, 10), z["app_sssp"])
    assert z["app_bfs"].max() > 2 and (z["app_sssp"] < 255).sum() > 100   # the fixture is not degenerate
