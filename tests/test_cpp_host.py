"""The C++ host mirror (include/graphlily: io, module and app classes with the reference's names)
exercised through its own test binaries (tests/cpp, built by tests/cpp/Makefile).

* CPU: ``test_io`` -- containers, npz loader, csr2csc, padding, normalisation, SSSP preprocessing
  against the reference's golden vectors and the oracle;
* GPU: ``test_module_spmv_spmspv``, ``test_module_apply``, ``test_app`` -- the module / app classes
  over the C ABI against the oracle (bit-exact for or-and / min-plus, 1e-5 relative for plus-times).
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "bin")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so", "libvaltype_model.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "all"], stdout=subprocess.DEVNULL)


def _run(name, **extra_env):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        _build()
    env = dict(os.environ, GLB_TEST_DATA=os.path.join(ROOT, "tests", "golden"), **extra_env)
    p = subprocess.run([exe], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, f"{name} failed:\n{p.stdout[-4000:]}\n{p.stderr[-2000:]}"
    assert "0 test(s) failed" in p.stdout
    return p.stdout


def test_cpp_io_layer():
    _build()   # the C++ host mirror must compile against the public headers
    out = _run("test_io")
    assert out.count("[  OK  ]") >= 7


def test_cpp_npz_reader_on_scipy_files(tmp_path):
    """the real datasets are scipy.sparse.save_npz files (compressed, int32 index
    arrays -- int64 for huge matrices -- and an int64 shape).  The in-repo C++ reader (io/npz.h over
    zlib, replacing cnpy) and the Python loader read every variant scipy writes."""
    import numpy as np
    import scipy.sparse as sp
    import sys
    sys.path.insert(0, ROOT)
    from graphlily_b200 import io
    _build()
    rng = np.random.default_rng(3)
    base = sp.random(300, 200, density=0.05, format="csr", dtype=np.float32, random_state=5)
    base.sort_indices()
    holes = base.copy().tolil()
    holes[10:40] = 0
    holes = holes.tocsr().astype(np.float32)
    cases = {"compressed_i32": (base, True, np.int32), "stored_i32": (base, False, np.int32),
             "compressed_i64": (base, True, np.int64), "stored_i64": (base, False, np.int64),
             "empty_rows": (holes, True, np.int32), "no_nnz": (sp.csr_matrix((7, 9), dtype=np.float32), True, np.int32)}
    for name, (m, compressed, idt) in cases.items():
        m = m.copy()
        m.indices, m.indptr = m.indices.astype(idt), m.indptr.astype(idt)
        path = str(tmp_path / f"{name}.npz")
        sp.save_npz(path, m, compressed=compressed)
        with open(tmp_path / f"{name}.txt", "w") as f:
            f.write(f"{m.shape[0]} {m.shape[1]} {m.nnz} {float(m.indices.astype(np.float64).sum())!r} "
                    f"{float(m.indptr.astype(np.float64).sum())!r} {float(m.data.astype(np.float64).sum())!r}\n")
        r = io.load_csr_matrix_from_float_npz(path)          # the Python loader reads the same files
        assert (r.num_rows, r.num_cols, r.nnz) == (m.shape[0], m.shape[1], m.nnz)
        assert r.indices.tolist() == m.indices.tolist() and r.indptr.tolist() == m.indptr.tolist()
        assert r.data.tobytes() == m.data.astype(np.float32).tobytes()
    out = _run("test_io", GLB_NPZ_CASES=str(tmp_path))
    assert "[  OK  ] DataLoader.ScipyWrittenNpzVariants" in out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_module_spmv_spmspv", "test_module_apply", "test_app"])
def test_cpp_modules_and_apps_on_gpu(name):
    out = _run(name)
    assert "FAILED" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_valtypes_unsigned", "test_valtypes_ufixed"])
def test_cpp_value_type_builds(name):
    """the C++ mirror compiled with val_t = unsigned / ap_ufixed<32,8,AP_RND,AP_SAT>
    (-DGRAPHLILY_VAL_T_UNSIGNED / _UFIXED, the choices of the reference's global.h:60-64): module classes and
    BFS / SSSP bit for bit against the sequential model of those types (oracle/valtype_model.cpp)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libvaltype_model.so"], stdout=subprocess.DEVNULL)
    out = _run(name)
    assert "FAILED" not in out and out.count("[  OK  ]") >= 5


@pytest.mark.gpu
def test_cpp_sharded_apps_two_processes():
    """Row-sharded BFS / PageRank / SSSP from the C++ mirror: two processes, one GPU each, vectors in a
    CUDA-IPC peer exchange (no torch anywhere).  The binary reports SKIPPED on a 1-GPU box."""
    out = _run("test_sharded")
    if "SKIPPED" in out:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    assert out.count("[  OK  ]") == 2


@pytest.mark.gpu
def test_reference_own_test_files_unmodified(tmp_path):
    """SURVEY.md 8b "Callers: same call sequences must compile and run": the reference's OWN gtest files
    (/root/reference/tests/test_app.cpp, test_module_spmv_spmspv.cpp, test_module_apply.cpp), compiled
    unmodified against include/graphlily (tests/cpp/Makefile: ref_* targets, built where /root/reference
    exists), run here on the datasets they name -- seeded synthetic matrices of those shapes -- and pass:
    every kernel result within their own 1e-4 of compute_reference_results."""
    import sys
    bins = [os.path.join(BIN, n) for n in ("ref_test_app", "ref_test_module_spmv_spmspv", "ref_test_module_apply")]
    if not all(os.path.exists(b) for b in bins):
        if not os.path.exists("/root/reference/tests/test_app.cpp"):
            pytest.skip("the ref_* binaries are built from /root/reference in the build container")
        _build()
    data = tmp_path / "sparse_matrix_graph"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_ref_test_data.py"), str(data)],
                          stdout=subprocess.DEVNULL)
    for exe in bins:
        env = dict(os.environ, GLB_DATASET_DIR=str(data))
        p = subprocess.run([exe], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, f"{exe} failed:\n{p.stdout[-4000:]}\n{p.stderr[-2000:]}"
        assert "0 test(s) failed" in p.stdout and "[  OK  ]" in p.stdout
