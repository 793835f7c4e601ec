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
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "all"], stdout=subprocess.DEVNULL)


def _run(name):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        _build()
    env = dict(os.environ, GLB_TEST_DATA=os.path.join(ROOT, "tests", "golden"))
    p = subprocess.run([exe], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, f"{name} failed:\n{p.stdout[-4000:]}\n{p.stderr[-2000:]}"
    assert "0 test(s) failed" in p.stdout
    return p.stdout


def test_cpp_io_layer():
    _build()   # the C++ host mirror must compile against the public headers
    out = _run("test_io")
    assert out.count("[  OK  ]") >= 7


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_module_spmv_spmspv", "test_module_apply", "test_app"])
def test_cpp_modules_and_apps_on_gpu(name):
    out = _run(name)
    assert "FAILED" not in out
