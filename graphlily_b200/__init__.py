"""graphlily_b200 -- B200-native SpMV / SpMSpV engine behind GraphLily's operator API.

Layout of the package (only what the hot path needs):

* ``csrc/``       hand-written sm_100a CUDA kernels + the C ABI (``include/graphlily_b200.h``)
* ``lib/``        the built ``libgraphlily_b200.so`` (git-ignored, travels to the GPU box)
* ``capi``        ctypes binding of that C ABI (no CPU fallback)
* ``io``          host containers / pre-processing mirroring ``graphlily::io``
* ``module``      ``SpMVModule`` ... operator classes mirroring ``graphlily::module``
* ``app``         ``BFS`` / ``PageRank`` / ``SSSP`` mirroring ``graphlily::app``
* ``datasets``    seeded synthetic graphs of the benchmark shapes

The C++ host mirror (same class and method names as the reference) is header-only under
``include/graphlily``.
"""
from . import io  # noqa: F401
from .io import CSRMatrix, CSCMatrix  # noqa: F401

ArithmeticSemiring = (0, 1.0, 0.0)     # {kMulAdd, one, zero}      global.h:96
LogicalSemiring = (1, 1.0, 0.0)        # {kLogicalAndOr, 1, 0}     global.h:97
TropicalSemiring = (2, 0.0, 255.0)     # {kAddMin, 0, UFIXED_INF}  global.h:99 (shipped variant)
kNoMask, kMaskWriteToZero, kMaskWriteToOne = 0, 1, 2   # global.h:103-107
