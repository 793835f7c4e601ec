"""GPU experiment: main-kernel time of ONE 1/N row shard of C2 (what a rank of an N-GPU run executes) for
different chunk sizes.  Usage: python tools/sweep_shard.py [N=8]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlily_b200 import capi, datasets  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows = 4_194_304
dev = torch.device("cuda", 0)
m = datasets.powerlaw_csr(rows, rows, rows * 32, seed=42, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
x = torch.from_numpy(np.random.default_rng(42).integers(0, 2, rows).astype(np.float32)).to(dev)
y = torch.zeros_like(x)
ip = np.asarray(m.indptr, dtype=np.int64)
cuts = [int(np.searchsorted(ip, ip[-1] * r // N)) // 32 * 32 for r in range(N + 1)]
cuts[0], cuts[-1] = 0, rows
for shard in (0, N // 2, N - 1):
    rb, re = cuts[shard], cuts[shard + 1]
    for groups in ("8", "4", "2"):
        os.environ["GLB_SPMV_MAX_GROUPS"] = groups
        A = capi.CsrMatrix(ctx, m, rb, re)
        info = A.info()
        for _ in range(5):
            A.spmv(0, 0.0, 0, x.data_ptr(), None, y.data_ptr())
        ctx.kernel_timing(True)
        for _ in range(30):
            A.spmv(0, 0.0, 0, x.data_ptr(), None, y.data_ptr())
        k_main, k_fix, n = ctx.kernel_timing_read()
        ctx.kernel_timing(False)
        alg = 8 * info["nnz"] + 4 * (re - rb + 1) + 4 * rows + 4 * (re - rb)
        print(f"shard {shard}/{N} rows {re - rb} nnz {info['nnz']} groups {groups}: chunks {info['chunks']} fixups {info['fixups']} "
              f"main {k_main / n * 1e3:.1f} us fix {k_fix / n * 1e3:.1f} us roofline {alg / (k_main / n) / 1e6 / 6551:.3f}", flush=True)
        A.close()
