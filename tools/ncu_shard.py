"""ncu target: a few SpMV launches on ONE 1/N row shard of C2 (what a rank of an N-GPU run executes).
Usage: ncu ... python tools/ncu_shard.py [N=8]"""
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
A = capi.CsrMatrix(ctx, m, cuts[N // 2], cuts[N // 2 + 1])
for _ in range(4):
    A.spmv(0, 0.0, 0, x.data_ptr(), None, y.data_ptr())
torch.cuda.synchronize()
print(A.info())
