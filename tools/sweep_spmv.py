"""GPU experiment: SpMV formatter / kernel variants on the C2 workload (one gpurun call).
Usage: python tools/sweep_spmv.py [rows] > gpurun_out/sweep.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlily_b200 import capi, datasets  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_194_304
dev = torch.device("cuda", 0)
m = datasets.powerlaw_csr(rows, rows, rows * 32, seed=42, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
x_host = np.random.default_rng(42).integers(0, 2, rows).astype(np.float32)
x = torch.from_numpy(x_host).to(dev)
y = torch.zeros_like(x)
alg_bytes = 8 * m.nnz + 4 * (rows + 1) + 4 * rows + 4 * rows

configs = [dict(GLB_SPMV_TILE_K=40960, GLB_SPMV_CARVEOUT=12, GLB_SPMV_TILE_THREADS=0)] + \
          [dict(GLB_SPMV_TILE_K=k, GLB_SPMV_CARVEOUT=12, GLB_SPMV_TILE_THREADS=t)
           for k, t in [(40960, 1024), (49152, 1024), (49152, 768), (40960, 768), (32768, 1024), (49152, 512)]]
if len(sys.argv) > 2:
    configs = [eval(sys.argv[2])]
ref = None
print(f"rows {rows} nnz {m.nnz} algorithmic bytes {alg_bytes}")
for cfg in configs:
    for k, v in cfg.items():
        os.environ[k] = str(v)
    t0 = time.time()
    A = capi.CsrMatrix(ctx, m)
    t_fmt = time.time() - t0
    for op, zero in ((0, 0.0),):
        for _ in range(5):
            A.spmv(op, zero, 0, x.data_ptr(), None, y.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50):
            A.spmv(op, zero, 0, x.data_ptr(), None, y.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        ctx.kernel_timing(True)
        for _ in range(20):
            A.spmv(op, zero, 0, x.data_ptr(), None, y.data_ptr())
        k_main, k_fix, n = ctx.kernel_timing_read()
        ctx.kernel_timing(False)
        got = y.cpu().numpy()
        if ref is None:
            ref = got.copy()
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)
        ok = bool(((err <= 2e-5) | (np.abs(got - ref) < 1e-12)).all())
        print(f"{cfg} format {t_fmt:.1f}s step {ms*1e3:.1f} us ({m.nnz/ms/1e6:.1f} GTEPS) main {k_main/n*1e3:.1f} us "
              f"fix {k_fix/n*1e3:.1f} us main-roofline {alg_bytes/(k_main/n)/1e6/6550.4:.3f} step-roofline "
              f"{alg_bytes/ms/1e6/6550.4:.3f} match {ok}", flush=True)
    A.close()
