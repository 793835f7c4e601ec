"""N-rank timeline of glb_spmv_host_batch_exchange on C2 (GLB_BATCH_TRACE=1): torchrun ... tools/trace_batch.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlily_b200 import capi, datasets  # noqa: E402
from graphlily_b200.exchange import open_exchange  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
rows = 4_194_304
m = datasets.powerlaw_csr(rows, rows, rows * 32, seed=42, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(rank, stream.cuda_stream)
ip = np.asarray(m.indptr, dtype=np.int64)
slot = rows // world
A = capi.CsrMatrix(ctx, m, rank * slot, (rank + 1) * slot)
xc, kind = open_exchange(ctx, rows, rank, world, n_vectors=4)
ring = 4
rng = np.random.default_rng(1)
xh = [torch.from_numpy(rng.integers(0, 2, rows).astype(np.float32)).pin_memory() for _ in range(ring)]
yh = [torch.zeros(rows, dtype=torch.float32).pin_memory() for _ in range(ring)]
steps = 24
xs = [xh[i % ring].data_ptr() for i in range(steps)]
ys = [yh[i % ring].data_ptr() for i in range(steps)]
xc.spmv_host_batch(A, 0, 0.0, 0, xs[:4], None, ys[:4])
for rep in range(2):
    if rep == 1:
        os.environ["GLB_BATCH_TRACE_ON"] = "1"
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    xc.spmv_host_batch(A, 0, 0.0, 0, xs, None, ys)
    dt = (time.perf_counter() - t0) * 1e3
    print(f"rank {rank}: {kind} batch of {steps}: {dt / steps:.4f} ms/vector", flush=True)
dist.barrier()
xc.close()
dist.destroy_process_group()
