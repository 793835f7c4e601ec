"""GPU experiment: per-level times of the push direction (fused SpMSpV launches) on a named shape.
Usage: python tools/profile_push.py [sssp|bfs] [scale]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlily_b200 import app, capi, datasets  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sssp"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = capi.Context(0, stream.cuda_stream)
g = (datasets.c5_orkut if name == "sssp" else datasets.c3_gplus)(scale, device=dev)
a = (app.SSSP if name == "sssp" else app.BFS)()
a.set_up_runtime(None, ctx=ctx)
a.load_and_format_matrix(g)
a.send_matrix_host_to_device()
iters = 6 if name == "sssp" else 7
a.use_graphs_ = False
for rep in range(2):
    a._push_setup(0)
    lists = a._frontier_lists() if name == "sssp" else [a.SpMSpV_.vector_buf, a.SpMSpV_.results_buf]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    sizes = []
    ev[0].record(stream)
    for level in range(1, iters + 1):
        sizes.append(capi.sparse_count(ctx, lists[(level - 1) & 1]))
        ev[level - 1].record(stream)
        a._push_level_fused(lists, level)
        ev[level].record(stream) if level == iters else None
    torch.cuda.synchronize()
ip = np.diff(a.SpMSpV_.csc_matrix_float_.indptr.astype(np.int64))
print(f"{name}: {a.matrix_num_rows_} vertices, nnz {a.get_nnz()}, columns > 512 nnz: {(ip > 512).sum()} holding {ip[ip > 512].sum()} nnz")
for level in range(1, iters + 1):
    print(f"level {level}: frontier {sizes[level - 1]:8d}  {ev[level - 1].elapsed_time(ev[level]) * 1e3:9.1f} us (includes one 8-byte host read)")
