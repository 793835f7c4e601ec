#!/usr/bin/env python
"""Apply operators (overlay modes 3-6) against the HBM roofline: python tools/bench_apply.py [len]

eWiseAdd, dense assign, sparse assign (BFS) and sparse relax (SSSP) on vectors larger than L2
(default 64 Mi floats = 256 MB), device-resident, CUDA events on the context's stream.  Algorithmic
bytes per SURVEY.md 8a: eWiseAdd 8 * len; dense assign 4 * len (mask) + 4 * len_assigned (writes; all
of them here); sparse assign 8 * nnz (list) + 4 * nnz (scattered writes); relax 8 * nnz + 8 * nnz
(read + conditional write of inout) + 8 * nnz_out.  One JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from graphlily_b200 import capi
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(0, stream.cuda_stream)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    a, b = ctx.zeros_f32(n, 1.0), ctx.zeros_f32(n, 0.0)
    rng = np.random.default_rng(0)
    k = n // 8
    idx = rng.choice(n, k, replace=False).astype(np.uint32)
    lst = ctx.to_device(capi.sparse_to_numpy(idx, np.ones(k, np.float32)))
    out = ctx.to_device(np.zeros(k + 1, capi.IDX_VAL))

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {}

    def report(name, ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak}

    report("ewise_add", timed(lambda: capi.ewise_add(ctx, a, b, n, 0.5)), 8 * n)
    report("assign_dense", timed(lambda: capi.assign_dense(ctx, a, b, n, 2.0, capi.MASK_WRITE_TO_ONE)), 8 * n)
    report("assign_sparse", timed(lambda: capi.assign_sparse(ctx, lst, b, 3.0)), 12 * k)
    capi.check(capi.lib.glb_buffer_fill_f32(ctx.handle, b.ptr, 5.0, n))

    def relax():   # every entry relaxes (inout = 5 > 1): worst case, all writes + full new frontier
        capi.check(capi.lib.glb_buffer_fill_f32(ctx.handle, b.ptr, 5.0, n))
        capi.assign_sparse_relax(ctx, lst, b, out)
    t_fill = timed(lambda: capi.check(capi.lib.glb_buffer_fill_f32(ctx.handle, b.ptr, 5.0, n)))
    report("fill_f32", t_fill, 4 * n)
    report("assign_sparse_relax", timed(relax) - t_fill, 24 * k)
    print(json.dumps({"len": n, "list_nnz": k, "hbm_peak_gbs": peak, "ops": res}), flush=True)


if __name__ == "__main__":
    main()
