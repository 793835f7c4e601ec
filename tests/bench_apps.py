#!/usr/bin/env python
"""App-level benchmarks on the BASELINE.json configs (bench_bfs / bench_pagerank / bench_sssp).

    python tests/bench_apps.py [bfs] [pagerank] [sssp] [--scale S] [--no-check]
    python -m torch.distributed.run --nproc-per-node N ... tests/bench_apps.py pagerank sssp

Shapes (the real datasets are not shipped, so seeded synthetic graphs of their shape):
  bfs       C3 gplus-shaped      107 648 vertices, ~13 M nnz, or-and semiring, 7 iterations (run_bfs.sh:20)
  pagerank  C4 ogbn-products-shaped 2 449 024 vertices, ~124 M nnz, fp32 plus-times, d = 0.9, 10 iterations
  sssp      C5 orkut-shaped     3 072 512 vertices, ~117 M nnz + diagonal, min-plus, 6 iterations
GTEPS follow the reference's definitions: BFS / SSSP nnz * iterations / t (bench_bfs.cpp:68-71),
PageRank nnz / t_iteration (bench_pagerank.cpp:59-65).  Timing: CUDA events around the whole app
call minus nothing -- uploads of the start vectors and the final read-back are inside, as in the
reference's wall-clock loops.  With N > 1 ranks the CSR is row-range sharded and every iteration ends
with one NCCL allgather (pull direction only).  Each result is checked against the oracle (the
reference's compute_reference_results restated): bit-exact for BFS / SSSP; PageRank within 1e-5
relative of the same iteration in fp64 and no further from the reference than the reference's own
sequential-fp32 rounding error (at these sizes the reference itself is ~3e-4 off the fp64 result).
One JSON line per app on stdout (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def log(*a):
    print("[bench_apps]", *a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("apps", nargs="*", default=["bfs", "pagerank", "sssp"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graphs (debugging)")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--copy-results", action="store_true",
                    help="return fresh host copies of the result vectors (the reference's semantics) instead of the "
                         "modules' page-locked mirrors")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from graphlily_b200 import app, capi, datasets, io

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank, stream.cuda_stream)
    # N > 1: "peer" (default) = SpMV write-back stores rows into every rank's vectors over NVLink
    # (glb_spmv_exchange); GLB_EXCHANGE=nccl = one in-place ncclAllGather after every SpMV
    from graphlily_b200.exchange import open_exchange
    exchange_kind = os.environ.get("GLB_EXCHANGE", "multicast") if world > 1 else "none"
    exchange_used = exchange_kind
    if exchange_kind == "nccl":
        uid = [capi.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    def all_gather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def timed(fn, reps):
        fn()   # warm-up (also the result that is checked)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # the iteration loop alone (device events around the app's launch sequence of the last call)
        loop_ms = loop_ev[0].elapsed_time(loop_ev[1]) if loop_ev[2] else float("nan")
        if world > 1:
            t = torch.tensor([ms, loop_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, loop_ms = float(t[0].item()), float(t[1].item())
        return ms, out, loop_ms

    loop_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), False]

    def instrument(a):
        """Bracket the app's iteration loop (ModuleCollection._replay) with events on the stream."""
        inner = a._replay

        def bracketed(key, launches):
            loop_ev[0].record(stream)
            inner(key, launches)
            loop_ev[1].record(stream)
            loop_ev[2] = True
        a._replay = bracketed

    def emit(name, cfg, nnz, iters, modes, check):
        if rank == 0:
            print(json.dumps({"app": name, "n_gpus": world, "config": cfg, "nnz": nnz, "iterations": iters, "modes": modes,
                              "matches_oracle": check}), flush=True)

    def pad128(n):
        return (n + 127) // 128 * 128

    s = args.scale
    for name in args.apps:
        t0 = time.time()
        if name == "bfs":
            n, nnz_target, iters = pad128(int(107_648 * s)), int(13_000_000 * s), 7
            g = datasets.powerlaw_graph(n, nnz_target, seed=3, device=dev)
            a = app.BFS()
        elif name == "pagerank":
            n, nnz_target, iters = pad128(int(2_449_024 * s)), int(124_000_000 * s), 10
            g = datasets.powerlaw_graph(n, nnz_target, seed=4, device=dev)
            a = app.PageRank()
        else:
            n, nnz_target, iters = pad128(int(3_072_512 * s)), int(117_000_000 * s), 6
            g = datasets.powerlaw_graph(n, nnz_target, seed=5, diagonal=True, device=dev)
            a = app.SSSP()
        a.set_up_runtime(None, ctx=ctx)
        a.set_pinned_results(not args.copy_results)
        instrument(a)
        if name == "pagerank":
            a.load_and_format_matrix(g, 0.9)
        else:
            a.load_and_format_matrix(g)
        xc = None
        if exchange_kind != "nccl" and world > 1:
            xc, got_kind = open_exchange(ctx, a.matrix_num_rows_, rank, world, n_vectors=3, kind=exchange_kind, device=dev, log=log)
            assert xc is not None, "exchange unavailable: rerun with GLB_EXCHANGE=nccl"
            exchange_used = got_kind
        a.set_sharding(rank, world, xc)
        a.send_matrix_host_to_device()
        nnz = a.get_nnz()
        log(f"{name}: {n} vertices, nnz {nnz}, generated + formatted + uploaded in {time.time() - t0:.1f}s")
        cfg = {"vertices": n, "nnz": nnz, "generator": "graphlily_b200.datasets.powerlaw_graph (symmetric, Zipf 0.9)",
               "sharding": "none" if world == 1 else f"row-range x{world}, " + (
                   {"peer": "rows stored into every rank's vector by the SpMV write-back over NVLink (peer-mapped memory)",
                    "multicast": "y slice sent once with multimem.st, replicated by the NVSwitch multicast",
                    "nccl": "NCCL allgather per iteration"}[exchange_used])}
        source = 0
        modes = {}
        results = {}
        if name == "pagerank":
            ms, out, loop_ms = timed(lambda: a.pull(0.9, iters), args.reps)
            modes["pull"] = {"ms_total": ms, "ms_per_iteration": ms / iters, "iterations_per_sec": iters / ms * 1e3,
                             "gteps": nnz / (ms / iters * 1e-3) / 1e9, "loop_only_ms_per_iteration": loop_ms / iters,
                             "loop_only_gteps": nnz / (loop_ms / iters * 1e-3) / 1e9}
            results["pull"] = out.copy()
        else:
            run_modes = [("pull", lambda: a.pull(source, iters)),
                         ("pull_push", lambda: a.pull_push(source, iters, 0.001 if name == "bfs" else 0.05)),
                         ("push", lambda: a.push(source, iters))]
            for mode, fn in run_modes:
                loop_ev[2] = False
                ms, out, loop_ms = timed(fn, args.reps)
                modes[mode] = {"ms_total": ms, "ms_per_iteration": ms / iters, "iterations_per_sec": iters / ms * 1e3,
                               "gteps": nnz * iters / (ms * 1e-3) / 1e9}
                if mode == "pull":
                    modes[mode].update(loop_only_ms_per_iteration=loop_ms / iters,
                                       loop_only_iterations_per_sec=iters / loop_ms * 1e3,
                                       loop_only_gteps=nnz * iters / (loop_ms * 1e-3) / 1e9)
                if mode == "pull_push":
                    modes[mode]["push_iterations"] = a.push_iterations_
                results[mode] = out.copy()   # the mirror is overwritten by the next mode
        # where the time goes: kernels only (events inside the C ABI around every SpMV main / fix-up
        # launch, this rank), next to the whole-call figure above (start vectors, exchange, read-back)
        a.use_graphs_ = False
        ctx.kernel_timing(True)
        (a.pull(0.9, iters) if name == "pagerank" else a.pull(source, iters))
        ms_main, ms_fix, launches = ctx.kernel_timing_read()
        ctx.kernel_timing(False)
        a.use_graphs_ = True
        kt = torch.tensor([ms_main + ms_fix], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(kt, op=dist.ReduceOp.MAX)
        modes["pull"]["spmv_kernels_ms_per_iteration_max_over_ranks"] = float(kt.item()) / max(launches, 1)
        check = None
        if not args.no_check and rank == 0:
            import oracle   # test infrastructure: the checker, never the thing measured
            m = a.csr_matrix_
            t1 = time.time()
            if name == "bfs":
                ref = oracle.port.bfs(m, source, iters)
            elif name == "pagerank":
                ref = oracle.port.pagerank(m, 0.9, iters)
            else:
                ref = oracle.port.sssp(m, source, iters, 255.0)
            cpu_s = time.time() - t1
            check = {}
            for mode, out in results.items():
                if name == "pagerank":
                    # fp32 sums are order dependent: beside the 1e-5 comparison with the reference's
                    # sequential order, measure both against the same iteration carried out in fp64
                    import scipy.sparse as sp
                    a64 = sp.csr_matrix((m.data.astype(np.float64), m.indices.astype(np.int64), m.indptr.astype(np.int64)),
                                        shape=(m.num_rows, m.num_cols))
                    tele = float((np.float32(1) - np.float32(0.9)) / np.float32(m.num_rows))
                    r64 = np.full(m.num_rows, float(np.float32(1.0 / m.num_rows)))
                    for _ in range(iters):
                        r64 = a64 @ r64 + tele
                    err = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-30)
                    bad = (err > 1e-5) & (np.abs(out - ref) >= 1e-12)
                    deg = np.diff(m.indptr.astype(np.int64))
                    # the gap to the reference must be explained by the reference's own rounding error:
                    # |gpu - ref| <= |ref - fp64| + 1e-5 |ref|, and the GPU itself is within 1e-5 of fp64
                    explained = np.abs(out - ref) <= np.abs(ref - r64) + 1e-5 * np.abs(ref)
                    check[mode] = bool(explained.all() and (np.abs(out - r64) <= 1e-5 * np.abs(r64)).all())
                    check["within_1e-5_of_reference_everywhere"] = bool(not bad.any())
                    check["max_rel_err_vs_reference"] = float(err.max())
                    check["rows_over_1e-5"] = int(bad.sum())
                    check["min_nnz_of_rows_over_1e-5"] = int(deg[bad].min()) if bad.any() else None
                    check["max_rel_err_gpu_vs_fp64"] = float((np.abs(out - r64) / np.abs(r64)).max())
                    check["max_rel_err_reference_vs_fp64"] = float((np.abs(ref - r64) / np.abs(r64)).max())
                else:
                    check[mode] = bool(out.tobytes() == ref.tobytes())
            check["oracle_cpu_seconds"] = cpu_s
            if name != "pagerank":
                check["reached_vertices"] = int((ref != (0 if name == "bfs" else 255)).sum())
        emit(name, cfg, nnz, iters, modes, check)
        del a
        if xc is not None:
            torch.cuda.synchronize()
            dist.barrier()
            xc.close()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
