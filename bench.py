#!/usr/bin/env python
"""bench.py -- SpMV GTEPS on the BASELINE.json configs[1] workload (bench_spmv), plus one record per
app (bench_bfs / bench_pagerank / bench_sssp on the C3 / C4 / C5 shapes) in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-apps]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload: synthetic power-law CSR, 4 194 304 x 4 194 304, 134 217 728 nnz
(32 / row), fp32 plus-times -- the shape of the reference's benchmark/bench_spmv.cpp:37-113
(GTEPS = nnz / seconds / 1e9, :106-112), with A[:] = 1/N and x in {0, 1} as in the reference's own
SpMV test (tests/test_module_spmv_spmspv.cpp:104-112).  The steps alternate between two fixed input
vectors x0 / x1 (outputs y0 / y1), so that the LAST timed step can be checked: round 1 fed y back as
x, which underflows to exactly 0 after nine steps with A = 1/N.  (With these inputs every fp32 sum is
exact -- all products equal 2^-22 -- so the 2^20-long rows of this matrix, on which the reference's
sequential fp32 sum would otherwise be 1e-3 off, compare at 1e-5 like every other row.)

One step = one SpMV over the whole matrix.  With N > 1 the CSR is row-range sharded (cuts balanced
by nnz), every rank computes its slice of y and the slices meet on every rank before the next step
(NVSwitch multicast / peer stores / one NCCL allgather): strong scaling, total work fixed.  The K
timed steps are ONE recorded launch sequence (CUDA graph) per rank, replayed once.

Reported on one JSON line by rank 0:
  value        device-timed GTEPS, inputs resident in HBM (matrix 1.07 GB >> 126 MB L2, so every
               step streams it from HBM; the 16 MB x is meant to live in L2)
  parity       the last timed step checked on EVERY rank: rows of the rank's shard against the
               reference's compute_reference_results, the checksum identity sum(y) = sum(colsum * x) over
               the whole exchanged vector, bit-identical vectors across ranks; rc 3 on mismatch
  e2e          GTEPS through glb_spmv_host_batch with pinned HOST x / y: H2D x, kernels, D2H y per step
  roofline     spmv_lane_kernel alone: algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline the reference's own compute_reference_results (oracle/_ref) on one host thread,
               on a bounded row sample of the same matrix (N = 1 only)
  bfs / pagerank / sssp   tests/bench_apps.run_app records (levels/s, GTEPS, parity vs the reference)
`--impl reference` runs only the CPU path (the reference has no threading: 1 thread).
"""
import argparse
import json
import os
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = 4_194_304
NNZ = 134_217_728
SEED = 42


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML while the measurement runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            log("NVML unavailable:", e)
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_matrix(device):
    from graphlily_b200 import datasets
    t0 = time.time()
    m = datasets.powerlaw_csr(ROWS, ROWS, NNZ, seed=SEED, device=device)       # A[:] = 1/N
    log(f"generated power-law CSR {m.num_rows} x {m.num_cols}, nnz {m.nnz} on {device} in {time.time() - t0:.1f}s")
    return m


def row_sample(m, begin, rows):
    from graphlily_b200.io import CSRMatrix
    ip = m.indptr.astype(np.int64)
    s, e = int(ip[begin]), int(ip[begin + rows])
    return CSRMatrix(rows, m.num_cols, m.data[s:e], m.indices[s:e], (ip[begin:begin + rows + 1] - s).astype(np.uint32))


def cpu_reference_backend():
    import oracle  # the ONLY use of oracle/ here: as the timed CPU baseline / reference arm and as the parity checker
    if oracle.ref is not None:
        return oracle.ref, "reference"
    return oracle.port, "port"


def workload_config(n_gpus):
    return {"workload": "bench_spmv: synthetic power-law CSR 4194304 x 4194304, 134217728 nnz (32/row), "
                        "fp32 plus-times SpMV, A=1/N, x in {0,1}; steps alternate two fixed input vectors (y0 = A x0, y1 = A x1, ...)",
            "generator": f"graphlily_b200.datasets.powerlaw_csr(seed={SEED}): Pareto(2.1) row degrees 1..2^20, "
                         "Zipf(0.9) column popularity, random column labels",
            "rows": ROWS, "nnz": NNZ, "semiring": "plus-times",
            "sharding": f"row ranges x{n_gpus} (balanced by nnz over the exchange, equal slots under NCCL)" if n_gpus > 1 else "none",
            "l2": "matrix streams (1.07 GB) exceed the 126 MB L2 every step; no flush needed",
            "layout": "lane-segment chunks (<=1024 nnz / warp), hot columns packed into an L1-resident vector"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU SpMV on a bounded row sample, 1 thread."""
    if rank != 0:
        return
    import torch
    backend, kind = cpu_reference_backend()
    m = make_matrix("cuda" if torch.cuda.is_available() else "cpu")
    x = np.random.default_rng(SEED).integers(0, 2, m.num_cols).astype(np.float32)
    probe = row_sample(m, 0, min(65536, m.num_rows))
    t_probe, _ = backend.spmv_timed(probe, 0, 0.0, x, reps=2)
    per_nnz = t_probe / max(probe.nnz, 1)
    budget = 60.0 / max(args.steps + args.warmup, 1)               # seconds per step
    rows = min(65536, m.num_rows)
    while rows * 2 <= min(1_048_576, m.num_rows) and per_nnz * int(m.indptr[rows * 2]) * 2.0 < budget:
        rows *= 2
    s = row_sample(m, 0, rows)
    for _ in range(args.warmup):
        backend.spmv_timed(s, 0, 0.0, x, reps=1)
    # each step times one compute_reference_results(vector) call (spmv_module.h:488-510); building the
    # module object around the borrowed arrays is setup, as the matrix upload is on the GPU arm
    dt = sum(backend.spmv_timed(s, 0, 0.0, x, reps=1)[0] for _ in range(args.steps)) / args.steps
    gteps = s.nnz / dt / 1e9
    sample = f"rows [0, {rows}) of the same matrix ({s.nnz} nnz) per step, 1 thread (the reference has no threading)"
    line = {"impl": "reference", "metric": "spmv_gteps", "value": gteps, "unit": "GTEPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": gteps, "unit": "GTEPS", "cores": 1, "kind": kind, "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": gteps, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-apps", action="store_true", help="skip the BFS / PageRank / SSSP records")
    ap.add_argument("--app-reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=ROWS, help=argparse.SUPPRESS)   # debugging only
    ap.add_argument("--app-scale", type=float, default=1.0, help=argparse.SUPPRESS)   # debugging only
    args = ap.parse_args()
    if args.rows != ROWS:   # scaled-down debugging run, never a bench line of record
        globals().update(ROWS=args.rows, NNZ=args.rows * 32)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from graphlily_b200 import capi  # fails loudly if the CUDA library is missing

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: graphlily_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    m = make_matrix(dev)
    n = m.num_rows
    assert n % (32 * world) == 0
    slot = n // world
    rb, re = rank * slot, (rank + 1) * slot

    # a real (non-default) stream shared by torch and the C ABI, so torch.cuda.Event brackets our kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = capi.Context(local_rank, stream.cuda_stream)
    # N > 1: how each rank's slice of y reaches the other ranks before the next step.
    #   "multicast" (default) symmetric-memory blocks with a multicast mapping: the finished slice is
    #           sent once with multimem.st, the NVSwitch replicates it to every rank
    #   "peer"  CUDA-IPC blocks: the SpMV write-back stores every row into all ranks' copies of the
    #           vector over NVLink; a 32-thread kernel publishes the epoch
    #   "nccl"  one in-place ncclAllGather after the kernels
    # (GLB_EXCHANGE selects; each falls back to the next when the system lacks it)
    from graphlily_b200.exchange import open_exchange
    xc, exchange = open_exchange(ctx, n, rank, world, n_vectors=4, device=dev, log=log)
    if world > 1 and (exchange == "nccl" or os.environ.get("GLB_EXCHANGE") == "nccl"):
        uid = [capi.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    if xc is not None:
        # the exchange needs no equal slots: cut the rows where the nnz prefix crosses r / N (32-row aligned)
        ip = np.asarray(m.indptr, dtype=np.int64)
        cuts = [int(np.searchsorted(ip, ip[-1] * r // world)) // 32 * 32 for r in range(world + 1)]
        cuts[0], cuts[-1] = 0, n
        rb, re = cuts[rank], cuts[rank + 1]
    t0 = time.time()
    A = capi.CsrMatrix(ctx, m, rb, re)
    info = A.info()
    log(f"rank {rank}: rows [{rb},{re}) nnz {info['nnz']} chunks {info['chunks']} fixups {info['fixups']} "
        f"layout {info['device_bytes'] / 1e9:.3f} GB, format+upload {time.time() - t0:.1f}s")

    x_host = np.random.default_rng(SEED).integers(0, 2, n).astype(np.float32)

    class Vec:   # a full-length device vector: a torch tensor, or a peer-mapped exchange vector
        def __init__(self, which):
            self.which = which
            self.t = None if xc is not None else torch.zeros(n, dtype=torch.float32, device=dev)

        def ptr(self):
            return xc.vector(self.which) if xc is not None else self.t.data_ptr()

        def load(self, host):
            capi.check(capi.lib.glb_buffer_h2d(ctx.handle, self.ptr(), host.ctypes.data, host.nbytes))

        def read(self):
            out = np.empty(n, np.float32)
            capi.check(capi.lib.glb_buffer_d2h(ctx.handle, out.ctypes.data, self.ptr(), out.nbytes))
            return out

    vecs = [Vec(0), Vec(1), Vec(2), Vec(3)]    # x0, x1 (inputs), y0, y1 (outputs)
    x_hosts = [x_host, np.random.default_rng(SEED + 7).integers(0, 2, n).astype(np.float32)]

    def enqueue_steps(k):
        """k steps: step j computes y[j & 1] = A x[j & 1]; over an exchange / allgather every rank ends with
        the complete y[j & 1], and the next step's first kernel opens with the acquire of that exchange."""
        if xc is not None:
            xc.spmv_iterate(A, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, 0, 2, k, plan=[(j & 1, 2 + (j & 1)) for j in range(k)])
        else:
            for j in range(k):
                A.spmv(capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, vecs[j & 1].ptr(), None, vecs[2 + (j & 1)].ptr())
                if world > 1:
                    ctx.allgather_f32(vecs[2 + (j & 1)].ptr(), slot)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- device-resident timing ------------------------------------------------------------
    vecs[0].load(x_hosts[0])
    vecs[1].load(x_hosts[1])
    barrier()
    enqueue_steps(args.warmup)                              # W warm-up steps, launch by launch
    recordable = world == 1 or xc is not None               # (the NCCL allgather is issued launch by launch)
    graph = ctx.record(lambda: enqueue_steps(args.steps)) if recordable else None
    if graph is not None:
        graph.launch()                                      # first replay (uploads the sequence): K more warm-up steps
    for v in vecs[2:]:                                      # the timed steps must produce the outputs that are checked
        v.load(np.zeros(n, np.float32))
    barrier()
    sampler.start()
    e0.record(stream)
    if graph is not None:
        graph.launch()
    else:
        enqueue_steps(args.steps)
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    gteps = m.nnz / (ms_step * 1e-3) / 1e9
    last = (args.steps - 1) & 1
    x_in, y_out = x_hosts[last], vecs[2 + last].read()     # the last timed step; y is complete on every rank
    # launch by launch from the host, for comparison (the same K steps, not recorded)
    barrier()
    e0.record(stream)
    enqueue_steps(args.steps)
    e1.record(stream)
    barrier()
    ms_step_host_launched = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    # ---- parity of the LAST step, on every rank --------------------------------------------
    backend, kind = cpu_reference_backend()
    rows_c = min(262_144, re - rb)
    s = row_sample(m, rb, rows_c)
    y_ref = backend.spmv(s, 0, 0.0, 0, x_in)
    got = y_out[rb:rb + rows_c]
    err = np.abs(got - y_ref) / np.maximum(np.abs(y_ref), 1e-30)
    rows_ok = bool(((err <= 1e-5) | (np.abs(got - y_ref) < 1e-12)).all())
    colsum = np.bincount(m.indices, weights=m.data.astype(np.float64), minlength=n)
    lhs, rhs = float(y_out.astype(np.float64).sum()), float(colsum @ x_in.astype(np.float64))
    checksum_rel = abs(lhs - rhs) / max(abs(rhs), 1e-300)
    del colsum
    sig = (zlib.crc32(vecs[last].read().tobytes()), zlib.crc32(y_out.tobytes()))
    sigs = [sig]
    if world > 1:
        sigs = [None] * world
        dist.all_gather_object(sigs, sig)
    mine = {"rows_ok": rows_ok, "max_rel_err": float(err.max()), "checksum_rel_err": checksum_rel,
            "nonzero_outputs": int((y_out != 0).sum())}
    alls = [mine]
    if world > 1:
        alls = [None] * world
        dist.all_gather_object(alls, mine)
    parity = {"ok": bool(all(a["rows_ok"] and a["checksum_rel_err"] <= 1e-5 and a["nonzero_outputs"] > n // 2 for a in alls)
                         and all(sg == sigs[0] for sg in sigs)),
              "checked": f"output of the last step: on every rank, rows [row_begin, row_begin + {rows_c}) of its shard vs "
                         f"compute_reference_results ({kind}) at 1e-5 relative; sum(y) = sum(colsum * x) over the whole "
                         "vector in float64 (covers every rank's exchanged slice); vectors bit-identical across ranks",
              "rows_checked_per_rank": rows_c, "max_rel_err_vs_reference": max(a["max_rel_err"] for a in alls),
              "checksum_rel_err_max": max(a["checksum_rel_err"] for a in alls),
              "nonzero_outputs_min": min(a["nonzero_outputs"] for a in alls),
              "ranks_agree_bitwise": bool(all(sg == sigs[0] for sg in sigs))}
    if not parity["ok"]:
        log(f"rank {rank}: PARITY FAILURE {parity} {alls}")

    # ---- dominant kernel alone (events inside the C ABI, same stream) --------------------------
    barrier()
    ctx.kernel_timing(True)
    enqueue_steps(args.steps)
    ms_main, ms_fix, launches = ctx.kernel_timing_read()
    ctx.kernel_timing(False)
    ms_kernel = ms_main / max(launches, 1)
    ms_fixup = ms_fix / max(launches, 1)
    rows_s = re - rb
    alg_bytes = 8 * info["nnz"] + 4 * (rows_s + 1) + 4 * m.num_cols + 4 * rows_s
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "spmv_traffic.json")))["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    secondary = None
    try:   # the unit that actually limits the kernel (committed ncu capture)
        secondary = json.load(open(os.path.join(ROOT, "profiles", "spmv_l1_port.json")))
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": "spmv_lane_kernel<plus-times>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_kernel, "fixup_kernel_ms": ms_fixup,
                "secondary": secondary if world == 1 else None}
    ms_kernels_max = max_over_ranks(ms_kernel + ms_fixup)
    step_breakdown = {"main_plus_fixup_ms_max_over_ranks": ms_kernels_max,
                      "other_ms": ms_step - ms_kernels_max,
                      "other_is": "hot-column pack + launch gaps" + ("" if world == 1 else " + slice exchange (push / signal) + acquire + rank skew"),
                      "host_launched_ms_per_step": ms_step_host_launched}

    # ---- end to end through the host-buffer entry points --------------------------------------
    # Every step uploads that step's x from pinned host memory and reads its y slice back into pinned
    # host memory.  Two public calls are timed: glb_spmv_host (one vector, returns when y has landed)
    # and glb_spmv_host_batch (the same per-vector work for a sequence of vectors, upload / kernels /
    # download of consecutive vectors overlapped on three streams).  The batch figure is `value`.
    ring = 4
    rng = np.random.default_rng(SEED + 1)
    xhs = [torch.from_numpy(x_host).pin_memory()] + \
          [torch.from_numpy(rng.integers(0, 2, n).astype(np.float32)).pin_memory() for _ in range(ring - 1)]
    yhs = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(ring)]
    e2e_steps = max(12, min(args.steps, 48)) // ring * ring
    xs = [xhs[i % ring].data_ptr() for i in range(e2e_steps)]
    ys = [yhs[i % ring].data_ptr() for i in range(e2e_steps)]

    def timed_host(fn):
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        fn()
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        return max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / e2e_steps

    def sync_calls():
        for i in range(e2e_steps):
            A.spmv_host(capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, xs[i], None, ys[i])

    for i in range(3):
        A.spmv_host(capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, xs[i], None, ys[i])
    sync_ms = timed_host(sync_calls)
    y_sync = [y[rb:re].clone() for y in yhs]
    for y in yhs:
        y.zero_()

    def batch_call(xv, yv):
        if xc is not None:   # sharded: every rank uploads its 1/N slice of x, NVLink completes it
            xc.spmv_host_batch(A, capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, xv, None, yv)
        else:
            A.spmv_host_batch(capi.OP_MUL_ADD, 0.0, capi.MASK_NONE, xv, None, yv)

    batch_call(xs[:ring], ys[:ring])
    batch_ms = timed_host(lambda: batch_call(xs, ys))
    batch_same = all(bool(torch.equal(yhs[i][rb:re], y_sync[i])) for i in range(ring))
    # the e2e results against the reference, on the same row sample of every rank's shard
    e2e_ok = True
    for i in (0, ring - 1):
        ref_i = backend.spmv(s, 0, 0.0, 0, xhs[i].numpy())
        got_i = yhs[i][rb:rb + rows_c].numpy()
        e_i = np.abs(got_i - ref_i) / np.maximum(np.abs(ref_i), 1e-30)
        e2e_ok = e2e_ok and bool(((e_i <= 1e-5) | (np.abs(got_i - ref_i) < 1e-12)).all())
    e2e_flags = [(batch_same, e2e_ok)]
    if world > 1:
        e2e_flags = [None] * world
        dist.all_gather_object(e2e_flags, (batch_same, e2e_ok))
    batch_same, e2e_ok = all(f[0] for f in e2e_flags), all(f[1] for f in e2e_flags)
    parity["e2e_matches_reference_on_sample"] = e2e_ok
    parity["ok"] = bool(parity["ok"] and e2e_ok and batch_same)
    sliced = xc is not None
    e2e = {"value": m.nnz / (batch_ms * 1e-3) / 1e9, "unit": "GTEPS",
           "h2d_bytes_per_step": 4 * n if sliced or world == 1 else 4 * n * world,     # whole job, all ranks
           "d2h_bytes_per_step": 4 * n, "ms_per_step": batch_ms, "steps": e2e_steps,
           "api": (f"glb_spmv_host_batch_exchange: {e2e_steps} vectors from a ring of {ring} pinned host x buffers; every rank "
                   "uploads its 1/N slice of x, the slices meet over NVLink, SpMV, y slice -> pinned host; "
                   "upload / kernels / download of consecutive vectors overlap (the copies of the neighbouring vectors start "
                   "once the slice push of the current one has left)" if sliced else
                   f"glb_spmv_host_batch: {e2e_steps} vectors from a ring of {ring} pinned host x buffers -> device, SpMV, "
                   "y slice -> pinned host y buffers; upload / kernels / download of consecutive vectors overlap"),
           "batch_matches_single_call_bitwise": batch_same,
           "single_call": {"value": m.nnz / (sync_ms * 1e-3) / 1e9, "ms_per_step": sync_ms,
                           "api": "glb_spmv_host per vector (returns when y has landed; no overlap between vectors)"}}
    clocks = sampler.stop()

    # ---- CPU baseline beside it (rank 0, N = 1) ----------------------------------------------
    cpu = None
    if world == 1:
        rows_b = min(1_048_576, m.num_rows)
        sb = row_sample(m, 0, rows_b)
        sec, _ = backend.spmv_timed(sb, 0, 0.0, x_host, reps=3)
        cpu = {"value": sb.nnz / sec / 1e9, "unit": "GTEPS", "cores": 1, "kind": kind,
               "sample": f"rows [0, {rows_b}) of the same matrix ({sb.nnz} nnz), best of 3, 1 thread "
                         "(the reference has no threading)",
               "host_cpus": os.cpu_count()}

    # kernels of ours per step: gather_hot (when hot columns are packed), spmv_lane, spmv_fixup, + the exchange's
    # push / signal kernel; one acquire kernel closes the recorded sequence
    hot = 1 if 0 < info["tile_k"] < m.num_cols else 0
    launches_per_step = 2 + hot + (0 if xc is None else 1)
    gpu_launches = launches_per_step * args.steps + (1 if xc is not None else 0)

    # ---- release the C2 matrix, then the app records ----------------------------------------
    A.close()
    del A, m, vecs, xhs, yhs, y_sync
    torch.cuda.empty_cache()
    apps = {}
    if not args.no_apps:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import bench_apps
        env = bench_apps.Env(ctx, stream, dev, rank, world, log)
        for name in bench_apps.APPS:
            try:
                apps[name] = bench_apps.run_app(name, env, scale=args.app_scale, reps=args.app_reps, check=True,
                                                hbm_peak_gbs=peak)
            except Exception as exc:  # noqa: BLE001 -- an app failure must not lose the SpMV line
                log(f"app {name} failed: {type(exc).__name__}: {exc}")
                apps[name] = {"app": name, "error": f"{type(exc).__name__}: {exc}", "parity": {"ok": False}}
            if rank == 0 and apps[name].get("parity") and not apps[name]["parity"]["ok"]:
                parity["ok"] = False

    if rank == 0:
        cfg = workload_config(world)
        if world > 1:
            cfg["exchange"] = {"multicast": "each rank's finished y slice sent once by a kernel of 16-byte multimem.st stores whose last CTA "
                                            "publishes the epoch; replicated to all ranks by the NVSwitch; the acquire sits at the head of "
                                            "the next step's first kernel",
                               "peer": "y rows stored into every rank's vector by the SpMV write-back over NVLink (peer-mapped memory), "
                                       "a 32-thread kernel publishes the epoch, acquire at the head of the next step's first kernel",
                               "nccl": "one in-place ncclAllGather of y per step"}[exchange]
        cfg["timed_region"] = ("one recorded launch sequence (CUDA graph) of K steps, replayed once" if graph is not None
                               else "K steps launched one by one")
        line = {"metric": "spmv_gteps", "value": gteps, "unit": "GTEPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
                "step_breakdown": step_breakdown, "parity": parity, "cpu_baseline": cpu, "nnz": NNZ}
        line.update(apps)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if not parity["ok"]:
        raise SystemExit(3)


if __name__ == "__main__":
    main()
