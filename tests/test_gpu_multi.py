"""Row-sharded runs on 2 GPUs (skipped on a 1-GPU box): the peer-mapped exchange
(``glb_spmv_exchange`` -- SpMV write-back stores rows into every rank's vectors over NVLink, CUDA IPC
between the per-GPU processes) and the apps on top of it, each against the oracle on the full
matrix.  One process per GPU; gloo carries the IPC handles."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    if kind == "multicast-pusher":   # the slice leaves through pusher CTAs beside the SpMV kernels (GLB_XCHG_MC=pusher)
        os.environ["GLB_XCHG_MC"] = "pusher"
        kind = "multicast-native"
    if kind.startswith("multicast"):   # (torch symmetric memory rendezvous runs over the default NCCL group)
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from graphlily_b200 import app, capi, datasets
    from graphlily_b200.exchange import open_exchange
    from graphlily_b200.io import CSRMatrix
    from util import assert_close_rel

    ctx = capi.Context(rank)

    def make_exchange(n_floats, n_vectors):
        xc, got = open_exchange(ctx, n_floats, rank, world, n_vectors=n_vectors, kind=kind,
                                log=lambda *a: print("[test]", *a, flush=True))
        assert xc is not None, "no peer-mapped exchange on this box"
        if kind.startswith("multicast"):   # an NVSwitch box: the requested flavour must be the one that opened
            assert got == "multicast" and xc.has_multicast(), (kind, got)
            assert bool(getattr(xc, "_keep", None)) == (kind == "multicast-torch"), "wrong flavour of the multicast exchange"
        if rank == 0:
            print(f"[test] exchange requested {kind}: got {got}, multicast mapping {xc.has_multicast()}", flush=True)
        return xc

    # ---- raw exchange: five ping-pong iterations x <- A x of each semiring ---------------------
    n = 8192
    m = datasets.powerlaw_csr(n, n, 1 << 18, seed=21, max_degree=5000, value=1.0 / 32)   # same seed on every rank
    slot = n // world
    A = capi.CsrMatrix(ctx, m, rank * slot, (rank + 1) * slot)
    xc = make_exchange(n, 3)
    rng = np.random.default_rng(5)
    for op, zero, x0 in ((0, 0.0, rng.random(n).astype(np.float32)),
                         (1, 0.0, (rng.random(n) < 0.001).astype(np.float32)),
                         (2, 255.0, np.where(rng.random(n) < 0.001, 0.0, 255.0).astype(np.float32))):
        mm = m if op != 2 else CSRMatrix(n, n, np.ones(m.nnz, np.float32), m.indices, m.indptr)
        Aop = A if op != 2 else capi.CsrMatrix(ctx, mm, rank * slot, (rank + 1) * slot)
        xc.barrier()
        xc.buffer(0).write(x0)
        ref = x0
        for it in range(5):
            xc.spmv(Aop, op, zero, capi.MASK_NONE, it % 2, (it + 1) % 2)
            ref = oracle.port.spmv(mm, op, zero, 0, ref)
        got = xc.buffer(1).read(np.float32, n)
        if op == 0:
            assert_close_rel(got, ref, 1e-4)      # five iterations of 1e-5-per-step differences
        else:
            assert got.tobytes() == ref.tobytes(), f"rank {rank} op {op}"
    # the same loop in ONE call (acquire folded into the head of the next SpMV), directly and as a
    # recorded launch sequence replayed twice (the epoch lives in device memory)
    x0 = rng.random(n).astype(np.float32)
    ref = x0
    for it in range(6):
        ref = oracle.port.spmv(m, 0, 0.0, 0, ref)
    xc.barrier()
    xc.buffer(0).write(x0)
    xc.barrier()
    xc.spmv_iterate(A, 0, 0.0, capi.MASK_NONE, 0, 1, 6)
    assert_close_rel(xc.buffer(0).read(np.float32, n), ref, 1e-4)
    g6 = ctx.record(lambda: xc.spmv_iterate(A, 0, 0.0, capi.MASK_NONE, 0, 1, 6))
    for _ in range(2):
        xc.barrier()
        xc.buffer(0).write(x0)
        xc.barrier()
        g6.launch()
        assert_close_rel(xc.buffer(0).read(np.float32, n), ref, 1e-4)
    # pipelined host-buffer batch: every rank uploads its slice of x, NVLink completes it
    xs = [capi.PinnedArray(n) for _ in range(5)]
    ys = [capi.PinnedArray(n) for _ in range(5)]
    mks = [capi.PinnedArray(n) for _ in range(5)]
    for k in range(5):
        xs[k].array[:] = np.random.default_rng(100 + k).random(n).astype(np.float32)   # same on every rank
        mks[k].array[:] = np.random.default_rng(200 + k).integers(0, 2, n).astype(np.float32)
        ys[k].array[:] = -3.0
    for mt in (capi.MASK_NONE, capi.MASK_WRITE_TO_ZERO):
        xc.spmv_host_batch(A, 0, 0.0, mt, [a.ptr for a in xs], [a.ptr for a in mks] if mt else None, [a.ptr for a in ys])
        for k in range(5):
            ref = oracle.port.spmv(m, 0, 0.0, mt, xs[k].array, mks[k].array if mt else None)
            assert_close_rel(ys[k].array[rank * slot:(rank + 1) * slot], ref[rank * slot:(rank + 1) * slot], 1e-5)
            outside = np.delete(ys[k].array, np.s_[rank * slot:(rank + 1) * slot])
            assert (outside == -3.0).all()
    # the slice gather
    v = np.full(n, -1.0, np.float32)
    v[rank * slot:(rank + 1) * slot] = rank + 1
    xc.barrier()
    xc.buffer(2).write(v)
    xc.barrier()     # every rank has placed its full vector before any slice arrives from a peer
    xc.allgather(2, rank * slot, slot)
    got = xc.buffer(2).read(np.float32, n)
    assert (got.reshape(world, slot) == np.arange(1, world + 1, dtype=np.float32)[:, None]).all()
    assert not xc.timed_out()
    dist.barrier()
    xc.close()

    # ---- the apps over the exchange -------------------------------------------------------------
    g = datasets.powerlaw_graph(4096, 60000, seed=2)
    for name in ("bfs", "pagerank", "sssp"):
        a = {"bfs": app.BFS, "pagerank": app.PageRank, "sssp": app.SSSP}[name]()
        a.set_up_runtime(None, ctx=ctx)
        a.load_and_format_matrix(*((g, 0.9) if name == "pagerank" else (g,)))
        x2 = make_exchange(a.matrix_num_rows_, 3)
        a.set_sharding(rank, world, x2)
        a.send_matrix_host_to_device()
        mat = a.csr_matrix_
        for rep in range(3):     # repeated runs reuse the vectors: the start-of-run barrier is under test
            if name == "bfs":
                got, ref = a.pull(rep, 5), oracle.port.bfs(mat, rep, 5)
            elif name == "pagerank":
                got, ref = a.pull(0.9, 3 + rep), oracle.port.pagerank(mat, 0.9, 3 + rep)
            else:
                got, ref = a.pull(rep, 4), oracle.port.sssp(mat, rep, 4, 255.0)
            if name == "pagerank":
                assert_close_rel(got, ref, 1e-5)
                continue
            assert got.tobytes() == ref.tobytes(), f"rank {rank} {name} run {rep}"
            # the push direction on row shards of the CSC (frontier exchanged as a dense vector), and the
            # direction switch, against the same oracle result
            iters = 5 if name == "bfs" else 4
            assert a.push(rep, iters).tobytes() == ref.tobytes(), f"rank {rank} {name} push run {rep}"
            for thr in (0.002, 0.2, 1.1):
                assert a.pull_push(rep, iters, thr).tobytes() == ref.tobytes(), f"rank {rank} {name} pull_push {thr} run {rep}"
        assert not x2.timed_out()
        dist.barrier()
        x2.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["peer", "multicast-native", "multicast-torch", "multicast-pusher"])
def test_exchange_and_apps_on_two_gpus(kind):
    from graphlily_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    mp.start_processes(_worker, args=(2, _free_port(), kind), nprocs=2, join=True, start_method="spawn")
