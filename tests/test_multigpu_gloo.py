"""N > 1 host logic on CPU (gloo, world_size 2): row-range sharding + one allgather per iteration.

No GPU here, so the per-rank SpMV is the sequential model of the kernel schedule
(tests/layout_model.py) run over the rank's own shard layout (``glb_csr_format_host`` with a row
range); what is under test is the partition (``ModuleCollection._row_range``), that shard layouts
keep GLOBAL row ids and write only their slice, and that an in-place allgather of equal slots
rebuilds the full next-iteration vector -- the protocol ``bench.py --gpus N`` and
``app.PageRank.set_sharding`` run with NCCL on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphlily_b200 import capi, datasets
    from graphlily_b200.app import ModuleCollection
    from layout_model import run_model

    m = datasets.powerlaw_csr(2048, 2048, 40000, seed=9, max_degree=3000)   # same seed on every rank
    mc = ModuleCollection()
    mc.set_sharding(rank, world)
    rb, re = mc._row_range(m.num_rows)
    assert (re - rb) * world == m.num_rows
    L = capi.format_host(m, rb, re, tile_k=64)
    assert L["nnz"] == int(m.indptr[re]) - int(m.indptr[rb])
    x = np.full(m.num_cols, np.float32(1.0 / m.num_cols), np.float32)
    for _ in range(3):   # three PageRank-shaped iterations: x <- A x + c
        y, written = run_model(L, x, 0, 0.0, m.num_rows)
        assert written[rb:re].all() and not written[:rb].any() and not written[re:].any()
        full = torch.zeros(m.num_rows, dtype=torch.float32)
        full[rb:re] = torch.from_numpy(y[rb:re] + np.float32(1e-4))
        slots = list(full.view(world, -1).unbind(0))           # in-place allgather of equal slots
        dist.all_gather(slots, full[rb:re].clone())
        x = full.numpy().copy()
    np.save(os.path.join(out_dir, f"x{rank}.npy"), x)
    dist.destroy_process_group()


def test_row_sharded_iterations_match_oracle(tmp_path, oracle):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from graphlily_b200 import datasets
    m = datasets.powerlaw_csr(2048, 2048, 40000, seed=9, max_degree=3000)
    x = np.full(m.num_cols, np.float32(1.0 / m.num_cols), np.float32)
    for _ in range(3):
        x = oracle.port.spmv(m, 0, 0.0, 0, x) + np.float32(1e-4)
    got = [np.load(tmp_path / f"x{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])                      # every rank holds the same full vector
    assert np.allclose(got[0], x, rtol=1e-5, atol=1e-9)


def test_row_range_partition():
    sys.path.insert(0, ROOT)
    from graphlily_b200.app import ModuleCollection
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            mc = ModuleCollection()
            mc.set_sharding(rank, world)
            seen.append(mc._row_range(2_449_024))
        assert seen[0][0] == 0 and seen[-1][1] == 2_449_024
        assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    mc = ModuleCollection()
    mc.set_sharding(1, 8)
    with pytest.raises(AssertionError):
        mc._row_range(1001)


def _worker_push(rank, world, port, out_dir):
    """The row-sharded PUSH protocol (app._exchange_frontier) with the per-rank SpMSpV played by the
    oracle on the rank's row shard of the CSC: own rows listed -> own slice of a dense vector ->
    allgather -> listed again on every rank -> sparse assign on the replicated distance vector."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from graphlily_b200 import datasets
    from graphlily_b200.app import ModuleCollection
    from graphlily_b200.io import CSRMatrix

    g = datasets.powerlaw_graph(1024, 12000, seed=4)
    g.data = np.ones(g.nnz, np.float32)
    n = g.num_rows
    ip, ix, d = oracle.port.csr2csc(g)
    mc = ModuleCollection()
    mc.set_sharding(rank, world)
    rb, re = mc._row_range(n)
    # row shard of the CSC: of every column the entries whose row the rank owns (glb_csc_create_rows)
    keep = (ix >= rb) & (ix < re)
    s_ip = np.concatenate([[0], np.cumsum(np.add.reduceat(keep, ip[:-1].astype(np.int64)) * (np.diff(ip.astype(np.int64)) > 0))])
    shard = CSRMatrix(n, g.num_cols, d[keep], ix[keep], s_ip.astype(np.uint32))
    distance = np.zeros(n, np.float32)
    distance[0] = 1.0
    f_idx, f_val = np.array([0], np.uint32), np.array([1.0], np.float32)
    for it in range(1, 6):
        y = oracle.port.spmspv(shard, 1, 0.0, 1, f_idx, f_val, distance)       # or-and, WriteToZero on the distance
        assert not y[:rb].any() and not y[re:].any()                          # a shard lists only its own rows
        dense = torch.zeros(n, dtype=torch.float32)
        dense[rb:re] = torch.from_numpy(y[rb:re])
        slots = list(dense.view(world, -1).unbind(0))
        dist.all_gather(slots, dense[rb:re].clone())
        f_idx = np.nonzero(dense.numpy())[0].astype(np.uint32)
        f_val = dense.numpy()[f_idx]
        distance[f_idx] = it + 1
    np.save(os.path.join(out_dir, f"d{rank}.npy"), distance)
    dist.destroy_process_group()


def test_row_sharded_push_protocol_matches_oracle(tmp_path, oracle):
    world, port = 2, _free_port()
    mp.spawn(_worker_push, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from graphlily_b200 import datasets
    g = datasets.powerlaw_graph(1024, 12000, seed=4)
    g.data = np.ones(g.nnz, np.float32)
    ref = oracle.port.bfs(g, 0, 5)
    got = [np.load(tmp_path / f"d{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], ref)


def _fd_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graphlily_b200.exchange import _share_fd_over_unix_socket
    share = _share_fd_over_unix_socket(rank, world)
    fd = None
    if rank == 0:   # stands in for the multicast object's descriptor: any fd must arrive as the SAME open file
        fd = os.open(os.path.join(out_dir, "object"), os.O_RDWR | os.O_CREAT)
        os.write(fd, b"multicast-object")
    got = share(fd)
    assert got is not None and got >= 0
    if rank != 0:   # a descriptor of this process onto rank 0's open file (positional I/O: the ranks share one file offset)
        assert os.pread(got, 16, 0) == b"multicast-object"
        os.pwrite(got, b"+seen-by-%d" % rank, 16 + 16 * rank)
    dist.barrier()
    if rank == 0:
        data = os.pread(fd, 256, 0)
        assert data.startswith(b"multicast-object") and data.count(b"+seen-by-") == world - 1, data
    os.close(got)
    dist.destroy_process_group()


def test_descriptor_travels_from_rank0_to_every_rank(tmp_path):
    """The host's only part in the library-made multicast exchange (capi.Exchange.open_multicast): ONE file
    descriptor goes from rank 0 to the other ranks' processes (SCM_RIGHTS over an abstract Unix socket whose name
    travels over torch.distributed)."""
    world = 3
    mp.spawn(_fd_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
