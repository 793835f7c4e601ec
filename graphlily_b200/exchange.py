"""Opening the peer-mapped exchange of a row-sharded run (host plumbing over torch.distributed).

One process per GPU.  Three ways for the ranks' slices of a vector to meet, best first:

* ``"multicast"``  an NVSwitch multicast object over the ranks' blocks: a finished slice is sent once with
                   ``multimem.st`` and the switch replicates it into every rank's block.  The library creates
                   and maps the object itself (``glb_xchg_mc_open`` / ``_bind``, kind ``"multicast-native"``;
                   the host only carries one file descriptor from rank 0 to the others over a Unix socket);
                   if that is not possible, blocks allocated as torch symmetric memory are adopted
                   (``glb_xchg_adopt``, kind ``"multicast-torch"``);
* ``"peer"``       blocks exchanged with CUDA IPC: the SpMV write-back stores every row into all
                   peers' blocks over NVLink (``glb_xchg_create`` / ``_export`` / ``_connect``);
* ``"nccl"``       no exchange object: one in-place ``ncclAllGather`` after every SpMV
                   (``glb_comm_init`` / ``glb_allgather_f32``).

``open_exchange`` is a collective: every rank calls it with the same arguments and all ranks end up
with the same kind."""
import os

from . import capi


def _agree(ok, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def _all_gather_bytes(world):
    import torch.distributed as dist

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    return gather


def _share_fd_over_unix_socket(rank, world):
    """-> share_fd(fd) for capi.Exchange.open_multicast: rank 0 serves its descriptor to the other ranks of this
    node over an abstract-namespace Unix socket (SCM_RIGHTS); the socket's name travels over torch.distributed."""
    import socket
    import uuid

    import torch.distributed as dist

    def share(fd):
        name = [("\0glb-mc-" + uuid.uuid4().hex) if rank == 0 else None]
        srv = None
        if rank == 0:
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(name[0])
            srv.listen(world)
        dist.broadcast_object_list(name, src=0)
        got = fd
        try:
            if rank == 0:
                srv.settimeout(60)
                for _ in range(world - 1):
                    conn, _addr = srv.accept()
                    with conn:
                        socket.send_fds(conn, [b"f"], [fd])
            else:
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                    c.settimeout(60)
                    c.connect(name[0])
                    _msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
                    got = fds[0] if fds else -1
        except OSError:
            got = -1 if rank != 0 else fd
        finally:
            if srv is not None:
                srv.close()
        return got
    return share


def open_exchange(ctx, n_floats, rank, world, n_vectors=2, kind=None, device=None, log=None):
    """-> (capi.Exchange or None, kind).  ``kind`` None reads GLB_EXCHANGE (default "multicast",
    falling back to "peer", then "nccl" -- for "nccl" the caller sets up ``ctx.comm_init``)."""
    import torch
    import torch.distributed as dist
    kind = kind or os.environ.get("GLB_EXCHANGE", "multicast")
    device = device if device is not None else torch.device("cuda", ctx.device)
    say = log or (lambda *_: None)
    if world == 1 or kind == "nccl":
        return None, ("none" if world == 1 else "nccl")
    if kind in ("multicast", "multicast-native"):
        xc = capi.Exchange.open_multicast(ctx, n_floats, rank, world, _share_fd_over_unix_socket(rank, world),
                                          lambda ok: _agree(ok, device), n_vectors)
        if xc is not None:
            dist.barrier()
            return xc, "multicast"
        say(f"rank {rank}: the library could not create a multicast object here; trying torch symmetric memory")
        kind = "multicast-torch" if kind == "multicast" else "peer"
    if kind == "multicast-torch":
        xc, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            words = capi.Exchange.block_bytes(n_floats, n_vectors) // 4
            t = symm_mem.empty(words, dtype=torch.float32, device=device)
            t.zero_()
            hdl = symm_mem.rendezvous(t, dist.group.WORLD)
            torch.cuda.synchronize(device)
            mc = int(hdl.multicast_ptr or 0)
            if mc:
                xc = capi.Exchange.adopt(ctx, n_floats, rank, world, [int(p) for p in hdl.buffer_ptrs], mc, n_vectors,
                                         keep=(t, hdl))
            else:
                err = "no multicast mapping on this system"
        except Exception as e:  # noqa: BLE001 -- any failure of the optional path selects the next one
            err = f"{type(e).__name__}: {e}"
        if _agree(xc is not None, device):
            dist.barrier()
            return xc, "multicast"
        if xc is not None:
            xc.close()
        say(f"rank {rank}: multicast exchange unavailable ({err}); trying CUDA IPC peer mapping")
        kind = "peer"
    xc, err = None, None
    try:
        xc = capi.Exchange(ctx, n_floats, rank, world, _all_gather_bytes(world), n_vectors=n_vectors)
    except capi.GlbError as e:
        err = e
    if _agree(xc is not None, device):
        return xc, "peer"
    if xc is not None:
        xc.close()
    say(f"rank {rank}: peer exchange unavailable ({err}); falling back to the NCCL allgather")
    return None, "nccl"
