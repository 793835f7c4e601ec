"""Opening the peer-mapped exchange of a row-sharded run (host plumbing over torch.distributed).

One process per GPU.  Three ways for the ranks' slices of a vector to meet, best first:

* ``"multicast"``  blocks allocated as torch symmetric memory with a multicast mapping: a finished
                   slice is sent once with ``multimem.st`` and the NVSwitch replicates it into every
                   rank's block (``glb_xchg_adopt`` with a multicast pointer);
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
    if kind == "multicast":
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
