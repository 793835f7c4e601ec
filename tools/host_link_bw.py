"""What the HOST side of the box sustains when N GPUs copy at once: every rank moves `mb` MB pinned host -> device
and device -> pinned host concurrently (two streams), all ranks at the same time; prints per-rank and aggregate GB/s.
This is the ceiling of the end-to-end (host-buffer) SpMV path: 32 MB cross the host link per C2 step whatever N.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29561 tools/host_link_bw.py [mb]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0 / world      # a rank's slice of the 16 MB vector
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("gloo")
n = int(mb * (1 << 20))
h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, seconds=0.5):
    reps = 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            if up:
                with torch.cuda.stream(s_up):
                    d[0].copy_(h[0], non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    h[1].copy_(d[1], non_blocking=True)
        torch.cuda.synchronize()
        reps += 8
    return reps * n / (time.perf_counter() - t0) / 1e9


run(True, True, 0.1)
for label, up, down in (("H2D only", True, False), ("D2H only", False, True), ("both ways", True, True)):
    mine = run(up, down)
    rates = [mine]
    if world > 1:
        rates = [None] * world
        dist.all_gather_object(rates, mine)
    if rank == 0:
        both = 2 if (up and down) else 1
        print(f"{world} GPU(s) at once, {mb:g} MB copies, {label}: per GPU {min(rates):.1f}..{max(rates):.1f} GB/s per direction, "
              f"aggregate {sum(rates) * both:.1f} GB/s over the host side", flush=True)
if world > 1:
    dist.destroy_process_group()
