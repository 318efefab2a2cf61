"""Aggregate host<->device copy bandwidth with every rank copying at once (torchrun): the floor of the N-GPU e2e leg.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe_mp.py

Each rank copies a pinned 512 MiB buffer to its GPU and 160 MiB back (the per-rank sizes of bench.py's e2e step at
assembly1024 on 8 GPUs), first alone (rank by rank), then all ranks together; prints GB/s per rank and in total."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
up_h = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
up_d = torch.empty_like(up_h, device="cuda")
dn_d = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
dn_h = torch.empty(160 << 20, dtype=torch.uint8).pin_memory()


def once():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    up_d.copy_(up_h, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dn_h.copy_(dn_d, non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


once()
alone = [None] * world
for r in range(world):
    dist.barrier()
    if r == rank:
        ts = [once() for _ in range(3)]
        alone[r] = (min(t[0] for t in ts), min(t[1] for t in ts))
dist.barrier()
ts = []
for _ in range(5):
    dist.barrier()
    ts.append(once())
tog = (sorted(t[0] for t in ts)[2], sorted(t[1] for t in ts)[2])
mine = {"rank": rank, "alone_h2d_gbs": up_h.numel() / alone[rank][0] / 1e9, "alone_d2h_gbs": dn_h.numel() / alone[rank][1] / 1e9,
        "together_h2d_gbs": up_h.numel() / tog[0] / 1e9, "together_d2h_gbs": dn_h.numel() / tog[1] / 1e9,
        "together_h2d_ms": tog[0] * 1e3, "together_d2h_ms": tog[1] * 1e3}
allr = [None] * world
dist.all_gather_object(allr, mine)
if rank == 0:
    print(json.dumps({"world": world, "per_rank": allr, "together_total_h2d_gbs": sum(r["together_h2d_gbs"] for r in allr),
                      "together_total_d2h_gbs": sum(r["together_d2h_gbs"] for r in allr),
                      "floor_ms_of_the_e2e_copies": max(r["together_h2d_ms"] + r["together_d2h_ms"] for r in allr)}))
dist.destroy_process_group()
