"""Host-side z-slab logic for the multi-GPU path (one process per GPU).

The grid is cut into contiguous z-slabs; every rank classifies and searches its slab plus one halo
plane (recomputed, never exchanged).  The only exchange on the data path is the boundary-sample
list: each rank detects the site corners of its own corner planes, the (key, corner) records are
all-gathered (variable length), and every rank sorts the union identically, so site ids agree
everywhere.  Works on NCCL (CUDA tensors, GPU box) and on gloo (CPU tensors, unit tests).
"""
from __future__ import annotations

import numpy as np


def slab_bounds(nz: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, near-equal split of the nz grid-vertex planes; earlier ranks take the remainder."""
    if not (0 <= rank < world) or world > nz:
        raise ValueError(f"bad slab request: nz={nz} world={world} rank={rank}")
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def balanced_bounds(weights, world: int) -> list[tuple[int, int]]:
    """Contiguous z-slabs of (nearly) equal total WEIGHT instead of equal height: weights[z] is an estimate of the work
    of vertex plane z (e.g. 1 + c * inside fraction: the measures of a plane cost more where it cuts solids).  Every slab
    gets at least one plane; returns [(z0, z1)] per rank.  Uniform weights reproduce slab_bounds up to the placement of
    the remainder.  Used by bench.py --balance."""
    w = np.asarray(weights, np.float64)
    nz = len(w)
    if world < 1 or world > nz or (w < 0).any():
        raise ValueError(f"bad request: {nz} planes, {world} slabs")
    if not w.sum() > 0:
        w = np.ones(nz)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        z = int(np.searchsorted(cum, target, side="left"))
        # the cut nearer to the target of the two around it
        if z > 0 and abs(cum[z - 1] - target) <= abs(cum[min(z, nz)] - target):
            z -= 1
        z = max(z, cuts[-1] + 1)             # at least one plane for the slab below ...
        z = min(z, nz - (world - r))         # ... and for every slab above
        cuts.append(z)
    cuts.append(nz)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def aligned_bounds(weights, world: int, align: int = 32) -> list[tuple[int, int]]:
    """balanced_bounds with the cuts snapped so that every slab below the top one has  height + 1 = 0 (mod align):
    pass X of the transform puts `align` = 32 consecutive planes of a slab into one warp, and a slab that is not the
    top one also transforms one halo plane -- 128 + 1 planes would cost a fifth group of warps for that one plane,
    127 + 1 planes cost none.  The top slab (no halo plane) takes whatever remains.  Falls back to balanced_bounds
    when the grid is too thin for aligned slabs."""
    nz = len(weights)
    base = balanced_bounds(weights, world)
    if world == 1 or nz < world * align:
        return base
    cuts = [0]
    for r in range(1, world):
        target = base[r][0]
        # heights h with (h + 1) % align == 0 around the balanced cut: the nearest that leaves room for the slabs above
        h = max(align - 1, int(round((target - cuts[-1] + 1) / align)) * align - 1)
        while cuts[-1] + h > nz - (world - r) * (align - 1) and h > align - 1:
            h -= align
        cuts.append(cuts[-1] + h)
    cuts.append(nz)
    if any(cuts[i + 1] <= cuts[i] for i in range(world)):
        return base
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def resident_planes(z0: int, z1: int, nz: int) -> tuple[int, int]:
    """Voxel planes a slab needs resident: its own planes plus one halo plane on each interior side
    (the lower one feeds the slab's first corner plane, the upper one the cells that reach up)."""
    return max(z0 - 1, 0), min(z1 + 1, nz)


def owned_corner_planes(z0: int, z1: int, nz: int) -> tuple[int, int]:
    """Corner planes whose sites this slab reports: [z0, z1), plus the top plane nz for the last slab."""
    return z0, (nz + 1 if z1 == nz else z1)


def exchange_site_records(keys, corners, group=None):
    """All-gather variable-length (key, corner) uint64 records over torch.distributed.

    keys / corners: 1-D torch int64 tensors (uint64 bit patterns) on the backend's device.  Returns
    the concatenation over ranks in rank order (any order would do: the keys are unique and the
    import sorts them)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=keys.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    pad = torch.zeros(2, cap, dtype=torch.int64, device=keys.device)
    pad[0, : keys.numel()] = keys
    pad[1, : corners.numel()] = corners
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    all_keys = torch.cat([b[0, :c] for b, c in zip(bufs, counts)])
    all_corners = torch.cat([b[1, :c] for b, c in zip(bufs, counts)])
    return all_keys.contiguous(), all_corners.contiguous()


class PeerExchange:
    """The same exchange over peer memory (csrc/vc_peer.cu): set up once per slab group, then
    `exchange()` per step = vc_sites_post_peers + vc_sites_collect_peers, no collective call and one
    host synchronisation.  torch.distributed only carries the 64-byte CUDA IPC handles here."""

    def __init__(self, ctx, cap: int, group=None):
        import torch.distributed as dist

        self.ctx = ctx
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handle = ctx.peer_create(world, rank, int(cap))
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        ctx.peer_open(handles)
        dist.barrier(group=group)  # every receive buffer is zeroed and mapped before the first post

    def exchange(self) -> int:
        self.ctx.sites_post_peers()
        return self.ctx.sites_collect_peers()

    def close(self):
        self.ctx.peer_close()


def unpack_corners(corners_u64: np.ndarray) -> np.ndarray:
    """corner record cx | cy<<21 | cz<<42  ->  int32 (n,3)."""
    c = np.asarray(corners_u64, np.uint64)
    m = np.uint64(0x1FFFFF)
    return np.stack([(c & m), ((c >> np.uint64(21)) & m), ((c >> np.uint64(42)) & m)], -1).astype(np.int32)
