"""BASELINE config[4] / config[3] at full size on N GPUs: the z-slab path with the peer-memory exchange, checked through
size-independent properties (the CPU oracle cannot hold these grids):
  (1) every rank numbers the same site set (count and a checksum of the site table agree across ranks);
  (2) 4d2 is exactly the squared distance to the reported site on whole planes of every slab;
  (3) at random vertices of every slab no site is closer and ties report the lowest id (brute force over ALL sites);
  (4) measures are 0 at outside anchors; cube >= faces >= edges where the cube cell is valid.
    torchrun --nproc-per-node 8 tools/stress_check.py stress 2048
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxel_ma_b200 import api, slabs, synth  # noqa: E402


def main():
    fam, side = sys.argv[1], int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = ny = nz = side
    z0, z1 = slabs.slab_bounds(nz, world, rank)
    lo, hi = slabs.resident_planes(z0, z1, nz)
    t0 = time.time()
    planes = synth.make(fam, side, z0=lo, z1=hi)
    gen_s = time.time() - t0
    ctx = api.Context(local)
    ctx.set_grid(nx, ny, nz, z0, z1)
    ctx.upload_volume(planes, zlo=lo)
    del planes
    # first step over NCCL (sizes the receive regions), then the peer exchange
    ctx.classify_grid(fetch=False)
    n = ctx.sites_detect_local()
    keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    corners = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    ctx.sites_export_local(keys.data_ptr(), corners.data_ptr())
    ak, ac = slabs.exchange_site_records(keys[:n], corners[:n])
    torch.cuda.current_stream().synchronize()
    ctx.sites_import_global(ak.data_ptr(), ac.data_ptr(), ak.numel())
    sites_nccl = ctx.get_sites()
    del ak, ac, keys, corners
    t = torch.tensor([n], dtype=torch.int64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    peers = slabs.PeerExchange(ctx, 2 * int(t[0]) + 4096)
    times = []
    for _ in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.classify_grid(fetch=False)
        ns = peers.exchange()
        ctx.closest_and_measures()
        ctx.synchronize()
        times.append(time.perf_counter() - t0)
    sites = ctx.get_sites()
    ok = {"sites_equal_nccl": bool(np.array_equal(sites, sites_nccl))}
    chk = torch.tensor([ns, int(np.frombuffer(sites.tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 62))], dtype=torch.int64, device="cuda")
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    ok["same_sites_on_all_ranks"] = all(bool((c == chk).all()) for c in allchk)
    # the result planes stay on the device: torch views over the library's buffers (__cuda_array_interface__)
    class DevArr:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

    def view(which, shape, typestr):
        return torch.as_tensor(DevArr(ctx.device_ptr(which), shape, typestr), device=f"cuda:{local}")

    sh = (z1 - z0, ny, nx)
    ids, d2 = view(api.ARR_ID, sh, "<i4"), view(api.ARR_D2X4, sh, "<u4").view(torch.int32)
    inside, cube = view(api.ARR_INSIDE, sh, "|u1"), view(api.ARR_CUBE, sh, "<f4")
    e, f = view(api.ARR_EDGE3, (3,) + sh, "<f4"), view(api.ARR_FACE3, (3,) + sh, "<f4")
    ts = torch.from_numpy(sites).to(f"cuda:{local}")
    # (2) whole planes
    good = True
    yy, xx = torch.meshgrid(torch.arange(ny, dtype=torch.float32, device=ts.device), torch.arange(nx, dtype=torch.float32, device=ts.device),
                            indexing="ij")
    for zz in sorted({0, (z1 - z0) // 2, z1 - z0 - 1}):
        s = ts[ids[zz].long()]
        q = 4 * ((s[..., 0] - xx) ** 2 + (s[..., 1] - yy) ** 2 + (s[..., 2] - float(z0 + zz)) ** 2)
        good &= bool((q.to(torch.int32) == d2[zz]).all())
    ok["d2_consistent_with_id"] = good
    # (3) brute force at random vertices (exact in float64: half-integer sites, integer vertices)
    rng = np.random.default_rng(100 + rank)
    m = 256
    vz, vy, vx = rng.integers(0, z1 - z0, m), rng.integers(0, ny, m), rng.integers(0, nx, m)
    s64 = ts.double()
    good = True
    for k in range(m):
        dd = 4 * ((s64[:, 0] - float(vx[k])) ** 2 + (s64[:, 1] - float(vy[k])) ** 2 + (s64[:, 2] - float(z0 + vz[k])) ** 2)
        mn = dd.min()
        j = int(torch.nonzero(dd == mn)[0, 0])  # lowest id among ties
        good &= (j == int(ids[vz[k], vy[k], vx[k]])) and (int(mn) == int(d2[vz[k], vy[k], vx[k]]))
    ok["brute_force_sample"] = bool(good)
    # (4) measures, plane by plane to bound the temporaries
    zero_out, mono = True, True
    for zz in range(0, z1 - z0, max(1, (z1 - z0) // 16)):
        out = inside[zz] == 0
        zero_out &= bool((e[:, zz][:, out] == 0).all()) and bool((cube[zz][out] == 0).all())
        val = cube[zz] > 0
        mono &= bool((cube[zz][val][None, :] >= f[:, zz][:, val]).all()) and bool((f[:, zz][:, val] >= e[:, zz][:, val].amin(0)[None, :]).all())
    ok["measures_zero_outside"] = zero_out
    ok["cube_ge_faces_where_valid"] = mono
    inside_frac = float(inside.float().mean())
    res = torch.tensor([int(all(ok.values()))], dtype=torch.int64, device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    tt = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"{fam}{side}", "n_gpus": world, "sites": ns, "all_ranks_ok": bool(res[0]), "rank0_checks": ok,
                          "ms_per_step": float(tt[0]) * 1e3, "vertices_per_s": nx * ny * nz / float(tt[0]), "gen_s": round(gen_s, 1),
                          "inside_fraction_rank0": inside_frac}))
    elif not all(ok.values()):
        print(f"rank {rank}: {ok}", file=sys.stderr)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
