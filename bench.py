#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: grid vertices/s through classify + closest-site + measures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--grid NX,NY,NZ]

One "step" = one pass of the hot path (classify -> boundary samples -> closest site per grid vertex
-> cell measures) over the synthetic volume, which is already resident in HBM when the timed region
starts.  N=1 runs BASELINE config[2] (twist512, the configuration the metric is quoted on); N>1
keeps 512^3 vertices per GPU (weak scaling): 512x512x1024, 512x1024x1024, 1024^3 (= config[3],
assembly1024) cut into z-slabs, one process per GPU, the site records all-gathered over NCCL.
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU operators instead
(oracle/_ref, bounded sample) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid vertices/sec (classify+closest-pt+measure)"
UNIT = "vertices/s"
# algorithmic bytes per grid vertex, SURVEY.md section 8(d) (stated again in DESIGN.md)
BYTES_CLASSIFY, BYTES_CLOSEST, BYTES_MEASURES = 5, 8, 32
BYTES_PIPELINE = BYTES_CLASSIFY + BYTES_CLOSEST + BYTES_MEASURES  # 45
# bytes one launch of a kernel has to move by its own contract, per unit it processes
WEAK_GRIDS = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (512, 1024, 1024), 8: (1024, 1024, 1024)}


def workload_for(n_gpus, args):
    if args.grid:
        g = tuple(int(v) for v in args.grid.split(","))
        name = args.workload or "assembly"
        return name, g
    if args.workload:
        fam = args.workload.rstrip("0123456789")
        side = int(args.workload[len(fam):] or 512)
        return fam, (side, side, side)
    if n_gpus == 1:
        return "twist", WEAK_GRIDS[1]
    return "assembly", WEAK_GRIDS.get(n_gpus, (512, 512, 512 * n_gpus))


def make_planes(fam, grid, z0, z1):
    from voxel_ma_b200 import synth
    nx, ny, nz = grid
    if fam in ("assembly", "stress"):
        return synth.make(fam, grid if not (nx == ny == nz) else nx, z0=z0, z1=z1)
    assert nx == ny == nz, "only the assembly family has non-cubic grids"
    return synth.make(fam, nx, z0=z0, z1=z1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.split(",") for l in open(self.f.name).read().strip().splitlines() if l.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].strip().replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in rows)]
        out.update(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                   samples=len(rows))
        return out


class quiet_stdout:
    """The reference prints progress to stdout (cout/printf); bench.py must print ONE JSON line, so
    fd 1 is pointed at stderr while the CPU legs run."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ------------------------------------------------------------------------------------------------
# the reference's CPU operators on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def _ann_worker(args):
    sites, q = args
    from oracle import bindings as ob
    ob.ref_ann(sites, q, brute=False)
    return ob.ref().ref_last_seconds()


def _sites_worker(block):
    from oracle import bindings as ob
    ob.ref_extract_sites(block)
    return ob.ref().ref_last_seconds()


def cpu_reference_rate(fam, grid, budget_s=20.0, cores=None):
    """vertices/s of the reference's CPU operators for this path, all host cores, bounded sample.

    classify + boundary samples: Surfacer::extractBoundaryVts (which calls voxTaggedAsInside 7x per
    voxel) on one sub-block per core; closest: ANNkd_tree build + annkSearch(k=1,eps=0) over the FULL
    site set, one forked process per core (ANN keeps search state in globals), a random sample of grid
    vertices as queries; measures: the lambdaForFace dictionary (C port) on a slab.  The three
    per-vertex costs add up exactly as they would in a whole-grid run."""
    import multiprocessing as mp

    from oracle import bindings as ob
    kind = "reference" if ob.have_ref() else "port"
    cores = cores or os.cpu_count() or 1
    nx, ny, nz = grid
    nvert = nx * ny * nz
    # a slab through the middle of the object, thick enough to hold the sub-blocks
    zs = min(nz, 64)
    zmid = nz // 2
    slab0 = max(0, zmid - zs // 2)
    planes = make_planes(fam, grid, slab0, slab0 + zs)
    t_total0 = time.perf_counter()
    # full site set (not timed: obtained with the C port; the ANN leg needs the true sites)
    vol_full_sites = None
    if nvert <= 512 ** 3:
        full = make_planes(fam, grid, 0, nz)
        inside_full = ob.classify_grid(full)
        vol_full_sites = ob.extract_sites(inside_full)
        del full
    else:  # very large grids: sites of the sampled slab's neighbourhood only (stated in `sample`)
        inside_full = ob.classify_grid(planes)
        vol_full_sites = ob.extract_sites(inside_full)
        vol_full_sites[:, 2] += slab0
    nsites = len(vol_full_sites)
    ctx = mp.get_context("fork")
    # --- stage A: site extraction rate (voxels/s), one sub-block per core
    bs = min(ny, 128)
    blocks = []
    for k in range(cores):
        y0 = (k * bs) % max(ny - bs + 1, 1)
        x0 = ((k * bs) // max(ny - bs + 1, 1) * bs) % max(nx - bs + 1, 1)
        blocks.append(np.ascontiguousarray(planes[:, y0:y0 + bs, x0:x0 + bs]))
    t0 = time.perf_counter()
    if kind == "reference":
        with ctx.Pool(cores) as pool:
            secs = pool.map(_sites_worker, blocks)
        wall_a = max(secs)
    else:
        for b in blocks:
            ob.extract_sites(ob.classify_grid(b))
        wall_a = (time.perf_counter() - t0) / cores
    rate_a = sum(b.size for b in blocks) / max(wall_a, 1e-9)
    # --- stage B: ANN 1-NN per grid vertex; size the sample from a short probe
    rng = np.random.default_rng(7)
    def queries(m):
        return np.stack([rng.integers(0, nx, m), rng.integers(0, ny, m), rng.integers(0, nz, m)], -1).astype(np.float64)
    s64 = vol_full_sites.astype(np.float64)
    if kind == "reference" and nsites > 0:
        probe = 4000
        t_probe = _ann_worker((s64, queries(probe)))
        per_q = max(t_probe / probe, 1e-7)
        m_per_core = int(min(max(budget_s * 0.6 / per_q, 2000), 3_000_000))
        with ctx.Pool(cores) as pool:
            secs = pool.map(_ann_worker, [(s64, queries(m_per_core)) for _ in range(cores)])
        rate_b = cores * m_per_core / max(max(secs), 1e-9)
        nq = cores * m_per_core
    else:
        m = 2000
        t0 = time.perf_counter()
        ob.closest_points(s64, queries(m)) if nsites else None
        rate_b = m / max(time.perf_counter() - t0, 1e-9)
        nq = m
    # --- stage C: measures on the slab (C port of the dictionary; OpenMP over the cores)
    ins = ob.classify_grid(planes)
    ids = np.zeros(planes.shape, np.int32)
    t0 = time.perf_counter()
    if nsites:
        ob.cell_measures_grid(vol_full_sites, ids, ins, nx, ny, planes.shape[0], 0, planes.shape[0] - 1)
    rate_c = planes[:-1].size / max(time.perf_counter() - t0, 1e-9)
    value = 1.0 / (1.0 / rate_a + 1.0 / rate_b + 1.0 / rate_c)
    sample = (f"{fam} {nx}x{ny}x{nz}, {nsites} sites; extractBoundaryVts on {cores} sub-blocks of {zs}x{bs}x{bs} voxels "
              f"({rate_a:.3g} voxels/s); ANN kd-tree build + {nq} annkSearch(k=1,eps=0) queries at random grid vertices over "
              f"the full site set, one forked process per core ({rate_b:.3g} q/s); lambda dictionary on {planes.shape[0]-1} "
              f"planes ({rate_c:.3g} v/s); per-vertex costs summed; {time.perf_counter()-t_total0:.1f}s of wall")
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fam, grid = workload_for(args.gpus, args)
    vals = []
    for _ in range(max(args.warmup, 0)):
        pass  # the CPU path has no warm-up state worth the minutes it would take
    t0 = time.perf_counter()
    cb = None
    for _ in range(max(1, min(args.steps, 3))):
        with quiet_stdout():
            cb = cpu_reference_rate(fam, grid, budget_s=12.0)
        vals.append(cb["value"])
    v = float(np.median(vals))
    cb["value"] = v
    nvert = grid[0] * grid[1] * grid[2]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": nvert / v * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{fam}{grid[0]}x{grid[1]}x{grid[2]}", "note": "reference CPU operators (Surfacer + ANN kd-tree + "
                   "lambdaForFace) on a bounded sample; ms_per_step is the whole-grid extrapolation"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from voxel_ma_b200 import api, slabs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dist = None
    torch.cuda.set_device(local)
    # stdout must carry ONE line (rank 0's JSON): anything a library prints there meanwhile (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) is sent to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fam, grid = workload_for(world, args)
    nx, ny, nz = grid
    z0, z1 = slabs.slab_bounds(nz, world, rank)
    lo, hi = slabs.resident_planes(z0, z1, nz)
    planes = make_planes(fam, grid, lo, hi)
    if world > 1 and args.balance > 0:
        # EXPERIMENTAL (off by default): slab heights from a per-plane work estimate 1 + c * inside fraction
        # (slabs.balanced_bounds) instead of equal heights; every rank counts its own planes, the counts are gathered
        frac = (planes[z0 - lo:z1 - lo] > 0).reshape(z1 - z0, -1).mean(axis=1)
        per = [None] * world
        dist.all_gather_object(per, (z0, frac.astype(np.float64)))
        full = np.zeros(nz)
        for zz, f in per:
            full[zz:zz + len(f)] = f
        z0, z1 = slabs.balanced_bounds(1.0 + args.balance * full, world)[rank]
        lo, hi = slabs.resident_planes(z0, z1, nz)
        planes = make_planes(fam, grid, lo, hi)
        print(f"[rank {rank}] balanced slab [{z0},{z1}) = {z1 - z0} planes", file=sys.stderr)
    ctx = api.Context(local)
    ctx.set_grid(nx, ny, nz, z0, z1)
    # pinned staging of the slab (the e2e leg copies from here every step)
    pin_vol = api.PinnedArray(planes.shape, np.float32)
    pin_vol.array[...] = planes
    del planes
    ctx.upload_volume(pin_vol.array, zlo=lo)
    nv_local = nx * ny * (z1 - z0)
    nv_total = nx * ny * nz
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    state = {"nsites": 0}

    def step():
        if world == 1:
            state["nsites"] = ctx.run_dense()
            return
        ctx.classify_grid(fetch=False)
        if state.get("peers") is not None:  # records stored straight into every rank's memory over NVLink
            state["nsites"] = state["peers"].exchange()
            ctx.closest_and_measures()
            return
        n = ctx.sites_detect_local()
        state["nlocal"] = n
        keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
        corners = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
        ctx.sites_export_local(keys.data_ptr(), corners.data_ptr())
        ak, ac = slabs.exchange_site_records(keys[:n], corners[:n])
        torch.cuda.current_stream().synchronize()
        ctx.sites_import_global(ak.data_ptr(), ac.data_ptr(), ak.numel())
        state["nsites"] = ak.numel()
        ctx.closest_and_measures()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    if world > 1 and args.exchange == "peers":
        # one step over NCCL first: it sizes the receive regions and is the cross-check of the peer path
        step()
        ctx.synchronize()
        ref_sites = ctx.get_sites()
        t = torch.tensor([state["nlocal"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        state["peers"] = slabs.PeerExchange(ctx, 2 * int(t[0]) + 4096)
        step()
        ctx.synchronize()
        if not np.array_equal(ctx.get_sites(), ref_sites):
            raise SystemExit("peer-memory exchange and NCCL all-gather disagree on the site numbering")
        del ref_sites
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed region: exactly K steps, CUDA events on the stream the kernels are launched on
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (ctx.launch_count() - launches0) // max(args.steps, 1)
    # ---- same K steps again with every launch bracketed by events: per-kernel durations (roofline)
    ctx.profile(True)
    ctx.profile_reset()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        step()
    p1.record(stream)
    barrier()
    ms_prof = p0.elapsed_time(p1)
    prof = ctx.profile_report()
    ctx.profile(False)
    if os.environ.get("VC_BENCH_RANKS"):  # development aid: every rank's own kernel times (load balance across slabs)
        print(f"[rank {rank}] ms/step {ms / args.steps:.3f} profiled {ms_prof / args.steps:.3f} " +
              " ".join(f"{k}={v['ms'] / args.steps:.3f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]), file=sys.stderr)
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        t = torch.tensor([ms, ms_prof], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_prof = float(t[0]), float(t[1])
    ms_per_step = ms / args.steps
    value = nv_total / (ms_per_step * 1e-3)

    if args.no_e2e:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "ms_per_step": ms_per_step,
                  "config": {"workload": f"{fam}{nx}x{ny}x{nz}", "sites": state["nsites"], "z_slabs": world},
                  "kernels": {k: round(v["ms"] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                  "ms_per_step_profiled": ms_prof / args.steps, "e2e": None, "note": "--no-e2e: device-resident legs only"})
        ctx.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- e2e: host buffers in, host buffers out, copies inside the timed region
    s = (z1 - z0, ny, nx)
    outs = {
        "inside": api.PinnedArray(s, np.uint8), "id": api.PinnedArray(s, np.int32), "d2": api.PinnedArray(s, np.uint32),
        "edge3": api.PinnedArray((3,) + s, np.float32), "face3": api.PinnedArray((3,) + s, np.float32),
        "cube": api.PinnedArray(s, np.float32), "radius": api.PinnedArray(s, np.float32),
    }

    def step_e2e():
        if world == 1:
            ctx.run_dense_host(pin_vol.array, outs["inside"].array, outs["id"].array, outs["d2"].array,
                               outs["edge3"].array, outs["face3"].array, outs["cube"].array, outs["radius"].array)
            return
        ctx.upload_volume(pin_vol.array, zlo=lo)
        step()
        for k, which in (("inside", api.ARR_INSIDE), ("id", api.ARR_ID), ("d2", api.ARR_D2X4), ("edge3", api.ARR_EDGE3),
                         ("face3", api.ARR_FACE3), ("cube", api.ARR_CUBE), ("radius", api.ARR_RADIUS)):
            ctx.lib.vc_download(ctx.h, which, outs[k].array.ctypes.data)

    def time_e2e(fn):
        """seconds per call: median of individually timed calls (each call ends with its own stream syncs), max over ranks"""
        reps = max(3, min(args.steps, 10))
        fn()
        barrier()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        barrier()
        dt = float(np.median(ts))
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return dt

    def total(*vals):
        if dist is None:
            return [int(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return [int(v) for v in t]

    e2e_planes_s = time_e2e(step_e2e)
    h2d, d2h_planes = total(pin_vol.array.nbytes, sum(o.array.nbytes for o in outs.values()))
    e2e_variants = {"all_planes": {"value": nv_total / e2e_planes_s, "ms_per_step": e2e_planes_s * 1e3, "h2d_bytes_per_step": h2d,
                                   "d2h_bytes_per_step": d2h_planes,
                                   "result": "inside u8, id i32, 4d2 u32, 7 lambda planes f32, radius f32 for every grid vertex"}}
    # ---- e2e, compact product (the headline): pinned host volume in; occupancy bit rows + one record
    # (vertex, id, 4d2, 7 lambda, radius) per INSIDE vertex out.  Measures anchored at outside vertices are 0
    # by definition, so nothing is lost; the dense planes are still computed in HBM.
    e2e_s, d2h = e2e_planes_s, d2h_planes
    if world == 1 or state.get("peers") is not None:
        del outs
        n_in = ctx.compact_count()
        cap = n_in + 1024
        wr = nx // 32 + 1
        # the record arrays are the rows of ONE pinned 11 x cap block (vert | id | 4d2 | 7 lambda | radius): a z chunk's
        # records then come back in a single 2-D copy
        rec = api.PinnedArray((11, cap), np.uint32)
        cb = {"bits": api.PinnedArray(((z1 - z0) * ny, wr), np.uint32), "vert": rec.array[0], "id": rec.array[1].view(np.int32),
              "d2": rec.array[2], "lam": rec.array[3:10].view(np.float32), "rad": rec.array[10].view(np.float32)}

        def step_compact(dense=None):
            # on a slab ctx the same call uploads the rank's planes and exchanges the site records over peer memory
            return ctx.run_dense_host_compact(pin_vol.array, cap, cb["bits"].array, cb["vert"], cb["id"], cb["d2"], cb["lam"], cb["rad"],
                                              *(dense or (None, None)))

        e2e_s = time_e2e(step_compact)
        (d2h,) = total(int(cb["bits"].array.nbytes + n_in * 44))
        (n_in_total,) = total(n_in)
        # the e2e result really is the device result: spot-check the records against the resident planes
        got_n, _ = step_compact()
        step()  # the dense planes (the compact step computes the records of few inside vertices directly)
        ctx.synchronize()
        ids_dev = ctx.download(api.ARR_ID).ravel()
        cube_dev = ctx.download(api.ARR_CUBE).ravel()
        v = cb["vert"][:got_n]
        if got_n != n_in or not (np.array_equal(cb["id"][:got_n], ids_dev[v]) and np.array_equal(cb["lam"][6, :got_n], cube_dev[v])):
            raise SystemExit("compact e2e records disagree with the dense planes")
        del ids_dev, cube_dev
        e2e_variants["compact"] = {"value": nv_total / e2e_s, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                   "inside_vertices": n_in_total,
                                   "result": "occupancy bit rows + (vertex u32, id i32, 4d2 u32, 7 lambda f32, radius f32) per inside vertex"}
        if world == 1:
            # the same pass from an MRC mode 0 (signed byte) volume -- a format the reference's reader takes as well
            # (isosurface_tao/reader.h:235-239); the synthetic field is quantised to 1/16, which moves the surface a
            # little, so this is another volume, timed for what the narrower upload buys
            v8 = api.PinnedArray(pin_vol.array.shape, np.int8)
            v8.array[...] = np.clip(np.rint(pin_vol.array * 16.0), -127, 127).astype(np.int8)
            ctx.upload_volume(v8.array)
            ctx.classify_grid(fetch=False)
            cap8 = ctx.compact_count() + 1024
            rec8 = api.PinnedArray((11, cap8), np.uint32)
            dt8 = time_e2e(lambda: ctx.run_dense_host_compact(v8.array, cap8, cb["bits"].array, rec8.array[0], rec8.array[1].view(np.int32),
                                                              rec8.array[2], rec8.array[3:10].view(np.float32),
                                                              rec8.array[10].view(np.float32)))
            e2e_variants["compact_int8_volume"] = {"value": nv_total / dt8, "ms_per_step": dt8 * 1e3, "h2d_bytes_per_step": int(v8.array.nbytes),
                                                   "d2h_bytes_per_step": int(cb["bits"].array.nbytes + (cap8 - 1024) * 44),
                                                   "result": "compact product from an MRC mode 0 (int8) volume (field quantised to 1/16)"}
            del v8, rec8
        dense = (api.PinnedArray((z1 - z0, ny, nx), np.int32), api.PinnedArray((z1 - z0, ny, nx), np.uint32))
        dt = time_e2e(lambda: step_compact((dense[0].array, dense[1].array)))
        e2e_variants["compact_plus_dense_ids"] = {"value": nv_total / dt, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
                                                  "d2h_bytes_per_step": d2h + 8 * nv_total,
                                                  "result": "compact product + id i32 and 4d2 u32 planes for every grid vertex"}
        del dense

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel of the profiled region
        steps = args.steps
        kern = {k: {"ms_per_launch": v["ms"] / max(v["launches"], 1), "launches_per_step": v["launches"] / steps,
                    "ms_per_step": v["ms"] / steps} for k, v in prof.items()}
        dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
        planes_c = (min(z1 + 1, nz) - z0)
        unit_bytes = {  # bytes one launch must move by the kernel's own contract
            "classify_f32": 5 * nx * ny * (hi - lo),
            "cell_measures": BYTES_MEASURES * nv_local,
            "edt_pass_z": 8 * (nx + 1) * (ny + 1) * planes_c,
            "edt_pass_x": 8 * ((nx + 1) * (ny + 1) + nx * (ny + 1)) * planes_c,
            "edt_pass_y": 8 * (nx * (ny + 1) + nx * ny) * planes_c,
        }
        roof = None
        traffic, traffic_src, sm_pct = None, None, None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel's whole-grid launch, from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("workload") == f"{fam}{nx}x{ny}x{nz}" and dom in tj["kernels"]:
                k = tj["kernels"][dom]
                traffic = (k["dram_read_gb"] + k["dram_write_gb"]) * 1e9
                traffic_src, sm_pct = tj["source"], k.get("sm_throughput_pct")
        except Exception:
            pass
        if dom in unit_bytes:
            ach = unit_bytes[dom] / (kern[dom]["ms_per_launch"] * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "sm_throughput_pct_ncu": sm_pct, "peak_source": peak_src, "bytes_per_launch": unit_bytes[dom],
                    "ms_per_launch": kern[dom]["ms_per_launch"], "share_of_step": kern[dom]["ms_per_step"] / (ms_prof / steps),
                    "note": "bytes_per_launch is the kernel's dense contract (8 B in + 8 B out per element of its planes); pass X reads "
                            "only the columns that hold sites (column bitmap), so its DRAM traffic can be below it"}
        pipe_ach = BYTES_PIPELINE * nv_local / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/f32",
            "data": "synthetic",
            "config": {"workload": f"{fam}{nx}x{ny}x{nz}", "sites": state["nsites"], "z_slabs": world, "slab_heights": "equal" if args.balance <= 0 else f"balanced (1 + {args.balance} x inside fraction)",
                       "vertices_per_gpu": nv_local,
                       "exchange": ("none (one slab)" if world == 1 else
                                    "peer memory: detection kernel stores records into every rank over NVLink (vc_peer.cu)"
                                    if state.get("peers") is not None else "NCCL all-gather"), "l2": "inputs larger than L2 (no flush needed)",
                       "outputs": "inside u8, id i32, 4d2 u32, 7 lambda planes f32, radius f32"},
            "e2e": {"value": nv_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3,
                    "result": e2e_variants.get("compact", e2e_variants["all_planes"])["result"]},
            "e2e_variants": e2e_variants,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "roofline_pipeline": {"bound": "hbm", "bytes_per_vertex": BYTES_PIPELINE, "achieved": pipe_ach, "peak": peak,
                                  "unit": "GB/s", "frac": pipe_ach / peak, "note": "45 B/vertex (SURVEY 8d) x vertices / step time, per GPU"},
            "kernels": {k: round(v["ms_per_step"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])},
            "ms_per_step_profiled": ms_prof / steps,
        }
        if world == 1 and fam == "twist":
            # stage 1' (new functionality, no reference counterpart): the same pass starting from the closed triangle
            # mesh of the twisted plate instead of a voxel volume -- parity classification, then sites / closest / measures
            try:
                from voxel_ma_b200 import synth
                mv, mt = synth.twist_mesh(nx)
                ctx.classify_mesh(mv, mt, fetch=False)
                ctx.run_dense()
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    ctx.classify_mesh(mv, mt, fetch=False)
                    ns_mesh = ctx.run_dense()
                    ts.append(time.perf_counter() - t0)
                mprof = {}
                ctx.profile(True)
                ctx.profile_reset()
                ctx.classify_mesh(mv, mt, fetch=False)
                mprof = {k: round(v["ms"], 4) for k, v in ctx.profile_report().items() if k.startswith("mesh_")}
                ctx.profile(False)
                line["mesh_path"] = {"ms_per_step": float(np.median(ts)) * 1e3, "value": nv_total / float(np.median(ts)), "unit": UNIT,
                                     "triangles": int(len(mt)), "sites": int(ns_mesh), "classify_kernels_ms": mprof,
                                     "note": "vc_classify_mesh (warp-ballot parity over the surface triangles, mesh uploaded from the host "
                                             "each step) + sites + closest + measures; wall clock per step"}
            except Exception as ex:
                line["mesh_path"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                with quiet_stdout():
                    line["cpu_baseline"] = cpu_reference_rate(fam, grid, budget_s=20.0)
            except Exception as ex:  # the checker missing must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        emit(line)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="sphere256 | torus256 | twist512 | assembly1024 ...")
    ap.add_argument("--grid", default=None, help="NX,NY,NZ (assembly family)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--balance", type=float, default=0.0, help="EXPERIMENTAL, N>1: weight c of the per-plane inside fraction in the "
                    "slab-height estimate 1 + c*fraction (0 = equal heights, the default)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (very large grids: the all-planes leg needs "
                    "41 B of pinned host memory per grid vertex)")
    ap.add_argument("--exchange", default="peers", choices=["peers", "nccl"],
                    help="N>1: how the site records travel between ranks (peer-memory kernel stores, or NCCL all-gather)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
