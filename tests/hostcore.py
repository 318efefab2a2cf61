"""ctypes access to tests/host_harness.cpp (the kernels' integer core driven on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libhostharness.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "host_harness.cpp")
        hdrs = [os.path.join(HERE, "..", "voxel_ma_b200", "csrc", h) for h in ("vc_core.h", "vc_mesh_core.h")]
        if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in [src, *hdrs]):
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", SO])
        L = C.CDLL(SO)
        L.hh_sites.restype = C.c_int64
        L.hh_site_records.restype = C.c_int64
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sites(inside):
    nz, ny, nx = inside.shape
    ins = np.ascontiguousarray(inside, np.uint8)
    n = lib().hh_sites(_p(ins), nx, ny, nz, None, C.c_int64(0))
    out = np.empty((n, 3), np.float32)
    lib().hh_sites(_p(ins), nx, ny, nz, _p(out), C.c_int64(n))
    return out


def closest_grid(sites_xyz, nx, ny, nz, z0=0, z1=None):
    z1 = nz if z1 is None else z1
    c = np.ascontiguousarray(np.round(np.asarray(sites_xyz) + 0.5).astype(np.int32))
    ids = np.empty((z1 - z0, ny, nx), np.int32)
    d2 = np.empty((z1 - z0, ny, nx), np.uint32)
    lib().hh_closest_grid(_p(c), C.c_int64(len(c)), nx, ny, nz, z0, z1, _p(ids), _p(d2))
    return ids, d2


def closest_grid2(sites_xyz, nx, ny, nz, z0=0, z1=None, want_stats=False):
    """round-2 data flow: compact columns, live rows, pruned (Meijster-style) envelopes"""
    z1 = nz if z1 is None else z1
    c = np.ascontiguousarray(np.round(np.asarray(sites_xyz) + 0.5).astype(np.int32))
    ids = np.empty((z1 - z0, ny, nx), np.int32)
    d2 = np.empty((z1 - z0, ny, nx), np.uint32)
    stats = np.zeros(8, np.int64)
    hist = np.zeros((2, 64), np.int64)
    lib().hh_closest_grid2(_p(c), C.c_int64(len(c)), nx, ny, nz, z0, z1, _p(ids), _p(d2), _p(stats), _p(hist))
    return (ids, d2, stats, hist) if want_stats else (ids, d2)


def sep_sweep():
    lib().hh_sep_sweep.restype = C.c_int64
    return int(lib().hh_sep_sweep())


def envelope(H, ntgt):
    H = np.ascontiguousarray(H, np.uint64)
    out = np.empty(ntgt, np.uint64)
    lib().hh_envelope(_p(H), len(H), ntgt, _p(out))
    return out


def envelope_pruned(H, ntgt):
    """the same line through the round-2 scan (vc_envelope_pruned) over the compacted live candidates"""
    H = np.ascontiguousarray(H, np.uint64)
    out = np.empty(ntgt, np.uint64)
    lib().hh_envelope_pruned(_p(H), len(H), ntgt, _p(out))
    return out


def envelope_masked(H, ntgt):
    """same line through the candidate-bitmap path of pass X (dead candidates hold garbage)"""
    H = np.ascontiguousarray(H, np.uint64)
    out = np.empty(ntgt, np.uint64)
    lib().hh_envelope_masked(_p(H), len(H), ntgt, _p(out))
    return out


def site_records(inside_planes, nx, ny, nz, zlo, czb, cze):
    ins = np.ascontiguousarray(inside_planes, np.uint8)
    zhi = zlo + ins.shape[0]
    n = lib().hh_site_records(_p(ins), nx, ny, nz, zlo, zhi, czb, cze, None, None, C.c_int64(0))
    assert n >= 0, "slab does not hold a needed plane"
    k = np.empty(n, np.uint64)
    c = np.empty(n, np.uint64)
    lib().hh_site_records(_p(ins), nx, ny, nz, zlo, zhi, czb, cze, _p(k), _p(c), C.c_int64(n))
    return k, c


def classify_mesh(verts, tris, nx, ny, nz):
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    out = np.empty((nz, ny, nx), np.uint8)
    rc = lib().hh_classify_mesh(_p(v), C.c_int64(len(v)), _p(t), C.c_int64(len(t)), nx, ny, nz, _p(out))
    return rc, out
