"""Deterministic synthetic volumes for the BASELINE.json configs (no network, no datasets).

Every generator returns a float32 array of shape ``(nz, ny, nx)`` -- x fastest, i.e. the MRC file
payload order the reference reads (3rdparty/isosurface_tao/reader.h:222-251).  Occupancy follows
the reference rule ``value > 0`` (include/spaceinfo.h:122-129).  The fields are continuous signed
functions, so exact zeros are measure-zero; dedicated tests add zeros / -0.0 / NaN explicitly.

SURVEY.md section 8(d) names the five workloads:
  sphere64      Tao's SphereVolumeGenerator(64,1,1,1) shifted by -3  (9 200 sites; probe input)
  torus256      major radius 88, minor radius 38, axis z             (~2e5 sites)
  twist512      6-voxel plate 400x200 twisted 180 deg + 3 rods       (dense ties / near ties)
  assembly1024  64 solids from mt19937_64(20181)                     (z-slab sharded)
  stress2048    the same generator scaled
All generators work slab-wise (``z0, z1``) so a rank can build only its own z-slab.
"""
from __future__ import annotations

import struct

import numpy as np

__all__ = [
    "sphere", "torus", "twist", "assembly", "make", "write_mrc", "read_mrc", "to_zfast_f64",
    "WORKLOADS",
]


def _grid(n, z0, z1, dtype=np.float32):
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    z0 = 0 if z0 is None else z0
    z1 = nz if z1 is None else z1
    z = np.arange(z0, z1, dtype=dtype)[:, None, None]
    y = np.arange(ny, dtype=dtype)[None, :, None]
    x = np.arange(nx, dtype=dtype)[None, None, :]
    return x, y, z, (nx, ny, nz)


def sphere(n: int, z0=None, z1=None) -> np.ndarray:
    """Tao's SphereVolumeGenerator(n,1,1,1) minus 3 (reader.h:1122-1151), with its INTEGER
    divisions ``(n-1)/2`` and ``(n-1)^2/4`` kept; radius ~0.35 n.  64 -> 9 200 sites,
    128 -> 37 328, 256 -> 150 200."""
    x, y, z, _ = _grid(n, z0, z1, np.float64)
    c = (n - 1) // 2
    dis = float(((n - 1) * (n - 1)) // 4)
    d = (x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2
    f = 10.0 - 10.0 * np.sqrt(d) / np.sqrt(dis) - 3.0
    return f.astype(np.float32)


def torus(n: int = 256, major: float = 88.0, minor: float = 38.0, z0=None, z1=None) -> np.ndarray:
    """Torus with axis z, centred; radii scale with n/256."""
    s = n / 256.0
    R, r = major * s, minor * s
    x, y, z, _ = _grid(n, z0, z1)
    c = np.float32((n - 1) / 2.0)
    q = np.sqrt((x - c) ** 2 + (y - c) ** 2) - np.float32(R)
    f = np.float32(r) - np.sqrt(q * q + (z - c) ** 2)
    return f.astype(np.float32)


def twist(n: int = 512, z0=None, z1=None) -> np.ndarray:
    """Thin plate (6 x 400 x 200 voxels at n=512) twisted 180 degrees about its long (x) axis,
    plus three rods of diameter 5: thin features give dense exact / near ties between closest
    samples."""
    s = np.float32(n / 512.0)
    x, y, z, _ = _grid(n, z0, z1)
    c = np.float32((n - 1) / 2.0)
    L, W, T = np.float32(200.0) * s, np.float32(100.0) * s, np.float32(3.0) * s
    u = x - c
    ang = np.float32(np.pi) * (u / (2 * L) + np.float32(0.5))
    ca, sa = np.cos(ang), np.sin(ang)
    v = (y - c) * ca + (z - c) * sa
    w = -(y - c) * sa + (z - c) * ca
    plate = np.minimum(np.minimum(L - np.abs(u), W - np.abs(v)), T - np.abs(w))
    rr = np.float32(2.5) * s
    rod_x = rr - np.sqrt((y - np.float32(0.2 * n)) ** 2 + (z - np.float32(0.8 * n)) ** 2)
    rod_x = np.minimum(rod_x, np.float32(0.45 * n) - np.abs(u))
    rod_y = rr - np.sqrt((x - np.float32(0.15 * n)) ** 2 + (z - np.float32(0.2 * n)) ** 2)
    rod_y = np.minimum(rod_y, np.float32(0.45 * n) - np.abs(y - c))
    rod_z = rr - np.sqrt((x - np.float32(0.85 * n)) ** 2 + (y - np.float32(0.85 * n)) ** 2)
    rod_z = np.minimum(rod_z, np.float32(0.45 * n) - np.abs(z - c))
    f = np.maximum(np.maximum(plate, rod_x), np.maximum(rod_y, rod_z))
    return np.broadcast_to(f, (f.shape[0], f.shape[1], f.shape[2])).astype(np.float32)


def assembly(n=1024, count: int = 64, seed: int = 20181, z0=None, z1=None) -> np.ndarray:
    """Union of `count` solids (spheres, tori, boxes) with centres / radii drawn from
    mt19937(seed).  ``n`` is a side length or an (nx, ny, nz) triple; solids are placed in the unit
    cube and scaled per axis extent, sized by the smallest side.  Built solid by solid on each
    solid's bounding box only, so a rank can generate just its own z-slab."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    m = min(nx, ny, nz)
    rng = np.random.Generator(np.random.MT19937(seed))  # mt19937; seeded deterministically
    nz0 = 0 if z0 is None else z0
    nz1 = nz if z1 is None else z1
    out = np.full((nz1 - nz0, ny, nx), -1.0, dtype=np.float32)
    dims = np.array([nx, ny, nz], dtype=np.float64)
    for _ in range(count):
        kind = int(rng.integers(0, 3))
        c = (rng.uniform(0.12, 0.88, size=3) * dims).astype(np.float32)
        r = np.float32(rng.uniform(0.03 * m, 0.10 * m))
        r2 = np.float32(rng.uniform(0.3, 0.5)) * r
        ext = int(np.ceil(r + r2 + 2))
        lo = np.maximum(np.floor(c).astype(int) - ext, 0)
        hi = np.minimum(np.floor(c).astype(int) + ext + 1, [nx, ny, nz])
        zlo, zhi = max(lo[2], nz0), min(hi[2], nz1)
        if zlo >= zhi:
            continue
        z = np.arange(zlo, zhi, dtype=np.float32)[:, None, None] - c[2]
        y = np.arange(lo[1], hi[1], dtype=np.float32)[None, :, None] - c[1]
        x = np.arange(lo[0], hi[0], dtype=np.float32)[None, None, :] - c[0]
        if kind == 0:
            f = r - np.sqrt(x * x + y * y + z * z)
        elif kind == 1:
            q = np.sqrt(x * x + y * y) - r
            f = r2 - np.sqrt(q * q + z * z)
        else:
            f = np.minimum(np.minimum(r - np.abs(x), r2 * 2 - np.abs(y)), r * np.float32(0.7) - np.abs(z))
        sub = out[zlo - nz0:zhi - nz0, lo[1]:hi[1], lo[0]:hi[0]]
        np.maximum(sub, f.astype(np.float32), out=sub)
    return out


WORKLOADS = {
    "sphere64": lambda z0=None, z1=None: sphere(64, z0, z1),
    "torus256": lambda z0=None, z1=None: torus(256, z0=z0, z1=z1),
    "twist512": lambda z0=None, z1=None: twist(512, z0, z1),
    "assembly1024": lambda z0=None, z1=None: assembly(1024, z0=z0, z1=z1),
    "stress2048": lambda z0=None, z1=None: assembly(2048, count=160, z0=z0, z1=z1),
}


def make(name: str, n=None, z0=None, z1=None) -> np.ndarray:
    """Build a named workload, optionally at another resolution ``n`` (same shape family; the
    assembly family also takes an (nx, ny, nz) triple)."""
    fam = name.rstrip("0123456789")
    if n is None:
        return WORKLOADS[name](z0, z1)
    if fam == "sphere":
        return sphere(n, z0, z1)
    if fam == "torus":
        return torus(n, z0=z0, z1=z1)
    if fam == "twist":
        return twist(n, z0, z1)
    if fam in ("assembly", "stress"):
        m = n if np.isscalar(n) else min(n)
        cnt = 64 if fam == "assembly" else 160
        if not np.isscalar(n):  # keep the solid density of the cubic workload
            cnt = max(1, int(round(cnt * (n[0] * n[1] * n[2]) / float(max(n)) ** 3)))
        return assembly(n, count=cnt, z0=z0, z1=z1)
    raise KeyError(name)


def to_zfast_f64(vol_xfast: np.ndarray) -> np.ndarray:
    """(nz,ny,nx) float32 -> Tao Volume order double[x*ny*nz + y*nz + z] (volume.h:217-224)."""
    return np.ascontiguousarray(vol_xfast.astype(np.float64).transpose(2, 1, 0))


def write_mrc(path: str, vol_xfast: np.ndarray) -> None:
    """MRC mode 2 (float32), 1024-byte header, x fastest -- the bytes MRCReader parses
    (reader.h:148-192: nx,ny,nz,mode, 3 offsets, 3 dims, 3 cell, 3 angles, 12 skipped bytes,
    dmin,dmax,dmean, 128 skipped bytes, rms)."""
    v = np.ascontiguousarray(vol_xfast, dtype=np.float32)
    nz, ny, nx = v.shape
    hdr = bytearray(1024)
    struct.pack_into("<10i", hdr, 0, nx, ny, nz, 2, 0, 0, 0, nx - 1, ny - 1, nz - 1)
    struct.pack_into("<6f", hdr, 40, float(nx - 1), float(ny - 1), float(nz - 1), 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    struct.pack_into("<3f", hdr, 76, float(v.min()), float(v.max()), float(v.mean()))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(v.tobytes())


def read_mrc(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        hdr = f.read(1024)
        nx, ny, nz, mode = struct.unpack_from("<4i", hdr, 0)
        assert mode == 2, "only MRC mode 2 (float32) is produced by this package"
        return np.frombuffer(f.read(), dtype=np.float32, count=nx * ny * nz).reshape(nz, ny, nx).copy()
