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


# SURVEY section 8d-4 fixes assembly1024 at 64 solids and S ~ 2-3e6 boundary samples; with radii drawn from
# [0.03, 0.10] x side the 64 solids give 5.5e6, so the named workloads shrink every radius by this factor
# (S ~ 2.5e6 at 1024^3: the same samples-per-vertex density as twist512).
ASSEMBLY_RSCALE = 0.67


def assembly(n=1024, count: int = 64, seed: int = 20181, z0=None, z1=None, rscale: float = 1.0, size_ref=None) -> np.ndarray:
    """Union of `count` solids (spheres, tori, boxes) with centres / radii drawn from
    mt19937(seed).  ``n`` is a side length or an (nx, ny, nz) triple; solids are placed in the unit
    cube and scaled per axis extent, sized by the smallest side; `rscale` scales every radius.  Built
    solid by solid on each solid's bounding box only, so a rank can generate just its own z-slab."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    m = min(nx, ny, nz) if size_ref is None else size_ref  # the side the radii are relative to
    rng = np.random.Generator(np.random.MT19937(seed))  # mt19937; seeded deterministically
    nz0 = 0 if z0 is None else z0
    nz1 = nz if z1 is None else z1
    out = np.full((nz1 - nz0, ny, nx), -1.0, dtype=np.float32)
    dims = np.array([nx, ny, nz], dtype=np.float64)
    for _ in range(count):
        kind = int(rng.integers(0, 3))
        c = (rng.uniform(0.12, 0.88, size=3) * dims).astype(np.float32)
        r = np.float32(rng.uniform(0.03 * m, 0.10 * m) * rscale)
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
    "assembly1024": lambda z0=None, z1=None: assembly(1024, z0=z0, z1=z1, rscale=ASSEMBLY_RSCALE),
    "stress2048": lambda z0=None, z1=None: assembly(2048, count=160, z0=z0, z1=z1, rscale=ASSEMBLY_RSCALE),
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
        # parts of the cube (the weak-scaling grids 512x512x1024, 512x1024x1024): solids as large as in the full cube, their
        # number scaled with the volume, so the boundary samples per grid vertex stay what they are at assembly1024
        ref = None if np.isscalar(n) else max(n)
        return assembly(n, count=cnt, z0=z0, z1=z1, rscale=ASSEMBLY_RSCALE if m >= 256 else 1.0, size_ref=ref)
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


# ---- closed triangle meshes (inputs of vc_classify_mesh, stage 1') ----------------------------------
def _param_mesh(fn, nu: int, nv: int, wrap_v: bool):
    """Triangles of a (u, v) parameter grid; u always wraps, v wraps for a torus.  Vertex (a, b) =
    fn(2 pi a / nu, b / nv')."""
    a = np.arange(nu)
    b = np.arange(nv if wrap_v else nv + 1)
    U, V = np.meshgrid(a, b, indexing="ij")
    verts = fn(U.ravel(), V.ravel()).astype(np.float32)
    nb = len(b)
    idx = lambda i, j: (i % nu) * nb + (j % nb if wrap_v else j)  # noqa: E731
    tris = []
    for i in range(nu):
        for j in range(nv):
            p00, p10, p01, p11 = idx(i, j), idx(i + 1, j), idx(i, j + 1), idx(i + 1, j + 1)
            tris.append((p00, p10, p11))
            tris.append((p00, p11, p01))
    return verts, np.asarray(tris, np.uint32)


def sphere_mesh(n: int, nu: int = 96, nv: int = 48, radius: float | None = None, center=None):
    """Closed UV sphere in voxel space (poles are rings of coincident vertices: degenerate triangles
    there exercise the zero-area rule).  Default radius 0.35 n, centred off-lattice."""
    r = 0.35 * n if radius is None else radius
    c = np.array([(n - 1) / 2.0 + 0.137, (n - 1) / 2.0 - 0.211, (n - 1) / 2.0 + 0.319] if center is None else center)

    def fn(a, b):
        th = 2 * np.pi * a / nu
        ph = np.pi * b / nv
        return np.stack([c[0] + r * np.sin(ph) * np.cos(th), c[1] + r * np.sin(ph) * np.sin(th), c[2] + r * np.cos(ph)], 1)

    return _param_mesh(fn, nu, nv, False)


def torus_mesh(n: int, major: float | None = None, minor: float | None = None, nu: int = 128, nv: int = 64, center=None):
    """Closed torus, axis z, radii scale like synth.torus (88 / 38 at n = 256)."""
    R = 88.0 * n / 256.0 if major is None else major
    r = 38.0 * n / 256.0 if minor is None else minor
    c = np.array([(n - 1) / 2.0 + 0.25, (n - 1) / 2.0 - 0.125, (n - 1) / 2.0 + 0.0625] if center is None else center)

    def fn(a, b):
        th = 2 * np.pi * a / nu
        ph = 2 * np.pi * b / nv
        q = R + r * np.cos(ph)
        return np.stack([c[0] + q * np.cos(th), c[1] + q * np.sin(th), c[2] + r * np.sin(ph)], 1)

    return _param_mesh(fn, nu, nv, True)


def box_mesh(lo, hi):
    """Axis-aligned box with 12 triangles.  With integer / half-integer corners every face is seen
    edge-on or passes exactly through voxel centres and column points: the tie rules decide."""
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    v = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    t = []
    for a, b, c, d in quads:
        t += [(a, b, c), (a, c, d)]
    return v, np.asarray(t, np.uint32)


def twist_mesh(n: int, nu: int = 256):
    """The twisted thin plate of synth.twist as a closed mesh: a 0.78 n x 0.39 n x 6 slab whose cross
    section rotates 180 degrees along x."""
    L, W, H = 0.78 * n, 0.39 * n, 6.0
    c = (n - 1) / 2.0
    ring = np.array([(-W / 2, -H / 2), (W / 2, -H / 2), (W / 2, H / 2), (-W / 2, H / 2)])
    verts = []
    for i in range(nu + 1):
        t = i / nu
        ang = np.pi * t
        ca, sa = np.cos(ang), np.sin(ang)
        for (y, z) in ring:
            verts.append((c - L / 2 + L * t + 0.031, c + ca * y - sa * z + 0.017, c + sa * y + ca * z - 0.043))
    tris = []
    for i in range(nu):
        for k in range(4):
            a, b = 4 * i + k, 4 * i + (k + 1) % 4
            tris += [(a, b, b + 4), (a, b + 4, a + 4)]
    e = 4 * nu
    tris += [(0, 2, 1), (0, 3, 2), (e, e + 1, e + 2), (e, e + 2, e + 3)]
    return np.asarray(verts, np.float32), np.asarray(tris, np.uint32)
