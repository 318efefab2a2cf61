"""Host side of the dense medial complex: dual quads (vc_medial_quads) -> a cubical 2-complex in the
form the reference's cellcomplex takes (vertices, edges, polygon faces; src/cellcomplex.cpp:364-491), plus the
statistics it is judged on (counts, Euler characteristic, lambda range) and a PLY writer with the reference's element
layout (src/plyall.cpp:98-170: vertex x y z, edge vertex1 vertex2, face vertex_indices).

A quad is the dual of a grid edge (v, v + e_axis): its corners are the centres of the 4 grid cubes around the edge,
cube (a, b, c) = vertices a..a+1 x b..b+1 x c..c+1 with centre (a + 1/2, b + 1/2, c + 1/2)."""
from __future__ import annotations

import numpy as np

# cube offsets (relative to the edge's lower end vertex) of the 4 cubes around an edge along +x / +y / +z, cyclic
_AROUND = {
    0: np.array([[0, -1, -1], [0, 0, -1], [0, 0, 0], [0, -1, 0]]),
    1: np.array([[-1, 0, -1], [-1, 0, 0], [0, 0, 0], [0, 0, -1]]),
    2: np.array([[-1, -1, 0], [0, -1, 0], [0, 0, 0], [-1, 0, 0]]),
}


def build_complex(anchor, axis, nx, ny, z0=0):
    """-> dict(vertices float32 [V,3] (cube centres, grid coordinates), edges int32 [E,2], faces int32 [F,4],
    face_of_quad = identity, cube ids).  Vertices / edges are the distinct corners / sides of the quads."""
    anchor = np.asarray(anchor, np.int64)
    axis = np.asarray(axis, np.int64)
    x = anchor % nx
    y = (anchor // nx) % ny
    z = anchor // (nx * ny) + z0
    v = np.stack([x, y, z], -1)
    corners = np.empty((len(anchor), 4, 3), np.int64)
    for a in range(3):
        m = axis == a
        corners[m] = v[m][:, None, :] + _AROUND[a][None, :, :]
    # cube id: the cubes of an emitted quad exist, so a, b, c >= 0
    cid = (corners[..., 2] * (ny + 1) + corners[..., 1]) * (nx + 1) + corners[..., 0]
    uniq, inv = np.unique(cid.ravel(), return_inverse=True)
    faces = inv.reshape(-1, 4).astype(np.int32)
    cz, rem = np.divmod(uniq, (ny + 1) * (nx + 1))
    cy, cx = np.divmod(rem, nx + 1)
    verts = np.stack([cx + 0.5, cy + 0.5, cz + 0.5], -1).astype(np.float32)
    e = np.stack([faces, np.roll(faces, -1, axis=1)], -1).reshape(-1, 2)
    e.sort(axis=1)
    edges = np.unique(e, axis=0).astype(np.int32)
    return {"vertices": verts, "edges": edges, "faces": faces, "cubes": uniq}


def statistics(cx, lam):
    """counts, Euler characteristic V - E + F, lambda range: what the complex is compared on"""
    V, E, F = len(cx["vertices"]), len(cx["edges"]), len(cx["faces"])
    return {"V": V, "E": E, "F": F, "euler": V - E + F, "lambda_min": float(lam.min()) if F else None,
            "lambda_max": float(lam.max()) if F else None}


def thin_by_threshold(cx, lam, t):
    """faces with lambda >= t only (the reference prunes by lambda threshold, src/ccthin.cpp:201-424; this is the plain cut,
    without the simple-pair collapse) -> the same dict for the remaining sub-complex"""
    keep = lam >= t
    faces = cx["faces"][keep]
    used = np.unique(faces.ravel())
    remap = np.full(len(cx["vertices"]), -1, np.int64)
    remap[used] = np.arange(len(used))
    faces = remap[faces].astype(np.int32)
    e = np.stack([faces, np.roll(faces, -1, axis=1)], -1).reshape(-1, 2)
    e.sort(axis=1)
    return {"vertices": cx["vertices"][used], "edges": np.unique(e, axis=0).astype(np.int32), "faces": faces}, lam[keep]


def write_ply(path, cx):
    """ASCII PLY with the reference writer's element layout (src/plyall.cpp:117-138)"""
    v, e, f = cx["vertices"], cx["edges"], cx["faces"]
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\n")
        fh.write(f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n")
        fh.write(f"element edge {len(e)}\nproperty int vertex1\nproperty int vertex2\n")
        fh.write(f"element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n")
        np.savetxt(fh, v, fmt="%g")
        np.savetxt(fh, e, fmt="%d")
        np.savetxt(fh, np.concatenate([np.full((len(f), 1), 4, np.int32), f], 1), fmt="%d")
