"""Seeded closed meshes shared by the mesh-classification tests (stage 1')."""
import numpy as np

from voxel_ma_b200 import synth


def mesh_cases():
    rng = np.random.default_rng(77)
    cases = {
        # name: (verts, tris, (nx, ny, nz), M)
        "sphere40": (*synth.sphere_mesh(40), (40, 40, 40), None),
        "torus48": (*synth.torus_mesh(48), (48, 48, 48), None),
        "twist56": (*synth.twist_mesh(56), (56, 56, 56), None),
        # faces exactly through voxel centres / column points, edges on the lattice: tie rules decide
        "box_integer": (*synth.box_mesh((4, 4, 4), (20, 10, 12)), (32, 24, 16), None),
        "box_half": (*synth.box_mesh((4.5, 3.5, 2.5), (20.5, 10.5, 12.5)), (33, 19, 17), None),
        # mesh larger than the grid on every side (clipping of rows, columns and the crossing count)
        "box_outside": (*synth.box_mesh((-9.25, -7.5, -3), (50.75, 30.5, 40)), (23, 14, 11), None),
        "sphere_clipped": (*synth.sphere_mesh(32, radius=22.0), (32, 20, 27), None),
        # faces whose (y,z) boxes hold more than 1024 columns: the block-per-triangle kernel (k_mesh_large)
        "box_big": (*synth.box_mesh((3.3, 2.2, 1.1), (60.5, 45.5, 38.5)), (64, 48, 40), None),
    }
    # two overlapping solids: even-odd rule (the overlap is outside)
    v1, t1 = synth.sphere_mesh(36, radius=9.0, center=(14.2, 17.1, 18.3))
    v2, t2 = synth.sphere_mesh(36, radius=8.0, center=(22.4, 18.2, 17.7))
    cases["two_spheres_xor"] = (np.concatenate([v1, v2]), np.concatenate([t1, t2 + len(v1)]), (36, 36, 36), None)
    # a model-space mesh with a model -> voxel transform (scale + translate, column-major 4x4)
    v, t = synth.torus_mesh(40)
    s = 0.37
    M = np.array([1 / s, 0, 0, 0, 0, 1 / s, 0, 0, 0, 0, 1 / s, 0, 2.5, -1.25, 0.75, 1], np.float64)
    vm = ((v.astype(np.float64) - np.array([2.5, -1.25, 0.75])) * s).astype(np.float32)
    cases["torus_transformed"] = (vm, t, (40, 40, 40), M)
    # random closed "soup": tetrahedra with random corners (self-intersecting union under the parity rule)
    tv, tt = [], []
    for k in range(40):
        p = rng.uniform(-3, 30, (4, 3)).astype(np.float32)
        tv.append(p)
        tt.append(np.array([(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)], np.uint32) + 4 * k)
    cases["tetra_soup"] = (np.concatenate(tv), np.concatenate(tt), (27, 26, 25), None)
    # vertices snapped to the lattice (many exact ties), ragged grid
    tv, tt = [], []
    for k in range(30):
        p = rng.integers(-2, 20, (4, 3)).astype(np.float32) * 0.5
        tv.append(p)
        tt.append(np.array([(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)], np.uint32) + 4 * k)
    cases["tetra_lattice"] = (np.concatenate(tv), np.concatenate(tt), (11, 9, 10), None)
    cases["empty_mesh"] = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), (8, 7, 6), None)
    return cases
