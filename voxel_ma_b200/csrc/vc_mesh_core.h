// vc_mesh_core.h -- the exact integer rules of the mesh parity classification (vc_mesh.cu).
// __host__ __device__ so tests/host_harness.cpp can drive the same lines on the CPU; the product
// calls them from CUDA kernels only.
//
// Coordinates are voxel-space positions snapped to 1/256 voxel (Q = 256 q), kept in [-2^18, 3*2^18):
// differences stay below 2^20, a 2-D cross product below 2^41, and the crossing numerator
// num = Ax*D + wB*(Bx-Ax) + wC*(Cx-Ax) below 2^63, so int64 is exact throughout.
#pragma once
#include <stdint.h>

#include "vc_core.h"

#define VC_MESH_SUB 256         // sub-voxel positions per voxel
#define VC_MESH_QMIN (-262144)  // -1024 voxels
#define VC_MESH_QMAX (786431)   //  3072 voxels - 1/256

// float voxel-space coordinate -> snapped integer; false when not finite / out of range
VC_HD bool vc_mesh_snap(float p, int* Q)
{
    double s = (double)p * 256.0 + 0.5; // exact: p is a float
    if (!(s >= (double)VC_MESH_QMIN && s <= (double)VC_MESH_QMAX))
    {
        *Q = 0;
        return false;
    }
    long long f = (long long)s; // truncation toward zero ...
    if ((double)f > s)
        --f; // ... corrected to floor for negative non-integers
    *Q = (int)f;
    return true;
}

// twice the signed area of (U, V, P) in the (y,z) plane
VC_HD long long vc_mesh_orient(int uy, int uz, int vy, int vz, int py, int pz)
{
    return (long long)(vy - uy) * (long long)(pz - uz) - (long long)(vz - uz) * (long long)(py - uy);
}

// makes (A,B,C) counter-clockwise in (y,z) by swapping B and C; false when the projection is degenerate
VC_HD bool vc_mesh_orient_ccw(int* ax, int* ay, int* az, int* bx, int* by, int* bz, int* cx, int* cy, int* cz)
{
    const long long D = vc_mesh_orient(*ay, *az, *by, *bz, *cy, *cz);
    if (D == 0)
        return false;
    if (D < 0)
    {
        int t;
        t = *bx, *bx = *cx, *cx = t;
        t = *by, *by = *cy, *cy = t;
        t = *bz, *bz = *cz, *cz = t;
    }
    (void)ax;
    return true;
}

// grid columns (j,k) whose centre can lie in the triangle's (y,z) box, clipped to the resident rows
VC_HD void vc_mesh_columns(int ay, int az, int by, int bz, int cy, int cz, int ny, int zlo, int zhi, int* j0, int* j1, int* k0,
                           int* k1)
{
    int ymin = ay < by ? ay : by, ymax = ay > by ? ay : by, zmin = az < bz ? az : bz, zmax = az > bz ? az : bz;
    ymin = ymin < cy ? ymin : cy, ymax = ymax > cy ? ymax : cy, zmin = zmin < cz ? zmin : cz, zmax = zmax > cz ? zmax : cz;
    int a = (ymin + 255) >> 8, b = ymax >> 8, c = (zmin + 255) >> 8, d = zmax >> 8; // ceil / floor of Q/256
    *j0 = a > 0 ? a : 0;
    *j1 = b < ny - 1 ? b : ny - 1;
    *k0 = c > zlo ? c : zlo;
    *k1 = d < zhi - 1 ? d : zhi - 1;
}

// edge U->V of a counter-clockwise triangle, w = orient(U,V,P): is P on the inner side?  A point ON the
// edge belongs to the triangle iff the edge runs against the canonical direction (increasing y, then z).
VC_HD bool vc_mesh_edge_in(long long w, int uy, int uz, int vy, int vz)
{
    return w > 0 || (w == 0 && (uy > vy || (uy == vy && uz > vz)));
}

// Does the counter-clockwise triangle cover column (j,k)?  If so *T = number of voxels i in [0,nx) with
// 256 i < x_c, the abscissa where the +x ray of the column meets the triangle's plane.
VC_HD bool vc_mesh_crossing(int ax, int ay, int az, int bx, int by, int bz, int cx, int cy, int cz, int j, int k, int nx, int* T)
{
    const int py = j * VC_MESH_SUB, pz = k * VC_MESH_SUB;
    const long long wa = vc_mesh_orient(by, bz, cy, cz, py, pz); // edge B->C
    const long long wb = vc_mesh_orient(cy, cz, ay, az, py, pz); // edge C->A
    const long long wc = vc_mesh_orient(ay, az, by, bz, py, pz); // edge A->B
    if (!vc_mesh_edge_in(wa, by, bz, cy, cz) || !vc_mesh_edge_in(wb, cy, cz, ay, az) || !vc_mesh_edge_in(wc, ay, az, by, bz))
        return false;
    const long long D = wa + wb + wc; // > 0
    const long long num = (long long)ax * D + wb * (long long)(bx - ax) + wc * (long long)(cx - ax);
    long long t = 0;
    if (num > 0)
    {
        const long long den = D * VC_MESH_SUB;
        t = (num + den - 1) / den; // ceil
    }
    *T = t > nx ? nx : (int)t;
    return true;
}
