// vc_circum.cu -- north_star stage 3, the two measures the reference only sketches: circumradius and object angle of a
// cell's closest-point set.
//
// The reference's measureforMA evaluates lambda only; a circumradius appears in comments and commented-out
// approximations (src/voroinfo.cpp:1441-1443, 1475-1479, 1526-1530) and an angle nowhere.  PARITY UNPINNED: there is no
// reference behaviour to match, the definition below is the builder's (SURVEY section 0 makes these optional outputs),
// restated independently in oracle/oracle.c (orc_cell_circum_angle_grid) and compared to 1e-12.
//
// Per cell of the dictionary of SURVEY section 0 (3 grid edges +x +y +z, 3 grid faces xy xz yz, 1 cube, anchored at vertex v;
// valid iff all its vertices are inside, else both measures are 0), with P = the DISTINCT closest sites of the cell's
// vertices (2 / 4 / 8 vertex sets) and m = the centre of the cell, all in float64 ("fp64 wherever the reference uses
// double": ANN distances are double, 3rdparty/ann/include/ANN/ANN.h:160-161):
//   circumradius = radius of the smallest ball that encloses P        (0 when |P| = 1)
//   object angle = max over pairs p, q in P of  angle(p - m, q - m) / 2   in [0, pi/2]   (0 when |P| = 1)
// The smallest enclosing ball of <= 8 points is found by enumeration: the smallest among the balls spanned by a pair
// (diametral), a triple (circumcircle) or a quadruple (circumsphere) of points that contains all of P.
#include <math.h>

#include "vc_internal.h"

struct D3
{
    double x, y, z;
};
__device__ __forceinline__ D3 d3sub(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double d3dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 d3cross(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

__device__ bool ball_holds(const D3* p, int n, D3 c, double r2)
{
    const double lim = r2 * (1.0 + 1e-12) + 1e-12;
    for (int i = 0; i < n; ++i)
    {
        const D3 d = d3sub(p[i], c);
        if (d3dot(d, d) > lim)
            return false;
    }
    return true;
}

__device__ double seb_radius(const D3* p, int n)
{
    if (n <= 1)
        return 0.0;
    double best = 1e300;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
        {
            const D3 c{0.5 * (p[i].x + p[j].x), 0.5 * (p[i].y + p[j].y), 0.5 * (p[i].z + p[j].z)};
            const D3 d = d3sub(p[i], p[j]);
            const double r2 = 0.25 * d3dot(d, d);
            if (r2 < best && ball_holds(p, n, c, r2))
                best = r2;
        }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            for (int k = j + 1; k < n; ++k)
            {
                const D3 a = d3sub(p[j], p[i]), b = d3sub(p[k], p[i]);
                const D3 ab = d3cross(a, b);
                const double den = 2.0 * d3dot(ab, ab);
                if (den < 1e-18)
                    continue; // collinear: a pair's ball covers it
                const double a2 = d3dot(a, a), b2 = d3dot(b, b);
                const D3 t{a2 * b.x - b2 * a.x, a2 * b.y - b2 * a.y, a2 * b.z - b2 * a.z};
                const D3 o = d3cross(t, ab);
                const D3 off{o.x / den, o.y / den, o.z / den};
                const double r2 = d3dot(off, off);
                if (r2 < best && ball_holds(p, n, D3{p[i].x + off.x, p[i].y + off.y, p[i].z + off.z}, r2))
                    best = r2;
            }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            for (int k = j + 1; k < n; ++k)
                for (int l = k + 1; l < n; ++l)
                {
                    const D3 a = d3sub(p[j], p[i]), b = d3sub(p[k], p[i]), c = d3sub(p[l], p[i]);
                    const D3 bc = d3cross(b, c), ca = d3cross(c, a), ab = d3cross(a, b);
                    const double det = d3dot(a, bc);
                    if (fabs(det) < 1e-12)
                        continue; // coplanar: a triple's ball covers it
                    const double a2 = d3dot(a, a), b2 = d3dot(b, b), c2 = d3dot(c, c), s = 0.5 / det;
                    const D3 off{s * (a2 * bc.x + b2 * ca.x + c2 * ab.x), s * (a2 * bc.y + b2 * ca.y + c2 * ab.y),
                                 s * (a2 * bc.z + b2 * ca.z + c2 * ab.z)};
                    const double r2 = d3dot(off, off);
                    if (r2 < best && ball_holds(p, n, D3{p[i].x + off.x, p[i].y + off.y, p[i].z + off.z}, r2))
                        best = r2;
                }
    return sqrt(best);
}

__device__ double object_angle(const D3* p, int n, D3 m)
{
    double best = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
        {
            const D3 u = d3sub(p[i], m), v = d3sub(p[j], m);
            const double lu = d3dot(u, u), lv = d3dot(v, v);
            if (lu <= 0.0 || lv <= 0.0)
                continue;
            double cs = d3dot(u, v) / sqrt(lu * lv);
            cs = cs > 1.0 ? 1.0 : (cs < -1.0 ? -1.0 : cs);
            const double ang = 0.5 * acos(cs);
            best = ang > best ? ang : best;
        }
    return best;
}

// cell c of the anchor's 7: the corner bits (bit0 +x, bit1 +y, bit2 +z) of its vertices
__constant__ unsigned char CA_CELL_MASK[7] = {0x03, 0x05, 0x11, 0x0F, 0x33, 0x55, 0xFF}; // bit c set = corner c belongs
// corners: 0=(0,0,0) 1=(1,0,0) 2=(0,1,0) 3=(1,1,0) 4=(0,0,1) 5=(1,0,1) 6=(0,1,1) 7=(1,1,1)
// edges +x {0,1}, +y {0,2}, +z {0,4}; faces xy {0,1,2,3}, xz {0,1,4,5}, yz {0,2,4,6}; cube all

__global__ void __launch_bounds__(128)
    k_circum_angle(const u32* __restrict__ bits, int wr, int nx, int ny, int nz, int z0, int zc, int zlo, int za, int zb,
                   const int* __restrict__ id, const float4* __restrict__ site, double* __restrict__ circ, double* __restrict__ ang)
{
    const size_t plane = (size_t)nx * ny, nv = plane * (size_t)(zb - za);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv)
        return;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = za + (int)(i / plane);
    int in[8], sid[8];
    D3 sp[8];
    for (int c = 0; c < 8; ++c)
    {
        const int xx = x + (c & 1), yy = y + ((c >> 1) & 1), zz = z + (c >> 2);
        const bool ok = xx < nx && yy < ny && zz < nz && zz < zc;
        in[c] = ok ? (int)((__ldg(bits + ((size_t)(zz - zlo) * ny + yy) * (size_t)wr + (xx >> 5)) >> (xx & 31)) & 1u) : 0;
        sid[c] = -1;
        if (in[c] && (c == 0 || in[0]))
        {
            sid[c] = __ldg(id + (size_t)(zz - z0) * plane + (size_t)yy * nx + xx);
            const float4 s = __ldg(site + sid[c]);
            sp[c] = D3{(double)s.x, (double)s.y, (double)s.z};
        }
    }
    for (int cell = 0; cell < 7; ++cell)
    {
        const unsigned mask = CA_CELL_MASK[cell];
        bool valid = true;
        D3 m{0, 0, 0};
        int nvert = 0;
        for (int c = 0; c < 8; ++c)
            if ((mask >> c) & 1u)
            {
                valid = valid && in[c];
                m.x += x + (c & 1);
                m.y += y + ((c >> 1) & 1);
                m.z += z + (c >> 2);
                ++nvert;
            }
        double r = 0.0, a = 0.0;
        if (valid)
        {
            D3 P[8];
            int ids[8], n = 0;
            for (int c = 0; c < 8; ++c)
                if ((mask >> c) & 1u)
                {
                    bool seen = false;
                    for (int k = 0; k < n; ++k)
                        seen = seen || ids[k] == sid[c];
                    if (!seen)
                    {
                        ids[n] = sid[c];
                        P[n++] = sp[c];
                    }
                }
            m.x /= nvert, m.y /= nvert, m.z /= nvert;
            r = seb_radius(P, n);
            a = object_angle(P, n, m);
        }
        circ[(size_t)cell * nv + i] = r;
        ang[(size_t)cell * nv + i] = a;
    }
}

extern "C" int vc_cell_circum_angle_grid(vc_ctx* c, int za, int zb, double* circum7, double* angle7)
{
    if (!c || !circum7 || !angle7)
        return VC_ERR_INVALID;
    VC_CUDA(c, cudaSetDevice(c->device));
    if (!c->have_closest || !c->have_inside || !c->lattice)
        return vc_fail(c, VC_ERR_STATE, "vc_cell_circum_angle_grid needs vc_classify_grid and the closest sites of a lattice site set");
    if (za < c->z0 || zb > c->z1 || za > zb)
        return vc_fail(c, VC_ERR_INVALID, "vc_cell_circum_angle_grid: plane range outside this ctx's owned planes");
    const size_t nv = (size_t)c->nx * c->ny * (size_t)(zb - za);
    if (nv == 0)
        return VC_OK;
    const bool dev = vc_is_device_ptr(circum7);
    if (dev != vc_is_device_ptr(angle7))
        return vc_fail(c, VC_ERR_INVALID, "vc_cell_circum_angle_grid: both outputs must be host or both device pointers");
    DevBuf tmp;
    double *dc = circum7, *da = angle7;
    if (!dev)
    {
        VC_CUDA(c, tmp.ensure(nv * 14 * sizeof(double)));
        dc = tmp.as<double>();
        da = dc + nv * 7;
    }
    VC_LAUNCH(c, "circum_angle", k_circum_angle, vc_blocks(nv, 128), 128, 0, c->bits.as<u32>(), c->wr, c->nx, c->ny, c->nz, c->z0,
              c->zc, c->zlo, za, zb, c->id.as<int>(), c->site_xyz.as<float4>(), dc, da);
    cudaError_t e = cudaGetLastError();
    if (!dev)
    {
        e = e == cudaSuccess ? cudaMemcpyAsync(circum7, dc, nv * 7 * sizeof(double), cudaMemcpyDeviceToHost, c->stream) : e;
        e = e == cudaSuccess ? cudaMemcpyAsync(angle7, da, nv * 7 * sizeof(double), cudaMemcpyDeviceToHost, c->stream) : e;
    }
    e = e == cudaSuccess ? cudaStreamSynchronize(c->stream) : e;
    tmp.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "vc_cell_circum_angle_grid", e);
    return VC_OK;
}
