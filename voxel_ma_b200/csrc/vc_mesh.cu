// vc_mesh.cu -- stage 1': inside/outside classification of the voxel grid straight from a closed
// triangle mesh (SURVEY section 8a/8f "K1'": new surface, no reference behaviour -- the reference only
// ever sees an already voxelised volume).  Parity rule: voxel centre (i,j,k) is inside iff the ray
// from it along +x crosses the surface an odd number of times.
//
// Everything that decides a flag is integer arithmetic (vc_mesh_core.h), so the CPU restatement in
// oracle/oracle.c reproduces the flags bit for bit:
//   1. vertices: q = M*p exactly like VoroInfo::tagVert's transform (double 4x4, homogeneous divide,
//      cast to float -- XForm.h:479-489), then snapped to 1/256 voxel: Q = floor(256 q + 0.5).
//   2. a triangle covers the grid column (j,k) iff the point (256 j, 256 k) lies inside its (y,z)
//      projection; points on an edge belong to exactly one side (vc_mesh_covers: the top-left rule,
//      equivalent to shifting every query by (+eps^2, -eps)), so a closed mesh is counted
//      consistently along shared edges and at vertices.
//   3. the crossing abscissa is the exact rational x_c = num / D; it flips every voxel i with
//      256 i < x_c, i.e. the T = ceil(num / (256 D)) first voxels of the row (clamped to [0,nx]).
//
// Two kernels:
//   k_mesh_toggles   one lane per triangle for the setup; triangles whose (y,z) box holds only a few
//                    columns are finished by their own lane, the large ones are elected with a warp
//                    ballot and rasterised by all 32 lanes, the very large ones are queued for
//                    k_mesh_large (one block each).  Each crossing is one atomicXor of bit T
//                    in the toggle row of its column (rows of nx+1 bits, the layout of ctx->bits).
//   k_mesh_parity    one warp per row: inside(i) = parity of the toggles above i = an exclusive
//                    suffix XOR -- inside a word by shifts, across the words of the row by a warp
//                    ballot of the word parities; writes the occupancy bit row in place and the byte
//                    flags with 128-bit stores.
#include "vc_internal.h"
#include "vc_mesh_core.h"

struct MeshXf
{
    double m[16];
};

// vertices -> quantised voxel-space coordinates (int x 3); flags[0] |= 1 when a vertex is out of range
__global__ void __launch_bounds__(256) k_mesh_quantise(const float* __restrict__ v, int64_t nv, MeshXf M, int* __restrict__ q, int* __restrict__ flags)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv)
        return;
    const double* xf = M.m;
    double v0 = v[3 * i], v1 = v[3 * i + 1], v2 = v[3 * i + 2];
#define ROW(a, b, cc, d) __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xf[a], v0), __dmul_rn(xf[b], v1)), __dmul_rn(xf[cc], v2)), xf[d])
    double h = __ddiv_rn(1.0, ROW(3, 7, 11, 15));
    float p[3];
    p[0] = __double2float_rn(__dmul_rn(h, ROW(0, 4, 8, 12)));
    p[1] = __double2float_rn(__dmul_rn(h, ROW(1, 5, 9, 13)));
    p[2] = __double2float_rn(__dmul_rn(h, ROW(2, 6, 10, 14)));
#undef ROW
    bool bad = false;
    for (int d = 0; d < 3; ++d)
    {
        int Q;
        bad |= !vc_mesh_snap(p[d], &Q);
        q[3 * i + d] = Q;
    }
    if (bad)
        atomicOr(flags, 1);
}

#define MESH_SMALL 6   // columns a lane finishes on its own before the warp takes the large triangles
#define MESH_HUGE 1024 // columns beyond which a triangle gets a whole block (k_mesh_large)

struct TriSetup
{
    int ax, ay, az, bx, by, bz, cx, cy, cz; // quantised, counter-clockwise in (y,z)
    int j0, j1, k0, k1;                     // covered grid columns (inclusive), empty when j0 > j1
};

__device__ __forceinline__ void mesh_column(const TriSetup& t, int j, int k, int nx, int ny, int zlo, int wr, u32* __restrict__ tog)
{
    int T;
    if (vc_mesh_crossing(t.ax, t.ay, t.az, t.bx, t.by, t.bz, t.cx, t.cy, t.cz, j, k, nx, &T) && T > 0)
        atomicXor(tog + ((size_t)(k - zlo) * ny + j) * (size_t)wr + (T >> 5), 1u << (T & 31));
}

__global__ void __launch_bounds__(256) k_mesh_toggles(const int* __restrict__ q, const u32* __restrict__ tris, int64_t nt, int64_t nv,
                                                      int nx, int ny, int zlo, int zhi, int wr, u32* __restrict__ tog,
                                                      int* __restrict__ flags, u32* __restrict__ big_list, u32* __restrict__ nbig)
{
    const int64_t ti = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    TriSetup t;
    t.j0 = t.k0 = 0;
    t.j1 = t.k1 = -1;
    if (ti < nt)
    {
        const u32 ia = tris[3 * ti], ib = tris[3 * ti + 1], ic = tris[3 * ti + 2];
        if (ia >= nv || ib >= nv || ic >= nv)
            atomicOr(flags, 2);
        else
        {
            t.ax = q[3 * ia], t.ay = q[3 * ia + 1], t.az = q[3 * ia + 2];
            t.bx = q[3 * ib], t.by = q[3 * ib + 1], t.bz = q[3 * ib + 2];
            t.cx = q[3 * ic], t.cy = q[3 * ic + 1], t.cz = q[3 * ic + 2];
            if (vc_mesh_orient_ccw(&t.ax, &t.ay, &t.az, &t.bx, &t.by, &t.bz, &t.cx, &t.cy, &t.cz))
                vc_mesh_columns(t.ay, t.az, t.by, t.bz, t.cy, t.cz, ny, zlo, zhi, &t.j0, &t.j1, &t.k0, &t.k1);
        }
    }
    const int nj = t.j1 - t.j0 + 1, nk = t.k1 - t.k0 + 1;
    const bool some = nj > 0 && nk > 0;
    const bool large = some && (long)nj * nk > MESH_SMALL;
    if (some && !large)
        for (int k = t.k0; k <= t.k1; ++k)
            for (int j = t.j0; j <= t.j1; ++j)
                mesh_column(t, j, k, nx, ny, zlo, wr, tog);
    // triangles whose column box holds more than a warp's worth of work are queued for k_mesh_large, which gives each a
    // whole block (a mesh of few, large triangles would otherwise leave most of the GPU idle: 2052 triangles = 65 warps)
    const bool huge = large && (long)nj * nk > MESH_HUGE;
    if (huge)
        big_list[atomicAdd(nbig, 1u)] = (u32)ti;
    // large triangles: one at a time, the whole warp over its column box
    unsigned todo = __ballot_sync(0xFFFFFFFFu, large && !huge);
    while (todo)
    {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        TriSetup s;
        s.ax = __shfl_sync(0xFFFFFFFFu, t.ax, src), s.ay = __shfl_sync(0xFFFFFFFFu, t.ay, src), s.az = __shfl_sync(0xFFFFFFFFu, t.az, src);
        s.bx = __shfl_sync(0xFFFFFFFFu, t.bx, src), s.by = __shfl_sync(0xFFFFFFFFu, t.by, src), s.bz = __shfl_sync(0xFFFFFFFFu, t.bz, src);
        s.cx = __shfl_sync(0xFFFFFFFFu, t.cx, src), s.cy = __shfl_sync(0xFFFFFFFFu, t.cy, src), s.cz = __shfl_sync(0xFFFFFFFFu, t.cz, src);
        s.j0 = __shfl_sync(0xFFFFFFFFu, t.j0, src), s.j1 = __shfl_sync(0xFFFFFFFFu, t.j1, src);
        s.k0 = __shfl_sync(0xFFFFFFFFu, t.k0, src), s.k1 = __shfl_sync(0xFFFFFFFFu, t.k1, src);
        const int w = s.j1 - s.j0 + 1;
        const long total = (long)w * (s.k1 - s.k0 + 1);
        for (long p = lane; p < total; p += 32)
        {
            const int k = s.k0 + (int)(p / w), j = s.j0 + (int)(p % w);
            mesh_column(s, j, k, nx, ny, zlo, wr, tog);
        }
    }
}

// the queued triangles: one block per triangle (block-stride over the queue), all threads over its column box
__global__ void __launch_bounds__(256) k_mesh_large(const int* __restrict__ q, const u32* __restrict__ tris, const u32* __restrict__ big_list,
                                                    const u32* __restrict__ nbig, int nx, int ny, int zlo, int zhi, int wr,
                                                    u32* __restrict__ tog)
{
    const u32 n = *nbig;
    for (u32 idx = blockIdx.x; idx < n; idx += gridDim.x)
    {
        const int64_t ti = big_list[idx];
        const u32 ia = tris[3 * ti], ib = tris[3 * ti + 1], ic = tris[3 * ti + 2]; // indices were validated when queued
        TriSetup t;
        t.ax = q[3 * ia], t.ay = q[3 * ia + 1], t.az = q[3 * ia + 2];
        t.bx = q[3 * ib], t.by = q[3 * ib + 1], t.bz = q[3 * ib + 2];
        t.cx = q[3 * ic], t.cy = q[3 * ic + 1], t.cz = q[3 * ic + 2];
        vc_mesh_orient_ccw(&t.ax, &t.ay, &t.az, &t.bx, &t.by, &t.bz, &t.cx, &t.cy, &t.cz);
        vc_mesh_columns(t.ay, t.az, t.by, t.bz, t.cy, t.cz, ny, zlo, zhi, &t.j0, &t.j1, &t.k0, &t.k1);
        const int w = t.j1 - t.j0 + 1;
        const long total = (long)w * (t.k1 - t.k0 + 1);
        for (long p = threadIdx.x; p < total; p += blockDim.x)
        {
            const int k = t.k0 + (int)(p / w), j = t.j0 + (int)(p % w);
            mesh_column(t, j, k, nx, ny, zlo, wr, tog);
        }
    }
}

// one warp per row; words are visited 32 at a time from the top of the row so the parity of everything
// above carries downwards
__global__ void __launch_bounds__(256) k_mesh_parity(u32* __restrict__ bits, u8* __restrict__ inside, size_t nrows, int nx, int wr)
{
    const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows)
        return;
    u32* r = bits + row * (size_t)wr;
    u8* out = inside + row * (size_t)nx;
    unsigned carry = 0; // parity of all toggles in the words above the current group
    for (int base = ((wr - 1) / 32) * 32; base >= 0; base -= 32)
    {
        const int w = base + lane;
        const u32 x = w < wr ? r[w] : 0u;
        const unsigned odd = __ballot_sync(0xFFFFFFFFu, __popc(x) & 1);
        // toggles above bit b: the higher bits of this word, the higher words of the group, the carry
        u32 s = x >> 1;
        s ^= s >> 1;
        s ^= s >> 2;
        s ^= s >> 4;
        s ^= s >> 8;
        s ^= s >> 16;
        const unsigned above = lane == 31 ? 0u : odd >> (lane + 1);
        if ((__popc(above) + carry) & 1)
            s = ~s;
        carry = (carry + __popc(odd)) & 1;
        if (w < wr)
        {
            const int cnt = nx - 32 * w; // voxels of the row in this word
            if (cnt <= 0)
                s = 0u;
            else if (cnt < 32)
                s &= (1u << cnt) - 1u;
            r[w] = s;
            if (cnt >= 32 && (nx & 15) == 0)
            {
                uint4 lo, hi;
                u32* pl = &lo.x;
                u32* ph = &hi.x;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                {
                    const u32 n0 = (s >> (4 * g)) & 15u, n1 = (s >> (16 + 4 * g)) & 15u;
                    pl[g] = (n0 & 1u) | ((n0 & 2u) << 7) | ((n0 & 4u) << 14) | ((n0 & 8u) << 21);
                    ph[g] = (n1 & 1u) | ((n1 & 2u) << 7) | ((n1 & 4u) << 14) | ((n1 & 8u) << 21);
                }
                *reinterpret_cast<uint4*>(out + 32 * (size_t)w) = lo;
                *reinterpret_cast<uint4*>(out + 32 * (size_t)w + 16) = hi;
            }
            else
                for (int b = 0; b < cnt && b < 32; ++b)
                    out[32 * (size_t)w + b] = (s >> b) & 1u;
        }
    }
}

int st_classify_mesh(vc_ctx* c, const float* verts, int64_t nv, const uint32_t* tris, int64_t nt, const double* M)
{
    if (!c->have_grid)
        return vc_fail(c, VC_ERR_STATE, "vc_classify_mesh: call vc_set_grid first");
    static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    MeshXf xf;
    for (int i = 0; i < 16; ++i)
        xf.m[i] = M ? M[i] : I[i];
    // resident voxel planes: the slab plus one halo plane each side, as for an uploaded volume
    c->zlo = c->z0 > 0 ? c->z0 - 1 : 0;
    c->zhi = c->z1 < c->nz ? c->z1 + 1 : c->nz;
    c->have_vol = false;
    c->have_inside = c->have_sites = c->have_closest = c->have_measures = false;
    const size_t nrows = (size_t)c->ny * (size_t)(c->zhi - c->zlo);
    c->wr = c->nx / 32 + 1;
    VC_CUDA(c, c->inside.ensure(nrows * (size_t)c->nx + 16));
    VC_CUDA(c, c->bits.ensure(nrows * (size_t)c->wr * 4 + 16));
    VC_CUDA(c, c->scratch.ensure(256));
    VC_CUDA(c, cudaMemsetAsync(c->bits.p, 0, nrows * (size_t)c->wr * 4, c->stream));
    VC_CUDA(c, cudaMemsetAsync(c->scratch.p, 0, 80, c->stream)); // flags at [0], the queue length of k_mesh_large at byte 64
    DevBuf dv, dq, dt, dl;
    cudaError_t e = cudaSuccess;
    const float* pv = verts;
    const u32* pt = tris;
    if (nv > 0 && !vc_is_device_ptr(verts))
    {
        if ((e = dv.ensure((size_t)nv * 12)) == cudaSuccess)
            e = cudaMemcpyAsync(dv.p, verts, (size_t)nv * 12, cudaMemcpyHostToDevice, c->stream);
        pv = dv.as<float>();
    }
    if (e == cudaSuccess && nt > 0 && !vc_is_device_ptr(tris))
    {
        if ((e = dt.ensure((size_t)nt * 12)) == cudaSuccess)
            e = cudaMemcpyAsync(dt.p, tris, (size_t)nt * 12, cudaMemcpyHostToDevice, c->stream);
        pt = dt.as<u32>();
    }
    if (e == cudaSuccess)
        e = dq.ensure((size_t)(nv > 0 ? nv : 1) * 12);
    if (e == cudaSuccess)
        e = dl.ensure((size_t)(nt > 0 ? nt : 1) * 4);
    int flags = 0;
    if (e == cudaSuccess)
    {
        int* dflags = c->scratch.as<int>();
        if (nv > 0)
            VC_LAUNCH(c, "mesh_quantise", k_mesh_quantise, vc_blocks((size_t)nv, 256), 256, 0, pv, nv, xf, dq.as<int>(), dflags);
        if (nt > 0 && nv > 0)
        {
            u32* nbig = (u32*)c->scratch.p + 16;
            VC_LAUNCH(c, "mesh_toggles", k_mesh_toggles, vc_blocks((size_t)nt, 256), 256, 0, dq.as<int>(), pt, nt, nv, c->nx, c->ny,
                      c->zlo, c->zhi, c->wr, c->bits.as<u32>(), dflags, dl.as<u32>(), nbig);
            VC_LAUNCH(c, "mesh_large", k_mesh_large, c->sm_count * 8, 256, 0, dq.as<int>(), pt, dl.as<u32>(), nbig, c->nx, c->ny, c->zlo,
                      c->zhi, c->wr, c->bits.as<u32>());
        }
        VC_LAUNCH(c, "mesh_parity", k_mesh_parity, vc_blocks(nrows * 32, 256), 256, 0, c->bits.as<u32>(), c->inside.as<u8>(), nrows,
                  c->nx, c->wr);
        e = cudaMemcpyAsync(&flags, dflags, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess)
        e = cudaGetLastError();
    dv.release();
    dq.release();
    dt.release();
    dl.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "vc_classify_mesh", e);
    if (flags & 2)
        return vc_fail(c, VC_ERR_INVALID, "vc_classify_mesh: a triangle refers to a vertex index >= nv");
    if (flags & 1)
        return vc_fail(c, VC_ERR_INVALID, "vc_classify_mesh: a vertex lies outside the supported range [-1024, 3072) voxels (or is not finite)");
    c->have_inside = true;
    return VC_OK;
}
