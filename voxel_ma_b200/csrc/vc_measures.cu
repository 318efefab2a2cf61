// vc_measures.cu -- stage 3: medial measures, plus the small per-element operators of the
// Voronoi-complex side of the boundary (tagVert, face lambda, vertex radii, max aggregation).
//
// Arithmetic follows the reference in float32, operation by operation, with explicit
// round-to-nearest intrinsics so nvcc cannot contract a multiply-add into an FMA:
//   MeasureForMA::lambdaForFace = trimesh::dist (include/measureforMA_imp.h:1-4,
//   3rdparty/trimesh2/include/Vec.h:1128-1143):  d2 = sqr(b0-a0); d2 += sqr(b1-a1); d2 += sqr(b2-a2); sqrt.
#include "vc_internal.h"

__device__ __forceinline__ float vc_dist2f(float ax, float ay, float az, float bx, float by, float bz)
{
    float t = __fsub_rn(bx, ax);
    float d2 = __fmul_rn(t, t);
    t = __fsub_rn(by, ay);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    t = __fsub_rn(bz, az);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    return d2;
}

// =============================================================================================
// K5  dense cell measures (dictionary of SURVEY section 0).  The 7 cells anchored at grid vertex v
// need the closest sites of the 8 vertices of v's cube.  sqrt is monotone and correctly rounded,
// so max over edges of sqrt(d2) == sqrt(max d2): the 12 edge terms are reduced as squared
// distances and only the 7 outputs (+ radius) take a square root.
//
// A block owns a 32 x 8 (x,y) tile and marches up a chunk of z planes.  Per plane it fetches the
// (32+1) x (8+1) ids once (id 4 B + flag 1 B per vertex: the algorithmic read), gathers each id's
// site from the L2-resident table once, and parks (site, flags) in shared memory; a thread keeps
// the 4 records of its own column for plane z in registers and reads the 4 of plane z+1 from
// shared memory, so the 8 x 4 B outputs per vertex are the only HBM traffic that scales.
// The id / flag loads of plane z+2 are issued before plane z+1 is consumed.
// id: planes [z0, zc) (zc = z1+1 halo when z1 < nz); inside: planes [zlo, zhi).
// =============================================================================================
#define CM_TX 32
#define CM_TY 8
#define CM_ENT ((CM_TX + 1) * (CM_TY + 1)) // 297 records per plane
#define CM_PER ((CM_ENT + CM_TX * CM_TY - 1) / (CM_TX * CM_TY)) // records a thread loads per plane (2)

struct CmRec
{
    float x, y, z;
    u32 f; // bit 0: vertex exists and has a site; bit 1: vertex is inside
};

__device__ __forceinline__ float cm_e2(const CmRec& a, const CmRec& b)
{
    return (a.f & b.f & 1u) ? vc_dist2f(a.x, a.y, a.z, b.x, b.y, b.z) : 0.0f;
}

__global__ void __launch_bounds__(CM_TX* CM_TY, 4)
    k_cell_measures(const int* __restrict__ id, const u8* __restrict__ inside, const float4* __restrict__ site,
                    int nx, int ny, int z0, int z1, int zc, int zlo, int zchunk, float* __restrict__ edge3,
                    float* __restrict__ face3, float* __restrict__ cube, float* __restrict__ radius)
{
    __shared__ float4 tile[2][CM_TY + 1][CM_TX + 1];
    const int tx = threadIdx.x & (CM_TX - 1), ty = threadIdx.x / CM_TX;
    const int x0 = blockIdx.x * CM_TX, y0 = blockIdx.y * CM_TY;
    const int zs = z0 + blockIdx.z * zchunk, ze = min(zs + zchunk, z1);
    const int x = x0 + tx, y = y0 + ty;
    const size_t plane = (size_t)nx * ny;
    const size_t nv = plane * (size_t)(z1 - z0);

    // the records this thread loads each plane: e = threadIdx.x + k*256 -> (hy, hx) of the halo tile
    size_t lo[CM_PER];
    int hyx[CM_PER];
    bool lok[CM_PER];
#pragma unroll
    for (int k = 0; k < CM_PER; ++k)
    {
        int e = threadIdx.x + k * CM_TX * CM_TY;
        int hy = e / (CM_TX + 1), hx = e - hy * (CM_TX + 1);
        lok[k] = e < CM_ENT && x0 + hx < nx && y0 + hy < ny;
        hyx[k] = e < CM_ENT ? hy * (CM_TX + 1) + hx : -1;
        lo[k] = (size_t)(x0 + hx) + (size_t)nx * (size_t)(y0 + hy);
    }
    int sid[CM_PER];
    u32 fin[CM_PER];
    auto fetch_ids = [&](int zz)
    {
#pragma unroll
        for (int k = 0; k < CM_PER; ++k)
        {
            sid[k] = -1;
            fin[k] = 0;
            if (lok[k] && zz < zc)
            {
                sid[k] = __ldcs(id + lo[k] + plane * (size_t)(zz - z0));
                fin[k] = __ldg(inside + lo[k] + plane * (size_t)(zz - zlo)) ? 2u : 0u;
            }
        }
    };
    auto park = [&](int buf)
    {
#pragma unroll
        for (int k = 0; k < CM_PER; ++k)
            if (hyx[k] >= 0)
            {
                float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                u32 f = fin[k];
                if (sid[k] >= 0)
                {
                    r = __ldg(site + sid[k]);
                    f |= 1u;
                }
                r.w = __uint_as_float(f);
                (&tile[buf][0][0])[hyx[k]] = r;
            }
    };
    auto rec = [&](int buf, int dy, int dx) -> CmRec
    {
        float4 v = tile[buf][ty + dy][tx + dx];
        CmRec r;
        r.x = v.x;
        r.y = v.y;
        r.z = v.z;
        r.f = __float_as_uint(v.w);
        return r;
    };

    fetch_ids(zs);
    park(0);
    fetch_ids(zs + 1);
    __syncthreads();
    CmRec a0 = rec(0, 0, 0), a1 = rec(0, 0, 1), a2 = rec(0, 1, 0), a3 = rec(0, 1, 1);
    // Every one of the 7 cells anchored at v contains v, so a vertex outside the solid reports 0 for
    // all of them and none of its edge terms is ever used: the distance arithmetic runs only for
    // inside anchors (in-plane edges of plane z+1 are computed when that plane's anchor is inside,
    // and carried over to the next iteration).
    float ex0 = 0.0f, ex1 = 0.0f, ey0 = 0.0f, ey1 = 0.0f;
    if (a0.f & 2u)
    {
        ex0 = cm_e2(a0, a1);
        ex1 = cm_e2(a2, a3);
        ey0 = cm_e2(a0, a2);
        ey1 = cm_e2(a1, a3);
    }
    const bool live = x < nx && y < ny;
    size_t o = (size_t)x + (size_t)nx * (size_t)y + plane * (size_t)(zs - z0);
    for (int z = zs; z < ze; ++z, o += plane)
    {
        const int buf = (z - zs + 1) & 1;
        park(buf);          // plane z+1 (ids fetched one iteration ago)
        fetch_ids(z + 2);   // in flight while plane z is computed
        __syncthreads();
        const CmRec b0 = rec(buf, 0, 0), b1 = rec(buf, 0, 1), b2 = rec(buf, 1, 0), b3 = rec(buf, 1, 1);
        // x edges {0,1},{2,3},{4,5},{6,7}; y edges {0,2},{1,3},{4,6},{5,7}; z edges {0,4},{1,5},{2,6},{3,7}
        float ex2 = 0.0f, ex3 = 0.0f, ey2 = 0.0f, ey3 = 0.0f;
        if ((a0.f | b0.f) & 2u)
        {
            ex2 = cm_e2(b0, b1);
            ex3 = cm_e2(b2, b3);
            ey2 = cm_e2(b0, b2);
            ey3 = cm_e2(b1, b3);
        }
        float le0 = 0.0f, le1 = 0.0f, le2 = 0.0f, lf0 = 0.0f, lf1 = 0.0f, lf2 = 0.0f, lc = 0.0f;
        if (a0.f & 2u)
        {
            const float w0 = cm_e2(a0, b0), w1 = cm_e2(a1, b1), w2 = cm_e2(a2, b2), w3 = cm_e2(a3, b3);
            const u32 i1 = a1.f & 2u, i2 = a2.f & 2u, i3 = a3.f & 2u;
            const u32 i4 = b0.f & 2u, i5 = b1.f & 2u, i6 = b2.f & 2u, i7 = b3.f & 2u;
            if (i1)
                le0 = __fsqrt_rn(ex0);
            if (i2)
                le1 = __fsqrt_rn(ey0);
            if (i4)
                le2 = __fsqrt_rn(w0);
            if (i1 & i2 & i3)
                lf0 = __fsqrt_rn(fmaxf(fmaxf(ex0, ex1), fmaxf(ey0, ey1)));
            if (i1 & i4 & i5)
                lf1 = __fsqrt_rn(fmaxf(fmaxf(ex0, ex2), fmaxf(w0, w1)));
            if (i2 & i4 & i6)
                lf2 = __fsqrt_rn(fmaxf(fmaxf(ey0, ey2), fmaxf(w0, w2)));
            if (i1 & i2 & i3 & i4 & i5 & i6 & i7)
            {
                float m = fmaxf(fmaxf(fmaxf(ex0, ex1), fmaxf(ex2, ex3)), fmaxf(fmaxf(ey0, ey1), fmaxf(ey2, ey3)));
                lc = __fsqrt_rn(fmaxf(m, fmaxf(fmaxf(w0, w1), fmaxf(w2, w3))));
            }
        }
        if (live)
        {
            if (edge3)
            {
                __stcs(edge3 + o, le0);
                __stcs(edge3 + nv + o, le1);
                __stcs(edge3 + 2 * nv + o, le2);
            }
            if (face3)
            {
                __stcs(face3 + o, lf0);
                __stcs(face3 + nv + o, lf1);
                __stcs(face3 + 2 * nv + o, lf2);
            }
            if (cube)
                __stcs(cube + o, lc);
            if (radius)
                __stcs(radius + o,
                       (a0.f & 1u) ? __fsqrt_rn(vc_dist2f(a0.x, a0.y, a0.z, (float)x, (float)y, (float)z)) : 0.0f);
        }
        a0 = b0;
        a1 = b1;
        a2 = b2;
        a3 = b3;
        ex0 = ex2;
        ex1 = ex3;
        ey0 = ey2;
        ey1 = ey3;
    }
}

int st_measures(vc_ctx* c, bool want_radius)
{
    if (!c->have_closest || !c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "vc_cell_measures_grid needs vc_classify_grid and vc_closest_grid");
    if (c->zhi < c->zc)
        return vc_fail(c, VC_ERR_STATE, "inside flags do not cover the halo plane");
    const size_t nv = (size_t)c->nx * c->ny * (size_t)(c->z1 - c->z0);
    VC_CUDA(c, c->edge3.ensure(nv * 3 * 4));
    VC_CUDA(c, c->face3.ensure(nv * 3 * 4));
    VC_CUDA(c, c->cube.ensure(nv * 4));
    if (want_radius)
        VC_CUDA(c, c->radius.ensure(nv * 4));
    // z chunks: long enough to amortise the one extra plane a chunk loads, short enough that the grid
    // is several waves of the SMs
    const int nplanes = c->z1 - c->z0;
    const unsigned gx = (c->nx + CM_TX - 1) / CM_TX, gy = (c->ny + CM_TY - 1) / CM_TY;
    int zchunk = 32;
    while (zchunk > 4 && (size_t)gx * gy * ((nplanes + zchunk - 1) / zchunk) < (size_t)c->sm_count * 16)
        zchunk >>= 1;
    dim3 grid(gx, gy, (nplanes + zchunk - 1) / zchunk);
    if (grid.y > 65535u || grid.z > 65535u)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "grid too large for the measures kernel");
    VC_LAUNCH(c, "cell_measures", k_cell_measures, grid, CM_TX * CM_TY, 0, c->id.as<int>(), c->inside.as<u8>(),
              c->site_xyz.as<float4>(), c->nx, c->ny, c->z0, c->z1, c->zc, c->zlo, zchunk, c->edge3.as<float>(),
              c->face3.as<float>(), c->cube.as<float>(), want_radius ? c->radius.as<float>() : nullptr);
    VC_CUDA(c, cudaGetLastError());
    c->have_measures = true;
    return VC_OK;
}

// =============================================================================================
// a4  VoroInfo::tagVert (src/voroinfo.cpp:447-454): q = M*p in double with the homogeneous divide
// of XForm.h:479-489, cast to float, round half away from zero, bounds, flag lookup.
// =============================================================================================
struct Mat16
{
    double m[16];
};

__global__ void k_classify_points(const float* __restrict__ xyz, int64_t n, Mat16 M, const u8* __restrict__ inside,
                                  int nx, int ny, int nz, u8* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double* xf = M.m;
    double v0 = xyz[3 * i], v1 = xyz[3 * i + 1], v2 = xyz[3 * i + 2];
#define ROW(a, b, cc, d) __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xf[a], v0), __dmul_rn(xf[b], v1)), __dmul_rn(xf[cc], v2)), xf[d])
    double h = __ddiv_rn(1.0, ROW(3, 7, 11, 15));
    float q0 = __double2float_rn(__dmul_rn(h, ROW(0, 4, 8, 12)));
    float q1 = __double2float_rn(__dmul_rn(h, ROW(1, 5, 9, 13)));
    float q2 = __double2float_rn(__dmul_rn(h, ROW(2, 6, 10, 14)));
#undef ROW
    // (int)std::round(float): half away from zero (include/spaceinfo.h:100-102)
    int x = (int)roundf(q0), y = (int)roundf(q1), z = (int)roundf(q2);
    bool in = x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz;
    out[i] = in ? inside[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z)] : (u8)0;
}

int st_classify_points(vc_ctx* c, const float* xyz, int64_t n, const double* M, uint8_t* out)
{
    if (!c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "vc_classify_points needs vc_classify_grid first");
    if (c->zlo != 0 || c->zhi != c->nz)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "vc_classify_points needs the whole grid resident");
    if (n == 0)
        return VC_OK;
    Mat16 m;
    static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int i = 0; i < 16; ++i)
        m.m[i] = M ? M[i] : I[i];
    DevBuf din, dout;
    VC_CUDA(c, din.ensure((size_t)n * 12));
    cudaError_t e = dout.ensure((size_t)n);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(din.p, xyz, (size_t)n * 12, cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "classify_points", k_classify_points, vc_blocks((size_t)n, 256), 256, 0, din.as<float>(), n, m,
                  c->inside.as<u8>(), c->nx, c->ny, c->nz, dout.as<u8>());
        e = cudaMemcpyAsync(out, dout.p, (size_t)n, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    din.release();
    dout.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "classify_points", e);
    return VC_OK;
}

// =============================================================================================
// a7 / a8 on a Voronoi complex
// =============================================================================================
__global__ void k_face_lambda(const float4* __restrict__ site, const int2* __restrict__ pairs, int64_t nf, int64_t ns,
                              float* __restrict__ out)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf)
        return;
    int2 p = pairs[f];
    float r = 0.0f;
    if (p.x >= 0 && p.x < ns && p.y >= 0 && p.y < ns)
    {
        float4 a = __ldg(site + p.x), b = __ldg(site + p.y);
        r = __fsqrt_rn(vc_dist2f(a.x, a.y, a.z, b.x, b.y, b.z));
    }
    out[f] = r;
}

__global__ void k_vertex_radii(const float4* __restrict__ site, const float* __restrict__ v, const int* __restrict__ sv,
                               int64_t nv, int64_t ns, float* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv)
        return;
    int s = sv[i];
    float r = 0.0f;
    if (s >= 0 && s < ns)
    {
        float4 a = __ldg(site + s);
        r = __fsqrt_rn(vc_dist2f(a.x, a.y, a.z, v[3 * i], v[3 * i + 1], v[3 * i + 2])); // dist(site_p, v_p)
    }
    out[i] = r;
}

__global__ void k_segment_max(const int* __restrict__ off, const int* __restrict__ items, int64_t n,
                              const float* __restrict__ value, const u8* __restrict__ valid, float* __restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n)
        return;
    float m = 0.0f;
    for (int k = off[e]; k < off[e + 1]; ++k)
    {
        int f = items[k];
        if (valid && !valid[f])
            continue;
        float v = value[f];
        m = v > m ? v : m; // std::max(m, v)
    }
    out[e] = m;
}

// small helper: temp device arrays for host-pointer operators
struct Tmp
{
    std::vector<DevBuf> bufs;
    ~Tmp()
    {
        for (auto& b : bufs)
            b.release();
    }
    void* up(vc_ctx* c, const void* host, size_t bytes, cudaError_t& e)
    {
        bufs.emplace_back();
        if (e == cudaSuccess)
            e = bufs.back().ensure(bytes);
        if (e == cudaSuccess && host)
            e = cudaMemcpyAsync(bufs.back().p, host, bytes, cudaMemcpyDefault, c->stream);
        return bufs.back().p;
    }
};

int st_face_lambda(vc_ctx* c, const int32_t* pairs, int64_t nf, float* out)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_face_lambda needs sites");
    if (nf == 0)
        return VC_OK;
    Tmp t;
    cudaError_t e = cudaSuccess;
    int2* dp = (int2*)t.up(c, pairs, (size_t)nf * 8, e);
    float* dout = (float*)t.up(c, nullptr, (size_t)nf * 4, e);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "face_lambda", k_face_lambda, vc_blocks((size_t)nf, 256), 256, 0, c->site_xyz.as<float4>(), dp, nf,
                  c->nsites, dout);
        e = cudaMemcpyAsync(out, dout, (size_t)nf * 4, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "face_lambda", e);
    return VC_OK;
}

int st_vertex_radii(vc_ctx* c, const float* v, int64_t nv, const int32_t* site_of_v, float* out)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_vertex_radii needs sites");
    if (nv == 0)
        return VC_OK;
    Tmp t;
    cudaError_t e = cudaSuccess;
    float* dv = (float*)t.up(c, v, (size_t)nv * 12, e);
    int* ds = (int*)t.up(c, site_of_v, (size_t)nv * 4, e);
    float* dout = (float*)t.up(c, nullptr, (size_t)nv * 4, e);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "vertex_radii", k_vertex_radii, vc_blocks((size_t)nv, 256), 256, 0, c->site_xyz.as<float4>(), dv,
                  ds, nv, c->nsites, dout);
        e = cudaMemcpyAsync(out, dout, (size_t)nv * 4, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "vertex_radii", e);
    return VC_OK;
}

int st_segment_max(vc_ctx* c, const int32_t* off, const int32_t* items, int64_t n, const float* value,
                   int64_t nvalue, const uint8_t* valid, float* out)
{
    if (n == 0)
        return VC_OK;
    int32_t nitems = off[n];
    Tmp t;
    cudaError_t e = cudaSuccess;
    int* doff = (int*)t.up(c, off, (size_t)(n + 1) * 4, e);
    int* dit = (int*)t.up(c, items, (size_t)(nitems > 0 ? nitems : 1) * 4, e);
    float* dval = (float*)t.up(c, value, (size_t)(nvalue > 0 ? nvalue : 1) * 4, e);
    u8* dvalid = valid ? (u8*)t.up(c, valid, (size_t)(nvalue > 0 ? nvalue : 1), e) : nullptr;
    float* dout = (float*)t.up(c, nullptr, (size_t)n * 4, e);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "segment_max", k_segment_max, vc_blocks((size_t)n, 256), 256, 0, doff, dit, n, dval, dvalid, dout);
        e = cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "segment_max", e);
    return VC_OK;
}
