// vc_compact.cu -- the compact product: one record per INSIDE grid vertex.
//
// Everything the reference keeps of this front end lives on inside elements: Voronoi vertices tagged
// outside are dropped while the diagram is loaded (src/voroinfo.cpp:128-139), cells with an outside
// vertex are invalid and report measure 0 (include/voroinfo_imp.h:26-34, src/voroinfo.cpp:1460-1461,
// 1513-1514).  On the dense grid that means: all 7 measures anchored at an outside vertex are 0 by
// definition, so a caller on the other side of PCIe only needs the occupancy bit rows plus, for each
// inside vertex, (linear index, closest site id, 4d^2, 7 lambda, radius) -- 44 B per inside vertex
// instead of 41 B per grid vertex.  The dense planes are still computed in HBM (same kernels, same
// values: tests compare the records with the planes); only what crosses the bus is compacted.
//
// Order: ascending linear vertex index, fixed by an exclusive prefix over the popcounts of the
// occupancy bit rows -- no atomics, so the record order is deterministic and a z chunk of the
// pipeline owns one contiguous range of records that can start its copy as soon as the chunk is done.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "vc_internal.h"

// inside count of every bit row of the owned planes; cnt[nrows] = 0 so the scan leaves the total there
__global__ void __launch_bounds__(256)
    k_row_popc(const u32* __restrict__ bits, int wr, size_t nrows, u32* __restrict__ cnt)
{
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp > nrows)
        return;
    u32 s = 0;
    if (warp < nrows)
        for (int w = lane; w < wr; w += 32)
            s += __popc(__ldg(bits + warp * (size_t)wr + w));
    for (int o = 16; o; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0)
        cnt[warp] = s;
}

// one warp per bit row of planes [za, zb): gathers the records of the row's inside vertices from
// the dense planes into the compact arrays at rowpre[row] + rank within the row
__global__ void __launch_bounds__(256)
    k_compact_records(const u32* __restrict__ bits, int wr, int nx, int ny, size_t row_first, size_t row_end, size_t bits_row_off,
                      const u32* __restrict__ rowpre, size_t nv, const int* __restrict__ id, const u32* __restrict__ d2,
                      const float* __restrict__ edge3, const float* __restrict__ face3, const float* __restrict__ cube,
                      const float* __restrict__ radius, size_t cap, u32* __restrict__ cvert, int* __restrict__ cid,
                      u32* __restrict__ cd2, float* __restrict__ clam, float* __restrict__ crad)
{
    const size_t row = row_first + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= row_end)
        return;
    u32 base = rowpre[row];
    if (rowpre[row + 1] == base)
        return;
    const u32* brow = bits + (row + bits_row_off) * (size_t)wr;
    const size_t vrow = row * (size_t)nx; // linear index of (x = 0, y, z - z0)
    for (int w0 = 0; w0 < wr; w0 += 32)
    {
        const int w = w0 + lane;
        u32 word = w < wr ? __ldg(brow + w) : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        size_t pos = (size_t)base + (size_t)(incl - cnt);
        base += (u32)__shfl_sync(0xffffffffu, incl, 31);
        while (word)
        {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            const size_t v = vrow + 32 * (size_t)w + b;
            cvert[pos] = (u32)v;
            if (cid)
                cid[pos] = __ldg(id + v);
            if (cd2)
                cd2[pos] = __ldg(d2 + v);
            if (clam)
            {
                clam[pos] = __ldg(edge3 + v);
                clam[cap + pos] = __ldg(edge3 + nv + v);
                clam[2 * cap + pos] = __ldg(edge3 + 2 * nv + v);
                clam[3 * cap + pos] = __ldg(face3 + v);
                clam[4 * cap + pos] = __ldg(face3 + nv + v);
                clam[5 * cap + pos] = __ldg(face3 + 2 * nv + v);
                clam[6 * cap + pos] = __ldg(cube + v);
            }
            if (crad)
                crad[pos] = __ldg(radius + v);
            ++pos;
        }
    }
}

// ---- measures of the inside vertices only, straight into the records ---------------------------------
// Same arithmetic and the same validity rules as k_cell_measures (vc_measures.cu), evaluated per inside
// anchor vertex from the id / d2x4 planes and the site table, without materialising the 8 dense float
// planes (36 B per grid vertex of stores that are 0 wherever the anchor is outside).  One warp per bit
// row; the row's non-zero words are taken one after the other and the 32 vertices of a word are the 32 lanes.
__device__ __forceinline__ float vc_dist2f_c(float4 a, float4 b)
{
    float t = __fsub_rn(b.x, a.x);
    float d2 = __fmul_rn(t, t);
    t = __fsub_rn(b.y, a.y);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    t = __fsub_rn(b.z, a.z);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    return d2;
}

__global__ void __launch_bounds__(256)
    k_sparse_records(const u32* __restrict__ bits, int wr, int nx, int ny, int z0, int zc, int zlo, size_t row_first, size_t row_end,
                     const u32* __restrict__ rowpre, const int* __restrict__ id, const u32* __restrict__ d2,
                     const float4* __restrict__ site, int radius_from_d2, size_t cap, u32* __restrict__ cvert,
                     int* __restrict__ cid, u32* __restrict__ cd2, float* __restrict__ clam, float* __restrict__ crad)
{
    const size_t row = row_first + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= row_end)
        return;
    u32 base = rowpre[row];
    if (rowpre[row + 1] == base)
        return;
    const int y = (int)(row % (size_t)ny), z = z0 + (int)(row / (size_t)ny);
    const size_t plane = (size_t)nx * ny;
    const bool yok = y + 1 < ny, zok = z + 1 < zc;
    // bit rows (y+dy, z+dz), j = dy + 2*dz
    const u32* brow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        brow[j] = bits + ((size_t)(z + (j >> 1) - zlo) * ny + (size_t)(y + (j & 1))) * (size_t)wr;
    const bool rok[4] = {true, yok, zok, yok && zok};
    // word by word (warp-uniform), lane = bit: the 32 vertices of a word are evaluated in parallel
    for (int w = 0; w < wr; ++w)
    {
        const u32 word = __ldg(brow[0] + w); // same address in every lane: one broadcast load
        if (word == 0u)
            continue;
        u32 W[4] = {word, 0, 0, 0}, N[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (rok[j])
            {
                if (j)
                    W[j] = __ldg(brow[j] + w);
                N[j] = (w + 1 < wr) ? __ldg(brow[j] + w + 1) : 0u;
            }
        const int b = lane;
        const size_t pos = (size_t)base + (size_t)__popc(word & ((1u << b) - 1u));
        base += (u32)__popc(word);
        if (!((word >> b) & 1u))
            continue;
        const int x = 32 * w + b;
        // inside flag and record (site position, has-site flag) of the cube vertex k = dx + 2*dy + 4*dz
        bool in[8], has[8];
        float4 r[8];
        int own_id = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const int j = k >> 1, dx = k & 1;
            in[k] = dx == 0 ? ((W[j] >> b) & 1u) : (b < 31 ? ((W[j] >> (b + 1)) & 1u) : (N[j] & 1u));
            has[k] = false;
            r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[k])
            {
                const int sid = __ldg(id + (size_t)(x + dx) + (size_t)nx * (size_t)(y + ((k >> 1) & 1)) +
                                      plane * (size_t)(z + (k >> 2) - z0));
                if (k == 0)
                    own_id = sid;
                if (sid >= 0)
                {
                    r[k] = __ldg(site + sid);
                    has[k] = true;
                }
            }
        }
        auto e2 = [&](int a, int c) -> float { return (has[a] && has[c]) ? vc_dist2f_c(r[a], r[c]) : 0.0f; };
        // x edges {0,1},{2,3},{4,5},{6,7}; y edges {0,2},{1,3},{4,6},{5,7}; z edges {0,4},{1,5},{2,6},{3,7}
        const float ex0 = e2(0, 1), ex1 = e2(2, 3), ex2 = e2(4, 5), ex3 = e2(6, 7);
        const float ey0 = e2(0, 2), ey1 = e2(1, 3), ey2 = e2(4, 6), ey3 = e2(5, 7);
        const float z0e = e2(0, 4), z1e = e2(1, 5), z2e = e2(2, 6), z3e = e2(3, 7);
        float l[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (in[1])
            l[0] = __fsqrt_rn(ex0);
        if (in[2])
            l[1] = __fsqrt_rn(ey0);
        if (in[4])
            l[2] = __fsqrt_rn(z0e);
        if (in[1] && in[2] && in[3])
            l[3] = __fsqrt_rn(fmaxf(fmaxf(ex0, ex1), fmaxf(ey0, ey1)));
        if (in[1] && in[4] && in[5])
            l[4] = __fsqrt_rn(fmaxf(fmaxf(ex0, ex2), fmaxf(z0e, z1e)));
        if (in[2] && in[4] && in[6])
            l[5] = __fsqrt_rn(fmaxf(fmaxf(ey0, ey2), fmaxf(z0e, z2e)));
        if (in[1] && in[2] && in[3] && in[4] && in[5] && in[6] && in[7])
        {
            float m = fmaxf(fmaxf(fmaxf(ex0, ex1), fmaxf(ex2, ex3)), fmaxf(fmaxf(ey0, ey1), fmaxf(ey2, ey3)));
            l[6] = __fsqrt_rn(fmaxf(m, fmaxf(fmaxf(z0e, z1e), fmaxf(z2e, z3e))));
        }
        const size_t v = row * (size_t)nx + (size_t)x;
        const u32 q = __ldg(d2 + v);
        float rad = 0.0f;
        if (radius_from_d2 && q < (1u << 24))
            rad = __fsqrt_rn(__fmul_rn((float)q, 0.25f));
        else if (own_id >= 0)
            rad = __fsqrt_rn(vc_dist2f_c(r[0], make_float4((float)x, (float)y, (float)z, 0.f)));
        cvert[pos] = (u32)v;
        cid[pos] = own_id;
        cd2[pos] = q;
#pragma unroll
        for (int k = 0; k < 7; ++k)
            clam[(size_t)k * cap + pos] = l[k];
        crad[pos] = rad;
    }
}

// P[z] = rowpre[z * ny]: the record index at which plane z starts (z = 0 .. nplanes), gathered on the device so that
// ONE small copy brings it to the host (a strided 2-D copy of 4-byte rows costs hundreds of microseconds)
__global__ void k_plane_starts(const u32* __restrict__ rowpre, int ny, int nplanes, u32* __restrict__ P)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z <= nplanes)
        P[z] = rowpre[(size_t)z * ny];
}

// row prefix of the owned planes, queued on c->stream; the plane-boundary values P[0..nplanes] are
// copied to host_P (pinned) by the same stream -- valid after the caller's next synchronisation
static int compact_prefix_async(vc_ctx* c, u32* host_P)
{
    if (!c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "the compact product needs vc_classify_grid");
    const size_t nrows = (size_t)c->ny * (size_t)(c->z1 - c->z0);
    if ((size_t)c->nx * nrows >= (1ull << 32))
        return vc_fail(c, VC_ERR_UNSUPPORTED, "compact records index vertices with 32 bits: slab too large");
    const int nplanes = c->z1 - c->z0;
    VC_CUDA(c, c->rowpre.ensure((nrows + 2 + (size_t)nplanes + 2) * 4));
    const u32* bits = c->bits.as<u32>() + (size_t)(c->z0 - c->zlo) * c->ny * (size_t)c->wr;
    VC_LAUNCH(c, "row_popc", k_row_popc, vc_blocks((nrows + 1) * 32, 256), 256, 0, bits, c->wr, nrows, c->rowpre.as<u32>());
    VC_TRY(vc_exclusive_scan_u32(c, c->rowpre.as<u32>(), (int64_t)nrows + 1));
    if (host_P)
    {
        u32* P = c->rowpre.as<u32>() + nrows + 2;
        VC_LAUNCH(c, "plane_starts", k_plane_starts, vc_blocks((size_t)nplanes + 1, 256), 256, 0, c->rowpre.as<u32>(), c->ny, nplanes, P);
        VC_CUDA(c, cudaMemcpyAsync(host_P, P, ((size_t)nplanes + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    return VC_OK;
}

// device records: ONE block of 11 rows x ccap words -- vert | id | d2x4 | lambda 0..6 | radius -- so that a caller
// whose host arrays are the rows of one such block gets a z chunk's records in a single 2-D copy
static int compact_alloc(vc_ctx* c, int64_t n, bool want_radius)
{
    (void)want_radius;
    const size_t m = (size_t)(n > 0 ? n : 1);
    VC_CUDA(c, c->crec.ensure(m * 11 * 4));
    c->ccap = (int64_t)m;
    return VC_OK;
}
static inline u32* rec_vert(vc_ctx* c) { return c->crec.as<u32>(); }
static inline int* rec_id(vc_ctx* c) { return (int*)(c->crec.as<u32>() + (size_t)c->ccap); }
static inline u32* rec_d2(vc_ctx* c) { return c->crec.as<u32>() + 2 * (size_t)c->ccap; }
static inline float* rec_lam(vc_ctx* c) { return (float*)(c->crec.as<u32>() + 3 * (size_t)c->ccap); }
static inline float* rec_rad(vc_ctx* c) { return (float*)(c->crec.as<u32>() + 10 * (size_t)c->ccap); }

// records of the owned planes [za, zb) on stream c->cur (needs the row prefix and the dense planes)
static int compact_range(vc_ctx* c, int za, int zb, bool want_radius)
{
    const size_t r0 = (size_t)(za - c->z0) * c->ny, r1 = (size_t)(zb - c->z0) * c->ny;
    const size_t nv = (size_t)c->nx * c->ny * (size_t)(c->z1 - c->z0);
    VC_LAUNCH(c, "compact_records", k_compact_records, vc_blocks((r1 - r0) * 32, 256), 256, 0, c->bits.as<u32>(), c->wr, c->nx, c->ny,
              r0, r1, (size_t)(c->z0 - c->zlo) * c->ny, c->rowpre.as<u32>(), nv, c->id.as<int>(), c->d2.as<u32>(),
              c->edge3.as<float>(), c->face3.as<float>(), c->cube.as<float>(), want_radius ? c->radius.as<float>() : nullptr,
              (size_t)c->ccap, rec_vert(c), rec_id(c), rec_d2(c), rec_lam(c), want_radius ? rec_rad(c) : nullptr);
    VC_CUDA(c, cudaGetLastError());
    return VC_OK;
}

// the same records for planes [za, zb) computed directly (k_sparse_records): needs ids / d2, not the dense measure planes
static int sparse_range(vc_ctx* c, int za, int zb)
{
    const size_t r0 = (size_t)(za - c->z0) * c->ny, r1 = (size_t)(zb - c->z0) * c->ny;
    VC_LAUNCH(c, "sparse_records", k_sparse_records, vc_blocks((r1 - r0) * 32, 256), 256, 0, c->bits.as<u32>(), c->wr, c->nx, c->ny,
              c->z0, c->zc, c->zlo, r0, r1, c->rowpre.as<u32>(), c->id.as<int>(), c->d2.as<u32>(), c->site_xyz.as<float4>(),
              c->lattice ? 1 : 0, (size_t)c->ccap, rec_vert(c), rec_id(c), rec_d2(c), rec_lam(c), rec_rad(c));
    VC_CUDA(c, cudaGetLastError());
    return VC_OK;
}

// copies of records [a, b) to the caller's arrays (host or device) on stream s
static int compact_copy_out(vc_ctx* c, cudaStream_t s, size_t a, size_t b, int64_t cap, uint32_t* vert, int32_t* id,
                            uint32_t* d2x4, float* lambda7, float* radius)
{
    if (b <= a)
        return VC_OK;
    const size_t n = b - a;
    const size_t hc = (size_t)cap;
    if (vert && id && d2x4 && lambda7 && radius && (uint32_t*)id == vert + hc && d2x4 == vert + 2 * hc &&
        (uint32_t*)lambda7 == vert + 3 * hc && (uint32_t*)radius == vert + 10 * hc)
    { // the caller's arrays are the rows of one 11 x cap block: one copy instead of five
        VC_CUDA(c, cudaMemcpy2DAsync(vert + a, hc * 4, rec_vert(c) + a, (size_t)c->ccap * 4, n * 4, 11, cudaMemcpyDefault, s));
        return VC_OK;
    }
    if (vert)
        VC_CUDA(c, cudaMemcpyAsync(vert + a, rec_vert(c) + a, n * 4, cudaMemcpyDefault, s));
    if (id)
        VC_CUDA(c, cudaMemcpyAsync(id + a, rec_id(c) + a, n * 4, cudaMemcpyDefault, s));
    if (d2x4)
        VC_CUDA(c, cudaMemcpyAsync(d2x4 + a, rec_d2(c) + a, n * 4, cudaMemcpyDefault, s));
    if (lambda7)
        VC_CUDA(c, cudaMemcpy2DAsync(lambda7 + a, (size_t)cap * 4, rec_lam(c) + a, (size_t)c->ccap * 4, n * 4, 7,
                                     cudaMemcpyDefault, s));
    if (radius)
        VC_CUDA(c, cudaMemcpyAsync(radius + a, rec_rad(c) + a, n * 4, cudaMemcpyDefault, s));
    return VC_OK;
}

extern "C"
{
    int vc_set_compact_mode(vc_ctx* c, int mode)
    {
        if (!c || mode < 0 || mode > 2)
            return VC_ERR_INVALID;
        c->compact_mode = mode;
        return VC_OK;
    }

    int vc_compact_count(vc_ctx* c, int64_t* n_inside)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        u32* P = (u32*)c->pinned + 16;
        VC_TRY(compact_prefix_async(c, P));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        c->ninside = (int64_t)P[c->z1 - c->z0];
        if (n_inside)
            *n_inside = c->ninside;
        return VC_OK;
    }

    int vc_compact_records(vc_ctx* c, int64_t cap, uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (!c->have_measures || !c->have_closest)
            return vc_fail(c, VC_ERR_STATE, "vc_compact_records needs the dense planes (vc_run_dense / vc_closest_and_measures)");
        if (radius && !c->radius.p)
            return vc_fail(c, VC_ERR_STATE, "vc_compact_records: the radius plane was not computed");
        int64_t n = 0;
        VC_TRY(vc_compact_count(c, &n));
        if (cap < n)
            return vc_fail(c, VC_ERR_NOMEM, "vc_compact_records: capacity below the inside count (vc_compact_count)");
        VC_TRY(compact_alloc(c, n, radius != nullptr));
        VC_TRY(compact_range(c, c->z0, c->z1, radius != nullptr));
        VC_TRY(compact_copy_out(c, c->stream, 0, (size_t)n, cap, vert, id, d2x4, lambda7, radius));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    // Host volume in, compact product out, every copy inside the call:
    //   upload in plane chunks, each classified as it lands -> sites -> per z chunk of the pipeline:
    //   transform, measures, compaction, and the chunk's records start their way back while the next
    //   chunk computes.  One host synchronisation in the middle (site count + inside count).
    static int host_compact(vc_ctx* c, const void* vol, bool i8, uint32_t* inside_bits, int64_t cap, int64_t* n_inside,
                            uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius, int32_t* id_dense,
                            uint32_t* d2x4_dense, int64_t* nsites)
    {
        if (!c || !vol || cap < 0)
            return VC_ERR_INVALID;
        const size_t esz = i8 ? 1 : 4; // bytes per voxel of the host volume
        VC_CUDA(c, cudaSetDevice(c->device));
        if (!c->have_grid)
            return vc_fail(c, VC_ERR_STATE, "vc_run_dense_host_compact: call vc_set_grid first");
        const bool slab = c->z0 != 0 || c->z1 != c->nz;
        if (slab && c->peer_world < 2)
            return vc_fail(c, VC_ERR_STATE, "vc_run_dense_host_compact on a slab needs the peer group of its ranks (vc_peer_create / "
                                            "vc_peer_open)");
        // resident voxel planes: the slab plus one halo plane on each interior side (what vc_volume_upload_f32 takes)
        const int zlo = c->z0 > 0 ? c->z0 - 1 : 0, zhi = c->z1 < c->nz ? c->z1 + 1 : c->nz;
        const size_t plane = (size_t)c->nx * c->ny, nv = plane * (size_t)(c->z1 - c->z0);
        VC_CUDA(c, c->vol.ensure(plane * (size_t)(zhi - zlo) * esz + 64));
        c->zlo = zlo;
        c->zhi = zhi;
        c->vol_i8 = i8;
        c->have_vol = true;
        c->have_inside = c->have_sites = c->have_closest = c->have_measures = false;
        const bool trace = getenv("VC_TRACE") != nullptr; // development aid: where the call's time goes
        cudaEvent_t tev[8] = {};
        double host_t[8] = {};
        const auto host_t0 = std::chrono::steady_clock::now();
        auto host_mark = [&](int i) { host_t[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
        auto mark = [&](int i, cudaStream_t st)
        {
            if (trace)
            {
                cudaEventCreate(&tev[i]);
                cudaEventRecord(tev[i], st);
            }
        };
        mark(0, c->s_h2d);
        VC_TRY(st_classify_begin(c));
        // upload in plane chunks on the copy stream; a chunk is classified as soon as it has landed
        const bool chunked = st_classify_chunkable(c);
        const int nch = zhi - zlo >= 32 ? 16 : 1;
        const int chunk = (zhi - zlo + nch - 1) / nch;
        cudaEvent_t ev = nullptr;
        for (int z = zlo; z < zhi; z += chunk)
        {
            const int ze = z + chunk < zhi ? z + chunk : zhi;
            const size_t off = plane * (size_t)(z - zlo), cnt = plane * (size_t)(ze - z);
            VC_CUDA(c, cudaMemcpyAsync(c->vol.as<char>() + off * esz, (const char*)vol + off * esz, cnt * esz, cudaMemcpyDefault, c->s_h2d));
            VC_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            VC_CUDA(c, cudaEventRecord(ev, c->s_h2d));
            VC_CUDA(c, cudaStreamWaitEvent(c->stream, ev, 0));
            VC_CUDA(c, cudaEventDestroy(ev));
            if (chunked)
            {
                VC_TRY(st_classify_planes(c, z, ze));
                // the chunk's occupancy bit rows go back while later chunks are still coming in (the device-to-host
                // direction is idle during the upload); left for later they would sit in front of the small
                // count read-back in the copy queue
                const int oa = z > c->z0 ? z : c->z0, ob = ze < c->z1 ? ze : c->z1; // owned planes of this chunk
                if (inside_bits && oa < ob)
                {
                    const size_t rw = (size_t)c->ny * (size_t)c->wr;
                    VC_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    VC_CUDA(c, cudaEventRecord(ev, c->stream));
                    VC_CUDA(c, cudaStreamWaitEvent(c->s_d2h, ev, 0));
                    VC_CUDA(c, cudaEventDestroy(ev));
                    VC_CUDA(c, cudaMemcpyAsync(inside_bits + (size_t)(oa - c->z0) * rw, c->bits.as<u32>() + (size_t)(oa - zlo) * rw,
                                               (size_t)(ob - oa) * rw * 4, cudaMemcpyDefault, c->s_d2h));
                }
            }
        }
        if (!chunked)
            VC_TRY(st_classify_planes(c, zlo, zhi));
        c->have_inside = true;
        mark(1, c->s_h2d);
        mark(2, c->stream);
        auto after_main = [&](cudaStream_t s) -> int
        { // s waits for everything queued on the main stream so far
            VC_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            VC_CUDA(c, cudaEventRecord(ev, c->stream));
            VC_CUDA(c, cudaStreamWaitEvent(s, ev, 0));
            VC_CUDA(c, cudaEventDestroy(ev));
            return VC_OK;
        };
        const bool bits_later = inside_bits && !chunked;
        u32* P = (u32*)c->pinned + 16;
        VC_TRY(compact_prefix_async(c, P));
        host_mark(0);
        // numbering the sites synchronises the main stream (site count), after which P is valid too
        int64_t n_all = 0;
        if (c->peer_world > 1)
        { // slab group: the detection kernel stores the records into every rank, the union is numbered everywhere
            VC_TRY(vc_sites_post_peers(c));
            VC_TRY(vc_sites_collect_peers(c, &n_all));
        }
        else
            VC_TRY(st_detect_sites(c));
        host_mark(1);
        mark(6, c->stream);
        if (bits_later)
        {
            VC_TRY(after_main(c->s_d2h));
            VC_CUDA(c, cudaMemcpyAsync(inside_bits, c->bits.as<u32>() + (size_t)(c->z0 - zlo) * c->ny * (size_t)c->wr,
                                       (size_t)c->ny * (size_t)(c->z1 - c->z0) * (size_t)c->wr * 4, cudaMemcpyDefault, c->s_d2h));
        }
        const int64_t n = (int64_t)P[c->z1 - c->z0];
        c->ninside = n;
        if (n_inside)
            *n_inside = n;
        if (cap < n && (vert || id || d2x4 || lambda7 || radius))
        {
            cudaStreamSynchronize(c->s_d2h);
            return vc_fail(c, VC_ERR_NOMEM, "vc_run_dense_host_compact: capacity below the inside count (returned in n_inside)");
        }
        if (c->peer_world < 2)
            VC_TRY(st_finalize_sites(c, c->cand_key.as<u64>(), c->cand_corner.as<u64>(), c->ncand, VC_SITES_SORT));
        mark(3, c->stream);
        host_mark(2);
        VC_TRY(compact_alloc(c, n, true));
        std::vector<u32> Ph(P, P + (c->z1 - c->z0) + 1); // the pinned scratch is reused by later stages
        // few inside vertices: their measures go straight into the records and the 8 dense float planes are
        // never written; many: the tiled dense kernel reuses neighbours better, records are gathered from it
        const bool sparse = c->compact_mode == 2 || (c->compact_mode == 0 && (size_t)n * 8 <= nv);
        c->skip_dense_measures = sparse;
        int hook_status = VC_OK;
        c->chunk_hook = [&](int za, int zb) -> int
        {
            VC_TRY(sparse ? sparse_range(c, za, zb) : compact_range(c, za, zb, true));
            cudaEvent_t e2 = nullptr;
            VC_CUDA(c, cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            VC_CUDA(c, cudaEventRecord(e2, c->cur));
            VC_CUDA(c, cudaStreamWaitEvent(c->s_d2h, e2, 0));
            VC_CUDA(c, cudaEventDestroy(e2));
            VC_TRY(compact_copy_out(c, c->s_d2h, Ph[za - c->z0], Ph[zb - c->z0], cap, vert, id, d2x4, lambda7, radius));
            const size_t o = plane * (size_t)(za - c->z0), m = plane * (size_t)(zb - za);
            if (id_dense)
                VC_CUDA(c, cudaMemcpyAsync(id_dense + o, c->id.as<int>() + o, m * 4, cudaMemcpyDefault, c->s_d2h));
            if (d2x4_dense)
                VC_CUDA(c, cudaMemcpyAsync(d2x4_dense + o, c->d2.as<u32>() + o, m * 4, cudaMemcpyDefault, c->s_d2h));
            return VC_OK;
        };
        hook_status = st_closest_measures_pipelined(c, true);
        c->chunk_hook = nullptr;
        c->skip_dense_measures = false;
        mark(4, c->stream);
        mark(5, c->s_d2h);
        host_mark(3);
        cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->s_d2h);
        if (trace && e1 == cudaSuccess && e2 == cudaSuccess)
        {
            float t[8] = {0};
            for (int i = 1; i < 7; ++i)
                cudaEventElapsedTime(&t[i], tev[0], tev[i]);
            host_mark(4);
            fprintf(stderr, "[vc trace] upload done %.3f | classified %.3f | sites detected %.3f | sites numbered %.3f | "
                            "transform+records %.3f | copied back %.3f ms   (host: upload queued %.3f, count known %.3f, numbering queued "
                            "%.3f, pipeline queued %.3f, all done %.3f)\n", t[1], t[2], t[6], t[3], t[4], t[5], host_t[0], host_t[1],
                    host_t[2], host_t[3], host_t[4]);
            for (int i = 0; i < 7; ++i)
                cudaEventDestroy(tev[i]);
        }
        VC_TRY(hook_status);
        VC_CUDA(c, e1);
        VC_CUDA(c, e2);
        if (nsites)
            *nsites = c->nsites;
        return VC_OK;
    }

    int vc_run_dense_host_compact(vc_ctx* c, const float* vol, uint32_t* inside_bits, int64_t cap, int64_t* n_inside,
                                  uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius, int32_t* id_dense,
                                  uint32_t* d2x4_dense, int64_t* nsites)
    {
        return host_compact(c, vol, false, inside_bits, cap, n_inside, vert, id, d2x4, lambda7, radius, id_dense, d2x4_dense, nsites);
    }

    int vc_run_dense_host_compact_i8(vc_ctx* c, const int8_t* vol, uint32_t* inside_bits, int64_t cap, int64_t* n_inside,
                                     uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius, int32_t* id_dense,
                                     uint32_t* d2x4_dense, int64_t* nsites)
    {
        return host_compact(c, vol, true, inside_bits, cap, n_inside, vert, id, d2x4, lambda7, radius, id_dense, d2x4_dense, nsites);
    }
}
