// vc_api.cu -- the extern "C" surface declared in include/voxcore_gpu.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "vc_internal.h"

#define VC_ABI_VERSION 1

int vc_fail(vc_ctx* c, int code, const char* what, cudaError_t e)
{
    if (c)
    {
        c->err = what ? what : "";
        if (e != cudaSuccess)
        {
            c->err += ": ";
            c->err += cudaGetErrorString(e);
        }
    }
    cudaGetLastError(); // clear the sticky-free error state
    return code;
}

ProfScope::ProfScope(vc_ctx* ctx, const char* name) : c(ctx)
{
    c->launches++;
    if (!c->profiling)
        return;
    for (auto& s : c->stats)
        if (s.name == name)
        {
            st = &s;
            break;
        }
    if (!st)
    {
        c->stats.emplace_back();
        c->stats.back().name = name;
        st = &c->stats.back();
    }
    auto get = [&]() -> cudaEvent_t
    {
        if (!c->ev_pool.empty())
        {
            cudaEvent_t e = c->ev_pool.back();
            c->ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, c->cur);
}

ProfScope::~ProfScope()
{
    if (!st)
        return;
    cudaEventRecord(b, c->cur);
    st->pending.emplace_back(a, b);
    st->launches++;
}

static void prof_resolve(vc_ctx* c)
{
    for (auto& s : c->stats)
    {
        for (auto& p : s.pending)
        {
            float ms = 0.f;
            if (cudaEventSynchronize(p.second) == cudaSuccess && cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess)
                s.ms += ms;
            c->ev_pool.push_back(p.first);
            c->ev_pool.push_back(p.second);
        }
        s.pending.clear();
    }
    cudaGetLastError();
}

extern "C"
{
    int vc_abi_version(void) { return VC_ABI_VERSION; }

    int vc_warmup(int device)
    { // driver + primary-context start-up only (seconds on a box without a persistence daemon): safe from any thread
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess ||
            cudaFree(nullptr) != cudaSuccess)
        {
            cudaGetLastError();
            return VC_ERR_CUDA;
        }
        return VC_OK;
    }

    int vc_ctx_create(int device, vc_ctx** out)
    {
        if (!out)
            return VC_ERR_INVALID;
        *out = nullptr;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        {
            cudaGetLastError();
            return VC_ERR_CUDA; // no device: there is no CPU fallback behind this ABI
        }
        if (cudaSetDevice(device) != cudaSuccess)
            return VC_ERR_CUDA;
        vc_ctx* c = new (std::nothrow) vc_ctx;
        if (!c)
            return VC_ERR_NOMEM;
        c->device = device;
        cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
            cudaHostAlloc(&c->pinned, 32768, cudaHostAllocDefault) != cudaSuccess || c->scratch.ensure(256) != cudaSuccess)
        {
            vc_ctx_destroy(c);
            return VC_ERR_CUDA;
        }
        c->cur = c->stream;
        if (const char* e = getenv("VC_WORKERS"))
            c->nworkers = atoi(e);
        if (const char* e = getenv("VC_ZCHUNK"))
            c->zchunk = atoi(e);
        c->nworkers = c->nworkers < 0 ? 0 : (c->nworkers > VC_MAX_WORKERS ? VC_MAX_WORKERS : c->nworkers);
        c->zchunk = c->zchunk < 0 ? 0 : c->zchunk;
        bool ok = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < c->nworkers; ++i)
            ok = cudaStreamCreateWithFlags(&c->workers[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok)
        {
            vc_ctx_destroy(c);
            return VC_ERR_CUDA;
        }
        *out = c;
        return VC_OK;
    }

    void vc_ctx_destroy(vc_ctx* c)
    {
        if (!c)
            return;
        cudaSetDevice(c->device);
        if (c->stream)
            cudaStreamSynchronize(c->stream);
        prof_resolve(c);
        for (auto e : c->ev_pool)
            cudaEventDestroy(e);
        vc_peer_release(c);
        DevBuf* bufs[] = {&c->vol, &c->inside, &c->bits, &c->cand_key, &c->cand_corner, &c->site_key, &c->site_corner,
                          &c->site_xyz, &c->line_ptr, &c->line_ent, &c->g1, &c->g2, &c->stk, &c->id, &c->d2, &c->edge3,
                          &c->face3, &c->cube, &c->radius, &c->sk0, &c->sk1, &c->sv0, &c->sv1, &c->shist,
                          &c->scratch, &c->cl_ptr, &c->cl_ent, &c->gsites, &c->colmask, &c->line_cur, &c->scan_sums, &c->rowpre, &c->crec,
                          &c->row_ptr, &c->live_row, &c->row_mask, &c->col_x, &c->col_line, &c->edt_meta, &c->medial_pre};
        for (auto e : c->ev_chunk)
            cudaEventDestroy(e);
        for (auto b : bufs)
            b->release();
        if (c->pinned)
            cudaFreeHost(c->pinned);
        if (c->stream)
            cudaStreamDestroy(c->stream);
        if (c->s_h2d)
            cudaStreamDestroy(c->s_h2d);
        if (c->s_d2h)
            cudaStreamDestroy(c->s_d2h);
        for (int i = 0; i < VC_MAX_WORKERS; ++i)
        {
            if (c->workers[i])
                cudaStreamDestroy(c->workers[i]);
            if (c->ev_join[i])
                cudaEventDestroy(c->ev_join[i]);
        }
        if (c->ev_fork)
            cudaEventDestroy(c->ev_fork);
        delete c;
    }

    const char* vc_last_error(const vc_ctx* c) { return c ? c->err.c_str() : "null context"; }
    void* vc_stream(vc_ctx* c) { return c ? (void*)c->stream : nullptr; }

    int vc_synchronize(vc_ctx* c)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    void* vc_host_alloc(size_t bytes)
    {
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return p;
    }
    void vc_host_free(void* p)
    {
        if (p)
            cudaFreeHost(p);
    }

    int vc_set_grid(vc_ctx* c, int nx, int ny, int nz, int z0, int z1)
    {
        if (!c)
            return VC_ERR_INVALID;
        if (nx < 1 || ny < 1 || nz < 1 || z0 < 0 || z1 > nz || z0 >= z1)
            return vc_fail(c, VC_ERR_INVALID, "vc_set_grid: bad dimensions or slab");
        if (nx > 2048 || ny > 2048 || nz > 2048)
            return vc_fail(c, VC_ERR_UNSUPPORTED, "vc_set_grid: sides above 2048 are not supported");
        VC_CUDA(c, cudaSetDevice(c->device));
        c->nx = nx;
        c->ny = ny;
        c->nz = nz;
        c->z0 = z0;
        c->z1 = z1;
        c->zc = z1 < nz ? z1 + 1 : z1;
        c->zlo = c->zhi = 0;
        c->have_grid = true;
        c->have_vol = c->have_inside = c->have_sites = c->have_closest = c->have_measures = false;
        return VC_OK;
    }

    static int upload_planes(vc_ctx* c, const void* planes, int zlo, int zhi, bool i8)
    {
        if (!c || !planes)
            return VC_ERR_INVALID;
        if (!c->have_grid)
            return vc_fail(c, VC_ERR_STATE, "vc_volume_upload: call vc_set_grid first");
        int need_lo = c->z0 > 0 ? c->z0 - 1 : 0, need_hi = c->z1 < c->nz ? c->z1 + 1 : c->nz;
        if (zlo < 0 || zhi > c->nz || zlo > need_lo || zhi < need_hi)
            return vc_fail(c, VC_ERR_INVALID, "vc_volume_upload: planes must cover the slab plus one halo plane each side");
        VC_CUDA(c, cudaSetDevice(c->device));
        const size_t n = (size_t)c->nx * c->ny * (size_t)(zhi - zlo), esz = i8 ? 1 : 4;
        VC_CUDA(c, c->vol.ensure(n * esz + 64));
        VC_CUDA(c, cudaMemcpyAsync(c->vol.p, planes, n * esz, cudaMemcpyDefault, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        c->zlo = zlo;
        c->zhi = zhi;
        c->vol_i8 = i8;
        c->have_vol = true;
        c->have_inside = c->have_sites = c->have_closest = c->have_measures = false;
        return VC_OK;
    }

    int vc_volume_upload_i8(vc_ctx* c, const int8_t* planes, int zlo, int zhi) { return upload_planes(c, planes, zlo, zhi, true); }

    int vc_volume_upload_f32(vc_ctx* c, const float* planes, int zlo, int zhi)
    {
        return upload_planes(c, planes, zlo, zhi, false);
    }


    int vc_volume_upload_f64_zfast(vc_ctx* c, const double* vol)
    {
        if (!c || !vol)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_upload_f64_zfast(c, vol);
    }

    static int copy_out(vc_ctx* c, void* dst, const void* src, size_t bytes)
    {
        if (!dst)
            return VC_OK;
        VC_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->stream));
        return VC_OK;
    }

    // owned planes of a per-vertex array whose first resident plane is `first_plane`
    static size_t owned_offset(const vc_ctx* c, int first_plane) { return (size_t)c->nx * c->ny * (size_t)(c->z0 - first_plane); }
    static size_t owned_count(const vc_ctx* c) { return (size_t)c->nx * c->ny * (size_t)(c->z1 - c->z0); }

    int vc_classify_grid(vc_ctx* c, uint8_t* inside_out)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (c->have_vol)
            VC_TRY(st_classify(c));
        else if (!c->have_inside)
            return vc_fail(c, VC_ERR_STATE, "vc_classify_grid: no volume uploaded");
        VC_TRY(copy_out(c, inside_out, c->inside.as<u8>() + owned_offset(c, c->zlo), owned_count(c)));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_classify_mesh(vc_ctx* c, const float* verts, int64_t nv, const uint32_t* tris, int64_t nt, const double* M,
                         uint8_t* inside_out)
    {
        if (!c || nv < 0 || nt < 0 || (nv > 0 && !verts) || (nt > 0 && !tris))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        VC_TRY(st_classify_mesh(c, verts, nv, tris, nt, M));
        VC_TRY(copy_out(c, inside_out, c->inside.as<u8>() + owned_offset(c, c->zlo), owned_count(c)));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_classify_points(vc_ctx* c, const float* xyz, int64_t n, const double* M, uint8_t* out)
    {
        if (!c || n < 0 || (n > 0 && (!xyz || !out)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_classify_points(c, xyz, n, M, out);
    }

    int vc_sites_detect_local(vc_ctx* c, int64_t* nlocal)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        VC_TRY(st_detect_sites(c));
        if (nlocal)
            *nlocal = c->ncand;
        return VC_OK;
    }

    int vc_sites_export_local(vc_ctx* c, uint64_t* keys_out, uint64_t* corners_out)
    {
        if (!c || !keys_out || !corners_out)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (c->ncand)
        {
            VC_TRY(copy_out(c, keys_out, c->cand_key.p, (size_t)c->ncand * 8));
            VC_TRY(copy_out(c, corners_out, c->cand_corner.p, (size_t)c->ncand * 8));
        }
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_sites_import_global(vc_ctx* c, const uint64_t* keys, const uint64_t* corners, int64_t n)
    {
        if (!c || n < 0 || (n > 0 && (!keys || !corners)))
            return VC_ERR_INVALID;
        if (!c->have_grid)
            return vc_fail(c, VC_ERR_STATE, "vc_sites_import_global: call vc_set_grid first");
        VC_CUDA(c, cudaSetDevice(c->device));
        DevBuf dk, dc;
        const u64* pk = (const u64*)keys;
        const u64* pc = (const u64*)corners;
        if (n > 0 && vc_is_device_ptr(keys) != vc_is_device_ptr(corners))
            return vc_fail(c, VC_ERR_INVALID, "vc_sites_import_global: keys and corners must both be host or both be device pointers");
        if (n > 0 && !vc_is_device_ptr(keys))
        {
            VC_CUDA(c, dk.ensure((size_t)n * 8));
            VC_CUDA(c, dc.ensure((size_t)n * 8));
            VC_CUDA(c, cudaMemcpyAsync(dk.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
            VC_CUDA(c, cudaMemcpyAsync(dc.p, corners, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
            pk = dk.as<u64>();
            pc = dc.as<u64>();
        }
        int s = st_finalize_sites(c, pk, pc, n, VC_SITES_SORT);
        cudaStreamSynchronize(c->stream);
        dk.release();
        dc.release();
        return s;
    }

    int vc_extract_sites(vc_ctx* c, int64_t* nsites)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (c->z0 != 0 || c->z1 != c->nz)
            return vc_fail(c, VC_ERR_STATE, "vc_extract_sites: slab contexts use vc_sites_detect_local / export / import_global");
        VC_TRY(st_detect_sites(c));
        VC_TRY(st_finalize_sites(c, c->cand_key.as<u64>(), c->cand_corner.as<u64>(), c->ncand, VC_SITES_SORT));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        if (nsites)
            *nsites = c->nsites;
        return VC_OK;
    }

    int64_t vc_num_sites(const vc_ctx* c) { return c && c->have_sites ? c->nsites : -1; }

    int vc_get_sites(vc_ctx* c, float* xyz_out)
    {
        if (!c || !xyz_out)
            return VC_ERR_INVALID;
        if (!c->have_sites)
            return vc_fail(c, VC_ERR_STATE, "vc_get_sites: no sites");
        VC_CUDA(c, cudaSetDevice(c->device));
        if (c->nsites == 0)
            return VC_OK;
        // float4 table -> packed xyz triples
        VC_CUDA(c, cudaMemcpy2DAsync(xyz_out, 12, c->site_xyz.p, 16, 12, (size_t)c->nsites, cudaMemcpyDefault, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_set_sites(vc_ctx* c, const float* xyz, int64_t n)
    {
        if (!c || n < 0 || (n > 0 && !xyz))
            return VC_ERR_INVALID;
        if (!c->have_grid)
            return vc_fail(c, VC_ERR_STATE, "vc_set_sites: call vc_set_grid first");
        VC_CUDA(c, cudaSetDevice(c->device));
        // lattice test on the host: every coordinate + 0.5 is an integer corner index of the grid
        bool lattice = true;
        std::vector<u64> corners((size_t)n);
        const int lim[3] = {c->nx, c->ny, c->nz};
        for (int64_t i = 0; i < n && lattice; ++i)
        {
            int ci[3];
            for (int d = 0; d < 3; ++d)
            {
                const float f = xyz[3 * i + d] + 0.5f;
                // range first: a cast of NaN / inf / a huge value to int is undefined
                const int k = (f >= 0.0f && f <= (float)lim[d]) ? (int)f : -1;
                if (k < 0 || !(f == (float)k) || xyz[3 * i + d] != (float)k - 0.5f)
                    lattice = false;
                ci[d] = k;
            }
            if (lattice)
                corners[i] = vc_pack_corner(ci[0], ci[1], ci[2]);
        }
        if (lattice)
        {
            DevBuf dc;
            VC_CUDA(c, dc.ensure((size_t)(n + 1) * 8));
            if (n)
                VC_CUDA(c, cudaMemcpyAsync(dc.p, corners.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
            int s = st_finalize_sites(c, nullptr, dc.as<u64>(), n, VC_SITES_EXTERNAL);
            cudaStreamSynchronize(c->stream);
            dc.release();
            if (s != VC_OK && s != VC_ERR_UNSUPPORTED) // more sites than the dense path's id fields hold: the cell list has no limit
                return s;
            if (s == VC_OK && c->lattice)
                return VC_OK;
        }
        return st_build_cell_list(c, xyz, n);
    }

    int vc_closest_grid(vc_ctx* c, int32_t* id_out, uint32_t* d2x4_out)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (!c->have_sites)
            return vc_fail(c, VC_ERR_STATE, "vc_closest_grid needs sites");
        if (c->lattice)
            VC_TRY(st_closest_lattice(c));
        else
            VC_TRY(st_closest_general_grid(c));
        VC_TRY(copy_out(c, id_out, c->id.p, owned_count(c) * 4));
        VC_TRY(copy_out(c, d2x4_out, c->d2.p, owned_count(c) * 4));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_closest_points(vc_ctx* c, const double* q, int64_t n, int32_t* id, double* d2)
    {
        if (!c || n < 0 || (n > 0 && (!q || !id)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_closest_points(c, q, n, id, d2);
    }

    int vc_closest_points_f32(vc_ctx* c, const float* q, int64_t n, float max_d2, int32_t* id, float* d2)
    {
        if (!c || n < 0 || (n > 0 && (!q || !id)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_closest_points_f32(c, q, n, max_d2, id, d2);
    }

    int vc_cell_measures_grid(vc_ctx* c, float* edge3, float* face3, float* cube, float* radius)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        VC_TRY(st_measures(c, true));
        size_t nv = owned_count(c);
        VC_TRY(copy_out(c, edge3, c->edge3.p, nv * 12));
        VC_TRY(copy_out(c, face3, c->face3.p, nv * 12));
        VC_TRY(copy_out(c, cube, c->cube.p, nv * 4));
        VC_TRY(copy_out(c, radius, c->radius.p, nv * 4));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_face_lambda(vc_ctx* c, const int32_t* site_pairs, int64_t nf, float* out)
    {
        if (!c || nf < 0 || (nf > 0 && (!site_pairs || !out)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_face_lambda(c, site_pairs, nf, out);
    }
    int vc_vertex_radii(vc_ctx* c, const float* v_xyz, int64_t nv, const int32_t* site_of_v, float* r_out)
    {
        if (!c || nv < 0 || (nv > 0 && (!v_xyz || !site_of_v || !r_out)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_vertex_radii(c, v_xyz, nv, site_of_v, r_out);
    }
    int vc_segment_max(vc_ctx* c, const int32_t* off, const int32_t* items, int64_t n, const float* value,
                       int64_t nvalue, const uint8_t* valid, float* out)
    {
        if (!c || n < 0 || (n > 0 && (!off || !items || !value || !out)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_segment_max(c, off, items, n, value, nvalue, valid, out);
    }

    int vc_radius_search(vc_ctx* c, const double* q, const double* sq_rad, int64_t n, const int64_t* off, int32_t* count, int32_t* idx,
                         double* d2)
    {
        if (!c || n < 0 || (n > 0 && (!q || !sq_rad)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_radius_search(c, q, sq_rad, n, off, count, idx, d2);
    }
    int vc_ref_counts(vc_ctx* c, const int32_t* idx, int64_t n, int64_t nbins, int32_t* out)
    {
        if (!c || n < 0 || nbins < 0 || (n > 0 && !idx) || (nbins > 0 && !out))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_ref_counts(c, idx, n, nbins, out);
    }
    int vc_simple_pairs(vc_ctx* c, const int32_t* edge_ref, const int32_t* edge_face0, int64_t ne, const float* face_measure,
                        const uint8_t* face_to_remove, int64_t nf, float f_t, const int32_t* vert_ref, const int32_t* vert_edge0,
                        int64_t nv, const float* edge_measure, float l_t, int32_t* pairs_out, int64_t cap, int64_t* npairs)
    {
        if (!c || ne < 0 || nf < 0 || nv < 0 || cap < 0 || (ne > 0 && (!edge_ref || !edge_face0 || !edge_measure)) ||
            (nf > 0 && !face_measure) || (nv > 0 && (!vert_ref || !vert_edge0)))
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return st_simple_pairs(c, edge_ref, edge_face0, ne, face_measure, face_to_remove, nf, f_t, vert_ref, vert_edge0, nv,
                               edge_measure, l_t, pairs_out, cap, npairs);
    }

    int vc_run_dense(vc_ctx* c, int64_t* nsites)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (c->z0 != 0 || c->z1 != c->nz)
            return vc_fail(c, VC_ERR_STATE, "vc_run_dense: slab contexts run the stages one by one around the site exchange");
        if (c->have_vol || !c->have_inside) // flags from vc_classify_mesh / the f64 upload stand in for a volume
            VC_TRY(st_classify(c));
        VC_TRY(st_detect_sites(c));
        VC_TRY(st_finalize_sites(c, c->cand_key.as<u64>(), c->cand_corner.as<u64>(), c->ncand, VC_SITES_SORT));
        VC_TRY(st_closest_measures_pipelined(c, true));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        if (nsites)
            *nsites = c->nsites;
        return VC_OK;
    }

    int vc_closest_and_measures(vc_ctx* c)
    {
        if (!c)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (!c->have_sites)
            return vc_fail(c, VC_ERR_STATE, "vc_closest_and_measures needs sites");
        VC_TRY(st_closest_measures_pipelined(c, true));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_set_pipeline(vc_ctx* c, int workers, int zchunk)
    {
        if (!c || workers < 0 || workers > VC_MAX_WORKERS || zchunk < 0)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        for (int i = 0; i < workers; ++i)
            if (!c->workers[i])
            {
                VC_CUDA(c, cudaStreamCreateWithFlags(&c->workers[i], cudaStreamNonBlocking));
                VC_CUDA(c, cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
            }
        c->nworkers = workers;
        c->zchunk = zchunk;
        return VC_OK;
    }

    static int result_array(vc_ctx* c, int which, void** p, size_t* bytes)
    {
        size_t nv = owned_count(c);
        switch (which)
        {
            case VC_ARR_INSIDE:
                if (!c->have_inside) break;
                *p = c->inside.as<u8>() + owned_offset(c, c->zlo);
                *bytes = nv;
                return VC_OK;
            case VC_ARR_ID:
                if (!c->have_closest) break;
                *p = c->id.p;
                *bytes = nv * 4;
                return VC_OK;
            case VC_ARR_D2X4:
                if (!c->have_closest) break;
                *p = c->d2.p;
                *bytes = nv * 4;
                return VC_OK;
            case VC_ARR_EDGE3:
                if (!c->have_measures) break;
                *p = c->edge3.p;
                *bytes = nv * 12;
                return VC_OK;
            case VC_ARR_FACE3:
                if (!c->have_measures) break;
                *p = c->face3.p;
                *bytes = nv * 12;
                return VC_OK;
            case VC_ARR_CUBE:
                if (!c->have_measures) break;
                *p = c->cube.p;
                *bytes = nv * 4;
                return VC_OK;
            case VC_ARR_RADIUS:
                if (!c->have_measures || !c->radius.p) break;
                *p = c->radius.p;
                *bytes = nv * 4;
                return VC_OK;
            default:
                return vc_fail(c, VC_ERR_INVALID, "unknown array");
        }
        return vc_fail(c, VC_ERR_STATE, "array not computed yet");
    }

    int vc_download(vc_ctx* c, int which, void* dst)
    {
        if (!c || !dst)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        void* p = nullptr;
        size_t bytes = 0;
        VC_TRY(result_array(c, which, &p, &bytes));
        VC_CUDA(c, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDefault, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    int vc_download_planes(vc_ctx* c, int which, int za, int zb, void* dst)
    {
        if (!c || !dst)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        void* p = nullptr;
        size_t bytes = 0;
        VC_TRY(result_array(c, which, &p, &bytes));
        // the closest-site planes also hold the recomputed halo plane z1 of a slab (planes [z0, zc))
        const int zend = (which == VC_ARR_ID || which == VC_ARR_D2X4) ? c->zc : c->z1;
        if (za < c->z0 || zb > zend || za > zb)
            return vc_fail(c, VC_ERR_INVALID, "vc_download_planes: plane range outside this ctx's planes");
        const size_t plane = (size_t)c->nx * c->ny;
        const int ncomp = (which == VC_ARR_EDGE3 || which == VC_ARR_FACE3) ? 3 : 1;
        const size_t esz = which == VC_ARR_INSIDE ? 1 : 4;
        const size_t row = plane * (size_t)(zb - za) * esz, comp_pitch = plane * (size_t)(c->z1 - c->z0) * esz;
        const char* src = (const char*)p + plane * (size_t)(za - c->z0) * esz;
        if (row && ncomp == 1)
            VC_CUDA(c, cudaMemcpyAsync(dst, src, row, cudaMemcpyDefault, c->stream));
        else if (row)
            VC_CUDA(c, cudaMemcpy2DAsync(dst, row, src, comp_pitch, row, ncomp, cudaMemcpyDefault, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        return VC_OK;
    }

    void* vc_device_ptr(vc_ctx* c, int which)
    {
        if (!c)
            return nullptr;
        void* p = nullptr;
        size_t bytes = 0;
        if (result_array(c, which, &p, &bytes) != VC_OK)
            return nullptr;
        return p;
    }

    int vc_run_dense_host(vc_ctx* c, const float* vol, uint8_t* inside, int32_t* id, uint32_t* d2x4, float* edge3,
                          float* face3, float* cube, float* radius, int64_t* nsites)
    {
        if (!c || !vol)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        if (!c->have_grid || c->z0 != 0 || c->z1 != c->nz)
            return vc_fail(c, VC_ERR_STATE, "vc_run_dense_host needs a ctx that owns the whole grid");
        const size_t plane = (size_t)c->nx * c->ny, nv = plane * c->nz;
        VC_CUDA(c, c->vol.ensure(nv * 4 + 64));
        VC_CUDA(c, c->inside.ensure(nv + 16));
        c->vol_i8 = false;
        c->zlo = 0;
        c->zhi = c->nz;
        // H2D in plane chunks on the copy stream; classification of a chunk starts as soon as it lands
        const int chunk = c->nz >= 16 ? (c->nz + 7) / 8 : c->nz;
        VcEvents events;
        cudaEvent_t ev[16];
        int nchunks = 0;
        for (int z = 0; z < c->nz; z += chunk, ++nchunks)
        {
            int ze = z + chunk < c->nz ? z + chunk : c->nz;
            size_t off = plane * z, cnt = plane * (size_t)(ze - z);
            VC_CUDA(c, cudaMemcpyAsync(c->vol.as<float>() + off, vol + off, cnt * 4, cudaMemcpyDefault, c->s_h2d));
            VC_CUDA(c, events.make(&ev[nchunks]));
            VC_CUDA(c, cudaEventRecord(ev[nchunks], c->s_h2d));
        }
        for (int i = 0; i < nchunks; ++i)
            VC_CUDA(c, cudaStreamWaitEvent(c->stream, ev[i], 0));
        c->have_vol = true;
        VC_TRY(st_classify(c));
        cudaEvent_t done;
        VC_CUDA(c, events.make(&done));
        auto d2h_after = [&](void* dst, const void* src, size_t bytes) -> int
        {
            if (!dst)
                return VC_OK;
            VC_CUDA(c, cudaEventRecord(done, c->stream));
            VC_CUDA(c, cudaStreamWaitEvent(c->s_d2h, done, 0));
            VC_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->s_d2h));
            return VC_OK;
        };
        VC_TRY(d2h_after(inside, c->inside.p, nv));
        VC_TRY(st_detect_sites(c));
        VC_TRY(st_finalize_sites(c, c->cand_key.as<u64>(), c->cand_corner.as<u64>(), c->ncand, VC_SITES_SORT));
        VC_TRY(st_closest_measures_pipelined(c, radius != nullptr));
        VC_TRY(d2h_after(id, c->id.p, nv * 4));
        VC_TRY(d2h_after(d2x4, c->d2.p, nv * 4));
        VC_TRY(d2h_after(edge3, c->edge3.p, nv * 12));
        VC_TRY(d2h_after(face3, c->face3.p, nv * 12));
        VC_TRY(d2h_after(cube, c->cube.p, nv * 4));
        if (radius)
            VC_TRY(d2h_after(radius, c->radius.p, nv * 4));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->s_d2h));
        if (nsites)
            *nsites = c->nsites;
        return VC_OK;
    }

    int vc_profile_enable(vc_ctx* c, int on)
    {
        if (!c)
            return VC_ERR_INVALID;
        c->profiling = on != 0;
        return VC_OK;
    }
    int vc_profile_reset(vc_ctx* c)
    {
        if (!c)
            return VC_ERR_INVALID;
        cudaSetDevice(c->device);
        prof_resolve(c);
        c->stats.clear();
        c->launches = 0;
        return VC_OK;
    }
    int vc_profile_count(vc_ctx* c)
    {
        if (!c)
            return 0;
        cudaSetDevice(c->device);
        prof_resolve(c);
        return (int)c->stats.size();
    }
    int vc_profile_get(vc_ctx* c, int i, const char** name, double* total_ms, int64_t* launches)
    {
        if (!c || i < 0 || i >= (int)c->stats.size())
            return VC_ERR_INVALID;
        if (name)
            *name = c->stats[i].name.c_str();
        if (total_ms)
            *total_ms = c->stats[i].ms;
        if (launches)
            *launches = c->stats[i].launches;
        return VC_OK;
    }
    int64_t vc_launch_count(const vc_ctx* c) { return c ? c->launches : 0; }
}
