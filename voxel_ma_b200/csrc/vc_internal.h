// vc_internal.h -- context, device buffers, launch/profile helper shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/voxcore_gpu.h"
#include "vc_core.h"

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

#define VC_MAX_WORKERS 16
#define VC_MAX_PEERS 8

// where a rank's site records go in the peer exchange (vc_peer.cu): rec[p] = start of THIS rank's
// region inside rank p's receive buffer (keys at [0,cap), corners at [cap,2cap))
struct VcPeerDst
{
    u64* rec[VC_MAX_PEERS] = {};
    int world = 0;
    u64 cap = 0;
};

struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    template <class T> T* as() const { return (T*)p; }
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap && p)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        if (bytes == 0)
            bytes = 256;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess)
            cap = bytes;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct KStat
{
    std::string name;
    double ms = 0.0;
    int64_t launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct vc_ctx
{
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    // z-chunk pipeline: the closest-site passes and the measures of different z chunks run on
    // worker streams so that the tail of one stage overlaps the next stage of another chunk
    cudaStream_t cur = nullptr;            // stream VC_LAUNCH uses (== stream outside the pipeline)
    cudaStream_t workers[VC_MAX_WORKERS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[VC_MAX_WORKERS] = {};
    int nworkers = 8, zchunk = 0; // zchunk 0 = automatic (vc_edt.cu)
    // grid: global size, owned vertex planes [z0,z1), closest planes [z0,zc), resident voxel planes [zlo,zhi)
    int nx = 0, ny = 0, nz = 0, z0 = 0, z1 = 0, zc = 0, zlo = 0, zhi = 0;
    bool have_grid = false, have_vol = false, have_inside = false, have_sites = false, have_closest = false,
         have_measures = false;
    bool vol_i8 = false;        // the resident volume is MRC mode 0 (signed bytes) instead of float32
    bool attr_measures = false; // dynamic shared memory opt-in done for this device
    bool lattice = true; // sites lie on the corner lattice -> dense transform
    DevBuf vol, inside;
    DevBuf bits; // occupancy bit rows: u32[(z - zlo) * ny + y][wr], bit x & 31 of word x >> 5; wr = nx / 32 + 1
    int wr = 0;
    // site candidates of this slab (unsorted) and the global, numbered site set
    DevBuf cand_key, cand_corner;
    int64_t ncand = 0;
    size_t cand_cap_hint = 0; // capacity the next single-pass detection starts with (grows with the counts seen)
    DevBuf site_key, site_corner, site_xyz; // u64, u64, float4 in id order
    int64_t nsites = 0;
    DevBuf line_ptr, line_ent; // z-line lists: int32[(nx+1)(ny+1)+1], u64 (cz<<32|id)
    DevBuf scan_sums;          // block totals of vc_exclusive_scan_u32
    DevBuf line_cur;           // per-column fill cursors while the lists are built
    DevBuf colmask;            // u32[ny+1][(nx+32)/32] bit rows: column (cx,cy) has sites (pass Z writes, pass X reads only those)
    // transform scratch + results
    DevBuf g1, g2, id, d2, edge3, face3, cube, radius;
    DevBuf stk; // spill space of the envelope stacks of the transform passes (16-byte entries per line and depth)
    // compact columns of the site set (vc_edt.cu): row_ptr[cy] = first column of row cy, col_x / col_line per column,
    // live_row = rows that hold columns (row_mask: the same as a bitmap), edt_meta = {#columns, #live rows}
    DevBuf row_ptr, live_row, row_mask, col_x, col_line, edt_meta;
    bool edt_cols_ready = false;
    std::vector<cudaEvent_t> ev_chunk; // "transform of z chunk k done" (the next chunk's measures wait for it)
    // sort scratch
    DevBuf sk0, sk1, sv0, sv1, shist;
    int coop_sort = -1; // blocks of the single-launch cooperative radix sort that fit the device (0: launch per phase)
    DevBuf scratch; // small device scratch (counters)
    void* pinned = nullptr; // small pinned host scratch
    // general (non-lattice) sites: uniform cell list
    DevBuf cl_ptr, cl_ent, gsites; // int32 offsets, int32 site ids per cell, double3 sites
    int cl_dim[3] = {0, 0, 0};
    double cl_org[3] = {0, 0, 0}, cl_h = 1.0;
    // peer exchange of the site records (vc_peer.cu): receive buffer of this rank, mapped buffers of the others
    int peer_world = 0, peer_rank = -1;
    int64_t peer_cap = 0, peer_timeout_ms = 0; // bound of the wait for the other ranks' posts (0: VC_PEER_TIMEOUT_MS or 10 s)
    u64 peer_seq = 0;
    bool peer_posted = false, peer_ipc = false;
    void* peer_base[VC_MAX_PEERS] = {};
    DevBuf peer_rx, peer_all; // own receive buffer; gathered (keys | corners) of all ranks
    // compact product (vc_compact.cu): exclusive prefix of the inside count per bit row of the owned planes,
    // records of the inside vertices; chunk_hook runs after the measures of each z chunk of the pipeline
    DevBuf rowpre, crec;
    DevBuf medial_pre; // per-row prefix of the dual-quad counts (vc_medial.cu)
    int64_t ninside = -1, ccap = 0;
    std::function<int(int, int)> chunk_hook;
    int compact_mode = 0;             // 0 automatic, 1 dense planes + gather, 2 records computed directly
    bool skip_dense_measures = false; // set by vc_run_dense_host_compact while its records are computed directly
    std::string err;
    bool profiling = false;
    std::vector<KStat> stats;
    std::vector<cudaEvent_t> ev_pool;
    int64_t launches = 0;
};

int vc_fail(vc_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess);

#define VC_CUDA(c, call)                                       \
    do                                                         \
    {                                                          \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess)                                 \
            return vc_fail((c), VC_ERR_CUDA, #call, _e);       \
    } while (0)

#define VC_TRY(call)          \
    do                        \
    {                         \
        int _s = (call);      \
        if (_s != VC_OK)      \
            return _s;        \
    } while (0)

// events that live for one API call: destroyed on every way out of it (error returns included)
struct VcEvents
{
    std::vector<cudaEvent_t> v;
    cudaError_t make(cudaEvent_t* e)
    {
        cudaError_t st = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
        if (st == cudaSuccess)
            v.push_back(*e);
        return st;
    }
    ~VcEvents()
    {
        for (auto e : v)
            cudaEventDestroy(e);
    }
};

// RAII launch bracket: counts the launch and, when profiling, records CUDA events on the ctx stream
struct ProfScope
{
    vc_ctx* c;
    KStat* st = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(vc_ctx* ctx, const char* name);
    ~ProfScope();
};

#define VC_LAUNCH(c, name, kern, grid, block, smem, ...)               \
    do                                                                 \
    {                                                                  \
        ProfScope _ps((c), (name));                                    \
        kern<<<(grid), (block), (smem), (c)->cur>>>(__VA_ARGS__);   \
    } while (0)

static inline unsigned vc_blocks(size_t n, unsigned per_block) { return (unsigned)((n + per_block - 1) / per_block); }
static inline bool vc_is_device_ptr(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- stage functions (vc_stages.cu / vc_sites.cu / vc_edt.cu / vc_measures.cu) -------------------
int st_classify(vc_ctx* c);
bool st_classify_chunkable(const vc_ctx* c);
int st_classify_begin(vc_ctx* c);
int st_classify_planes(vc_ctx* c, int za, int zb);
int vc_exclusive_scan_u32(vc_ctx* c, u32* a, int64_t len);
int st_detect_sites(vc_ctx* c);
void vc_peer_release(vc_ctx* c);
extern "C" int vc_sites_post_peers(vc_ctx* c);
extern "C" int vc_sites_collect_peers(vc_ctx* c, int64_t* n_all);
enum
{
    VC_SITES_EXTERNAL = 0,
    VC_SITES_SORT = 1,
    VC_SITES_PRESORTED = 2
};
int st_finalize_sites(vc_ctx* c, const u64* keys_dev, const u64* corners_dev, int64_t n, int mode);
int st_closest_lattice(vc_ctx* c);
int st_measures(vc_ctx* c, bool want_radius);
// the two stages above for the whole slab, z chunk by z chunk on the worker streams
int st_closest_measures_pipelined(vc_ctx* c, bool want_radius);
// building blocks: planes [zb,ze) of the slab on stream c->cur
int edt_range(vc_ctx* c, int zb, int ze, int region, int region_planes);
int measures_range(vc_ctx* c, int za, int zb, bool want_radius, bool alone);
int measures_alloc(vc_ctx* c, bool want_radius);
int st_classify_points(vc_ctx* c, const float* xyz, int64_t n, const double* M, uint8_t* out);
int st_face_lambda(vc_ctx* c, const int32_t* pairs, int64_t nf, float* out);
int st_vertex_radii(vc_ctx* c, const float* v, int64_t nv, const int32_t* site_of_v, float* out);
int st_segment_max(vc_ctx* c, const int32_t* off, const int32_t* items, int64_t n, const float* value,
                   int64_t nvalue, const uint8_t* valid, float* out);
int st_ref_counts(vc_ctx* c, const int32_t* idx, int64_t n, int64_t nbins, int32_t* out);
int st_simple_pairs(vc_ctx* c, const int32_t* edge_ref, const int32_t* edge_face0, int64_t ne, const float* face_measure,
                    const uint8_t* face_to_remove, int64_t nf, float f_t, const int32_t* vert_ref, const int32_t* vert_edge0,
                    int64_t nv, const float* edge_measure, float l_t, int32_t* pairs_out, int64_t cap, int64_t* npairs);
int st_upload_f64_zfast(vc_ctx* c, const double* vol);
int st_classify_mesh(vc_ctx* c, const float* verts, int64_t nv, const uint32_t* tris, int64_t nt, const double* M);
// general sites
int st_build_cell_list(vc_ctx* c, const float* xyz_host, int64_t n);
int st_closest_points(vc_ctx* c, const double* q, int64_t n, int32_t* id, double* d2);
int st_closest_points_f32(vc_ctx* c, const float* q, int64_t n, float max_d2, int32_t* id, float* d2);
int st_closest_general_grid(vc_ctx* c);
int st_radius_search(vc_ctx* c, const double* q, const double* sq_rad, int64_t n, const int64_t* off, int32_t* count, int32_t* idx,
                     double* d2);

// radix sort of (u64 key, u32 value) pairs on bits [0,nbits); result pointers returned
int vc_radix_sort_pairs(vc_ctx* c, int64_t n, int nbits, u64** keys_io, u32** vals_io);
int vc_sort_records_by_key(vc_ctx* c, const u64* keys_dev, int64_t n, u64** keys_io, u32** vals_io);
