// vc_thin.cu -- K6: the data-parallel part of CellComplexThinning::prune (src/ccthin.cpp:201-271).
//
// The thinning itself is a FIFO worklist whose order decides the result, so it
// stays on the host, in the reference's code.  What is data-parallel is everything prune does before
// the first pop:
//   - the reference counts  cellcomplex::refCntPerVert / refCntPerEdge (src/cellcomplex.cpp:315-332):
//     how many edges use a vertex, how many faces use an edge          -> k_histogram
//   - the seeding scan (src/ccthin.cpp:246-270): every edge with exactly one face whose face is below
//     the face threshold (or marked to-remove) gives a face-edge pair, every vertex with exactly one
//     edge below the edge threshold gives an edge-vertex pair; pairs enter the queue in ascending
//     edge index, then ascending vertex index                          -> k_sp_flags / k_sp_emit
// The emit kernel is an order-preserving compaction: a warp ballot ranks the candidates inside a warp,
// per-block totals are scanned by one block, so the queue content comes out in exactly the reference's
// push order.
#include "vc_internal.h"

__global__ void __launch_bounds__(256) k_histogram(const int* __restrict__ idx, int64_t n, int64_t nbins, int* __restrict__ out,
                                                   int* __restrict__ flags)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int b = idx[i];
    if (b < 0 || b >= nbins)
        atomicOr(flags, 1);
    else
        atomicAdd(out + b, 1);
}

// item i < ne : edge i   -> candidate iff ref[i]==1 and (to_remove[f] or face_measure[f] < f_t), f = first[i]
// item i >= ne: vertex v -> candidate iff ref[v]==1 and edge_measure[first[v]] < l_t
struct SpArgs
{
    const int *edge_ref, *edge_face0, *vert_ref, *vert_edge0;
    const float *face_measure, *edge_measure;
    const u8* face_to_remove;
    int64_t ne, nv, nf;
    float f_t, l_t;
};

__device__ __forceinline__ bool sp_candidate(const SpArgs& a, int64_t i, int* partner, int* bad)
{
    if (i < a.ne)
    {
        if (a.edge_ref[i] != 1)
            return false;
        const int f = a.edge_face0[i];
        if (f < 0 || f >= a.nf)
        {
            *bad = 1;
            return false;
        }
        *partner = f;
        // face_edge_pair_below_threshold (src/ccthin.cpp:514-521): the edge threshold is not consulted
        return (a.face_to_remove && a.face_to_remove[f]) || a.face_measure[f] < a.f_t;
    }
    const int64_t v = i - a.ne;
    if (v >= a.nv || a.vert_ref[v] != 1)
        return false;
    const int e = a.vert_edge0[v];
    if (e < 0 || e >= a.ne)
    {
        *bad = 1;
        return false;
    }
    *partner = e;
    return a.edge_measure[e] < a.l_t; // edge_vert_pair_below_threshold (src/ccthin.cpp:508-512)
}

#define SP_BLOCK 256
__global__ void __launch_bounds__(SP_BLOCK) k_sp_count(SpArgs a, int* __restrict__ block_count, int* __restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * SP_BLOCK + threadIdx.x;
    int partner = 0, bad = 0;
    const bool c = i < a.ne + a.nv && sp_candidate(a, i, &partner, &bad);
    if (bad)
        atomicOr(flags, 2);
    const int total = __syncthreads_count(c);
    if (threadIdx.x == 0)
        block_count[blockIdx.x] = total;
}

// exclusive scan of the block totals, one block; [nblocks] receives the grand total
__global__ void __launch_bounds__(1024) k_sp_scan(int* __restrict__ block_count, int nblocks)
{
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024)
    {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? block_count[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int t = __shfl_up_sync(0xFFFFFFFFu, s, d);
            if (lane >= d)
                s += t;
        }
        if (lane == 31)
            warp_sum[w] = s;
        __syncthreads();
        if (w == 0)
        {
            int ws = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                int t = __shfl_up_sync(0xFFFFFFFFu, ws, d);
                if (lane >= d)
                    ws += t;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const int before = carry + (w ? warp_sum[w - 1] : 0) + s - v;
        if (i < nblocks)
            block_count[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023)
            carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        block_count[nblocks] = carry;
}

// pairs_out: int32 triples (type, idx0, idx1): FE_PAIR = 1 (face, edge), EV_PAIR = 0 (edge, vertex) -- include/ccthin.h:22-28
__global__ void __launch_bounds__(SP_BLOCK) k_sp_emit(SpArgs a, const int* __restrict__ block_base, int* __restrict__ pairs_out, int64_t cap)
{
    __shared__ int warp_cnt[SP_BLOCK / 32];
    const int64_t i = (int64_t)blockIdx.x * SP_BLOCK + threadIdx.x;
    int partner = 0, bad = 0;
    const bool c = i < a.ne + a.nv && sp_candidate(a, i, &partner, &bad);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, c);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        warp_cnt[w] = __popc(m);
    __syncthreads();
    int before = block_base[blockIdx.x];
    for (int k = 0; k < w; ++k)
        before += warp_cnt[k];
    if (c)
    {
        const int64_t slot = before + __popc(m & ((1u << lane) - 1u));
        if (slot < cap)
        {
            const bool fe = i < a.ne;
            pairs_out[3 * slot] = fe ? 1 : 0;
            pairs_out[3 * slot + 1] = partner;
            pairs_out[3 * slot + 2] = (int)(fe ? i : i - a.ne);
        }
    }
}

static cudaError_t to_device(vc_ctx* c, DevBuf& buf, const void* host, size_t bytes, const void** dev)
{
    *dev = nullptr;
    if (!host || bytes == 0)
        return cudaSuccess;
    if (vc_is_device_ptr(host))
    {
        *dev = host;
        return cudaSuccess;
    }
    cudaError_t e = buf.ensure(bytes);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, c->stream);
    *dev = buf.p;
    return e;
}

int st_ref_counts(vc_ctx* c, const int32_t* idx, int64_t n, int64_t nbins, int32_t* out)
{
    DevBuf di, dout;
    const void* pi = nullptr;
    cudaError_t e = to_device(c, di, idx, (size_t)n * 4, &pi);
    if (e == cudaSuccess)
        e = dout.ensure((size_t)(nbins > 0 ? nbins : 1) * 4);
    if (e == cudaSuccess)
        e = c->scratch.ensure(256);
    int flags = 0;
    if (e == cudaSuccess)
    {
        cudaMemsetAsync(dout.p, 0, (size_t)(nbins > 0 ? nbins : 1) * 4, c->stream);
        cudaMemsetAsync(c->scratch.p, 0, 8, c->stream);
        if (n > 0)
            VC_LAUNCH(c, "ref_counts", k_histogram, vc_blocks((size_t)n, 256), 256, 0, (const int*)pi, n, nbins, dout.as<int>(),
                      c->scratch.as<int>());
        if (nbins > 0)
            e = cudaMemcpyAsync(out, dout.p, (size_t)nbins * 4, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&flags, c->scratch.p, 4, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    di.release();
    dout.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "vc_ref_counts", e);
    if (flags)
        return vc_fail(c, VC_ERR_INVALID, "vc_ref_counts: an index lies outside [0, nbins)");
    return VC_OK;
}

int st_simple_pairs(vc_ctx* c, const int32_t* edge_ref, const int32_t* edge_face0, int64_t ne, const float* face_measure,
                    const uint8_t* face_to_remove, int64_t nf, float f_t, const int32_t* vert_ref, const int32_t* vert_edge0,
                    int64_t nv, const float* edge_measure, float l_t, int32_t* pairs_out, int64_t cap, int64_t* npairs)
{
    DevBuf b[7], dcount, dpairs;
    SpArgs a{};
    a.ne = ne, a.nv = nv, a.nf = nf, a.f_t = f_t, a.l_t = l_t;
    cudaError_t e = cudaSuccess;
    const void* p = nullptr;
#define UP(k, field, host, bytes, T)                              \
    if (e == cudaSuccess)                                         \
    {                                                             \
        e = to_device(c, b[k], host, bytes, &p);                  \
        a.field = (const T*)p;                                    \
    }
    UP(0, edge_ref, edge_ref, (size_t)ne * 4, int)
    UP(1, edge_face0, edge_face0, (size_t)ne * 4, int)
    UP(2, vert_ref, vert_ref, (size_t)nv * 4, int)
    UP(3, vert_edge0, vert_edge0, (size_t)nv * 4, int)
    UP(4, face_measure, face_measure, (size_t)nf * 4, float)
    UP(5, edge_measure, edge_measure, (size_t)ne * 4, float)
    UP(6, face_to_remove, face_to_remove, (size_t)nf, u8)
#undef UP
    const int64_t n = ne + nv;
    const unsigned nblocks = vc_blocks((size_t)(n > 0 ? n : 1), SP_BLOCK);
    if (e == cudaSuccess)
        e = dcount.ensure(((size_t)nblocks + 1) * 4);
    if (e == cudaSuccess)
        e = c->scratch.ensure(256);
    int total = 0, flags = 0;
    if (e == cudaSuccess)
    {
        cudaMemsetAsync(c->scratch.p, 0, 8, c->stream);
        VC_LAUNCH(c, "simple_pairs_count", k_sp_count, nblocks, SP_BLOCK, 0, a, dcount.as<int>(), c->scratch.as<int>());
        VC_LAUNCH(c, "simple_pairs_scan", k_sp_scan, 1, 1024, 0, dcount.as<int>(), (int)nblocks);
        e = cudaMemcpyAsync(&total, dcount.as<int>() + nblocks, 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(&flags, c->scratch.p, 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->stream);
    }
    if (e == cudaSuccess && !flags && pairs_out && cap > 0 && total > 0)
    {
        const int64_t m = total < cap ? total : cap;
        const bool dev_out = vc_is_device_ptr(pairs_out);
        int* dst = (int*)pairs_out;
        if (!dev_out)
        {
            e = dpairs.ensure((size_t)m * 12);
            dst = dpairs.as<int>();
        }
        if (e == cudaSuccess)
        {
            VC_LAUNCH(c, "simple_pairs_emit", k_sp_emit, nblocks, SP_BLOCK, 0, a, dcount.as<int>(), dst, m);
            if (!dev_out)
                e = cudaMemcpyAsync(pairs_out, dst, (size_t)m * 12, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess)
                e = cudaStreamSynchronize(c->stream);
        }
    }
    if (e == cudaSuccess)
        e = cudaGetLastError();
    for (auto& x : b)
        x.release();
    dcount.release();
    dpairs.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "vc_simple_pairs", e);
    if (flags)
        return vc_fail(c, VC_ERR_INVALID, "vc_simple_pairs: a first-neighbour index lies outside its table");
    if (npairs)
        *npairs = total;
    return VC_OK;
}
