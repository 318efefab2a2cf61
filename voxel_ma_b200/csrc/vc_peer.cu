// vc_peer.cu -- the one exchange of the multi-GPU path, written over peer memory.
//
// Every rank (one process per GPU, one z-slab each) needs the UNION of the boundary-sample records
// (first-encounter key, corner) of all slabs, because site ids are ranks in the global x-major scan
// order (src/surfacing.cpp:240-285).  Round 2: the numbering itself is distributed.  A rank detects the
// records of its own corner planes, sorts THEM by key (1/world of the records), and one kernel stores the
// sorted run into the receive buffer of every rank over NVLink (coalesced 8-byte stores straight into peer
// memory); one warp then posts (sequence number, count) into every rank's header with a system-scope
// release store.  A rank collects by spinning (acquire loads, bounded in time) on its OWN header lines
// until all ranks have posted the current sequence number; the id of a record is then its position in
// the merge of the world sorted runs = its index in its own run + the number of smaller keys in every
// other run (keys are unique: a corner plane belongs to one slab), found per record by binary searches
// that a warp narrows for its 32 consecutive records (k_merge_rank) -- no rank ever sorts the union
// (round 1 did: 0.47 ms per step for 2.7e6 records, replicated on every GPU).  No collective library on
// the data path -- torch.distributed only carries the 64-byte IPC handles once, at set-up.
//
// Receive buffer of a rank (u64 units), double-buffered on the parity of the sequence number so a
// fast rank's next exchange cannot overwrite what a slow rank is still packing:
//     header  [2][VC_MAX_PEERS][8]      line (parity, source): [0] = sequence, [1] = record count,
//                                       [2] = the source's region capacity, [3] = its world size (checked by the receiver:
//                                       region offsets are computed from them, ranks that disagree would overwrite each other)
//     records [2][world][2 * cap]       region (parity, source): keys[cap] | corners[cap]
// A rank may start exchange k+2 only after collecting k+1, i.e. after every rank posted k+1, which
// each does after packing k (stream order): two buffers are enough.
#include <cstring>

#include "vc_internal.h"

#define PEER_HDR_U64 (2 * VC_MAX_PEERS * 8 + 8) // the post lines, then one line of set-up facts: [0] = cap, [1] = world
#define PEER_CFG_OFF (2 * VC_MAX_PEERS * 8)
#define PEER_TIMEOUT_MS_DEFAULT 10000 // a missing rank becomes an error, not a hang (VC_PEER_TIMEOUT_MS / vc_peer_set_timeout)

static size_t peer_rx_u64(int world, int64_t cap) { return (size_t)PEER_HDR_U64 + 2ull * world * 2ull * (size_t)cap; }
static __host__ __device__ inline size_t peer_hdr_off(int parity, int src) { return ((size_t)parity * VC_MAX_PEERS + src) * 8; }
static inline size_t peer_rec_off(int world, int64_t cap, int parity, int src)
{
    return (size_t)PEER_HDR_U64 + ((size_t)parity * world + src) * 2ull * (size_t)cap;
}

struct VcPeerHdr
{
    u64* line[VC_MAX_PEERS]; // header line (parity, my rank) inside rank p's receive buffer
};

// one warp: lane p tells rank p how many records this rank wrote (after they are visible system-wide)
__global__ void k_peer_post(VcPeerHdr hdr, int world, u64 n, u64 seq, u64 cap)
{
    const int p = threadIdx.x;
    if (p >= world)
        return;
    __threadfence_system();
    u64* line = hdr.line[p];
    line[1] = n;
    line[2] = cap;
    line[3] = (u64)world;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(line), "l"(seq) : "memory");
}

// one warp: lane p waits for rank p's post of `seq` in MY header, then the warp turns the counts into
// offsets.  res[0] = total, res[1] = status (0 ok, 1 timeout, 2 a region overflowed, 3 a rank was set up with another
// capacity / world size), off[p] = start of p.  The wait is bounded in TIME (%globaltimer, nanoseconds).
__device__ __forceinline__ u64 vc_globaltimer()
{
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_peer_wait(const u64* __restrict__ my_hdr, int world, u64 seq, u64 cap, u64 timeout_ns, u64* __restrict__ off,
                            u64* __restrict__ res)
{
    const int p = threadIdx.x;
    u64 n = 0;
    int bad = 0;
    if (p < world)
    {
        const u64* line = my_hdr + p * 8;
        const u64 t0 = vc_globaltimer();
        for (;;)
        {
            u64 s;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(s) : "l"(line) : "memory");
            if (s == seq)
                break;
            if (vc_globaltimer() - t0 > timeout_ns)
            {
                bad = 1;
                break;
            }
            __nanosleep(200);
        }
        if (!bad)
        {
            u64 pcap, pworld;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(n) : "l"(line + 1) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(pcap) : "l"(line + 2) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(pworld) : "l"(line + 3) : "memory");
            if (pcap != cap || pworld != (u64)world)
                bad = 3;
            else if (n > cap)
                bad = 2;
        }
    }
    u64 incl = bad ? 0 : n;
    const u64 mine = incl;
    for (int o = 1; o < 32; o <<= 1)
    {
        u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (p >= o)
            incl += t;
    }
    int worst = bad;
    for (int o = 16; o; o >>= 1)
        worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (p < world)
        off[p] = incl - mine;
    // res[2 + k] = the run that is k-th shortest (k_merge_rank starts with the short ones)
    int place = 0;
    for (int t = 0; t < world; ++t)
    {
        const u64 nt = __shfl_sync(0xffffffffu, mine, t);
        place += nt < mine || (nt == mine && t < p);
    }
    if (p < world)
        res[2 + place] = (u64)p;
    if (p == world - 1)
    {
        off[world] = incl;
        res[0] = incl;
        res[1] = (u64)worst;
    }
}

// this rank's sorted run -> the region (parity, my rank) of every rank's receive buffer: keys[0, n) | corners[cap, cap + n)
__global__ void __launch_bounds__(256)
    k_peer_store_run(const u64* __restrict__ ksorted, const u32* __restrict__ order, const u64* __restrict__ corners, u64 n, VcPeerDst dst)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    {
        const u64 k = ksorted[i], pc = corners[order[i]];
        for (int p = 0; p < dst.world; ++p)
        {
            dst.rec[p][i] = k;
            dst.rec[p][dst.cap + i] = pc;
        }
    }
    __threadfence_system(); // the records are in the peers' memory before the count is posted
}

__device__ __forceinline__ u64 peer_lower_bound(const u64* __restrict__ k, u64 lo, u64 hi, u64 key)
{ // first index in [lo, hi) whose key is >= `key`; __ldcg: L2 is where the peers' stores land, never a stale L1 line
    while (lo < hi)
    {
        const u64 mid = (lo + hi) >> 1;
        if (__ldcg(k + mid) < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// id of a record = its index in its own run + the number of smaller keys in every other run.  One warp takes
// 32 * MR_SUB consecutive records of one run: lane q brackets the block's first key in run q and lane 16 + q its last key
// (two full binary searches per other run, done by different lanes at once).  The keys of run q between the two
// brackets -- about as many as the warp's own, consecutive keys of a run are close in the other runs as well -- are then
// staged through shared memory with coalesced loads (unless the bracket is long: where this run has a gap, another
// run may have 1e5 keys between two of this run's), and every lane ranks its keys there: per-lane searches in global
// memory cost a 32-byte sector per 8-byte probe, and their volume, not their latency, is what bounded this kernel
// (0.23 ms for 2.7e6 records whichever way the searches were arranged).  Writes the merged tables in id order.
#ifndef MR_SUB
#define MR_SUB 4
#endif
#ifndef MR_ORDER
#define MR_ORDER 1
#endif
#define MR_CH 256 // keys staged per round and warp
#ifndef MR_STAGE_MAX
#define MR_STAGE_MAX 256 // longer brackets are searched, not staged
#endif
__global__ void __launch_bounds__(256)
    k_merge_rank(const u64* __restrict__ rec, int world, u64 cap, const u64* __restrict__ off, const u64* __restrict__ order,
                 u64* __restrict__ keys_out, u64* __restrict__ corners_out)
{
    __shared__ u64 stage[8][MR_CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // block of 32 * MR_SUB records, numbered run after run, the SHORTEST run first: a short run is a sparse slab, whose
    // consecutive keys have long brackets in every other run -- the slowest warps, so they start first (MR_ORDER)
    u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int p = -1;
    u64 np = 0;
#if MR_ORDER
    for (int k = 0; k < world && p < 0; ++k)
    {
        const int cand = (int)order[k];
        np = off[cand + 1] - off[cand];
        const u64 nb = (np + 32 * MR_SUB - 1) / (32 * MR_SUB);
        if (w < nb)
            p = cand;
        else
            w -= nb;
    }
    if (p < 0)
        return;
#else
    for (p = 0; p < world; ++p)
    {
        np = off[p + 1] - off[p];
        const u64 nb = (np + 32 * MR_SUB - 1) / (32 * MR_SUB);
        if (w < nb)
            break;
        w -= nb;
    }
    if (p >= world)
        return;
#endif
    const u64* kp = rec + (size_t)p * 2ull * cap;
    const u64 i0 = w * 32 * MR_SUB;
    const u64 iend = i0 + 32 * MR_SUB < np ? i0 + 32 * MR_SUB : np;
    const int q = lane & 15;
    u64 bracket = 0;
    if (q < world && q != p)
    {
        const u64* kq = rec + (size_t)q * 2ull * cap;
        bracket = peer_lower_bound(kq, 0, off[q + 1] - off[q], __ldcg(kp + (lane < 16 ? i0 : iend - 1)));
    }
    u64 K[MR_SUB];
    u32 below[MR_SUB];
#pragma unroll
    for (int s = 0; s < MR_SUB; ++s)
    {
        const u64 i = i0 + (u64)s * 32 + lane;
        K[s] = i < iend ? __ldcg(kp + i) : ~0ull;
        below[s] = 0;
    }
    u64* st = stage[warp];
    for (int r = 0; r < world; ++r)
    {
        const u64 a = __shfl_sync(0xffffffffu, bracket, r), b = __shfl_sync(0xffffffffu, bracket, 16 + r);
        if (r == p)
            continue;
        const u64* kr = rec + (size_t)r * 2ull * cap;
        if (b - a > MR_STAGE_MAX)
        { // a gap in this run's coverage facing a dense part of run r: few warps, searched per lane in global memory
            // (the lane's MR_SUB searches advance together: a warp with long brackets in every run is the kernel's tail)
            u32 lo[MR_SUB], len[MR_SUB];
#pragma unroll
            for (int s = 0; s < MR_SUB; ++s)
            {
                lo[s] = (u32)a;
                len[s] = (u32)(b - a);
            }
            for (u32 m = (u32)(b - a); m > 0; m >>= 1)
            {
                u64 probe[MR_SUB];
#pragma unroll
                for (int s = 0; s < MR_SUB; ++s)
                    probe[s] = len[s] ? __ldcg(kr + lo[s] + (len[s] >> 1)) : 0ull;
#pragma unroll
                for (int s = 0; s < MR_SUB; ++s)
                {
                    const u32 half = len[s] >> 1;
                    const bool right = len[s] && probe[s] < K[s];
                    lo[s] = right ? lo[s] + half + 1 : lo[s];
                    len[s] = right ? len[s] - half - 1 : half;
                }
            }
#pragma unroll
            for (int s = 0; s < MR_SUB; ++s)
                below[s] += lo[s];
            continue;
        }
#pragma unroll
        for (int s = 0; s < MR_SUB; ++s)
            below[s] += (u32)a;
        for (u64 c0 = a; c0 < b; c0 += MR_CH)
        {
            const u32 n = b - c0 < MR_CH ? (u32)(b - c0) : MR_CH;
            for (u32 t = lane; t < n; t += 32)
                st[t] = __ldcg(kr + c0 + t);
            __syncwarp();
            u32 lo[MR_SUB], len[MR_SUB];
#pragma unroll
            for (int s = 0; s < MR_SUB; ++s)
            {
                lo[s] = 0;
                len[s] = n;
            }
            for (u32 m = n; m > 0; m >>= 1) // lower_bound of every key among the staged ones, branch-free
            {
#pragma unroll
                for (int s = 0; s < MR_SUB; ++s)
                {
                    const u32 half = len[s] >> 1;
                    const bool right = len[s] && st[lo[s] + half] < K[s];
                    lo[s] = right ? lo[s] + half + 1 : lo[s];
                    len[s] = right ? len[s] - half - 1 : half;
                }
            }
#pragma unroll
            for (int s = 0; s < MR_SUB; ++s)
                below[s] += lo[s];
            __syncwarp();
        }
    }
#pragma unroll
    for (int s = 0; s < MR_SUB; ++s)
    {
        const u64 i = i0 + (u64)s * 32 + lane;
        if (i < iend)
        {
            const u64 gid = i + below[s];
            keys_out[gid] = K[s];
            corners_out[gid] = __ldcg(kp + cap + i);
        }
    }
}

void vc_peer_release(vc_ctx* c)
{
    if (c->peer_ipc)
        for (int p = 0; p < c->peer_world; ++p)
            if (p != c->peer_rank && c->peer_base[p])
                cudaIpcCloseMemHandle(c->peer_base[p]);
    for (int p = 0; p < VC_MAX_PEERS; ++p)
        c->peer_base[p] = nullptr;
    c->peer_rx.release();
    c->peer_all.release();
    c->peer_world = 0;
    c->peer_rank = -1;
    c->peer_cap = 0;
    c->peer_ipc = c->peer_posted = false;
    cudaGetLastError();
}

extern "C"
{
    int vc_peer_create(vc_ctx* c, int world, int rank, int64_t cap, void* handle_out)
    {
        if (!c || world < 1 || world > VC_MAX_PEERS || rank < 0 || rank >= world || cap < 1)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        vc_peer_release(c);
        const size_t bytes = peer_rx_u64(world, cap) * 8;
        VC_CUDA(c, c->peer_rx.ensure(bytes));
        VC_CUDA(c, cudaMemset(c->peer_rx.p, 0, bytes)); // sequence numbers start at 0; the first exchange posts 1
        const u64 cfg[2] = {(u64)cap, (u64)world}; // read by every rank that maps this buffer (peer_check_cfg)
        VC_CUDA(c, cudaMemcpy((u64*)c->peer_rx.p + PEER_CFG_OFF, cfg, sizeof cfg, cudaMemcpyHostToDevice));
        VC_CUDA(c, c->peer_all.ensure((size_t)world * 2ull * (size_t)cap * 8));
        c->peer_world = world;
        c->peer_rank = rank;
        c->peer_cap = cap;
        c->peer_seq = 0;
        if (c->peer_timeout_ms <= 0)
        {
            const char* e = getenv("VC_PEER_TIMEOUT_MS");
            c->peer_timeout_ms = e && atoll(e) > 0 ? atoll(e) : PEER_TIMEOUT_MS_DEFAULT;
        }
        c->peer_base[rank] = c->peer_rx.p;
        if (handle_out)
        {
            cudaIpcMemHandle_t h;
            VC_CUDA(c, cudaIpcGetMemHandle(&h, c->peer_rx.p));
            static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
            memcpy(handle_out, &h, sizeof(h));
        }
        return VC_OK;
    }

    // a peer's receive buffer must have been created with this rank's capacity and world size: the offsets of the
    // regions this rank stores into are computed from them
    static int peer_check_cfg(vc_ctx* c, int p)
    {
        u64 cfg[2] = {0, 0};
        VC_CUDA(c, cudaMemcpy(cfg, (const u64*)c->peer_base[p] + PEER_CFG_OFF, sizeof cfg, cudaMemcpyDefault));
        if (cfg[0] != (u64)c->peer_cap || cfg[1] != (u64)c->peer_world)
        {
            if (c->peer_ipc)
                cudaIpcCloseMemHandle(c->peer_base[p]);
            c->peer_base[p] = nullptr;
            return vc_fail(c, VC_ERR_INVALID, "a rank of the slab group was created with another capacity or world size "
                                              "(vc_peer_create arguments must agree on every rank)");
        }
        return VC_OK;
    }

    int vc_peer_open(vc_ctx* c, const void* handles)
    {
        if (!c || !handles || c->peer_world < 1)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        for (int p = 0; p < c->peer_world; ++p)
        {
            if (p == c->peer_rank)
                continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles + (size_t)p * 64, 64);
            void* base = nullptr;
            VC_CUDA(c, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            c->peer_base[p] = base;
            c->peer_ipc = true;
            VC_TRY(peer_check_cfg(c, p));
        }
        c->peer_ipc = true;
        return VC_OK;
    }

    int vc_peer_open_ptrs(vc_ctx* c, void* const* bases)
    {
        if (!c || !bases || c->peer_world < 1)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        for (int p = 0; p < c->peer_world; ++p)
        {
            if (p == c->peer_rank)
                continue;
            if (!bases[p])
                return vc_fail(c, VC_ERR_INVALID, "vc_peer_open_ptrs: null receive buffer");
            cudaPointerAttributes a;
            VC_CUDA(c, cudaPointerGetAttributes(&a, bases[p]));
            if (a.device != c->device)
            {
                cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return vc_fail(c, VC_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
                cudaGetLastError();
            }
            c->peer_base[p] = bases[p];
            VC_TRY(peer_check_cfg(c, p));
        }
        c->peer_ipc = false;
        return VC_OK;
    }

    void* vc_peer_buffer(vc_ctx* c) { return c ? c->peer_rx.p : nullptr; }

    int vc_peer_set_timeout(vc_ctx* c, int64_t milliseconds)
    {
        if (!c || milliseconds < 1)
            return VC_ERR_INVALID;
        c->peer_timeout_ms = milliseconds;
        return VC_OK;
    }

    int vc_peer_close(vc_ctx* c)
    {
        if (!c)
            return VC_ERR_INVALID;
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        vc_peer_release(c);
        return VC_OK;
    }

    int vc_sites_post_peers(vc_ctx* c)
    {
        if (!c)
            return VC_ERR_INVALID;
        if (c->peer_world < 1)
            return vc_fail(c, VC_ERR_STATE, "vc_sites_post_peers: no peer group (vc_peer_create / vc_peer_open)");
        for (int p = 0; p < c->peer_world; ++p)
            if (!c->peer_base[p])
                return vc_fail(c, VC_ERR_STATE, "vc_sites_post_peers: a peer's receive buffer is not mapped");
        if (c->peer_posted)
            return vc_fail(c, VC_ERR_STATE, "vc_sites_post_peers: the previous exchange was not collected");
        VC_CUDA(c, cudaSetDevice(c->device));
        const u64 seq = c->peer_seq + 1;
        const int parity = (int)(seq & 1);
        VcPeerDst dst;
        VcPeerHdr hdr;
        dst.world = c->peer_world;
        dst.cap = (u64)c->peer_cap;
        for (int p = 0; p < c->peer_world; ++p)
        {
            u64* base = (u64*)c->peer_base[p];
            dst.rec[p] = base + peer_rec_off(c->peer_world, c->peer_cap, parity, c->peer_rank);
            hdr.line[p] = base + peer_hdr_off(parity, c->peer_rank);
        }
        // the records of this slab's corner planes, sorted by key HERE (1/world of the union) ...
        VC_TRY(st_detect_sites(c));
        const int64_t n = c->ncand;
        if (n > c->peer_cap)
            return vc_fail(c, VC_ERR_NOMEM, "vc_sites_post_peers: this rank produced more site records than the capacity given to "
                                            "vc_peer_create");
        VC_CUDA(c, c->sk0.ensure((size_t)(n + 1) * 8));
        VC_CUDA(c, c->sk1.ensure((size_t)(n + 1) * 8));
        VC_CUDA(c, c->sv0.ensure((size_t)(n + 1) * 4));
        VC_CUDA(c, c->sv1.ensure((size_t)(n + 1) * 4));
        u64* k = c->sk0.as<u64>();
        u32* v = c->sv0.as<u32>();
        if (n > 0)
        {
            VC_TRY(vc_sort_records_by_key(c, c->cand_key.as<u64>(), n, &k, &v));
            // ... and stored as one sorted run into every rank's receive region over NVLink
            unsigned blocks = vc_blocks((size_t)n, 256);
            blocks = blocks > (unsigned)c->sm_count * 8 ? (unsigned)c->sm_count * 8 : blocks;
            VC_LAUNCH(c, "peer_store_run", k_peer_store_run, blocks, 256, 0, k, v, c->cand_corner.as<u64>(), (u64)n, dst);
        }
        VC_LAUNCH(c, "peer_post", k_peer_post, 1, 32, 0, hdr, c->peer_world, (u64)n, seq, (u64)c->peer_cap);
        VC_CUDA(c, cudaGetLastError());
        c->peer_seq = seq;
        c->peer_posted = true;
        return VC_OK;
    }

    int vc_sites_collect_peers(vc_ctx* c, int64_t* n_all)
    {
        if (!c)
            return VC_ERR_INVALID;
        if (!c->peer_posted)
            return vc_fail(c, VC_ERR_STATE, "vc_sites_collect_peers: nothing posted");
        VC_CUDA(c, cudaSetDevice(c->device));
        const int world = c->peer_world;
        const int parity = (int)(c->peer_seq & 1);
        u64* base = (u64*)c->peer_rx.p;
        u64* off = c->scratch.as<u64>() + 4; // world + 1 offsets, then res[2 + world] (total, status, runs by length)
        u64* res = off + VC_MAX_PEERS + 1;
        VC_LAUNCH(c, "peer_wait", k_peer_wait, 1, 32, 0, base + peer_hdr_off(parity, 0), world, c->peer_seq,
                  (u64)c->peer_cap, (u64)c->peer_timeout_ms * 1000000ull, off, res);
        u64* keys = c->peer_all.as<u64>();
        u64* corners = keys + (size_t)world * (size_t)c->peer_cap;
        u64* h = (u64*)c->pinned;
        VC_CUDA(c, cudaMemcpyAsync(h, res, 16, cudaMemcpyDeviceToHost, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        if (h[1] == 1) // the post stays pending: the records of this sequence are still collectable by calling again
            return vc_fail(c, VC_ERR_STATE, "vc_sites_collect_peers: timed out waiting for a rank of the slab group (call again to "
                                            "keep waiting; VC_PEER_TIMEOUT_MS / vc_peer_set_timeout set the bound)");
        c->peer_posted = false;
        if (h[1] == 3)
            return vc_fail(c, VC_ERR_INVALID, "vc_sites_collect_peers: a rank of the slab group was created with another capacity or "
                                              "world size (vc_peer_create arguments must agree on every rank)");
        if (h[1] == 2)
            return vc_fail(c, VC_ERR_NOMEM, "vc_sites_collect_peers: a rank produced more site records than the capacity given "
                                            "to vc_peer_create");
        const int64_t n = (int64_t)h[0];
        if (n_all)
            *n_all = n;
        if (n > 0)
        { // merge the sorted runs by ranking: every record straight to its id
            const size_t warps = (size_t)(n + 32 * MR_SUB - 1) / (32 * MR_SUB) + (size_t)world;
            VC_LAUNCH(c, "merge_rank", k_merge_rank, vc_blocks(warps * 32, 256), 256, 0, base + peer_rec_off(world, c->peer_cap, parity, 0),
                      world, (u64)c->peer_cap, off, res + 2, keys, corners);
        }
        return st_finalize_sites(c, keys, corners, n, VC_SITES_PRESORTED);
    }
}
