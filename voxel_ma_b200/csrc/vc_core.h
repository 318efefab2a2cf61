// vc_core.h -- the exact integer arithmetic shared by the kernels (host+device inline functions).
//
// Everything the closest-site transform decides is decided here, in integers, so that "bit-exact
// ids" has one definition.  The functions are __host__ __device__ only so that tests/ can drive
// the very same code line by line on the CPU against the oracle (tests/host_harness.cpp); the
// product calls them from CUDA kernels only.
//
// Geometry.  Grid vertices are the integer lattice points v in [0,n)^3 (voxel centres,
// include/surfacing.h:97-119).  Sites are voxel corners at half-integer coordinates c - 0.5 with
// corner index c in [0,n]^3 (include/surfacing.h:121-136).  In doubled coordinates a vertex is 2v
// (even), a site is 2c-1 (odd), so along one axis the offset is  delta(v,c) = 2(v-c)+1  (odd) and
//     4*d^2 = delta_x^2 + delta_y^2 + delta_z^2      is an exact integer (< 2^26 for n <= 2048).
//
// Contract (SURVEY section 7-1, 3rdparty/ann/src/brute.cpp:56-82): the winner at a vertex is the
// lexicographic minimum of (d^2, site id).  A candidate is carried as one 64-bit word
//     H = (4*d^2_so_far << 32) | id
// so a lexicographic compare is one unsigned compare, and the minimum over all sites separates
// exactly over the three axes: min_s (dx^2+dy^2+dz^2, id) = min_cx [dx^2 + min_cy [dy^2 + min_cz (dz^2, id)]].
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

typedef unsigned long long vc_u64; // (uint64_t is unsigned long on LP64; CUDA atomics want unsigned long long)
#define VC_INF 0xFFFFFFFFFFFFFFFFull

// value of candidate (position p on the corner axis, word H) at target vertex t on that axis
VC_HD vc_u64 vc_eval(vc_u64 H, int p, int t)
{
    int d = 2 * (t - p) + 1;
    return H + ((vc_u64)(uint32_t)(d * d) << 32);
}

// One line of the separable transform: lower envelope of the candidates in[j*stride], j in
// [0,ncand), evaluated at targets t in [0,ntgt); `emit(t, value)` is called for t = ntgt-1 .. 0.
//
// In the doubled coordinate x = 2t+1 a candidate at position p is the parabola
//     f_p(x) = D_p + (x - 2p)^2 ,      g_p := D_p + 4 p^2 ,
// all of the same curvature, so two of them cross exactly once, at x_ab = (g_b - g_a) / (4 (b - a)),
// and the lower envelope visits the surviving candidates in position order.
//
// Forward scan (Maurer-style, integers only -- no division, no stored break points): the stack
// holds the candidates of the envelope so far.  With u < v the two upper entries and w the new
// candidate, v is STRICTLY hidden (above min(f_u, f_w) everywhere) iff x_uv > x_vw, i.e.
//     (g_v - g_u) (w - v)  >  (g_w - g_v) (v - u)          (64-bit products, |.| < 2^40)
// and only then is it dropped; a candidate that merely touches the envelope in one point stays,
// because at that point it may hold the lowest site id.
// Backward scan: at target t the winner is the top unless the entry below is at least as close;
// while  D_sec(t) <= D_best(t)  the top is popped (left of the crossing it never wins again) and
// the minimum is taken on the full (distance, id) words -- which is exactly where the
// reference's tie rule (3rdparty/ann/src/brute.cpp:56-82: equal distance -> lowest id) lives.
//
// Memory behaviour is what bounds this scan (one thread per line, a dependent chain per step), so:
//   - stack: the two upper entries live in registers; the rest is a per-line array of packed
//     8-byte entries  g:27 | id:25 | p:12, CONTIGUOUS per line, so
//     consecutive pops of a thread hit the same 32-byte sector.  (A thread-local CUDA array would
//     interleave the 32 lanes of a warp word by word: lanes at different depths then touch 32
//     different sectors per access.)
//   - candidates: fetched VC_PF at a time into two register banks that alternate (the loads of the
//     next bank are in flight while one is consumed), and requested into L2 VC_PF_L2 candidates
//     ahead.
#ifndef VC_PF
#define VC_PF 4
#endif
#ifndef VC_PF_L2
#define VC_PF_L2 48
#endif

#if defined(__CUDA_ARCH__)
#define VC_LOAD_STREAM(p) __ldcs(p) // read once: evict-first
#define VC_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define VC_LOAD_STREAM(p) (*(p))
#define VC_PREFETCH_L2(p) ((void)0)
#endif

#define VC_MAX_SITE_ID ((1 << 25) - 1) // ids must fit the packed stack entry

// packed stack entry  g:27 | id:25 | p:12   (g = D + 4p^2 < 2^26 + 2^24, p <= 2048), laid out so that both
// 32-bit halves are cheap to build:  hi = g << 5 | id >> 20,  lo = id << 12 | p
VC_HD vc_u64 vc_ent_pack(int g, uint32_t id, int p)
{
    uint32_t hi = ((uint32_t)g << 5) | (id >> 20), lo = (id << 12) | (uint32_t)p;
    return ((vc_u64)hi << 32) | lo;
}
VC_HD void vc_ent_unpack(vc_u64 e, int& g, uint32_t& id, int& p)
{
    uint32_t hi = (uint32_t)(e >> 32), lo = (uint32_t)e;
    p = (int)(lo & 0xFFFu);
    id = (lo >> 12) | ((hi & 31u) << 20);
    g = (int)(hi >> 5);
}

// Storage of the stack entries below the two in registers (depth d = 0 .. q-2).  The scan is written
// against this small interface so the device can keep the upper entries in shared memory
// (vc_edt.cu: StackRing) while the CPU test harness uses a plain array.
//   store(d, e)      entry at depth d := e          (forward scan, d = current storage top + 1)
//   load(d)          entry at depth d               (forward scan pops)
//   begin_drain(d)   the backward scan starts: depths d, d-1, ... 0 will be read in that order
//   drain(d)         entry at depth d, in drain order
struct vc_stack_array
{
    vc_u64* a;
    VC_HD void store(int d, vc_u64 e) { a[d] = e; }
    VC_HD vc_u64 load(int d) { return a[d]; }
    VC_HD void begin_drain(int) {}
    VC_HD vc_u64 drain(int d) { return a[d]; }
};

// emit(t, V, id): the winner at target t has 4*d^2-so-far V and site id `id` (both 0xFFFFFFFF: no candidate)
// mask (nullable): this line's candidate bitmap, bit (j & 31) of mask[j >> 5]; a clear bit means candidate j is
// known to be VC_INF and in[j*stride] is neither read nor prefetched (pass X: columns without sites, whose G1
// entries pass Z never writes).  One word serves 32 candidates and the next one is requested a word ahead.
template <class Stack, class Emit>
VC_HD void vc_envelope_line(const vc_u64* __restrict__ in, long stride, int ncand, int ntgt, Stack& stk, Emit emit,
                            const uint32_t* __restrict__ mask = nullptr)
{
    // two upper entries in registers: top (gt, pt, idt) and second (gs, ps, ids); a = gt - gs, b = pt - ps
    int q = -1, gt = 0, pt = 0, gs = 0, ps = 0, a = 0, b = 1;
    uint32_t idt = 0, ids = 0;
    auto push = [&](vc_u64 H, int j)
    {
        const uint32_t D = (uint32_t)(H >> 32);
        if (D == 0xFFFFFFFFu)
            return; // VC_INF: this position has no candidate
        const int g = (int)D + 4 * j * j;
        // top strictly hidden by (second, new):  (gt-gs)(j-pt) > (g-gt)(pt-ps)
        while (q >= 1 && (long long)a * (long long)(j - pt) > (long long)(g - gt) * (long long)b)
        {
            --q;
            gt = gs;
            pt = ps;
            idt = ids;
            if (q >= 1)
            {
                vc_ent_unpack(stk.load(q - 1), gs, ids, ps);
                a = gt - gs;
                b = pt - ps;
            }
        }
        if (q >= 1)
            stk.store(q - 1, vc_ent_pack(gs, ids, ps));
        if (q >= 0)
        {
            a = g - gt;
            b = j - pt;
            gs = gt;
            ps = pt;
            ids = idt;
        }
        gt = g;
        pt = j;
        idt = (uint32_t)H;
        ++q;
    };
    vc_u64 bankA[VC_PF], bankB[VC_PF];
    // candidate bitmap words: mw covers the 32 candidates of the FETCH in hand (VC_PF divides 32), mw_next the
    // following 32; without a mask every candidate is live
    const int nmw = (ncand + 31) >> 5;
    uint32_t mw = 0xFFFFFFFFu, mw_next = 0xFFFFFFFFu;
    if (mask)
    {
        mw_next = nmw > 0 ? mask[0] : 0u;
    }
#define VC_FETCH(bank, j0)                                                                         \
    if (mask && (((j0) & 31) == 0))                                                                \
    {                                                                                              \
        mw = mw_next;                                                                              \
        mw_next = (((j0) >> 5) + 1 < nmw) ? mask[((j0) >> 5) + 1] : 0u;                             \
    }                                                                                              \
    _Pragma("unroll") for (int k = 0; k < VC_PF; ++k)                                              \
    {                                                                                              \
        bank[k] = ((j0) + k < ncand && ((mw >> (((j0) + k) & 31)) & 1u)) ? VC_LOAD_STREAM(in + (long)((j0) + k) * stride) \
                                                                       : (vc_u64)VC_INF;         \
        if (VC_PF_L2 > 0 && !mask && (j0) + k + VC_PF_L2 < ncand)                                   \
            VC_PREFETCH_L2(in + (long)((j0) + k + VC_PF_L2) * stride);                             \
    }
#define VC_CONSUME(bank, j0) _Pragma("unroll") for (int k = 0; k < VC_PF; ++k) push(bank[k], (j0) + k);
    if (VC_PF_L2 > 0 && !mask)
        for (int j = 0; j < VC_PF_L2 && j < ncand; ++j)
            VC_PREFETCH_L2(in + (long)j * stride);
    VC_FETCH(bankA, 0)
    for (int j0 = 0; j0 < ncand; j0 += 2 * VC_PF)
    {
        VC_FETCH(bankB, j0 + VC_PF)
        VC_CONSUME(bankA, j0)
        VC_FETCH(bankA, j0 + 2 * VC_PF)
        VC_CONSUME(bankB, j0 + VC_PF)
    }
#undef VC_FETCH
#undef VC_CONSUME
    // Backward scan.  With x = 2t+1 a candidate's value is V(t) = g + x (x - 4p); stepping t -> t-1
    // changes it by -m with m = 4 (x - 2p - 1), and m itself by -8: two additions per entry per target.
    // One uniform loop over t (a line without candidates emits all-ones) so that a warp whose lanes each
    // own a line stays convergent at the emit() call and may synchronise inside it.
    stk.begin_drain(q - 2);
    const bool empty = q < 0;
    int x = 2 * (ntgt - 1) + 1;
    uint32_t Vt = (uint32_t)(gt + x * (x - 4 * pt)), Vs = (uint32_t)(gs + x * (x - 4 * ps));
    int mt = 4 * (x - 2 * pt - 1), ms = 4 * (x - 2 * ps - 1);
    for (int t = ntgt - 1; t >= 0; --t)
    {
        uint32_t bV = 0xFFFFFFFFu, bid = 0xFFFFFFFFu;
        if (!empty)
        {
            bV = Vt;
            bid = idt;
            while (q > 0 && Vs <= bV)
            { // the entry below is at least as close: the top never wins again left of here
                if (Vs < bV || ids < bid)
                {
                    bV = Vs;
                    bid = ids;
                }
                --q;
                Vt = Vs;
                mt = ms;
                idt = ids;
                if (q >= 1)
                {
                    vc_ent_unpack(stk.drain(q - 1), gs, ids, ps);
                    Vs = (uint32_t)(gs + x * (x - 4 * ps));
                    ms = 4 * (x - 2 * ps - 1);
                }
            }
        }
        emit(t, bV, bid);
        Vt -= (uint32_t)mt;
        mt -= 8;
        Vs -= (uint32_t)ms;
        ms -= 8;
        x -= 2;
    }
}

// ---- pruned envelope (round 2) ---------------------------------------------------------------------
// The same 1-D pass as a Meijster-style scan on the TOTAL order of the words (4d^2, id): a candidate v "beats" u
// at target t iff (V_v(t), id_v) < (V_u(t), id_u) lexicographically.  For u left of v that holds exactly for
// t >= sep(u, v): V_v(t) - V_u(t) = (g_v - g_u) - 4 (v - u) x with x = 2t + 1 falls monotonically, so with
// c = (id_v < id_u ? 0 : 1) the condition is 4 w x >= dg + c  (w = v - u, dg = g_v - g_u), i.e.
//     sep(u, v) = floor((dg + c - 1 + 4w) / (8w)).
// The stack keeps only candidates that are the strict minimum at >= 1 integer target of [0, ntgt): each entry
// carries the first target `start` at which it beats the entry below it, starts increase strictly up the stack,
// a new candidate pops the top while it beats it at the top's own start and is dropped when its start would lie
// beyond the last target.  The backward scan is then a walk: the top is the winner until t == start, no compares.
// Distance ties never need a second look (the id is part of the order), and the stack depth is bounded by the
// number of distinct winners of the line instead of the number of parabolas touching the real-valued envelope.
//
// The one division needs no division instruction and no table: a candidate whose start would lie beyond the last
// target is recognised by a multiplication (N >= ntgt * 8w) and dropped; for the others N > 0, the quotient is
// below ntgt <= 2048 and M = N >> 3 < 2^22 is exact in float, so trunc(float(M) * rcp(w)) is off by at most one
// (relative error of the reciprocal and of the product <= 2^-22, quotient < 2^11) and one remainder check makes
// it exact -- with the approximate reciprocal of the device (rcp.approx, 1 ulp) as with the host's 1.0f / w, so
// both produce the same integers (tests/test_host_core.py sweeps every divisor).
// Returns the start, or a value >= ntgt when the candidate never wins inside [0, ntgt).
VC_HD int vc_sep(int dg, int c, int w, int ntgt)
{
    const int N = dg + c - 1 + 4 * w; // > 0: the caller's top survived at its own start, so the new start is >= 1
    if (N >= ntgt * 8 * w)
        return ntgt;
    const int M = N >> 3;
#if defined(__CUDA_ARCH__) && defined(WHATIF_NODIV) // timing experiment only: wrong results
    return M;
#else
#if defined(__CUDA_ARCH__)
    float rw;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rw) : "f"((float)w));
    int q = __float2int_rz(__fmul_rn((float)M, rw));
#else
    int q = (int)((float)M * (1.0f / (float)w));
#endif
    const int r = M - q * w;
    q += (r >= w ? 1 : 0) - (r < 0 ? 1 : 0);
    return q;
#endif
}

struct
#if defined(__CUDACC__)
    __align__(16)
#endif
    vc_ent
{
    int g;       // D + 4 p^2
    uint32_t id; // site id
    int p;       // position on the candidate axis
    int start;   // first target at which this entry beats the one below it (0 for the bottom)
};

// Storage of the stack entries below the top (which lives in registers), depths d = 0 .. q-1.  The scan is
// written against this interface so that the device keeps the upper entries in a shared-memory ring that spills
// to global memory (vc_edt.cu: StackRing) while the CPU harness uses a plain array:
//   store(d, e)      entry at depth d := e          (forward scan, d = current depth of the top)
//   load(d)          entry at depth d               (forward scan pops)
//   begin_drain(d)   the backward scan starts: depths d, d-1, ... 0 will be read in exactly that order
//   drain(d)         entry at depth d, in drain order
struct vc_pstack_array
{
    vc_ent* a;
    int maxdepth = 0, npop = 0;
    VC_HD void store(int d, const vc_ent& e)
    {
        a[d] = e;
        if (d + 1 > maxdepth)
            maxdepth = d + 1;
    }
    VC_HD vc_ent load(int d)
    {
        ++npop;
        return a[d];
    }
    VC_HD void begin_drain(int) {}
    VC_HD vc_ent drain(int d) { return a[d]; }
};

// Candidate source of a line: get(k, H, p) hands out the k-th LIVE candidate (k ascending, every k exactly once): its
// word H = 4 D << 32 | id (finite) and its position p on the candidate axis (ascending).  On the device the
// positions are the same for every lane of a warp, so the scan's control flow only diverges in the pops, and the
// words are staged through shared memory several candidates ahead (vc_edt.cu: CandStage); the CPU harness reads
// plain arrays.
struct vc_psource_array
{
    const vc_u64* src; // word of candidate k at src[k * stride]
    long stride;
    const int* pos;
    VC_HD void get(int k, vc_u64& H, int& p) const
    {
        H = src[(long)k * stride];
        p = pos[k];
    }
};

#if defined(__CUDA_ARCH__) && !defined(VC_NO_RECONVERGE)
#define VC_LANES(m) const unsigned m = __ballot_sync(__activemask(), ncand > 0)
#define VC_RECONVERGE(m) __syncwarp(m)
#else
#define VC_LANES(m) (void)0
#define VC_RECONVERGE(m) (void)0
#endif

// One line.  emit(t, V, id) is called for t = ntgt-1 .. 0 (all-ones when the line has no candidate).
template <class Source, class Stack, class Emit>
VC_HD void vc_envelope_pruned(Source& src, int ncand, int ntgt, Stack& stk, Emit emit)
{
    int q = -1; // depth of the top (registers); depths 0 .. q-1 are in stk
    int gt = 0, pt = 0, st = 0, at = 0; // at = 4 (2 st + 1)
    uint32_t idt = 0;
    VC_LANES(lanes); // the lanes that scan a line with candidates (same count for all of them)
    for (int k = 0; k < ncand; ++k)
    {
        vc_u64 H;
        int j;
        src.get(k, H, j);
        const int g = (int)(uint32_t)(H >> 32) + 4 * j * j;
        const uint32_t id = (uint32_t)H;
        while (q >= 0)
        { // new - top at the top's own start; the top survives iff it still wins there
            const int diff = (g - gt) - (j - pt) * at;
            if (diff > 0 || (diff == 0 && id >= idt))
                break;
            --q;
            if (q >= 0)
            {
                const vc_ent e = stk.load(q);
                gt = e.g;
                idt = e.id;
                pt = e.p;
                st = e.start;
                at = 8 * st + 4;
            }
        }
        VC_RECONVERGE(lanes); // device: the lanes that stopped popping early wait here instead of running ahead alone
        int s = 0;
        if (q >= 0)
        {
            s = vc_sep(g - gt, id < idt ? 0 : 1, j - pt, ntgt);
            if (s >= ntgt)
                continue; // never the winner inside the line
            vc_ent e;
            e.g = gt;
            e.id = idt;
            e.p = pt;
            e.start = st;
            stk.store(q, e);
        }
        gt = g;
        idt = id;
        pt = j;
        st = s;
        at = 8 * s + 4;
        ++q;
    }
    // Backward scan: the top is the winner until t == its start.  V(t) = g + x (x - 4p) with x = 2t + 1;
    // V(t-1) = V(t) - m with m = 4 (x - 2p - 1), and m itself falls by 8 per step.
#if defined(__CUDA_ARCH__) && defined(WHATIF_NOBACK) // timing experiment only
    if (q > -5)
    {
        emit(0, (uint32_t)gt, idt);
        return;
    }
#endif
    stk.begin_drain(q - 1);
    int x = 2 * (ntgt - 1) + 1;
    uint32_t V = (uint32_t)(gt + x * (x - 4 * pt));
    int m = 4 * (x - 2 * pt - 1);
    for (int t = ntgt - 1; t >= 0; --t)
    {
        emit(t, q < 0 ? 0xFFFFFFFFu : V, q < 0 ? 0xFFFFFFFFu : idt);
        x -= 2;
        if (t == st && t > 0 && q > 0)
        {
            const vc_ent e = stk.drain(--q);
            gt = e.g;
            idt = e.id;
            pt = e.p;
            st = e.start;
            V = (uint32_t)(gt + x * (x - 4 * pt));
            m = 4 * (x - 2 * pt - 1);
        }
        else
        {
            V -= (uint32_t)m;
            m -= 8;
        }
    }
}

// First pass (along z) straight from a line's sorted site list: entries e[i] = (cz << 32) | id,
// ascending cz.  `lo` = index of the last entry with cz <= vz (or first-1), maintained by the caller.
VC_HD vc_u64 vc_nearest_on_zline(const vc_u64* __restrict__ e, int first, int last, int lo, int vz)
{
    vc_u64 H = VC_INF;
    if (lo >= first)
    {
        vc_u64 en = e[lo];
        int d = 2 * (vz - (int)(en >> 32)) + 1;
        H = ((vc_u64)(uint32_t)(d * d) << 32) | (uint32_t)en;
    }
    if (lo + 1 < last)
    {
        vc_u64 en = e[lo + 1];
        int d = 2 * ((int)(en >> 32) - vz) - 1;
        vc_u64 H2 = ((vc_u64)(uint32_t)(d * d) << 32) | (uint32_t)en;
        H = H2 < H ? H2 : H;
    }
    return H;
}

// ---- site numbering -----------------------------------------------------------------------------
// Surfacer::extractBoundaryVts (src/surfacing.cpp:240-285) numbers a corner by its first encounter
// in the scan: voxels x outer / y / z inner, the 6 neighbours in the order -x,+x,-y,+y,-z,+z
// (include/surfacing.h:170-178), the 4 corners of the shared face in the order of
// include/surfacing.h:184-190.  So  id(corner) = rank of  min over emitters of
//     key = ((x*ny + y)*nz + z) * 24 + o*4 + ii          (validated in SURVEY section 7-2)
// and every emitter of a corner is one of its 8 incident voxels looking at an in-block neighbour.
//
// occ: bit (a*4+b*2+c) = occupancy of voxel (cx-1+a, cy-1+b, cz-1+c), 0 when out of bounds
// inb: same indexing, 1 when that voxel is inside the volume bounds.
// Returns the key, or VC_INF when the corner is not a site.
VC_HD vc_u64 vc_site_key(uint32_t occ, uint32_t inb, int cx, int cy, int cz, int ny, int nz)
{
    if (occ == 0u || (occ & inb) == 0xFFu)
        return VC_INF; // all 8 equal (out-of-bounds voxels read as 0)
    // position of corner slot ci inside m_cornersWRTNbOffset[o], packed 4 bits per (o, ci); 15 = absent
    // o=0:(2,3,7,6) o=1:(0,1,5,4) o=2:(4,6,2,0) o=3:(1,3,7,5) o=4:(0,2,3,1) o=5:(5,7,6,4)
    const uint32_t slot_of[6] = {
        0x23FF10FFu, // o=0: ci2->0 ci3->1 ci7->2 ci6->3        (nibble ci from the right)
        0xFF23FF10u, // o=1: ci0->0 ci1->1 ci5->2 ci4->3
        0xF1F0F2F3u, // o=2: ci4->0 ci6->1 ci2->2 ci0->3
        0x2F3F1F0Fu, // o=3: ci1->0 ci3->1 ci7->2 ci5->3
        0xFFFF2130u, // o=4: ci0->0 ci2->1 ci3->2 ci1->3
        0x1203FFFFu  // o=5: ci5->0 ci7->1 ci6->2 ci4->3
    };
    for (int bit = 0; bit < 8; ++bit)
    { // (a,b,c) lexicographic = increasing x, then y, then z = increasing scan index
        if (!((inb >> bit) & 1u))
            continue;
        int a = bit >> 2, b = (bit >> 1) & 1, c = bit & 1;
        uint32_t me = (occ >> bit) & 1u;
        int o = 6;
        if (((occ >> (bit ^ 1)) & 1u) != me)
            o = c == 0 ? 5 : 4;
        if (((occ >> (bit ^ 2)) & 1u) != me)
            o = b == 0 ? 3 : 2;
        if (((occ >> (bit ^ 4)) & 1u) != me)
            o = a == 0 ? 1 : 0;
        if (o == 6)
            continue;
        // the corner seen from this voxel: +0.5 along an axis when the voxel is the lower one
        int ci = (c == 0 ? 4 : 0) + (a == 0 ? 0 : 2) + (b == 0 ? 1 : 0);
        uint32_t ii = (slot_of[o] >> (4 * ci)) & 0xFu;
        vc_u64 L = ((vc_u64)(cx - 1 + a) * (vc_u64)ny + (vc_u64)(cy - 1 + b)) * (vc_u64)nz +
                     (vc_u64)(cz - 1 + c);
        return L * 24ull + (vc_u64)(o * 4) + ii;
    }
    return VC_INF;
}

// corner record exchanged between ranks: cx | cy<<21 | cz<<42
VC_HD vc_u64 vc_pack_corner(int cx, int cy, int cz)
{
    return (vc_u64)cx | ((vc_u64)cy << 21) | ((vc_u64)cz << 42);
}
VC_HD void vc_unpack_corner(vc_u64 p, int& cx, int& cy, int& cz)
{
    cx = (int)(p & 0x1FFFFFu);
    cy = (int)((p >> 21) & 0x1FFFFFu);
    cz = (int)((p >> 42) & 0x1FFFFFu);
}
