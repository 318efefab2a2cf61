// ref_thinspy.cpp -- TEST INFRASTRUCTURE (never linked into the product): the queue that the UNMODIFIED, compiled
// CellComplexThinning::prune (src/ccthin.cpp:201-270 in libvoxref.so) seeds, captured without restating its scans.
//
// prune() hands its queue to prune_while_iteration(), a call that goes through the PLT of libvoxref.so (both are
// default-visibility members, -fPIC).  This library defines that one member itself; loaded RTLD_GLOBAL before
// libvoxref.so it is the definition the reference's prune() reaches, and all it does is empty the queue into a
// vector.  ref_thin_seed() builds the reference's own cellcomplex / CellComplexThinning from plain arrays the way
// src/highlevelalgo.cpp:819-837 does, calls prune(), and returns that queue together with the state the seeding read
// (reference counts, first incident face / edge, measures after preprocess()) -- the inputs of K6
// (orc_simple_pairs, vc_simple_pairs).  Golden vectors: tests/golden/make_thin_golden.py.
#include <cstdint>
#include <map>
#include <queue>
#include <set>
#include <vector>

#include <trimesh/XForm.h>
#include <voxelcore/cellcomplex.h> // everything ccthin.h includes comes first, so that only ITS class is opened up
#include <voxelcore/commondefs.h>
#define private public // state of the thinning object (the class layout does not depend on access)
#include <voxelcore/ccthin.h>
#undef private

static std::vector<int32_t> g_pairs;

void CellComplexThinning::prune_while_iteration(const set<unsigned>&, float, float, std::queue<simple_pair>& q)
{
    g_pairs.clear();
    while (!q.empty())
    {
        const simple_pair p = q.front();
        q.pop();
        g_pairs.push_back((int32_t)p.type);
        g_pairs.push_back((int32_t)p.idx0);
        g_pairs.push_back((int32_t)p.idx1);
    }
}

extern "C"
{
    // sizes of the complex the reference builds from (vts, edges, tris): it may add the triangles' edges
    // returns the number of seeded pairs, or -1 when an output is too small (capacities in elements)
    int64_t ref_thin_seed(const float* vts, int64_t nv, const int32_t* edges, int64_t ne, const int32_t* tris, int64_t nf,
                          const float* v_m, const float* e_m, const float* f_m, float f_t, float l_t, int32_t* pairs, int64_t pair_cap,
                          int32_t* edge_ref, int32_t* edge_face0, float* edge_measure, int64_t edge_cap, int32_t* vert_ref,
                          int32_t* vert_edge0, float* face_measure, uint8_t* face_to_remove, int32_t* edge_ends, int32_t* face_edges,
                          int64_t* sizes)
    {
        vector<point> V((size_t)nv);
        for (int64_t i = 0; i < nv; ++i)
            V[i] = point(vts[3 * i], vts[3 * i + 1], vts[3 * i + 2]);
        vector<ivec2> E((size_t)ne);
        for (int64_t i = 0; i < ne; ++i)
            E[i] = ivec2(edges[2 * i], edges[2 * i + 1]);
        vector<uTriFace> F((size_t)nf);
        for (int64_t i = 0; i < nf; ++i)
            F[i] = uTriFace(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
        cellcomplex cc(V, E, F);
        sizes[0] = cc.numVts(), sizes[1] = cc.numEdges(), sizes[2] = cc.numFaces();
        if ((int64_t)cc.numEdges() > edge_cap || (int64_t)cc.numEdges() != ne || (int64_t)cc.numFaces() != nf)
            return -1;
        vector<float> vm(v_m, v_m + nv), em(e_m, e_m + ne), fm(f_m, f_m + nf);
        CellComplexThinning th;
        th.setup(&cc);
        th.assignElementValues(vm, em, fm);
        th.preprocess();
        g_pairs.clear();
        th.prune(f_t, l_t, false); // stops after the seeding: prune_while_iteration above
        const int64_t np = (int64_t)g_pairs.size() / 3;
        if (np > pair_cap)
            return -1;
        for (size_t i = 0; i < g_pairs.size(); ++i)
            pairs[i] = g_pairs[i];
        for (int64_t e = 0; e < ne; ++e)
        {
            edge_ref[e] = th.m_ref_edge_per_prune[e];
            edge_face0[e] = edge_ref[e] > 0 ? (int32_t)cc.nbFaceofEdge(e, 0) : 0;
            edge_measure[e] = th.m_measure[CellComplexThinning::EDGE][e];
        }
        for (int64_t v = 0; v < nv; ++v)
        {
            vert_ref[v] = th.m_ref_vert_per_prune[v];
            vert_edge0[v] = vert_ref[v] > 0 ? (int32_t)cc.nbEdgeofVert(v, 0) : 0;
        }
        vector<int> fe;
        for (int64_t f = 0; f < nf; ++f)
        {
            face_measure[f] = th.m_measure[CellComplexThinning::FACE][f];
            face_to_remove[f] = th.m_to_remove_face[f] ? 1 : 0;
            cc.getFaceERep(f, fe); // the incidence lists the reference counts are histograms of
            if (fe.size() != 3)
                return -1;
            for (int k = 0; k < 3; ++k)
                face_edges[3 * f + k] = fe[k];
        }
        for (int64_t e = 0; e < ne; ++e)
        {
            const auto ed = cc.getEdge(e);
            edge_ends[2 * e] = ed[0];
            edge_ends[2 * e + 1] = ed[1];
        }
        return np;
    }
}
