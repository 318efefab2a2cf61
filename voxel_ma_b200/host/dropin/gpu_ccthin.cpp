// gpu_ccthin.cpp -- drop-in for CellComplexThinning::prune (reference: src/ccthin.cpp:201-405).
//
// Same signature, same stdout, same result.  What prune does before its first queue pop is
// data-parallel and goes to the GPU (include/voxcore_gpu.h, K6):
//   reference counts   cellcomplex::refCntPerVert / refCntPerEdge  -> vc_ref_counts (incidence histogram)
//   queue seeding      the two scans at src/ccthin.cpp:246-270      -> vc_simple_pairs (ordered compaction:
//                                                                      the pairs arrive in the reference's push order)
// The FIFO loop is order-dependent and stays the reference's own
// prune_while_iteration, called unchanged on the queue filled from the GPU result.  The closing
// "what is left" report (src/ccthin.cpp:348-404) is a host-side sanity pass over the final state.
#include <iostream>
#include <queue>
#include <set>
#include <vector>

#include <voxelcore/ccthin.h>

#include "../voxcore_session.hpp"

namespace
{
[[noreturn]] void die(const char* where)
{
    std::cout << "Error: " << where << " failed on the GPU front end; aborting." << std::endl;
    std::exit(1);
}
} // namespace

void CellComplexThinning::prune(float _f_t, float _l_t, bool _remove_small_components)
{
    (void)_remove_small_components; // unused by the reference as well (the block is commented out there)
    std::cout << "f_t, e_t: " << _f_t << ", " << _l_t << std::endl;
    vcgpu::Session& s = vcgpu::Session::get();
    if (!s.ok())
        die("prune");
    const int64_t nE = (int64_t)m_cc->numEdges(), nF = (int64_t)m_cc->numFaces(), nV = (int64_t)m_cc->numVts();

    // 01. reset remove tags, reference counts from the incidence lists
    m_removed[EDGE].assign(nE, false);
    m_removed[FACE].assign(nF, false);
    m_removed[VERTEX].assign(nV, false);
    m_to_remove_face.assign(nF, false);
    vcgpu::TraceScope tr_all("CellComplexThinning::prune (GPU seeding + the reference's loop)");
    std::vector<int32_t> ends, face_edges, f;
    vcgpu::TraceScope* tr = new vcgpu::TraceScope("prune: incidence lists to flat arrays");
    ends.reserve(2 * nE);
    for (int64_t e = 0; e < nE; ++e)
    {
        const auto& ed = m_cc->getEdge(e);
        ends.push_back(ed[0]);
        ends.push_back(ed[1]);
    }
    for (int64_t fi = 0; fi < nF; ++fi)
    {
        m_cc->getFaceERep(fi, f);
        face_edges.insert(face_edges.end(), f.begin(), f.end());
    }
    m_ref_vert_per_prune.assign(nV, 0);
    m_ref_edge_per_prune.assign(nE, 0);
    delete tr;
    tr = new vcgpu::TraceScope("prune: 2 x vc_ref_counts");
    if (!s.check(vc_ref_counts(s.ctx(), ends.data(), (int64_t)ends.size(), nV, m_ref_vert_per_prune.data()), "vc_ref_counts") ||
        !s.check(vc_ref_counts(s.ctx(), face_edges.data(), (int64_t)face_edges.size(), nE, m_ref_edge_per_prune.data()), "vc_ref_counts"))
        die("prune");

    delete tr;
    // 02. seed the queue: first incident face of every edge / first incident edge of every vertex
    std::cout << "init.ing q ..." << std::endl;
    tr = new vcgpu::TraceScope("prune: first-neighbour arrays + vc_simple_pairs + queue fill");
    std::vector<int32_t> edge_face0(nE, 0), vert_edge0(nV, 0);
    for (int64_t e = 0; e < nE; ++e)
        if (m_ref_edge_per_prune[e] > 0)
            edge_face0[e] = m_cc->nbFaceofEdge(e, 0);
    for (int64_t v = 0; v < nV; ++v)
        if (m_ref_vert_per_prune[v] > 0)
            vert_edge0[v] = m_cc->nbEdgeofVert(v, 0);
    std::vector<int32_t> pairs((size_t)(nE + nV) * 3 + 3);
    int64_t np = 0;
    if (!s.check(vc_simple_pairs(s.ctx(), m_ref_edge_per_prune.data(), edge_face0.data(), nE, m_measure[FACE].data(), nullptr, nF, _f_t,
                                 m_ref_vert_per_prune.data(), vert_edge0.data(), nV, m_measure[EDGE].data(), _l_t, pairs.data(),
                                 nE + nV, &np),
                 "vc_simple_pairs"))
        die("prune");
    std::queue<simple_pair> q;
    for (int64_t i = 0; i < np; ++i)
        q.push(simple_pair((simple_pair::spairtype)pairs[3 * i], (unsigned)pairs[3 * i + 1], (unsigned)pairs[3 * i + 2]));
    std::cout << "after init, q size: " << q.size() << "" << std::endl;
    delete tr;

    // 03. iterative retraction: the reference's own loop
    std::set<unsigned> vts_to_debug;
    prune_while_iteration(vts_to_debug, _f_t, _l_t, q);

    // report what is left that should have gone (never prints on a consistent complex)
    for (int64_t ei = 0; ei < nE; ++ei)
    {
        if (m_removed[EDGE][ei] || m_ref_edge_per_prune[ei] != 1 || !edge_vert_pair_below_threshold(ei, _l_t))
            continue;
        const auto& e = m_cc->getEdge(ei);
        for (int k = 0; k < 2; ++k)
            if (m_ref_vert_per_prune[e[k]] == 1)
                std::cout << "edge-vert pair " << ei << "-" << e[k] << "should be removed!" << std::endl;
    }
    for (int64_t fi = 0; fi < nF; ++fi)
    {
        if (m_removed[FACE][fi])
            continue;
        m_cc->getFaceERep(fi, f);
        for (int ei : f)
            if (m_ref_edge_per_prune[ei] == 1 && !m_removed[EDGE][ei] && face_edge_pair_below_threshold(fi, ei, _f_t, _l_t))
                std::cout << "face-edge pair " << fi << "-" << ei << "should have been removed!" << std::endl
                          << "(check simple pair removal logic, or the logic used to perform this test.)" << std::endl;
    }
    // (the reference's third report, faces marked to-remove but kept, cannot fire: nothing sets the mark
    //  since mark_components is commented out at src/ccthin.cpp:296-331)
}
