// gpu_voxelapps.cpp -- drop-ins for the two ANN consumers inside voxelvoro::apps::assignScalarToSites
// (reference: src/voxelapps.cpp:184-228 match_voro_with_medialcurve, :309-371 tag_stable_subset_with_skel).
//
// Same signatures, same stdout, same out-parameters.  The kd-tree build + one annkSearch / annkFRSearch per
// query point become batched calls into the C ABI:
//   ANNkd_tree(pts) + annkSearch(q, 1, ..)       -> vc_set_sites + vc_closest_points  (double, squared L2)
//   annkFRSearch(q, r2, 0) ; annkFRSearch(q, r2, k, idx)  -> vc_radius_search (count, then every site in range)
// Nearest-neighbour ids: a query with several medial-curve vertices at EXACTLY the same double distance gets
// the lowest index here (ANNbruteForce's rule) where the kd-tree returns whichever its traversal meets first
// (SURVEY section 7-1); distances and every non-tied id are identical.  The fixed-radius result is used by
// the reference as a set (src/voxelapps.cpp:354-361), so it is identical without caveat, including the
// radius it passes: d2_nearest + eps with eps a LENGTH (src/voxelapps.cpp:335,346) -- reproduced, not fixed.
#include <iostream>
#include <vector>

#include <voxelcore/voroinfo.h>
#include <voxelcore/voxelapps.h>

#include "../voxcore_session.hpp"

namespace
{
[[noreturn]] void die(const char* where)
{
    std::cout << "Error: " << where << " failed on the GPU front end; aborting." << std::endl;
    std::exit(1);
}

// the medial-curve vertices as the resident sample set (ANN widens the same floats to double)
void make_mc_resident(const std::vector<point>& mc_vts, const char* who)
{
    vcgpu::Session& s = vcgpu::Session::get();
    if (!s.ok())
        die(who);
    std::vector<float> xyz(mc_vts.size() * 3);
    for (size_t i = 0; i < mc_vts.size(); ++i)
        for (int d = 0; d < 3; ++d)
            xyz[3 * i + d] = mc_vts[i][d];
    if (!s.set_sites(xyz.data(), (int64_t)mc_vts.size()))
        die(who);
}
} // namespace

namespace voxelvoro
{
namespace apps
{
void match_voro_with_medialcurve(const VoroInfo& _voro, const vector<point>& _mc_vts, vector<int>& _voro_v_to_mc_v)
{
    const int nv = (int)_voro.geom().numVts();
    _voro_v_to_mc_v.assign(nv, -1);
    make_mc_resident(_mc_vts, "match_voro_with_medialcurve");
    std::cout << "Done: inserting MC vts to kdtree" << std::endl;
    // valid Voronoi vertices are the queries, in vertex order
    std::vector<int> which;
    std::vector<double> q;
    for (int i = 0; i < nv; ++i)
    {
        if (!_voro.isVertexValid(i))
            continue;
        const auto& v = _voro.geom().getVert(i);
        which.push_back(i);
        q.push_back(v[0]);
        q.push_back(v[1]);
        q.push_back(v[2]);
    }
    std::vector<int32_t> nn(which.size());
    vcgpu::Session& s = vcgpu::Session::get();
    if (!which.empty() &&
        !s.check(vc_closest_points(s.ctx(), q.data(), (int64_t)which.size(), nn.data(), nullptr), "vc_closest_points"))
        die("match_voro_with_medialcurve");
    for (size_t k = 0; k < which.size(); ++k)
        _voro_v_to_mc_v[which[k]] = nn[k];
    std::cout << "Done: building correspondence between voro vts and MC vts." << std::endl;
}

void tag_stable_subset_with_skel(const VoroInfo& _voro, const vector<point>& _mc_vts, const vector<point>& _skel_vts,
                                 vector<bool>& _is_stable, vector<int>& _mc_to_skel)
{
    make_mc_resident(_mc_vts, "tag_stable_subset_with_skel");
    vcgpu::Session& s = vcgpu::Session::get();
    _mc_to_skel.assign(_mc_vts.size(), -1);
    const float eps = _voro.getInsidePartSize() * 0.00000001f;
    const int64_t nq = (int64_t)_skel_vts.size();
    std::vector<double> q((size_t)nq * 3), d2((size_t)nq), r2((size_t)nq);
    for (int64_t i = 0; i < nq; ++i)
        for (int d = 0; d < 3; ++d)
            q[3 * i + d] = _skel_vts[i][d];
    std::vector<int32_t> nn((size_t)nq), cnt((size_t)nq);
    if (nq)
    {
        if (!s.check(vc_closest_points(s.ctx(), q.data(), nq, nn.data(), d2.data()), "vc_closest_points"))
            die("tag_stable_subset_with_skel");
        for (int64_t i = 0; i < nq; ++i)
            r2[i] = d2[i] + eps; // double + float, as the reference forms the radius argument
        if (!s.check(vc_radius_search(s.ctx(), q.data(), r2.data(), nq, nullptr, cnt.data(), nullptr, nullptr), "vc_radius_search"))
            die("tag_stable_subset_with_skel");
    }
    std::vector<int64_t> off((size_t)nq + 1, 0);
    for (int64_t i = 0; i < nq; ++i)
        off[i + 1] = off[i] + cnt[i];
    std::vector<int32_t> idx((size_t)off[nq] + 1);
    if (off[nq] &&
        !s.check(vc_radius_search(s.ctx(), q.data(), r2.data(), nq, off.data(), nullptr, idx.data(), nullptr), "vc_radius_search"))
        die("tag_stable_subset_with_skel");
    for (int64_t i = 0; i < nq; ++i)
    {
        if (cnt[i] == 0)
            cout << "Warning: skel vertex " << i << " has labeled 0 MC vts as stable!" << endl;
        for (int64_t j = off[i]; j < off[i + 1]; ++j)
        {
            _is_stable[idx[j]] = true;
            _mc_to_skel[idx[j]] = (int)i;
        }
    }
    std::cout << "Done: labeling stable subset of MC." << std::endl;
}
} // namespace apps
} // namespace voxelvoro
