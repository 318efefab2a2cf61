// gpu_voroinfo.cpp -- drop-ins for the data-parallel member functions of voxelvoro::VoroInfo.
//
// Each definition below has the reference's signature and observable behaviour (out-parameters,
// return value, text on stdout); the loop over vertices / faces / incidences the reference runs on
// one host thread is one call into libvoxcore_gpu.so (include/voxcore_gpu.h):
//
//   VoroInfo::loadFromTetgenFiles(tetgenio)   src/voroinfo.cpp:113-279  tagVert per TetGen vertex
//                                             -> one vc_classify_points, then the reference's loader
//   VoroInfo::tagVert                         src/voroinfo.cpp:447-454  -> prefetched tag / 1-point query
//   VoroInfo::tagVtsUsingUniformVol           src/voroinfo.cpp:456-494  -> vc_classify_points
//   VoroInfo::computeInfoRelatedtoSites()     src/voroinfo.cpp:286-325  -> vc_vertex_radii
//   VoroInfo::computeFacesMeasure             src/voroinfo.cpp:1552-1574 -> vc_face_lambda
//   VoroInfo::computeEdgesMeasure(LAMBDA)     src/voroinfo.cpp:1490-1538 -> vc_face_lambda + vc_segment_max
//   VoroInfo::computeVertexMeasure            src/voroinfo.cpp:1432-1488 -> vc_face_lambda + vc_segment_max
//
// The host-side graph walking (which faces touch an edge, which vertices bound a face) stays on the
// host exactly as in the reference: it is pointer chasing through cellcomplex, not arithmetic.
// Built only into the GPU CLI (voxel_ma_b200/host/Makefile.dropin), where the reference's own
// definitions of these symbols are weakened / renamed at object level.
#include <algorithm>
#include <iostream>
#include <memory>
#include <vector>

#include <voxelcore/densevolume.h>
#include <voxelcore/geomalgo.h>
#include <voxelcore/voroinfo.h>

#include "../voxcore_session.hpp"

namespace vcgpu
{
bool make_resident(const std::shared_ptr<Volume3DScalar>& vol); // gpu_surfacing.cpp

static bool resident(const std::shared_ptr<Volume3DScalar>& vol)
{
    return Session::get().volume_is(vol.get()) || make_resident(vol);
}
[[noreturn]] static void die(const char* where)
{ // functions whose reference signature cannot carry an error: the CLI convention is exit code 1
    std::cout << "Error: " << where << " failed on the GPU front end; aborting." << std::endl;
    std::exit(1);
}
// lambda of every face of the complex (vc_face_lambda over m_face_sites), resident sites = _sites
static void all_face_lambdas(const std::vector<point>& sites, const std::vector<ivec2>& face_sites, std::vector<float>& lam)
{
    Session& s = Session::get();
    static_assert(sizeof(point) == 12 && sizeof(ivec2) == 8, "trimesh::Vec is a plain array");
    TraceScope tr("all_face_lambdas (set_sites + vc_face_lambda)");
    {
        TraceScope t1("  set_sites (fingerprint, upload when the set changed)");
        if (!s.set_sites(sites.empty() ? nullptr : &sites[0][0], (int64_t)sites.size()))
            die("vc_set_sites");
    }
    {
        TraceScope t2("  result vector");
        lam.assign(face_sites.size(), 0.0f);
    }
    TraceScope t3("  vc_face_lambda");
    if (!face_sites.empty() &&
        !s.check(vc_face_lambda(s.ctx(), &face_sites[0][0], (int64_t)face_sites.size(), lam.data()), "vc_face_lambda"))
        die("vc_face_lambda");
}
} // namespace vcgpu

namespace voxelvoro
{
// the reference's loader, renamed at object level (a member function is a function whose first argument is `this`)
bool ref_loadFromTetgenFiles(VoroInfo* self, const tetgenio& tetio, shared_ptr<Volume3DScalar> vol) asm(
    "vcref_VoroInfo_loadFromTetgenFiles");

bool VoroInfo::loadFromTetgenFiles(const tetgenio& _tetio, shared_ptr<Volume3DScalar> _vol)
{
    vcgpu::Session& s = vcgpu::Session::get();
    if (_vol && std::dynamic_pointer_cast<DenseVolume>(_vol))
    {
        if (!vcgpu::resident(_vol))
            return false;
        // the loader builds each vertex as point(double, double, double): the same narrowing here
        std::vector<float> xyz((size_t)_tetio.numberofvpoints * 3);
        for (size_t i = 0; i < xyz.size(); ++i)
            xyz[i] = (float)_tetio.vpointlist[i];
        if (!s.prefetch_tags(xyz))
            return false;
    }
    const bool ok = ref_loadFromTetgenFiles(this, _tetio, _vol);
    s.drop_tags();
    return ok;
}

bool VoroInfo::tagVert(const point& _p, const shared_ptr<Volume3DScalar>& _vol) const
{
    if (!std::dynamic_pointer_cast<DenseVolume>(_vol))
    { // not a dense grid: the reference's scalar rule on the volume's own accessor
        ivec3 vox;
        SpaceConverter::fromModelToVox(_p, vox, trimesh::xform::identity());
        return SpaceConverter::voxTaggedAsInside(vox, _vol);
    }
    if (!vcgpu::resident(_vol))
        vcgpu::die("tagVert");
    bool inside = false;
    if (!vcgpu::Session::get().tag(&_p[0], inside))
        vcgpu::die("tagVert");
    return inside;
}

bool VoroInfo::tagVtsUsingUniformVol(const shared_ptr<Volume3DScalar>& _vol)
{
    const size_t nv = m_geom.numVts();
    m_vts_valid.resize(nv, false);
    if (std::dynamic_pointer_cast<DenseVolume>(_vol))
    {
        if (!vcgpu::resident(_vol))
            return false;
        std::vector<float> xyz(nv * 3);
        for (size_t i = 0; i < nv; ++i)
        {
            const point& p = m_geom.getVert(i);
            xyz[3 * i] = p[0], xyz[3 * i + 1] = p[1], xyz[3 * i + 2] = p[2];
        }
        std::vector<uint8_t> tags(nv, 0);
        vcgpu::Session& s = vcgpu::Session::get();
        if (nv && !s.check(vc_classify_points(s.ctx(), xyz.data(), (int64_t)nv, nullptr, tags.data()), "vc_classify_points"))
            return false;
        for (size_t i = 0; i < nv; ++i)
            m_vts_valid[i] = tags[i] != 0;
    }
    else
        for (size_t i = 0; i < nv; ++i)
            m_vts_valid[i] = tagVert(m_geom.getVert(i), _vol);
    // derived flags and the bbox of the inside part, as src/voroinfo.cpp:472-491
    m_edge_valid.resize(m_geom.numEdges(), false);
    for (size_t e = 0; e < m_geom.numEdges(); ++e)
        m_edge_valid[e] = computeEdgeValidity((int)e);
    m_face_valid.resize(m_geom.numFaces(), false);
    for (size_t f = 0; f < m_geom.numFaces(); ++f)
        m_face_valid[f] = computeFaceValidity((int)f);
    m_bbox.clear();
    for (size_t i = 0; i < nv; ++i)
        if (isVertexValid((int)i))
            m_bbox += m_geom.getVert(i);
    return true;
}

void VoroInfo::computeInfoRelatedtoSites()
{
    m_face_sites_valid = true;
    m_r_per_v.resize(m_geom.numVts(), 0.0f);
    // one (vertex, site) incidence per face corner, in the reference's visiting order
    std::vector<int> vts_f, inc_v, inc_s;
    for (size_t fi = 0; fi < m_face_sites.size(); ++fi)
    {
        m_geom.getFaceVRep((int)fi, vts_f);
        for (int vi : vts_f)
        {
            inc_v.push_back(vi);
            inc_s.push_back(m_face_sites[fi][0]);
        }
    }
    std::vector<float> xyz(inc_v.size() * 3), r(inc_v.size(), 0.0f);
    for (size_t k = 0; k < inc_v.size(); ++k)
    {
        const point& p = m_geom.getVert(inc_v[k]);
        xyz[3 * k] = p[0], xyz[3 * k + 1] = p[1], xyz[3 * k + 2] = p[2];
    }
    vcgpu::Session& s = vcgpu::Session::get();
    if (!s.set_sites(m_site_positions.empty() ? nullptr : &m_site_positions[0][0], (int64_t)m_site_positions.size()))
        vcgpu::die("computeInfoRelatedtoSites");
    if (!inc_v.empty() &&
        !s.check(vc_vertex_radii(s.ctx(), xyz.data(), (int64_t)inc_v.size(), inc_s.data(), r.data()), "vc_vertex_radii"))
        vcgpu::die("computeInfoRelatedtoSites");
    // last writer wins, and the reference's consistency report (src/voroinfo.cpp:296-321)
    const float eps = getInsidePartSize() * 1.0e-5f;
    int inconsistent = 0;
    for (size_t k = 0; k < inc_v.size(); ++k)
    {
        float& slot = m_r_per_v[inc_v[k]];
        if (slot > 0.0f && !util::is_equal(r[k], slot, eps))
            ++inconsistent;
        slot = r[k];
    }
    if (inconsistent)
        std::cout << "Potential bug!! inconsistent radius (at " << inconsistent << " voro vts)" << std::endl;
    std::cout << "radii estimated." << std::endl;
    m_r_valid = true;
}

void VoroInfo::computeFacesMeasure(MeasureForMA::meassuretype _mssure_tp, const vector<int>& _faces_indices,
                                   vector<float>& _faces_msure) const
{
    _faces_msure.clear();
    if (_mssure_tp != MeasureForMA::LAMBDA)
        return;
    vcgpu::TraceScope tr("computeFacesMeasure");
    std::vector<ivec2> pairs;
    pairs.reserve(_faces_indices.size());
    for (int fi : _faces_indices)
        pairs.push_back(getSitesOfFace(fi));
    vcgpu::all_face_lambdas(m_site_positions, pairs, _faces_msure);
}

// out[k] = max over the valid faces listed for element k of lambda(face), 0 when there is none
static void max_lambda_over_faces(const VoroInfo& voro, const std::vector<point>& sites, const std::vector<ivec2>& face_sites,
                                  const std::vector<int32_t>& off, const std::vector<int32_t>& items, vector<float>& out)
{
    vcgpu::TraceScope tr("max_lambda_over_faces (lambdas + validity + vc_segment_max)");
    std::vector<float> lam;
    vcgpu::all_face_lambdas(sites, face_sites, lam);
    std::vector<uint8_t> valid(face_sites.size());
    for (size_t f = 0; f < valid.size(); ++f)
        valid[f] = voro.isFaceValid((int)f) ? 1 : 0;
    const int64_t n = (int64_t)off.size() - 1;
    out.assign((size_t)n, 0.0f);
    vcgpu::Session& s = vcgpu::Session::get();
    static const int32_t none = 0;
    if (n > 0 && !lam.empty() &&
        !s.check(vc_segment_max(s.ctx(), off.data(), items.empty() ? &none : items.data(), n, lam.data(), (int64_t)lam.size(),
                                valid.data(), out.data()),
                 "vc_segment_max"))
        vcgpu::die("vc_segment_max");
}

void VoroInfo::computeEdgesMeasure(MeasureForMA::meassuretype _mssure_tp, const vector<int>& _edges_indices,
                                   vector<float>& _edges_msure) const
{
    _edges_msure.clear();
    if (_mssure_tp != MeasureForMA::LAMBDA)
        return;
    vcgpu::TraceScope tr("computeEdgesMeasure");
    std::vector<int32_t> off(1, 0), items;
    for (int ei : _edges_indices)
    {
        const int nf = geom().cntNbFacesofEdge(ei);
        for (int j = 0; j < nf; ++j)
            items.push_back(geom().nbFaceofEdge(ei, j));
        off.push_back((int32_t)items.size());
    }
    max_lambda_over_faces(*this, m_site_positions, m_face_sites, off, items, _edges_msure);
}

void VoroInfo::computeVertexMeasure(MeasureForMA::meassuretype _mssure_tp, const vector<int>& _vts_indices,
                                    vector<float>& _vts_msure) const
{
    _vts_msure.clear();
    if (_mssure_tp != MeasureForMA::LAMBDA)
        return;
    vcgpu::TraceScope tr("computeVertexMeasure");
    std::vector<int32_t> off(1, 0), items;
    for (int vi : _vts_indices)
    {
        const int ne = geom().cntNbEdgesofVert(vi);
        for (int a = 0; a < ne; ++a)
        {
            const int ei = geom().nbEdgeofVert(vi, a);
            const int nf = geom().cntNbFacesofEdge(ei);
            for (int b = 0; b < nf; ++b)
                items.push_back(geom().nbFaceofEdge(ei, b));
        }
        off.push_back((int32_t)items.size());
    }
    max_lambda_over_faces(*this, m_site_positions, m_face_sites, off, items, _vts_msure);
}
} // namespace voxelvoro
