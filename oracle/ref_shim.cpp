// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" wrappers over the REAL reference implementation (danielyan86129/voxel_ma, compiled
// from /root/reference by oracle/Makefile.ref into oracle/_ref/libvoxref.so).  Every function
// here only marshals plain arrays into the reference's own types and calls the reference's own
// functions; no algorithm is restated in this file.  It is used (a) to pin oracle/oracle.c (the C
// restatement) and (b) to generate the golden fixtures under tests/golden/.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the resulting library.

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>
#include <trimesh/KDtree.h>
#include <unistd.h>
#include <chrono>

#include <ANN/ANN.h>
#include <isosurface/volume.h>
#include <voxelcore/densevolume.h>
#include <voxelcore/highlevelalgo.h>
#include <voxelcore/measureforMA.h>
#include <voxelcore/spaceinfo.h>
#include <voxelcore/surfacing.h>
#include <voxelcore/voroinfo.h>

using std::shared_ptr;
using std::vector;

namespace
{
// Tao's Volume stores double[x*sy*sz + y*sz + z]  (3rdparty/isosurface_tao/volume.h:217-224)
shared_ptr<Volume3DScalar> make_volume(const double* zfast, int nx, int ny, int nz)
{
    auto v = std::make_shared<Volume>(nx, ny, nz);
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z)
                v->setDataAt(x, y, z, zfast[((size_t)x * ny + y) * nz + z]);
    return std::make_shared<DenseVolume>(v);
}

// Surfacer::init() opens "cycle8.txt" relative to the CWD (src/surfacing.cpp:462).
struct CwdGuard
{
    char old[4096];
    bool ok;
    explicit CwdGuard(const char* dir)
    {
        ok = getcwd(old, sizeof old) != nullptr && dir && chdir(dir) == 0;
    }
    ~CwdGuard()
    {
        if (ok)
            (void)!chdir(old);
    }
};

char g_data_dir[4096] = ".";
// seconds spent inside the reference call proper (marshalling excluded), std::chrono like the
// reference's own struct timer (include/commondefs.h:110-168)
double g_last_seconds = 0.0;
struct Stopwatch
{
    std::chrono::high_resolution_clock::time_point t0 = std::chrono::high_resolution_clock::now();
    ~Stopwatch()
    {
        g_last_seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    }
};

// gives the shim read access to VoroInfo's protected state without touching the reference
struct VoroProbe : public voxelvoro::VoroInfo
{
    const vector<ivec2>& faceSites() const { return m_face_sites; }
    const vector<bool>& vtsValid() const { return m_vts_valid; }
};

struct Pipeline
{
    shared_ptr<Volume3DScalar> vol;
    VoroProbe voro;
    // raw TetGen Voronoi vertices and the reference's tagVert() verdict on each
    vector<float> tet_vpts;
    vector<uint8_t> tet_vtag;
    // extractInsideWithMeasure() outputs
    vector<point> out_vts;
    vector<ivec2> out_edges;
    vector<uTriFace> out_tris;
    vector<int> from_fi;
    vector<float> v_msure, e_msure, f_msure;
    // per-vertex site chosen by computeInfoRelatedtoSites (last writer wins) and its radius
    vector<int> site_of_v;
};
} // namespace

extern "C"
{
    // directory that holds cycle8.txt (oracle/_ref)
    void ref_set_data_dir(const char* dir) { snprintf(g_data_dir, sizeof g_data_dir, "%s", dir); }
    double ref_last_seconds(void) { return g_last_seconds; }

    // a2: SpaceConverter::voxTaggedAsInside for every voxel (include/spaceinfo.h:53-58).
    // out is x-fastest: out[x + nx*(y + ny*z)].
    void ref_classify_grid(const double* zfast, int nx, int ny, int nz, uint8_t* out)
    {
        auto vol = make_volume(zfast, nx, ny, nz);
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x)
                    out[x + (size_t)nx * (y + (size_t)ny * z)] =
                        SpaceConverter::voxTaggedAsInside(ivec3(x, y, z), vol) ? 1 : 0;
    }

    // a4: VoroInfo::tagVert (src/voroinfo.cpp:447-454)
    void ref_tag_points(const double* zfast, int nx, int ny, int nz, const float* xyz, int64_t n,
                        uint8_t* out)
    {
        auto vol = make_volume(zfast, nx, ny, nz);
        voxelvoro::VoroInfo voro;
        for (int64_t i = 0; i < n; ++i)
            out[i] = voro.tagVert(point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), vol) ? 1 : 0;
    }

    // a3: Surfacer::extractBoundaryVts (src/surfacing.cpp:223-321). Returns the site count;
    // writes min(count, cap) points.
    int64_t ref_extract_sites(const double* zfast, int nx, int ny, int nz, float* out_xyz,
                              int64_t cap)
    {
        auto vol = make_volume(zfast, nx, ny, nz);
        CwdGuard g(g_data_dir);
        Surfacer surf;
        vector<point> sites;
        {
            Stopwatch sw;
            if (surf.extractBoundaryVts(vol, sites) != SurfacerErrCode::SUCCESS)
                return -1;
        }
        int64_t n = (int64_t)sites.size();
        for (int64_t i = 0; i < n && i < cap; ++i)
        {
            out_xyz[3 * i] = sites[i][0];
            out_xyz[3 * i + 1] = sites[i][1];
            out_xyz[3 * i + 2] = sites[i][2];
        }
        return n;
    }

    // a5: ANNkd_tree::annkSearch(k=1, eps=0) (3rdparty/ann/src/kd_search.cpp:88-216), built the
    // way the reference builds it (src/voroinfo.cpp:344: default bucket size / split rule).
    void ref_ann_kd_search(const double* data, int n, int dim, const double* q, int64_t nq,
                           int32_t* idx, double* d2)
    {
        ANNpointArray pa = annAllocPts(n, dim);
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d)
                pa[i][d] = data[(size_t)i * dim + d];
        {
            Stopwatch sw; // tree build + queries, as the reference pays for both (src/voroinfo.cpp:344-362)
            ANNkd_tree tree(pa, n, dim);
            ANNpoint qq = annAllocPt(dim);
            for (int64_t i = 0; i < nq; ++i)
            {
                for (int d = 0; d < dim; ++d)
                    qq[d] = q[(size_t)i * dim + d];
                ANNidx id;
                ANNdist dd;
                tree.annkSearch(qq, 1, &id, &dd, 0.0);
                idx[i] = id;
                d2[i] = dd;
            }
            annDeallocPt(qq);
        }
        annDeallocPts(pa);
        annClose();
    }

    // a5 at full size: the same ANNkd_tree::annkSearch(k=1, eps=0), one query per grid vertex (x, y, z) of
    // the planes [z0, z1) -- the dense query set of BASELINE's metric, with no query array to marshal.
    // id_out / d2_out are [z1-z0][ny][nx], x fastest.  One tree per call (per forked worker: ANN keeps its
    // search state in globals, 3rdparty/ann/src/kd_search.cpp:78-82).
    void ref_ann_kd_grid(const double* data, int n, int nx, int ny, int z0, int z1, int32_t* id_out, double* d2_out)
    {
        ANNpointArray pa = annAllocPts(n, 3);
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d)
                pa[i][d] = data[(size_t)i * 3 + d];
        {
            Stopwatch sw;
            ANNkd_tree tree(pa, n, 3);
            ANNpoint qq = annAllocPt(3);
            size_t o = 0;
            for (int z = z0; z < z1; ++z)
                for (int y = 0; y < ny; ++y)
                    for (int x = 0; x < nx; ++x, ++o)
                    {
                        qq[0] = x;
                        qq[1] = y;
                        qq[2] = z;
                        ANNidx id;
                        ANNdist dd;
                        tree.annkSearch(qq, 1, &id, &dd, 0.0);
                        id_out[o] = id;
                        d2_out[o] = dd;
                    }
            annDeallocPt(qq);
        }
        annDeallocPts(pa);
        annClose();
    }

    // f-2: trimesh::KDtree::closest_to_pt as estimateRadiiField calls it (src/exporters.cpp:626-637):
    // idx = index of the returned point (-1 when NULL), d = trimesh::dist(point(closest), v)
    void ref_kdtree_closest(const float* pts, int64_t n, const float* q, int64_t nq, float maxdist2, int32_t* idx,
                            float* d)
    {
        Stopwatch sw;
        trimesh::KDtree tree(pts, (size_t)n);
        for (int64_t i = 0; i < nq; ++i)
        {
            const float* c = tree.closest_to_pt(q + 3 * i, maxdist2);
            idx[i] = c ? (int32_t)((c - pts) / 3) : -1;
            d[i] = c ? trimesh::dist(trimesh::point(c), trimesh::point(q + 3 * i)) : -1.0f;
        }
    }

    // a5 contract: ANNbruteForce::annkSearch (3rdparty/ann/src/brute.cpp:56-82): (d2, lowest id)
    void ref_ann_brute_search(const double* data, int n, int dim, const double* q, int64_t nq,
                              int32_t* idx, double* d2)
    {
        ANNpointArray pa = annAllocPts(n, dim);
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d)
                pa[i][d] = data[(size_t)i * dim + d];
        {
            ANNbruteForce bf(pa, n, dim);
            ANNpoint qq = annAllocPt(dim);
            for (int64_t i = 0; i < nq; ++i)
            {
                for (int d = 0; d < dim; ++d)
                    qq[d] = q[(size_t)i * dim + d];
                ANNidx id;
                ANNdist dd;
                bf.annkSearch(qq, 1, &id, &dd, 0.0);
                idx[i] = id;
                d2[i] = dd;
            }
            annDeallocPt(qq);
        }
        annDeallocPts(pa);
    }

    // fixed-radius search as the reference calls it (src/voxelapps.cpp:346-353): count with k=0,
    // then fetch. Returns the count for query i in cnt[i]; if idx_out != NULL writes up to kmax ids
    // per query (row-major, padded with -1).
    void ref_ann_kd_fr_search(const double* data, int n, int dim, const double* q, int64_t nq,
                              const double* sq_rad, int kmax, int32_t* cnt, int32_t* idx_out,
                              double* d2_out)
    {
        ANNpointArray pa = annAllocPts(n, dim);
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d)
                pa[i][d] = data[(size_t)i * dim + d];
        {
            ANNkd_tree tree(pa, n, dim);
            ANNpoint qq = annAllocPt(dim);
            vector<ANNidx> ids(kmax > 0 ? kmax : 1);
            vector<ANNdist> dds(kmax > 0 ? kmax : 1);
            for (int64_t i = 0; i < nq; ++i)
            {
                for (int d = 0; d < dim; ++d)
                    qq[d] = q[(size_t)i * dim + d];
                int c = tree.annkFRSearch(qq, sq_rad[i], 0, nullptr, nullptr, 0.0);
                cnt[i] = c;
                if (idx_out && kmax > 0)
                {
                    int k = c < kmax ? c : kmax;
                    for (int j = 0; j < kmax; ++j)
                    {
                        idx_out[i * kmax + j] = -1;
                        d2_out[i * kmax + j] = -1.0;
                    }
                    if (k > 0)
                    {
                        tree.annkFRSearch(qq, sq_rad[i], k, ids.data(), dds.data(), 0.0);
                        for (int j = 0; j < k; ++j)
                        {
                            idx_out[i * kmax + j] = ids[j];
                            d2_out[i * kmax + j] = dds[j];
                        }
                    }
                }
            }
            annDeallocPt(qq);
        }
        annDeallocPts(pa);
        annClose();
    }

    // a6: MeasureForMA::lambdaForFace (include/measureforMA_imp.h:1-4)
    void ref_lambda_for_face(const float* a, const float* b, int64_t n, float* out)
    {
        for (int64_t i = 0; i < n; ++i)
            out[i] = MeasureForMA::lambdaForFace(point(a[3 * i], a[3 * i + 1], a[3 * i + 2]),
                                                 point(b[3 * i], b[3 * i + 1], b[3 * i + 2]));
    }

    // ---- the whole reference pipeline on one volume (= -md=vol2ma up to measures) ------------
    // computeVD (TetGen) -> [preprocessVoro] -> extractInsideWithMeasure.
    void* ref_pipeline_run(const double* zfast, int nx, int ny, int nz, int do_preprocess)
    {
        auto* p = new Pipeline;
        p->vol = make_volume(zfast, nx, ny, nz);
        CwdGuard g(g_data_dir);

        // raw TetGen Voronoi vertices + the reference's verdict, for the a4 fixture.  Same calls
        // as computeVD (src/highlevelalgo.cpp:487-529), made here only to see TetGen's output.
        {
            vector<point> sites;
            Surfacer surf;
            surf.extractBoundaryVts(p->vol, sites);
            tetgenio in, out;
            voxelvoro::pts2tetgen(sites, in);
            tetgenbehavior b;
            b.nonodewritten = 1;
            b.noelewritten = 1;
            b.nofacewritten = 1;
            b.voroout = 1;
            b.quiet = 1;
            tetrahedralize(&b, &in, &out);
            p->tet_vpts.resize((size_t)out.numberofvpoints * 3);
            p->tet_vtag.resize(out.numberofvpoints);
            for (int i = 0; i < out.numberofvpoints; ++i)
            {
                point v(out.vpointlist[i * 3], out.vpointlist[i * 3 + 1],
                        out.vpointlist[i * 3 + 2]);
                p->tet_vpts[3 * i] = v[0];
                p->tet_vpts[3 * i + 1] = v[1];
                p->tet_vpts[3 * i + 2] = v[2];
                p->tet_vtag[i] = p->voro.tagVert(v, p->vol) ? 1 : 0;
            }
        }

        voxelvoro::computeVD(p->vol, p->voro);
        // site chosen per vertex by computeInfoRelatedtoSites (src/voroinfo.cpp:298-318):
        // faces in order, first site of the face, last writer wins.
        {
            const auto& fs = p->voro.faceSites();
            p->site_of_v.assign(p->voro.geom().numVts(), -1);
            vector<int> vts_f;
            for (int fi = 0; fi < (int)fs.size(); ++fi)
            {
                p->voro.geom().getFaceVRep(fi, vts_f);
                for (auto vi : vts_f)
                    p->site_of_v[vi] = fs[fi][0];
            }
        }
        if (do_preprocess)
        {
            if (!voxelvoro::preprocessVoro(p->voro, p->vol, false))
            {
                delete p;
                return nullptr;
            }
            p->voro.extractInsideWithMeasure(MeasureForMA::LAMBDA, p->out_vts, p->out_edges,
                                             p->out_tris, p->from_fi, p->v_msure, p->e_msure,
                                             p->f_msure);
        }
        return p;
    }
    void ref_pipeline_free(void* h) { delete (Pipeline*)h; }

    int64_t ref_pipeline_count(void* h, int what)
    {
        auto* p = (Pipeline*)h;
        switch (what)
        {
            case 0: return (int64_t)p->voro.getSitesPosition().size();
            case 1: return (int64_t)p->voro.geom().numVts();
            case 2: return (int64_t)p->voro.geom().numEdges();
            case 3: return (int64_t)p->voro.geom().numFaces();
            case 4: return (int64_t)p->tet_vtag.size();
            case 5: return (int64_t)p->out_vts.size();
            case 6: return (int64_t)p->out_edges.size();
            case 7: return (int64_t)p->out_tris.size();
            case 8: return (int64_t)p->voro.faceSites().size();
        }
        return -1;
    }
    void ref_pipeline_sites(void* h, float* xyz)
    {
        auto* p = (Pipeline*)h;
        const auto& s = p->voro.getSitesPosition();
        for (size_t i = 0; i < s.size(); ++i)
            for (int d = 0; d < 3; ++d)
                xyz[3 * i + d] = s[i][d];
    }
    // state right after computeVD (do_preprocess=0): vertices, radii, per-vertex site
    void ref_pipeline_vts(void* h, float* xyz, float* radii, int32_t* site_of_v)
    {
        auto* p = (Pipeline*)h;
        size_t n = p->voro.geom().numVts();
        const auto& r = p->voro.getRadii();
        for (size_t i = 0; i < n; ++i)
        {
            const auto& v = p->voro.geom().getVert(i);
            for (int d = 0; d < 3; ++d)
                xyz[3 * i + d] = v[d];
            if (radii && i < r.size())
                radii[i] = r[i];
            if (site_of_v && i < p->site_of_v.size())
                site_of_v[i] = p->site_of_v[i];
        }
    }
    // per-face site pair and lambda (computeFacesMeasure over all faces, src/voroinfo.cpp:1552)
    void ref_pipeline_faces(void* h, int32_t* site_pairs, float* lambda)
    {
        auto* p = (Pipeline*)h;
        const auto& fs = p->voro.faceSites();
        vector<int> all(fs.size());
        for (size_t i = 0; i < fs.size(); ++i)
        {
            all[i] = (int)i;
            site_pairs[2 * i] = fs[i][0];
            site_pairs[2 * i + 1] = fs[i][1];
        }
        vector<float> m;
        p->voro.computeFacesMeasure(MeasureForMA::LAMBDA, all, m);
        for (size_t i = 0; i < m.size(); ++i)
            lambda[i] = m[i];
    }
    void ref_pipeline_tet_vpts(void* h, float* xyz, uint8_t* tag)
    {
        auto* p = (Pipeline*)h;
        memcpy(xyz, p->tet_vpts.data(), p->tet_vpts.size() * sizeof(float));
        memcpy(tag, p->tet_vtag.data(), p->tet_vtag.size());
    }
    // extractInsideWithMeasure outputs (after preprocess): the inside complex handed to cellcomplex / CellComplexThinning
    // (src/highlevelalgo.cpp:738-741, 819-822); counts = ref_pipeline_count 5 / 6 / 7
    void ref_pipeline_inside(void* h, float* vts, int32_t* edges, int32_t* tris)
    {
        auto* p = (Pipeline*)h;
        for (size_t i = 0; i < p->out_vts.size(); ++i)
            for (int k = 0; k < 3; ++k)
                vts[3 * i + k] = p->out_vts[i][k];
        for (size_t i = 0; i < p->out_edges.size(); ++i)
            for (int k = 0; k < 2; ++k)
                edges[2 * i + k] = p->out_edges[i][k];
        for (size_t i = 0; i < p->out_tris.size(); ++i)
            for (int k = 0; k < 3; ++k)
                tris[3 * i + k] = (int32_t)p->out_tris[i][k];
    }
    // extractInsideWithMeasure outputs (after preprocess): V/E/F measures
    void ref_pipeline_measures(void* h, float* v_m, float* e_m, float* f_m, int32_t* from_fi)
    {
        auto* p = (Pipeline*)h;
        if (v_m) memcpy(v_m, p->v_msure.data(), p->v_msure.size() * sizeof(float));
        if (e_m) memcpy(e_m, p->e_msure.data(), p->e_msure.size() * sizeof(float));
        if (f_m) memcpy(f_m, p->f_msure.data(), p->f_msure.size() * sizeof(float));
        if (from_fi) memcpy(from_fi, p->from_fi.data(), p->from_fi.size() * sizeof(int));
    }
}
