// densecore_gpu.cpp -- SURVEY section 8(f-4): the voxel core of the DENSE product, without TetGen.
//
//   densecore_gpu <volume.mrc> <out_base> [t1,t2,...]
//
// The reference's vol2ma (src/voroUtility.cpp:454-525) spends 45-50 % of its time in TetGen building the Voronoi
// diagram of the boundary samples (src/highlevelalgo.cpp:503-529) before it filters, measures and thins it.  This tool
// runs the same back end -- the reference's own cellcomplex (src/cellcomplex.cpp:364-491), CellComplexThinning
// (src/ccthin.cpp:201-424, here with the GPU seeding of host/dropin/gpu_ccthin.cpp), Euler characteristic
// (src/geomalgo.cpp) and PLY writer (src/exporters.cpp:342-476), all linked from the reference tree, none copied --
// on a complex that comes straight off the GPU's closest-site grid: vc_medial_quads (csrc/vc_medial.cu) emits the dual
// quad of every grid edge whose end vertices have different closest sites, with lambda = lambdaForFace of the two.
// The complex is cubical, not TetGen's, so its files are compared with the reference's on statistics (counts, Euler
// characteristic, measure range), not on bytes.  Output: <out_base>.ply and <out_base>_thinned<t>.ply, and on stdout
// the same "face measure range" / Euler lines the reference prints.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include <voxelcore/ccthin.h>
#include <voxelcore/cellcomplex.h>
#include <voxelcore/exporters.h>
#include <voxelcore/geomalgo.h>
#include <voxelcore/importers.h>

#include "../voxcore_session.hpp"

namespace vcgpu
{
bool make_resident(const std::shared_ptr<Volume3DScalar>& vol); // host/dropin/gpu_surfacing.cpp
}

using std::cout;
using std::endl;
using std::vector;

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        cout << "usage: densecore_gpu <volume.mrc> <out_base> [t1,t2,...]" << endl;
        return 1;
    }
    vector<double> tt;
    if (argc > 3)
    {
        std::stringstream ss(argv[3]);
        for (std::string tok; std::getline(ss, tok, ',');)
            tt.push_back(atof(tok.c_str()));
    }
    const double t0 = now();
    std::shared_ptr<Volume3DScalar> vol;
    if (voxelvoro::readVolume(argv[1], vol) != voxelvoro::ImportErrCode::SUCCESS || !vol)
    {
        cout << "Error: cannot read volume " << argv[1] << endl;
        return 1;
    }
    const int nx = vol->getSizeX(), ny = vol->getSizeY(), nz = vol->getSizeZ();
    const double t1 = now();
    vcgpu::Session& s = vcgpu::Session::get();
    if (!s.ok() || !vcgpu::make_resident(vol))
    {
        cout << "Error: no GPU context; exiting." << endl;
        return 1;
    }
    int64_t nsites = 0, nq = 0;
    if (!s.check(vc_run_dense(s.ctx(), &nsites), "vc_run_dense") || !s.check(vc_medial_quads_count(s.ctx(), &nq), "vc_medial_quads_count"))
        return 1;
    vector<uint32_t> anchor((size_t)nq + 1);
    vector<uint8_t> axis((size_t)nq + 1);
    vector<int32_t> sa((size_t)nq + 1), sb((size_t)nq + 1);
    vector<float> lam((size_t)nq + 1);
    if (!s.check(vc_medial_quads(s.ctx(), nq + 1, anchor.data(), axis.data(), sa.data(), sb.data(), lam.data(), &nq), "vc_medial_quads"))
        return 1;
    const double t2 = now();
    cout << "time(GPU front end: classify + sites + closest + measures + dual quads) -> " << (t2 - t1) * 1000 << " ms  (" << nsites
         << " sites, " << nq << " quads)" << endl;
    // quads -> vertices (cube centres), unique edges, triangles (a quad = 2 triangles, as the reference triangulates its
    // polygon faces before thinning, src/voroinfo.cpp:624-729)
    static const int AROUND[3][4][3] = {{{0, -1, -1}, {0, 0, -1}, {0, 0, 0}, {0, -1, 0}},
                                        {{-1, 0, -1}, {-1, 0, 0}, {0, 0, 0}, {0, 0, -1}},
                                        {{-1, -1, 0}, {0, -1, 0}, {0, 0, 0}, {-1, 0, 0}}};
    std::unordered_map<uint64_t, int> vid;
    vector<point> vts;
    vector<uTriFace> tris;
    vector<float> tri_msure;
    std::map<std::pair<int, int>, float> emap; // edge -> max lambda of its faces (computeEdgesMeasure, src/voroinfo.cpp:1490-1538)
    auto add_edge = [&](int a, int b, float m)
    {
        auto k = std::make_pair(std::min(a, b), std::max(a, b));
        auto it = emap.find(k);
        if (it == emap.end())
            emap.emplace(k, m);
        else
            it->second = std::max(it->second, m);
    };
    for (int64_t q = 0; q < nq; ++q)
    {
        const int x = anchor[q] % nx, y = (anchor[q] / nx) % ny, z = anchor[q] / ((uint32_t)nx * ny);
        int c[4];
        for (int k = 0; k < 4; ++k)
        {
            const int a = x + AROUND[axis[q]][k][0], b = y + AROUND[axis[q]][k][1], cc = z + AROUND[axis[q]][k][2];
            const uint64_t key = ((uint64_t)cc * (ny + 1) + b) * (nx + 1) + a;
            auto it = vid.find(key);
            if (it == vid.end())
            {
                it = vid.emplace(key, (int)vts.size()).first;
                vts.emplace_back(a + 0.5f, b + 0.5f, cc + 0.5f);
            }
            c[k] = it->second;
        }
        tris.emplace_back(c[0], c[1], c[2]);
        tris.emplace_back(c[0], c[2], c[3]);
        tri_msure.push_back(lam[q]);
        tri_msure.push_back(lam[q]);
        for (int k = 0; k < 4; ++k)
            add_edge(c[k], c[(k + 1) & 3], lam[q]);
        add_edge(c[0], c[2], lam[q]);
    }
    vector<ivec2> edges;
    vector<float> edge_msure, vert_msure(vts.size(), 0.0f);
    for (auto& e : emap)
    {
        edges.emplace_back(e.first.first, e.first.second);
        edge_msure.push_back(e.second);
        vert_msure[e.first.first] = std::max(vert_msure[e.first.first], e.second); // computeVertexMeasure, :1432-1488
        vert_msure[e.first.second] = std::max(vert_msure[e.first.second], e.second);
    }
    if (vts.empty())
    {
        cout << "dense voxel core: 0 quads (no inside part thick enough to hold a grid cube)" << endl;
        return 0;
    }
    eulerchar ec;
    util::computeEulerChar(vts, edges, tris, ec);
    ec.logToConsole("dense voxel core");
    cout << "face measure range: [" << *std::min_element(tri_msure.begin(), tri_msure.end()) << ","
         << *std::max_element(tri_msure.begin(), tri_msure.end()) << "]" << endl;
    // the analogue of VoroInfo::getInsidePartSize (src/voroinfo.cpp:496): largest side of the bounding box of the kept vertices
    point lo = vts[0], hi = vts[0];
    for (auto& p : vts)
        for (int d = 0; d < 3; ++d)
            lo[d] = std::min(lo[d], p[d]), hi[d] = std::max(hi[d], p[d]);
    const float size = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    const std::string base = argv[2];
    voxelvoro::writeToPLY((base + ".ply").c_str(), vts, edges, tris, vert_msure, edge_msure, tri_msure);
    const double t3 = now();
    CellComplexThinning ccthin;
    cellcomplex cc(vts, edges, tris);
    ccthin.setup(&cc);
    ccthin.assignElementValues(vert_msure, edge_msure, tri_msure);
    ccthin.preprocess();
    for (double t : tt)
    {
        vector<point> v2;
        vector<ivec2> e2;
        vector<uTriFace> f2;
        const float thresh = size * (float)t;
        ccthin.prune(thresh, thresh, false);
        ccthin.remainingCC(v2, e2, f2, nullptr);
        eulerchar ec2;
        util::computeEulerChar(v2, e2, f2, ec2);
        std::stringstream ss;
        ss << "dense voxel core after thinning t=" << t;
        ec2.logToConsole(ss.str().c_str());
        if (!v2.empty())
        {
            std::stringstream fn;
            fn << base << "_thinned" << t << ".ply";
            vector<float> none_v(v2.size(), 0.0f), none_e(e2.size(), 0.0f), none_f(f2.size(), 0.0f);
            voxelvoro::writeToPLY(fn.str().c_str(), v2, e2, f2, none_v, none_e, none_f);
        }
    }
    const double t4 = now();
    cout << "time(I/O read) -> " << (t1 - t0) * 1000 << " ms; time(complex + full .ply) -> " << (t3 - t2) * 1000 << " ms; time(THIN) -> "
         << (t4 - t3) * 1000 << " ms; total " << (t4 - t0) << " s" << endl;
    return 0;
}
