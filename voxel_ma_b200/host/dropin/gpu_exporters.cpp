// gpu_exporters.cpp -- drop-in for voxelvoro::estimateRadiiField, the body of `-md=r`
// (reference: src/exporters.cpp:554-654; CLI branch src/voroUtility.cpp:332-359).
//
// Same signature, same files in and out.  The reference builds a trimesh::KDtree over the boundary points
// and asks closest_to_pt(v, bbox.radius()^2) once per medial-axis vertex, keeping only the DISTANCE of the
// answer; here the boundary points become the resident sample set and all vertices are answered by one
// vc_closest_points_f32 call (float32 distances in the tree's own operation order, include/voxcore_gpu.h),
// so the radii -- and therefore the output file -- are identical.
#include <cmath>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <trimesh/TriMesh.h>
#include <voxelcore/exporters.h>

#include "../voxcore_session.hpp"

namespace voxelvoro
{
int estimateRadiiField(const char* _ma_file_name, const char* _bndry_pts_file_name, const char* _radii_file_name)
{
    std::unique_ptr<trimesh::TriMesh> ma(trimesh::TriMesh::read(_ma_file_name));
    if (!ma)
    {
        std::cout << "Error reading file " << _ma_file_name << std::endl;
        return -1;
    }
    std::ifstream in(_bndry_pts_file_name);
    if (!in.is_open())
    {
        std::cout << "Error: couldn't open file " << _bndry_pts_file_name << std::endl;
        return -1;
    }
    // header: the first line that is not a '#' comment starts with the point count
    std::string ln;
    while (std::getline(in, ln))
        if (ln.empty() || ln[0] != '#')
            break;
    long n = 0;
    {
        std::istringstream hs(ln);
        hs >> n;
    }
    if (n < 0)
        n = 0;
    // body: "id x y z" per line, blank lines and comments skipped, stored in file order
    std::vector<float> pts((size_t)n * 3, 0.0f);
    long got = 0;
    while (std::getline(in, ln))
    {
        if (ln.empty() || ln[0] == '#')
            continue;
        std::istringstream ls(ln);
        int id;
        float x = 0, y = 0, z = 0;
        ls >> id >> x >> y >> z;
        if (got < n)
        {
            pts[3 * got] = x;
            pts[3 * got + 1] = y;
            pts[3 * got + 2] = z;
        }
        ++got;
    }
    in.close();

    vcgpu::Session& s = vcgpu::Session::get();
    if (!s.ok())
        return -1;
    ma->need_bbox();
    const size_t nv = ma->vertices.size();
    // search limit: the medial axis' bounding-sphere radius squared, as the reference passes it; a
    // non-positive limit means "the tree's own root radius" there (3rdparty/trimesh2/libsrc/KDtree.cc:533-534,
    // :166-173)
    float lim = ma->bbox.radius() * ma->bbox.radius();
    if (!(lim > 0.0f) && n > 0)
    {
        float lo[3] = {pts[0], pts[1], pts[2]}, hi[3] = {pts[0], pts[1], pts[2]};
        for (long i = 1; i < n; ++i)
            for (int d = 0; d < 3; ++d)
            {
                lo[d] = std::min(lo[d], pts[3 * i + d]);
                hi[d] = std::max(hi[d], pts[3 * i + d]);
            }
        const float rx = 0.5f * (hi[0] - lo[0]), ry = 0.5f * (hi[1] - lo[1]), rz = 0.5f * (hi[2] - lo[2]);
        const float r = std::sqrt(rx * rx + ry * ry + rz * rz);
        lim = r * r;
    }
    std::vector<float> q(nv * 3), d2(nv);
    std::vector<int32_t> id(nv);
    for (size_t i = 0; i < nv; ++i)
        for (int d = 0; d < 3; ++d)
            q[3 * i + d] = ma->vertices[i][d];
    if (nv)
    {
        if (!s.set_sites(pts.data(), n) ||
            !s.check(vc_closest_points_f32(s.ctx(), q.data(), (int64_t)nv, lim, id.data(), d2.data()), "vc_closest_points_f32"))
            return -1;
    }
    std::ofstream out(_radii_file_name);
    out << nv << std::endl;
    for (size_t i = 0; i < nv; ++i)
    {
        if (id[i] < 0)
        { // the reference dereferences a null pointer here; say what happened instead
            std::cout << "Error: no boundary point within the search radius of medial-axis vertex " << i << std::endl;
            return -1;
        }
        const auto& v = ma->vertices[i];
        out << v[0] << " " << v[1] << " " << v[2] << " " << std::sqrt(d2[i]) << std::endl;
    }
    out.close();
    return 0;
}
} // namespace voxelvoro
