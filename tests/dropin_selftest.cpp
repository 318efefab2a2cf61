// dropin_selftest.cpp -- TEST INFRASTRUCTURE.  Exercises the two voxelapps drop-ins that the CLI cannot reach
// without a skeleton file: voxelvoro::apps::tag_stable_subset_with_skel and
// match_voro_with_medialcurve.  The binary is linked like main_voroUtility_gpu (reference objects + GPU
// drop-ins), so the strong definitions are the GPU ones; the REFERENCE's own definitions are fetched at run
// time from the unmodified oracle/_ref/libvoxref.so (dlopen + mangled name) and both are run on the same
// VoroInfo (built here from a sphere volume through computeVD) and the same synthetic medial curve.
// Exit code 0 and "selftest OK" when every output agrees.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <vector>

#include <isosurface/volume.h>
#include <voxelcore/densevolume.h>
#include <voxelcore/highlevelalgo.h>
#include <voxelcore/voroinfo.h>
#include <voxelcore/voxelapps.h>

using std::vector;

namespace voxelvoro
{
namespace apps
{ // defined in the drop-in (strong) -- not declared in voxelapps.h
void match_voro_with_medialcurve(const VoroInfo&, const vector<point>&, vector<int>&);
void tag_stable_subset_with_skel(const VoroInfo&, const vector<point>&, const vector<point>&, vector<bool>&, vector<int>&);
} // namespace apps
} // namespace voxelvoro

typedef void (*match_fn)(const voxelvoro::VoroInfo&, const vector<point>&, vector<int>&);
typedef void (*tag_fn)(const voxelvoro::VoroInfo&, const vector<point>&, const vector<point>&, vector<bool>&, vector<int>&);

static unsigned long long g_state = 88172645463325252ull;
static double urand()
{ // xorshift64: the same stream on every machine
    g_state ^= g_state << 13;
    g_state ^= g_state >> 7;
    g_state ^= g_state << 17;
    return (double)(g_state >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv)
{
    const char* reflib = argc > 1 ? argv[1] : "oracle/_ref/libvoxref.so";
    void* h = dlopen(reflib, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
    if (!h)
    {
        std::printf("selftest: cannot open %s: %s\n", reflib, dlerror());
        return 2;
    }
    auto ref_match = (match_fn)dlsym(
        h, "_ZN9voxelvoro4apps27match_voro_with_medialcurveERKNS_8VoroInfoERKSt6vectorIN7trimesh3VecILm3EfEESaIS7_EERS4_IiSaIiEE");
    auto ref_tag = (tag_fn)dlsym(h, "_ZN9voxelvoro4apps27tag_stable_subset_with_skelERKNS_8VoroInfoERKSt6vectorIN7trimesh3VecILm3EfEESaIS7_"
                                    "EESB_RS4_IbSaIbEERS4_IiSaIiEE");
    if (!ref_match || !ref_tag)
    {
        std::printf("selftest: reference symbols not found\n");
        return 2;
    }
    // a 28^3 ball through the (GPU drop-in) front end
    const int n = 28;
    auto tao = std::make_shared<Volume>(n, n, n);
    for (int x = 0; x < n; ++x)
        for (int y = 0; y < n; ++y)
            for (int z = 0; z < n; ++z)
            {
                double dx = x - 13.3, dy = y - 13.6, dz = z - 13.1;
                tao->setDataAt(x, y, z, 95.0 - (dx * dx + dy * dy + dz * dz));
            }
    std::shared_ptr<Volume3DScalar> vol = std::make_shared<DenseVolume>(tao);
    voxelvoro::VoroInfo voro;
    voxelvoro::computeVD(vol, voro);
    if (!voxelvoro::preprocessVoro(voro, vol, false))
    {
        std::printf("selftest: preprocessVoro failed\n");
        return 2;
    }
    const int nv = (int)voro.geom().numVts();
    // medial curve: every 7th valid Voronoi vertex (exact hits and exact ties for the radius query) + random points
    vector<point> mc, skel;
    for (int i = 0; i < nv; i += 7)
        if (voro.isVertexValid(i))
            mc.push_back(voro.geom().getVert(i));
    for (int i = 0; i < 300; ++i)
        mc.emplace_back((float)(6 + 16 * urand()), (float)(6 + 16 * urand()), (float)(6 + 16 * urand()));
    for (size_t i = 0; i < mc.size(); i += 5)
        skel.push_back(mc[i]); // a skeleton vertex sitting ON a medial-curve vertex: nearest d2 = 0, radius = eps
    for (int i = 0; i < 200; ++i)
        skel.emplace_back((float)(6 + 16 * urand()), (float)(6 + 16 * urand()), (float)(6 + 16 * urand()));
    for (int i = 0; i < 60; ++i) // integer lattice points: several medial-curve vertices at exactly the same distance
        skel.emplace_back((float)(int)(8 + 12 * urand()), (float)(int)(8 + 12 * urand()), (float)(int)(8 + 12 * urand()));

    int bad = 0;
    {
        vector<bool> sa(mc.size(), false), sb(mc.size(), false);
        vector<int> ma, mb;
        voxelvoro::apps::tag_stable_subset_with_skel(voro, mc, skel, sa, ma);
        ref_tag(voro, mc, skel, sb, mb);
        size_t nst = 0;
        for (size_t i = 0; i < mc.size(); ++i)
        {
            bad += sa[i] != sb[i];
            nst += sa[i];
        }
        bad += ma != mb;
        std::printf("selftest: tag_stable_subset_with_skel  mc=%zu skel=%zu stable=%zu  %s\n", mc.size(), skel.size(), nst,
                    (ma == mb && sa == sb) ? "identical" : "MISMATCH");
    }
    {
        vector<int> a, b;
        voxelvoro::apps::match_voro_with_medialcurve(voro, mc, a);
        ref_match(voro, mc, b);
        // ids must agree wherever the nearest medial-curve vertex is unique; on exact ties the kd-tree's pick is
        // traversal-dependent, the GPU's is the lowest index, and both must be at the same distance
        size_t diff = 0, tie_ok = 0;
        for (int i = 0; i < nv; ++i)
            if (a[i] != b[i])
            {
                ++diff;
                if (a[i] >= 0 && b[i] >= 0)
                {
                    const point& v = voro.geom().getVert(i);
                    double da = 0, db = 0;
                    for (int d = 0; d < 3; ++d)
                    {
                        double t = (double)v[d] - (double)mc[a[i]][d];
                        da += t * t;
                        t = (double)v[d] - (double)mc[b[i]][d];
                        db += t * t;
                    }
                    if (da == db && a[i] < b[i])
                        ++tie_ok;
                }
            }
        bad += diff != tie_ok;
        std::printf("selftest: match_voro_with_medialcurve  voro vts=%d  differing ids=%zu (all exact ties resolved to the lower "
                    "index: %s)\n", nv, diff, diff == tie_ok ? "yes" : "NO");
    }
    std::printf(bad ? "selftest FAILED\n" : "selftest OK\n");
    return bad ? 1 : 0;
}
