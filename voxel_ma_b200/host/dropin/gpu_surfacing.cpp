// gpu_surfacing.cpp -- drop-in for Surfacer::extractBoundaryVts (reference: src/surfacing.cpp:223-321).
//
// Same signature, same out-parameter, same error codes; the body of the dense branch -- the x-major
// scan over every voxel with a hash map of corners -- becomes: upload the volume, classify it,
// extract the sites on the GPU in the reference's first-encounter numbering, copy them out.
// Any other volume type (the octree walker branch, src/surfacing.cpp:287-311) is not part of the
// data-parallel path and is forwarded to the reference's own code, which the build keeps under the
// name vcref_Surfacer_extractBoundaryVts (voxel_ma_b200/host/Makefile.dropin renames the symbol in
// the reference object; nothing of the reference is copied).
#include <memory>
#include <vector>

#include <voxelcore/densevolume.h>
#include <voxelcore/surfacing.h>

#include "../voxcore_session.hpp"

// the reference's own definition, renamed at object level (a member function is a function whose first argument is `this`)
SurfacerErrCode ref_extractBoundaryVts(Surfacer* self, const std::shared_ptr<Volume3DScalar>& vol, std::vector<point>& vts) asm(
    "vcref_Surfacer_extractBoundaryVts");

#include <thread>

namespace
{
// The CLI reads its volume first (src/voroUtility.cpp:454-470; one fread per voxel, seconds at 256^3) and only then reaches
// the first GPU call.  Starting the CUDA driver / primary context takes seconds on a cold GPU, so it is started here, at
// load time, on a helper thread: by the time computeVD asks for the session the device is up.  Silent: a mode that never
// touches the GPU (or a box without one) sees no message from this; the session reports a missing device when it is used.
struct WarmUp
{
    std::thread t;
    WarmUp()
    {
        t = std::thread([] {
            vcgpu::TraceScope tr("vc_warmup (helper thread, from process start)");
            const char* dev = std::getenv("VC_DEVICE");
            vc_warmup(dev ? std::atoi(dev) : 0);
        });
    }
    ~WarmUp()
    {
        if (t.joinable())
            t.join();
    }
} g_warm_up;
} // namespace

namespace vcgpu
{
// Pulls a dense volume through the reference's accessor into Tao's in-memory order and makes it
// the resident classified volume of the session.
bool make_resident(const std::shared_ptr<Volume3DScalar>& vol)
{
    Session& s = Session::get();
    if (!s.ok())
        return false;
    const int nx = vol->getSizeX(), ny = vol->getSizeY(), nz = vol->getSizeZ();
    TraceScope tr("make_resident (volume pull + upload + classify)");
    std::vector<double> zfast((size_t)nx * ny * nz);
    const Volume3DScalar* v = vol.get(); // the const accessor: a plain array read, safe from several threads
#pragma omp parallel for schedule(static)
    for (int x = 0; x < nx; ++x)
    {
        size_t i = (size_t)x * ny * nz;
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z)
                zfast[i++] = v->getDataAt(x, y, z);
    }
    return s.set_volume(vol.get(), zfast.data(), nx, ny, nz);
}
} // namespace vcgpu

SurfacerErrCode Surfacer::extractBoundaryVts(const shared_ptr<Volume3DScalar>& _vol, vector<point>& _vts)
{
    if (!_vol)
        return SurfacerErrCode::EMPTY_VOL;
    if (!std::dynamic_pointer_cast<DenseVolume>(_vol))
        return ref_extractBoundaryVts(this, _vol, _vts);
    _vts.clear();
    if (!vcgpu::Session::get().ok())
    { // computeVD (src/highlevelalgo.cpp:495) ignores the return code and would hand TetGen an empty set
        std::cout << "Error: no GPU context; exiting." << std::endl;
        std::exit(1);
    }
    vcgpu::TraceScope tr("Surfacer::extractBoundaryVts (GPU)");
    if (!vcgpu::make_resident(_vol))
        return SurfacerErrCode::FAILURE;
    std::vector<float> xyz;
    if (!vcgpu::Session::get().extract_sites(xyz))
        return SurfacerErrCode::FAILURE;
    _vts.reserve(xyz.size() / 3);
    for (size_t i = 0; i + 2 < xyz.size(); i += 3)
        _vts.emplace_back(xyz[i], xyz[i + 1], xyz[i + 2]);
    return SurfacerErrCode::SUCCESS;
}
