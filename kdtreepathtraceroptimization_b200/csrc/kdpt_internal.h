// kdpt_internal.h — context object behind the C ABI (include/kdpt.h). Not installed.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/kdpt.h"
#include "kdpt_core.cuh"
#include "kdpt_flatten.h"

namespace kdpt {

void setError(const std::string& msg);
int cudaFail(cudaError_t e, const char* what);

#define KDPT_CUDA(call)                                              \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return ::kdpt::cudaFail(_e, #call);   \
    } while (0)

struct DeviceBuffer {
    void* p = nullptr; size_t bytes = 0;
    int alloc(size_t n);
    int ensure(size_t n);     // grow-only
    void release();
    template <class T> T* as() const { return (T*)p; }
};

// device copies of the analytic scene (render loop)
struct SceneDev {
    DeviceBuffer geoms, materials, offsets, geomBounds;   // geomBounds: two float4 (padded world AABB lo, hi) per geom
    int numGeoms = 0, numMaterials = 0, numShapes = 0;
    kdpt_camera camera{};
    int traceDepth = 0;
    bool hasCamera = false;
};

} // namespace kdpt

struct kdpt_ctx {
    int device = 0;
    int numSMs = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    kdpt::Tree tree{};                 // device pointers: the reference view (reference-compatible modes)
    kdpt::Tree xtree{};                // the refined view (exact modes); equals `tree` when refinement is off
    kdpt::DeviceBuffer treeMem[25];
    kdpt::FlatTree info;               // host copy of the sizes only (vectors released after upload)
    uint64_t deviceBytes = 0;
    int xNumNodes = 0;                 // nodes of the refined view, 0 = refinement off
    kdpt::SceneDev scene;
    // KDPT_KDVIZ: the node boxes (mins, maxs: 6 floats per reference node) and, from the first render in that mode on, the
    // scene's geoms followed by one cube geom per node, with their padded world boxes
    std::vector<float> nodeBoxes;
    kdpt::DeviceBuffer vizGeoms, vizBounds;
    int vizCount = 0;
    // traversal scratch
    kdpt::DeviceBuffer rays, tmin, hits, counter, binTmp;
    // render loop state
    kdpt::DeviceBuffer image, pathsA, pathsB, isects, scanTmp, keysTmp, hist, accum, gqWork, gqHit;
    int pixelCount = 0;
    float* boundImage = nullptr;        // caller-owned radiance buffer (kdpt_bind_image)
    cudaStream_t ownStream = nullptr;   // `stream` may point at a caller's stream (kdpt_set_stream)
    cudaStream_t copyIn = nullptr, copyOut = nullptr;   // kdpt_trace: upload / download streams of the sliced pipeline
    cudaEvent_t evCopy = nullptr;
    uint64_t launches = 0;              // kernels launched so far
    std::vector<cudaEvent_t> eventPool; // TESTINGMODE event pairs
};

namespace kdpt {
// rays per slice of kdpt_trace's upload / trace / download pipeline (batches of up to two slices go through whole).
// A slice gains copy overlap and loses in the leaf binning, which needs many rays per leaf. Measured on the 10 M-triangle
// soup, 16.7 M rays from pinned host buffers, end to end: on a box with fast host copies (~25 GB/s) 2 M-ray slices 66.6
// Mrays/s, 4 M 83.6, unsliced 84-92; on a box with slow ones (~6.5 GB/s) unsliced drops to 57. 4 M is the robust choice.
constexpr int kTraceSlice = 1 << 22;
// kdpt_scene.cpp — the cube geom boxIntersectionTestBox (intersections.h:51-66) makes of a box: scale = extent, translation =
// centre, matrices as glm builds them (KDPT_KDVIZ)
void nodeBoxGeom(const float* mins, const float* maxs, int materialid, kdpt_geom* out);
// kdpt_api.cu — padded world-space AABB of the unit cube under a geom's transform: (lo, -), (hi, -), 8 floats
void geomWorldBounds(const float* transform, float* lohi8);
// kdpt_trace.cu
int launchTrace(kdpt_ctx* ctx, const kdpt_ray* dRays, const float* dTmin, int n, kdpt_hit* dHits, int mode,
                cudaStream_t stream);
// kdpt_build.cu — KD build on the GPU; same arrays as kdpt_scene.cpp:buildTree. deviceMs (may be NULL): device time from
// the end of the triangle upload to the last kernel, including the per-level host decisions
int buildTreeDevice(const float* tri18, const int* mtl, int n, int maxdepth, int device,
                    std::vector<kdpt_node_bare>& nodes, kdpt::BulkVector<kdpt_tri_bare>& tris, kdpt::BulkVector<int>& triSource,
                    float* deviceMs);
} // namespace kdpt
