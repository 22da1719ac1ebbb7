// kdpt_api.cu — the C ABI of include/kdpt.h: context life cycle (pathtraceInit / pathtraceFree,
// src/pathtrace.cu:201-313) and the traversal-only entry points. The render loop lives in kdpt_render.cu.
#include "kdpt_internal.h"
#include <string.h>
#include <stdlib.h>
#include <math.h>

namespace kdpt {

static thread_local std::string g_lastError;
void setError(const std::string& msg) { g_lastError = msg; }
int cudaFail(cudaError_t e, const char* what)
{
    g_lastError = std::string(what) + ": " + cudaGetErrorString(e);
    return KDPT_ERR_CUDA;
}

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold lifted, so
// that pathtraceFree + pathtraceInit — which the reference runs on every camera move (main.cpp:1134-1137) — reuses the
// previous context's memory instead of paying cudaMalloc/cudaFree for ~1 GB of path state each time.
static cudaStream_t allocStream()
{
    static thread_local cudaStream_t streams[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (!streams[dev]) {
        cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return streams[dev];
}

int DeviceBuffer::alloc(size_t n)
{
    release();
    if (n == 0) n = 16;
    cudaStream_t as = allocStream();
    cudaError_t e = cudaMallocAsync(&p, n, as);
    if (e == cudaSuccess) e = cudaStreamSynchronize(as);       // usable from any stream afterwards
    if (e != cudaSuccess) { p = nullptr; return cudaFail(e, "cudaMallocAsync"); }
    bytes = n;
    return KDPT_OK;
}
int DeviceBuffer::ensure(size_t n) { return n <= bytes && p ? KDPT_OK : alloc(n + n / 4); }
void DeviceBuffer::release()
{
    if (p) {
        cudaDeviceSynchronize();                                // nothing enqueued anywhere may still touch it (what cudaFree implies)
        cudaFreeAsync(p, allocStream());
    }
    p = nullptr; bytes = 0;
}

template <class V, class T = typename V::value_type>
static int upload(DeviceBuffer& buf, const V& v, const T** devPtr, cudaStream_t s)
{
    int rc = buf.alloc(v.size() * sizeof(T));
    if (rc) return rc;
    if (!v.empty()) KDPT_CUDA(cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    *devPtr = (const T*)buf.p;
    return KDPT_OK;
}

} // namespace kdpt

namespace kdpt {
void geomWorldBounds(const float* m, float* lohi8)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
    for (int c = 0; c < 8; c++) {
        double x = (c & 1) ? 0.5 : -0.5, y = (c & 2) ? 0.5 : -0.5, z = (c & 4) ? 0.5 : -0.5;
        for (int a = 0; a < 3; a++) {
            double v = (double)m[a] * x + (double)m[4 + a] * y + (double)m[8 + a] * z + (double)m[12 + a];
            if (!(v == v) || !(fabs(v) < 1e30)) bad = true;
            lo[a] = v < lo[a] ? v : lo[a]; hi[a] = v > hi[a] ? v : hi[a];
        }
    }
    for (int a = 0; a < 3; a++) {
        double pad = 2e-3 + 1e-4 * (hi[a] - lo[a]) + 1e-5 * (fabs(lo[a]) > fabs(hi[a]) ? fabs(lo[a]) : fabs(hi[a]));
        lohi8[a] = bad ? -INFINITY : (float)(lo[a] - pad);         // degenerate transform: never cull
        lohi8[4 + a] = bad ? INFINITY : (float)(hi[a] + pad);
    }
    lohi8[3] = lohi8[7] = 0.f;
}
} // namespace kdpt

using namespace kdpt;

extern "C" {

int kdpt_abi_version(void) { return KDPT_ABI_VERSION; }
const char* kdpt_last_error(void) { return g_lastError.c_str(); }

void kdpt_camera_first_frame(kdpt_camera* cam)
{
    if (!cam) return;
    // separate binary32 operations in glm's order (this file is built with -ffp-contract=off / -fmad=false)
    auto dot3 = [](const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; };
    auto norm3 = [&](const float* a, float* o) { float s = 1.0f / sqrtf(dot3(a, a)); o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; };
    auto cross3 = [](const float* x, const float* y, float* o) {
        float r0 = x[1] * y[2] - y[1] * x[2], r1 = x[2] * y[0] - y[2] * x[0], r2 = x[0] * y[1] - y[0] * x[1];
        o[0] = r0; o[1] = r1; o[2] = r2;
    };
    float viewXZ[3] = {cam->view[0], 0.0f, cam->view[2]}, viewZY[3] = {0.0f, cam->view[1], cam->view[2]};
    float nxz[3], nzy[3];
    norm3(viewXZ, nxz); norm3(viewZY, nzy);
    const float fwd[3] = {0, 0, -1}, upAxis[3] = {0, 1, 0};
    float phi = acosf(dot3(nxz, fwd));
    float theta = acosf(dot3(nzy, upAxis));
    float dlt[3] = {cam->position[0] - cam->lookAt[0], cam->position[1] - cam->lookAt[1], cam->position[2] - cam->lookAt[2]};
    float zoom = sqrtf(dot3(dlt, dlt));
    float cp[3] = {zoom * sinf(phi) * sinf(theta) + 0.f, zoom * cosf(theta) + 0.f, zoom * cosf(phi) * sinf(theta) + 0.f};
    float n[3]; norm3(cp, n);
    float v[3] = {-n[0], -n[1], -n[2]}, r[3], u[3];
    cross3(v, upAxis, r); cross3(r, v, u);
    for (int k = 0; k < 3; k++) {
        cam->view[k] = v[k]; cam->up[k] = u[k]; cam->right[k] = r[k];
        cam->position[k] = cp[k] + (cam->lookAt[k] + 0.f);
    }
}

int kdpt_destroy(kdpt_ctx* ctx)
{
    if (!ctx) return KDPT_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->eventPool) cudaEventDestroy(e);
    for (auto& b : ctx->treeMem) b.release();
    DeviceBuffer* bufs[] = {&ctx->scene.geoms, &ctx->scene.materials, &ctx->scene.offsets, &ctx->scene.geomBounds, &ctx->rays, &ctx->tmin,
                            &ctx->hits, &ctx->counter, &ctx->image, &ctx->pathsA, &ctx->pathsB, &ctx->isects,
                            &ctx->scanTmp, &ctx->keysTmp, &ctx->hist, &ctx->accum, &ctx->binTmp, &ctx->gqWork, &ctx->gqHit,
                            &ctx->vizGeoms, &ctx->vizBounds};
    for (auto* b : bufs) b->release();
    cudaEvent_t evs[] = {ctx->ev0, ctx->ev1, ctx->ev2, ctx->ev3};
    for (auto e : evs) if (e) cudaEventDestroy(e);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    if (ctx->copyIn) cudaStreamDestroy(ctx->copyIn);
    if (ctx->copyOut) cudaStreamDestroy(ctx->copyOut);
    if (ctx->evCopy) cudaEventDestroy(ctx->evCopy);
    delete ctx;
    return KDPT_OK;
}


int kdpt_create(const kdpt_scene* scene, int device, kdpt_ctx** out) { return kdpt_create_opts(scene, device, nullptr, out); }

int kdpt_create_opts(const kdpt_scene* scene, int device, const kdpt_build_opts* bopts, kdpt_ctx** out)
{
    static_assert(sizeof(kdpt_geom) == 236 && sizeof(kdpt_material) == 56 && sizeof(kdpt_camera) == 84, "reference POD ABI");
    static_assert(sizeof(kdpt_ray) == 24 && sizeof(kdpt_hit) == 32, "kdpt ABI");
    if (!scene || !out) { setError("kdpt_create: null argument"); return KDPT_ERR_ARG; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        setError(std::string("kdpt_create: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                 "); this library has no CPU path");
        return KDPT_ERR_CUDA;
    }
    if (device < 0 || device >= count) { setError("kdpt_create: device ordinal out of range"); return KDPT_ERR_ARG; }
    FlatTree flat; std::string err;
    FlattenOpts fopts;
    if (bopts) {
        if (bopts->refineLeafMax >= 0) fopts.refineLeafMax = bopts->refineLeafMax;
        if (bopts->refineTraversalCost > 0.0f) fopts.traversalCost = bopts->refineTraversalCost;
        if (bopts->nodeForm == 1 || bopts->nodeForm == 2) fopts.wideNodes = bopts->nodeForm == 2;
    }
    if (!flatten(scene->nodes, scene->numNodes, scene->tris, scene->numTris, flat, err, fopts)) { setError(err); return KDPT_ERR_ARG; }
    if (scene->numGeoms < 0 || scene->numMaterials < 0 || scene->numShapes < 0 ||
        (scene->numGeoms > 0 && !scene->geoms) || (scene->numMaterials > 0 && !scene->materials) ||
        (scene->numShapes > 0 && !scene->objMaterialOffsets)) {
        setError("kdpt_create: bad geoms/materials/offsets");
        return KDPT_ERR_ARG;
    }
    KDPT_CUDA(cudaSetDevice(device));
    kdpt_ctx* ctx = new kdpt_ctx();
    ctx->device = device;
    int rc = KDPT_OK;
    auto fail = [&](int code) { kdpt_destroy(ctx); return code; };
    // (cudaGetDeviceProperties costs about a millisecond; the SM count is all that is needed)
    if ((e = cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail(cudaFail(e, "cudaDeviceGetAttribute"));
    if ((e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cudaFail(e, "cudaStreamCreate"));
    ctx->stream = ctx->ownStream;
    cudaEvent_t* evs[] = {&ctx->ev0, &ctx->ev1, &ctx->ev2, &ctx->ev3};
    for (auto pe : evs) if ((e = cudaEventCreate(pe)) != cudaSuccess) return fail(cudaFail(e, "cudaEventCreate"));

    Tree& T = ctx->tree;
    cudaStream_t s = ctx->stream;
    if ((rc = upload(ctx->treeMem[0], flat.nodes, (const HostNode**)&T.nodes, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[1], flat.boxLo, (const HostF4**)&T.boxLo, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[2], flat.boxHi, (const HostF4**)&T.boxHi, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[3], flat.triA, (const HostF4**)&T.triA, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[4], flat.triB, (const HostF4**)&T.triB, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[5], flat.triC, (const HostF2**)&T.triC, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[6], flat.nrmA, (const HostF4**)&T.nrmA, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[7], flat.nrmB, (const HostF4**)&T.nrmB, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[8], flat.nrmC, (const HostF2**)&T.nrmC, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[9], flat.triOrig, &T.triOrig, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[10], flat.chunkLo, (const HostF4**)&T.chunkLo, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[11], flat.chunkHi, (const HostF4**)&T.chunkHi, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[12], flat.superLo, (const HostF4**)&T.superLo, s))) return fail(rc);
    if ((rc = upload(ctx->treeMem[13], flat.superHi, (const HostF4**)&T.superHi, s))) return fail(rc);
    T.entryOf = nullptr; T.top = nullptr; T.numTop = 0; T.wide = nullptr; T.wideRoot = 0;
    T.numNodes = flat.numNodes; T.numTris = flat.numTris; T.numRefNodes = flat.numNodes;
    for (int a = 0; a < 3; a++) { T.rootLo[a] = flat.rootLo[a]; T.rootHi[a] = flat.rootHi[a]; }
    for (int k = 0; k < 4; k++) T.tagCand[k] = flat.tagCand[k];
    T.materialSize = scene->numMaterials;
    // lane-private leaf steps are cheap while the hot triangle arrays (40 B/entry) stay L1-resident, ~3x a coalesced
    // cooperative step once every lane streams its own lines from L2 (measured: bunny/dragon_5 vs the 1M soup)
    { size_t hot = (size_t)flat.numTris * 40; T.laneCost32 = hot > ((size_t)4 << 20) ? 192 : (hot > ((size_t)512 << 10) ? 64 : 32); }
    if (const char* e = getenv("KDPT_LANE_COST32")) { int v = atoi(e); if (v > 0) T.laneCost32 = v; }
    ctx->deviceBytes = flat.bytes();
    // the refined view of the same scene for the exact modes (kdpt_flatten.cpp:refineLeaves)
    ctx->xtree = T;
    if (!flat.xnodes.empty()) {
        Tree& X = ctx->xtree;
        if ((rc = upload(ctx->treeMem[14], flat.xnodes, (const HostNode**)&X.nodes, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[15], flat.xtriA, (const HostF4**)&X.triA, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[16], flat.xtriB, (const HostF4**)&X.triB, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[17], flat.xtriC, (const HostF2**)&X.triC, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[18], flat.xchunkLo, (const HostF4**)&X.chunkLo, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[19], flat.xchunkHi, (const HostF4**)&X.chunkHi, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[20], flat.xsuperLo, (const HostF4**)&X.superLo, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[21], flat.xsuperHi, (const HostF4**)&X.superHi, s))) return fail(rc);
        if ((rc = upload(ctx->treeMem[22], flat.entryOf, &X.entryOf, s))) return fail(rc);
        if (!flat.xtop.empty()) {
            if ((rc = upload(ctx->treeMem[23], flat.xtop, (const HostNode**)&X.top, s))) return fail(rc);
            X.numTop = (int)flat.xtop.size();
        }
        if (!flat.xwide.empty()) {
            if ((rc = upload(ctx->treeMem[24], flat.xwide, (const HostNode**)&X.wide, s))) return fail(rc);
            X.wideRoot = flat.xwideRoot;
        }
        X.numNodes = (int)flat.xnodes.size(); X.numTris = flat.xNumTris;
        { size_t hot = (size_t)flat.xNumTris * 40; X.laneCost32 = hot > ((size_t)4 << 20) ? 192 : (hot > ((size_t)512 << 10) ? 64 : 32); }
        if (const char* e = getenv("KDPT_LANE_COST32")) { int v = atoi(e); if (v > 0) X.laneCost32 = v; }
    }

    SceneDev& S = ctx->scene;
    S.numGeoms = scene->numGeoms; S.numMaterials = scene->numMaterials; S.numShapes = scene->numShapes;
    if ((rc = S.geoms.alloc(sizeof(kdpt_geom) * (size_t)S.numGeoms))) return fail(rc);
    if ((rc = S.materials.alloc(sizeof(kdpt_material) * (size_t)S.numMaterials))) return fail(rc);
    if ((rc = S.offsets.alloc(sizeof(int) * (size_t)S.numShapes))) return fail(rc);
    if (S.numGeoms && (e = cudaMemcpyAsync(S.geoms.p, scene->geoms, sizeof(kdpt_geom) * (size_t)S.numGeoms, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(cudaFail(e, "upload geoms"));
    if (S.numMaterials && (e = cudaMemcpyAsync(S.materials.p, scene->materials, sizeof(kdpt_material) * (size_t)S.numMaterials, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(cudaFail(e, "upload materials"));
    if (S.numShapes && (e = cudaMemcpyAsync(S.offsets.p, scene->objMaterialOffsets, sizeof(int) * (size_t)S.numShapes, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(cudaFail(e, "upload offsets"));
    {
        // World-space AABB of every analytic geom (the unit cube under `transform`; it also contains the radius-0.5
        // sphere), padded: a ray that misses the padded box provably misses the geom, so the kernel can skip the exact
        // (and expensive) object-space test without changing any result. Two float4 per geom: (lo, -), (hi, -).
        std::vector<float> gb(8 * (size_t)(S.numGeoms > 0 ? S.numGeoms : 1), 0.f);
        for (int g = 0; g < S.numGeoms; g++) geomWorldBounds(scene->geoms[g].transform, &gb[8 * (size_t)g]);
        if ((rc = S.geomBounds.alloc(sizeof(float) * gb.size()))) return fail(rc);
        if ((e = cudaMemcpyAsync(S.geomBounds.p, gb.data(), sizeof(float) * gb.size(), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(cudaFail(e, "upload geom bounds"));
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(cudaFail(e, "cudaStreamSynchronize"));
    }
    ctx->nodeBoxes.resize(6 * (size_t)scene->numNodes);
    for (int i = 0; i < scene->numNodes; i++)
        for (int a = 0; a < 3; a++) { ctx->nodeBoxes[6 * (size_t)i + a] = scene->nodes[i].mins[a]; ctx->nodeBoxes[6 * (size_t)i + 3 + a] = scene->nodes[i].maxs[a]; }
    S.camera = scene->camera; S.traceDepth = scene->traceDepth;
    S.hasCamera = scene->camera.resolution[0] > 0 && scene->camera.resolution[1] > 0;
    if ((rc = ctx->counter.alloc(256))) return fail(rc);
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(cudaFail(e, "cudaStreamSynchronize"));
    // keep only the sizes on the host
    flat.nodes.clear(); flat.nodes.shrink_to_fit(); flat.boxLo.clear(); flat.boxLo.shrink_to_fit();
    flat.boxHi.clear(); flat.boxHi.shrink_to_fit(); flat.triA.clear(); flat.triA.shrink_to_fit();
    flat.triB.clear(); flat.triB.shrink_to_fit(); flat.triC.clear(); flat.triC.shrink_to_fit();
    flat.nrmA.clear(); flat.nrmA.shrink_to_fit(); flat.nrmB.clear(); flat.nrmB.shrink_to_fit();
    flat.nrmC.clear(); flat.nrmC.shrink_to_fit(); flat.triOrig.clear(); flat.triOrig.shrink_to_fit();
    flat.chunkLo.clear(); flat.chunkLo.shrink_to_fit(); flat.chunkHi.clear(); flat.chunkHi.shrink_to_fit();
    flat.superLo.clear(); flat.superLo.shrink_to_fit(); flat.superHi.clear(); flat.superHi.shrink_to_fit();
    ctx->xNumNodes = (int)flat.xnodes.size();
    flat.xnodes.clear(); flat.xnodes.shrink_to_fit(); flat.xwide.clear(); flat.xwide.shrink_to_fit(); flat.xtriA.clear(); flat.xtriA.shrink_to_fit();
    flat.xtriB.clear(); flat.xtriB.shrink_to_fit(); flat.xtriC.clear(); flat.xtriC.shrink_to_fit();
    flat.xchunkLo.clear(); flat.xchunkLo.shrink_to_fit(); flat.xchunkHi.clear(); flat.xchunkHi.shrink_to_fit();
    flat.xsuperLo.clear(); flat.xsuperLo.shrink_to_fit(); flat.xsuperHi.clear(); flat.xsuperHi.shrink_to_fit();
    flat.entryOf.clear(); flat.entryOf.shrink_to_fit();
    ctx->info = flat;
    *out = ctx;
    return KDPT_OK;
}

int kdpt_node_box_geom(const float mins[3], const float maxs[3], int materialid, kdpt_geom* out)
{
    if (!mins || !maxs || !out) { setError("kdpt_node_box_geom: null argument"); return KDPT_ERR_ARG; }
    nodeBoxGeom(mins, maxs, materialid, out);
    return KDPT_OK;
}

int kdpt_tree_info(const kdpt_ctx* ctx, int* numNodes, int* numLeafTris, int* maxDepth, int* maxLeaf, uint64_t* deviceBytes)
{
    if (!ctx) { setError("kdpt_tree_info: null context"); return KDPT_ERR_ARG; }
    if (numNodes) *numNodes = ctx->info.numNodes;
    if (numLeafTris) *numLeafTris = ctx->info.numTris;
    if (maxDepth) *maxDepth = ctx->info.maxDepth;
    if (maxLeaf) *maxLeaf = ctx->info.maxLeaf;
    if (deviceBytes) *deviceBytes = ctx->deviceBytes;
    return KDPT_OK;
}

int kdpt_refined_info(const kdpt_ctx* ctx, int* numNodes, int* numLeafTris, int* maxDepth, int* maxLeaf)
{
    if (!ctx) { setError("kdpt_refined_info: null context"); return KDPT_ERR_ARG; }
    const bool on = ctx->xNumNodes > 0;
    if (numNodes) *numNodes = on ? ctx->xNumNodes : ctx->info.numNodes;
    if (numLeafTris) *numLeafTris = on ? ctx->info.xNumTris : ctx->info.numTris;
    if (maxDepth) *maxDepth = on ? ctx->info.xMaxDepth : ctx->info.maxDepth;
    if (maxLeaf) *maxLeaf = on ? ctx->info.xMaxLeaf : ctx->info.maxLeaf;
    return KDPT_OK;
}

int kdpt_set_stream(kdpt_ctx* ctx, void* stream)
{
    if (!ctx) { setError("kdpt_set_stream: null context"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = stream ? (cudaStream_t)stream : ctx->ownStream;
    return KDPT_OK;
}

int kdpt_sync(kdpt_ctx* ctx)
{
    if (!ctx) { setError("kdpt_sync: null context"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    return KDPT_OK;
}

int kdpt_trace_device(kdpt_ctx* ctx, const kdpt_ray* d_rays, const float* d_t_min_in, int n, kdpt_hit* d_hits, int mode, void* stream)
{
    if (!ctx || n < 0 || (n > 0 && (!d_rays || !d_hits))) { setError("kdpt_trace_device: bad argument"); return KDPT_ERR_ARG; }
    if (mode < 0 || mode >= KDPT_MODE_COUNT) { setError("kdpt_trace_device: unknown mode"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    ctx->launches++;
    return launchTrace(ctx, d_rays, d_t_min_in, n, d_hits, mode, stream ? (cudaStream_t)stream : ctx->stream);
}

int kdpt_trace(kdpt_ctx* ctx, const kdpt_ray* rays, const float* t_min_in, int n, kdpt_hit* hits, int mode)
{
    if (!ctx || n < 0 || (n > 0 && (!rays || !hits))) { setError("kdpt_trace: bad argument"); return KDPT_ERR_ARG; }
    if (mode < 0 || mode >= KDPT_MODE_COUNT) { setError("kdpt_trace: unknown mode"); return KDPT_ERR_ARG; }
    if (n == 0) return KDPT_OK;
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc;
    size_t rb = sizeof(kdpt_ray) * (size_t)n, tb = sizeof(float) * (size_t)n, hb = sizeof(kdpt_hit) * (size_t)n;
    if ((rc = ctx->rays.ensure(rb)) || (rc = ctx->hits.ensure(hb)) || (t_min_in && (rc = ctx->tmin.ensure(tb)))) return rc;
    cudaStream_t s = ctx->stream;
    // whatever goes wrong below, no copy from / into the caller's buffers may still be in flight when this call returns
    struct SyncUnlessDone {
        kdpt_ctx* c; cudaStream_t s; bool done;
        ~SyncUnlessDone() { if (!done) { if (c->copyIn) cudaStreamSynchronize(c->copyIn); if (c->copyOut) cudaStreamSynchronize(c->copyOut); cudaStreamSynchronize(s); } }
    } guard{ctx, s, false};
    kdpt_ray* dRays = ctx->rays.as<kdpt_ray>();
    kdpt_hit* dHits = ctx->hits.as<kdpt_hit>();
    float* dTmin = t_min_in ? ctx->tmin.as<float>() : nullptr;
    // Big batches go through in slices on three streams, so that the upload of slice k+1 and the download of slice k-1
    // overlap the traversal of slice k (full overlap needs page-locked host buffers; pageable ones still work, the
    // copies then simply serialise). A slice is still large enough for the leaf binning to find coherent packets.
    static const int envSlice = getenv("KDPT_TRACE_SLICE") ? atoi(getenv("KDPT_TRACE_SLICE")) : 0;     // development override
    const int sliceRays = envSlice > 0 ? envSlice : kTraceSlice;
    const int slice = n > 2 * sliceRays ? sliceRays : n;
    if (slice < n && !ctx->copyIn) {
        KDPT_CUDA(cudaStreamCreateWithFlags(&ctx->copyIn, cudaStreamNonBlocking));
        KDPT_CUDA(cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking));
        KDPT_CUDA(cudaEventCreateWithFlags(&ctx->evCopy, cudaEventDisableTiming));
    }
    if (slice < n) {        // work already queued on the context's stream comes first
        KDPT_CUDA(cudaEventRecord(ctx->evCopy, s));
        KDPT_CUDA(cudaStreamWaitEvent(ctx->copyIn, ctx->evCopy, 0));
        KDPT_CUDA(cudaStreamWaitEvent(ctx->copyOut, ctx->evCopy, 0));
    }
    for (int first = 0; first < n; first += slice) {
        const int m = n - first < slice ? n - first : slice;
        cudaStream_t in = slice < n ? ctx->copyIn : s, out = slice < n ? ctx->copyOut : s;
        KDPT_CUDA(cudaMemcpyAsync(dRays + first, rays + first, sizeof(kdpt_ray) * (size_t)m, cudaMemcpyHostToDevice, in));
        if (t_min_in) KDPT_CUDA(cudaMemcpyAsync(dTmin + first, t_min_in + first, sizeof(float) * (size_t)m, cudaMemcpyHostToDevice, in));
        if (in != s) { KDPT_CUDA(cudaEventRecord(ctx->evCopy, in)); KDPT_CUDA(cudaStreamWaitEvent(s, ctx->evCopy, 0)); }
        if ((rc = launchTrace(ctx, dRays + first, dTmin ? dTmin + first : nullptr, m, dHits + first, mode, s))) return rc;
        ctx->launches++;
        if (out != s) { KDPT_CUDA(cudaEventRecord(ctx->evCopy, s)); KDPT_CUDA(cudaStreamWaitEvent(out, ctx->evCopy, 0)); }
        KDPT_CUDA(cudaMemcpyAsync(hits + first, dHits + first, sizeof(kdpt_hit) * (size_t)m, cudaMemcpyDeviceToHost, out));
    }
    if (slice < n) KDPT_CUDA(cudaStreamSynchronize(ctx->copyOut));
    KDPT_CUDA(cudaStreamSynchronize(s));
    guard.done = true;
    return KDPT_OK;
}

} // extern "C"
