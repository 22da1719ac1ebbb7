// pathtrace_kdpt.cpp — the drop-in for the reference's src/pathtrace.cu that INTEGRATION.md describes: the three functions
// of src/pathtrace.h (6-21) implemented on libkdpt.so's C ABI. This is REFERENCE-SIDE binding code (what a maintainer adds
// to the reference), kept here so that it is compiled against the reference's own headers and run (oracle/Makefile target
// `dropin`, tests/test_dropin_gpu.py). It is not part of the product library.
#include <vector_types.h>   // uchar4 (the reference gets it through its CUDA includes)
#include "pathtrace.h"
#include "scene.h"
#include <kdpt.h>
#include <cstdio>
#include <cstring>
#include <vector>

static kdpt_ctx* g_ctx = nullptr;
static Scene* g_scene = nullptr;
// what g_ctx was built from: the tree arrays by address and size (Scene::loadObj allocates them once and never writes them
// again), everything small by value
static const void* g_nodes = nullptr; static const void* g_tris = nullptr;
static int g_numNodes = -1, g_numTris = -1, g_traceDepth = -1;
static std::vector<char> g_small;

void pathtraceInit(Scene* scene, bool /*enablekd*/) {                       // pathtrace.cu:201-272
    g_scene = scene;
    kdpt_scene s{};
    s.nodes = (const kdpt_node_bare*)scene->newNodesBare;     s.numNodes = scene->hasObj ? scene->numNodes : 0;
    s.tris  = (const kdpt_tri_bare*)scene->newTrianglesBare;  s.numTris  = scene->hasObj ? scene->numTriangles : 0;
    s.geoms = (const kdpt_geom*)scene->geoms.data();          s.numGeoms = (int)scene->geoms.size();
    s.materials = (const kdpt_material*)scene->materials.data(); s.numMaterials = (int)scene->materials.size();
    s.objMaterialOffsets = scene->obj_materialOffsets;        s.numShapes = scene->obj_numshapes;
    std::memcpy(&s.camera, &scene->state.camera, sizeof(kdpt_camera));       // already recomputed by runCuda()
    s.traceDepth = scene->state.traceDepth;
    // runCuda() frees and re-inits on every camera move (main.cpp:1134-1137) with the arrays it already gave us:
    // keep the flattened tree on the GPU and only swap the camera (the radiance buffer is cleared either way)
    std::vector<char> small;
    auto add = [&](const void* p, size_t n) { small.insert(small.end(), (const char*)p, (const char*)p + n); };
    add(s.geoms, sizeof(kdpt_geom) * s.numGeoms); add(s.materials, sizeof(kdpt_material) * s.numMaterials);
    add(s.objMaterialOffsets, sizeof(int) * s.numShapes);
    if (g_ctx && s.nodes == g_nodes && s.tris == g_tris && s.numNodes == g_numNodes && s.numTris == g_numTris &&
        s.traceDepth == g_traceDepth && small == g_small) {
        if (kdpt_set_camera(g_ctx, &s.camera) != KDPT_OK) std::fprintf(stderr, "kdpt: %s\n", kdpt_last_error());
        return;
    }
    kdpt_destroy(g_ctx); g_ctx = nullptr;
    if (kdpt_create(&s, /*device*/0, &g_ctx) != KDPT_OK) { std::fprintf(stderr, "kdpt: %s\n", kdpt_last_error()); g_ctx = nullptr; }
    g_nodes = s.nodes; g_tris = s.tris; g_numNodes = s.numNodes; g_numTris = s.numTris; g_traceDepth = s.traceDepth;
    g_small.swap(small);
}

void pathtraceFree(Scene*, bool) {}      // pathtrace.cu:274-313; the context lives until the next scene (see pathtraceInit)

void pathtrace(uchar4* /*pbo*/, int /*frame*/, int iter, float focalLength, float dofAngle, bool cacherays,
               bool antialias, float softness, bool enableSss, bool testingmode, bool compaction, bool enablekd,
               bool vizkd, bool usebbox, bool shortstack) {              // pathtrace.cu:2405-2635
    if (!g_ctx) return;                                    // pathtraceInit failed and said so
    kdpt_params p{};
    p.focalLength = focalLength; p.dofAngle = dofAngle; p.softness = softness;
    p.antialias = antialias; p.enableSss = enableSss; p.compaction = compaction; p.testingmode = testingmode;
    p.cacheRays = cacherays;
    p.mode = !enablekd ? (usebbox ? KDPT_BRUTE_BBOX : KDPT_BRUTE)            // key K / key B: pathTraceOneBounce
           : vizkd ? KDPT_KDVIZ                             // key V: pathTraceOneBounceKDbareBoxes
           : shortstack ? KDPT_REF_T2 : KDPT_REF_T1;       // bit-compatible with traverseKDbare(ShortHybrid);
                                                           // KDPT_SHORTSTACK / KDPT_FULLSTACK give the exact closest hit
    p.sortByMaterial = 1;                                  // the reference sorts in iteration 2 only
    kdpt_render_stats st{};
    if (kdpt_render(g_ctx, iter, 1, 1, &p, &st) != KDPT_OK) { std::fprintf(stderr, "kdpt: %s\n", kdpt_last_error()); return; }
    if (testingmode) std::printf("   pathTrace time = %f\n", st.intersectMs);
    kdpt_get_image(g_ctx, (float*)g_scene->state.image.data());             // pathtrace.cu:2631-2632
}
