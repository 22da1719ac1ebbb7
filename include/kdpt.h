/* kdpt.h — C ABI of the B200-native KD-tree traversal / path-tracing hot path.
 *
 * Drop-in boundary for the render-loop interface of reddeupenn/kdtreePathTracerOptimization:
 *     void pathtraceInit(Scene*, bool enablekd);                       src/pathtrace.h:6   (pathtrace.cu:201-272)
 *     void pathtraceFree(Scene*, bool enablekd);                       src/pathtrace.h:7   (pathtrace.cu:274-313)
 *     void pathtrace(uchar4* pbo, int frame, int iteration, ...);      src/pathtrace.h:8-21 (pathtrace.cu:2405-2635)
 * The reference passes a C++ `Scene*`; this ABI takes exactly the flat arrays pathtraceInit reads from it
 * (pathtrace.cu:203-258) with the reference's own POD layouts, so a maintainer binds it with a 20-line shim
 * (INTEGRATION.md). Plain pointers and sizes only; no C++ or torch types.
 *
 * All functions return 0 on success and a negative kdpt_status otherwise; kdpt_last_error() gives the message of
 * the calling thread's last failure. Nothing here ever calls exit() (the reference's checkCUDAError does,
 * pathtrace.cu:38-60). There is no CPU fallback: without a usable CUDA device every entry point fails with
 * KDPT_ERR_CUDA.
 *
 * Process-wide effects a host that drives several GPUs from one thread should know: every call that takes a context makes
 * that context's device the calling thread's current device (cudaSetDevice) and leaves it so — restoring a "previous" device
 * would create a primary context on device 0 in every process that never chose one; device memory comes from the device's
 * default stream-ordered pool, whose release threshold the library lifts (freed blocks stay cached for the next
 * pathtraceInit; cudaMemPoolTrimTo on that pool gives them back).
 */
#ifndef KDPT_H
#define KDPT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KDPT_ABI_VERSION 1

typedef enum {
    KDPT_OK = 0,
    KDPT_ERR_ARG = -1,     /* bad argument / malformed tree */
    KDPT_ERR_CUDA = -2,    /* CUDA runtime error or no device */
    KDPT_ERR_ALLOC = -3,
    KDPT_ERR_STATE = -4    /* call not valid for this context (e.g. render without camera) */
} kdpt_status;

/* ---- reference POD layouts, accepted verbatim (x86-64 sizeof/offsetof as in SURVEY.md §8b) ------------------ */
typedef struct {            /* KDN::NodeBare, src/KDnode.h:64-82 — 64 B */
    int axis; float splitPos; float mins[3]; float maxs[3];
    int ID, parentID, leftID, rightID; int triIdStart, triIdSize; float tmin, tmax;
} kdpt_node_bare;

typedef struct {            /* KDN::TriBare, src/KDnode.h:51-62 — 76 B, component-major */
    float x1, x2, x3, y1, y2, y3, z1, z2, z3;
    float nx1, nx2, nx3, ny1, ny2, ny3, nz1, nz2, nz3;
    int mtlIdx;
} kdpt_tri_bare;

typedef struct {            /* Geom, src/sceneStructs.h:22-31 — 236 B; mat4 are column-major (glm) */
    int type;               /* 0 = SPHERE, 1 = CUBE (src/sceneStructs.h:10-13) */
    int materialid;
    float translation[3], rotation[3], scale[3];
    float transform[16], inverseTransform[16], invTranspose[16];
} kdpt_geom;

typedef struct {            /* Material, src/sceneStructs.h:33-44 — 56 B */
    float color[3]; float specular_exponent; float specular_color[3];
    float hasReflective, hasRefractive, indexOfRefraction, emittance; float transmittance[3];
} kdpt_material;

typedef struct {            /* Camera, src/sceneStructs.h:46-55 — 84 B */
    int resolution[2]; float position[3], lookAt[3], view[3], up[3], right[3], fov[2], pixelLength[2];
} kdpt_camera;

/* What pathtraceInit(Scene*) reads (pathtrace.cu:203-258): hst_scene->geoms, materials, newNodesBare/numNodes,
 * newTrianglesBare/numTriangles, obj_materialOffsets/obj_numshapes, state.camera, state.traceDepth. */
typedef struct {
    const kdpt_node_bare* nodes; int numNodes;
    const kdpt_tri_bare* tris; int numTris;
    const kdpt_geom* geoms; int numGeoms;
    const kdpt_material* materials; int numMaterials;
    const int* objMaterialOffsets; int numShapes;
    kdpt_camera camera;
    int traceDepth;
} kdpt_scene;

/* ---- traversal-only entry point (what the reference's pathTraceOneBounceKDbare does per ray, minus scatter) -- */
typedef struct { float origin[3]; float direction[3]; } kdpt_ray;          /* 24 B */

typedef struct {            /* 32 B */
    float t;                /* the reference's t_min: |o - (o + d*z + n*eps)|, or the incoming t_min when no hit */
    float z, u, v;          /* glm::intersectRayTriangle baryPosition .z / .x / .y of the accepted triangle */
    int tri;                /* canonical TriBare index (first entry with identical 76 bytes) or -1 */
    float normal[3];        /* normalised interpolated normal of the accepted triangle */
} kdpt_hit;

typedef enum {
    /* exact closest hit (== brute force over TriBare[] with the reference's arithmetic and acceptance rule) */
    KDPT_FULLSTACK = 0,     /* near-first traversal, full per-ray stack; nudge 1e-5 (traverseKDbare's) */
    KDPT_SHORTSTACK = 1,    /* short register stack + restart;           nudge 1e-4 (traverseKDbareShortHybrid's) */
    /* bit-compatible with the reference's own (lossy) functions, including the hits they drop */
    KDPT_REF_T1 = 2,        /* traverseKDbare,            pathtrace.cu:881-1020 */
    KDPT_REF_T2 = 3,        /* traverseKDbareShortHybrid, pathtrace.cu:1023-1235 */
    /* the reference's KD-less modes (ENABLEKD off, key K; USEBBOX, key B): every triangle against every ray, nudge 1e-4,
     * pathTraceOneBounce (pathtrace.cu:402-628), with the mesh bounding-box reject of :497-511 in the second form. Same
     * results as KDPT_SHORTSTACK; they exist for the reference's brute force / bounding box / kd-tree comparison */
    KDPT_BRUTE = 4,
    KDPT_BRUTE_BBOX = 5,
    /* the reference's view of the tree itself (VIZKD, key V; pathTraceOneBounceKDbareBoxes, pathtrace.cu:1738-1885): no
     * triangles — every node box is drawn as a cube (boxIntersectionTestBox, intersections.h:51-104) of the LAST material,
     * after the scene's own geoms. Render loop only (kdpt_render / kdpt_bounce); kdpt_trace refuses it */
    KDPT_KDVIZ = 6,
    KDPT_MODE_COUNT = 7
} kdpt_mode;

/* ---- render loop toggles: the trailing arguments of pathtrace(), src/pathtrace.h:8-21, defaults main.cpp:35-60 */
typedef struct {
    float focalLength;      /* dofDistance, default 6 */
    float dofAngle;         /* default 0 */
    float softness;         /* default 0 */
    int antialias;          /* default 1 */
    int enableSss;          /* default 0 */
    int compaction;         /* default 1 */
    int mode;               /* kdpt_mode; the reference's SHORTSTACK=true is KDPT_REF_T2, false KDPT_REF_T1 */
    int sortByMaterial;     /* 0 never, 1 only in iteration 2 (reference, pathtrace.cu:2600), 2 every bounce */
    int testingmode;        /* TESTINGMODE (pathtrace.cu:2478-2481, 2565-2574): time every intersect kernel with CUDA
                               events and report the sum in kdpt_render_stats.intersectMs */
    int laneQueue;          /* tuning, no effect on results: how a bounce traverses the lanes of a 32-path packet that enter
                               the KD tree. 0 = the kernel decides per launch from the measured entering fraction
                               (default), 1 = always at once, 2 = always pooled in the per-warp queue */
    int cacheRays;          /* rayCaching (pathtrace.cu:2448-2456, default 0): every iteration starts from iteration 1's camera
                               rays (no per-iteration antialiasing / depth-of-field jitter); here that is a choice of jitter
                               constants, the rays are generated inside the first bounce either way */
    int reserved[1];
} kdpt_params;

typedef struct {            /* filled by kdpt_render */
    uint64_t rayBounces;    /* paths traced summed over bounces and iterations */
    uint64_t launches;      /* kernels launched */
    float intersectMs;      /* device time of the intersect kernels (the reference's TESTINGMODE number) */
    float totalMs;          /* device time of the whole call */
    uint64_t intersectLaunches; /* number of intersect (bounce) kernel launches behind intersectMs */
} kdpt_render_stats;

/* the reference's per-path records, for the stage-level entry points below */
typedef struct {            /* PathSegment (with its Ray), src/sceneStructs.h:15-20, 65-71 — 56 B */
    float origin[3], direction[3]; unsigned char isinside, _pad[3]; float sdepth;
    float color[3]; int pixelIndex, remainingBounces, materialIdHit;
} kdpt_path_segment;
typedef struct {            /* ShadeableIntersection, src/sceneStructs.h:81-85 — 20 B */
    float t; float surfaceNormal[3]; int materialId;
} kdpt_intersection;

typedef enum {
    KDPT_STAGE_INTERSECT = 1,   /* pathTraceOneBounceKDbare: geoms + KD traversal + scatterRay, pathtrace.cu:1571-1734 */
    KDPT_STAGE_SHADE = 2        /* shadeMaterial, pathtrace.cu:2304-2369 */
} kdpt_stage;

typedef struct kdpt_ctx kdpt_ctx;

int kdpt_abi_version(void);
const char* kdpt_last_error(void);

/* Host helper: what the reference's main.cpp does to the loaded camera before the first pathtrace() call
 * (orbit angles from the view vector, main.cpp:1053-1073, then runCuda's recompute of position/view/up/right,
 * main.cpp:1111-1129). Pure host arithmetic; call it on the camera the scene loader produced. */
void kdpt_camera_first_frame(kdpt_camera* camera);

/* ---- host side: the reference's scene / OBJ loading and KD build (src/scene.cpp, KDtree.cpp, KDnode.cpp, objmesh.cpp +
 * tinyobjloader v1.0.0) — pure host code, produces exactly the arrays above, byte-identical to the reference loader's. */
typedef struct kdpt_host_scene kdpt_host_scene;
/* Scene(txt) + Scene::loadObj(obj, mtl_dir) (main.cpp:1031-1033). obj_path may be NULL (no mesh); mtl_dir NULL = the
 * OBJ's directory (main.cpp:1033); kd_depth <= 0 = the reference's literal 13 (scene.cpp:872). synth_normals != 0
 * generates area-weighted vertex normals for OBJs without one `vn` per `v` (the reference reads out of bounds there). */
int kdpt_scene_load(const char* scene_txt, const char* obj_path, const char* mtl_dir, int kd_depth, int synth_normals,
                    kdpt_host_scene** out);
/* KDtree(triangles) -> updateBbox -> split(kd_depth) -> flatten, for a caller-supplied triangle list:
 * tri18 = x1 y1 z1 x2 y2 z2 x3 y3 z3 nx1 ny1 nz1 nx2 ny2 nz2 nx3 ny3 nz3 per triangle, mtl = shape index per triangle or NULL */
int kdpt_scene_from_triangles(const float* tri18, const int* mtl, int n, int kd_depth, kdpt_host_scene** out);
/* The same build on the GPU (KDnode::split is the start-up bottleneck of the 1 M / 10 M-triangle
 * configs): level-synchronous segmented partition on `device`, node numbering on the host. Output arrays are
 * byte-identical to kdpt_scene_from_triangles's (and so to the reference builder's). device_ms (may be NULL) receives the
 * device time between the end of the triangle upload and the last kernel. No CPU fallback: KDPT_ERR_CUDA without a device. */
int kdpt_scene_from_triangles_gpu(const float* tri18, const int* mtl, int n, int kd_depth, int device, float* device_ms,
                                  kdpt_host_scene** out);
/* Binary scene cache: every array of a host scene in one checksummed file (same machine ABI), so
 * that a 1 M / 10 M-triangle config starts without re-parsing the OBJ and re-building the tree. A truncated, corrupt or
 * foreign file is refused with KDPT_ERR_ARG, never half-loaded. */
int kdpt_scene_save(const kdpt_host_scene* scene, const char* path);
int kdpt_scene_load_cache(const char* path, kdpt_host_scene** out);
/* kdpt_scene_load through a content-addressed cache: the file name under cache_dir is a hash of the bytes of the scene
 * file, the OBJ, the material libraries it names, and the build options; *cache_hit (may be NULL) says which way it went. */
int kdpt_scene_load_cached(const char* scene_txt, const char* obj_path, const char* mtl_dir, int kd_depth, int synth_normals,
                           const char* cache_dir, kdpt_host_scene** out, int* cache_hit);
/* borrow the arrays (valid until kdpt_scene_free) */
int kdpt_scene_view(const kdpt_host_scene* scene, kdpt_scene* view, int* iterations);
/* source triangle (OBJ face order) of every TriBare entry */
const int* kdpt_scene_triangle_sources(const kdpt_host_scene* scene);
void kdpt_scene_free(kdpt_host_scene* scene);

/* pathtraceInit: flatten + upload. device = CUDA ordinal. The host arrays are not retained. */
int kdpt_create(const kdpt_scene* scene, int device, kdpt_ctx** out);
/* The same with build options. The exact modes (KDPT_FULLSTACK / KDPT_SHORTSTACK) do not have to walk the reference's
 * leaves in the reference's order — they return the brute-force closest hit — so for them pathtraceInit splits every
 * reference leaf of more than refineLeafMax triangles further, inside its own cell (the reference builder stops at depth 14
 * or when a split keeps all triangles on one side: leaves of hundreds to tens of thousands of triangles). Results do not
 * depend on these options; the reference-compatible modes always walk the tree as given. NULL = defaults. */
typedef struct {
    int refineLeafMax;          /* < 0: default (8); 0: no refinement (exact modes walk the reference tree, fat leaves are
                                   tested warp-cooperatively through chunk boxes); > 0: leaf size at which splitting stops */
    float refineTraversalCost;  /* <= 0: default (2.0); cost of one node step relative to one triangle test in the builder's
                                   surface-area heuristic */
    int nodeForm;               /* how the exact modes' tree is laid out for their kernel. 1: one 16-byte record per node; 2: two
                                   levels per 32-byte record, small leaves named inline by their parents (half the dependent
                                   loads of a descent; faster once the tree exceeds the L2, a little slower below); 0: chosen by size
                                   (two-level from 256 MB of device arrays on) */
    int reserved[5];
} kdpt_build_opts;
int kdpt_create_opts(const kdpt_scene* scene, int device, const kdpt_build_opts* opts, kdpt_ctx** out);
/* pathtraceFree */
int kdpt_destroy(kdpt_ctx* ctx);

/* sizes of the flattened device representation (bytes), for roofline bookkeeping */
int kdpt_tree_info(const kdpt_ctx* ctx, int* numNodes, int* numLeafTris, int* maxDepth, int* maxLeaf,
                   uint64_t* deviceBytes);

/* the tree the exact modes walk (== kdpt_tree_info's numbers when refinement is off) */
int kdpt_refined_info(const kdpt_ctx* ctx, int* numNodes, int* numLeafTris, int* maxDepth, int* maxLeaf);

/* KDPT_KDVIZ draws node i as this geom: the cube boxIntersectionTestBox (src/intersections.h:51-66) makes of nodes[i].mins /
 * .maxs — scale = extent, translation = centre, transform / inverse as glm builds them — with the given material. Host-only
 * helper (no GPU needed): a caller that wants the boxes of some levels only can append such geoms to its own scene. */
int kdpt_node_box_geom(const float mins[3], const float maxs[3], int materialid, kdpt_geom* out);

/* n rays in, n hits out, HOST buffers (H2D + kernel + D2H). t_min_in may be NULL (FLT_MAX). */
int kdpt_trace(kdpt_ctx* ctx, const kdpt_ray* rays, const float* t_min_in, int n, kdpt_hit* hits, int mode);
/* same with DEVICE buffers on ctx's device; asynchronous on `stream` (a cudaStream_t, NULL = ctx's own stream,
 * in which case the call returns after the kernel has been enqueued; use kdpt_sync). */
int kdpt_trace_device(kdpt_ctx* ctx, const kdpt_ray* d_rays, const float* d_t_min_in, int n, kdpt_hit* d_hits,
                      int mode, void* stream);
int kdpt_sync(kdpt_ctx* ctx);

/* pathtrace(): iterations iter_begin .. iter_begin+iter_count-1 (the reference counts from 1), accumulating into
 * the context's radiance buffer. iter_stride > 1 renders iter_begin, iter_begin+stride, ... (sample sharding). */
int kdpt_render(kdpt_ctx* ctx, int iter_begin, int iter_count, int iter_stride, const kdpt_params* params,
                kdpt_render_stats* stats /* NULL: enqueue only, do not synchronise */);
/* stage-level operators on caller-owned HOST arrays in the reference's own layouts (parity tests, integration):
 * generateRayFromCamera (pathtrace.cu:315-397): W*H segments out */
int kdpt_generate_rays(kdpt_ctx* ctx, int iter, const kdpt_params* params, kdpt_path_segment* out);
/* one bounce over n segments (in/out), `stages` = OR of kdpt_stage; isects (n, may be NULL) receives what the
 * intersect stage wrote. path_index == array index, as in the reference. */
int kdpt_bounce(kdpt_ctx* ctx, int iter, int depth, const kdpt_params* params, kdpt_path_segment* paths, int n,
                kdpt_intersection* isects, int stages);
/* the loop's other two stages on a caller-owned HOST array, through the same kernels the render loop launches:
 * thrust::remove_if(paths, is_zero_bounce) (pathtrace.cu:2591-2597, predicate 103-110): stable, in place, *n_out survivors;
 * thrust::sort(paths) on materialIdHit (pathtrace.cu:2600-2606, operator< 123-131): stable, in place, keys >= -1 */
int kdpt_compact_paths(kdpt_ctx* ctx, kdpt_path_segment* paths, int n, int* n_out);
int kdpt_sort_paths(kdpt_ctx* ctx, kdpt_path_segment* paths, int n);
int kdpt_clear_image(kdpt_ctx* ctx);
/* A camera move without rebuilding anything. The reference answers every camera change with pathtraceFree +
 * pathtraceInit (main.cpp:1134-1137): a re-upload of the whole tree, and here a re-flatten as well (1 s for 10 M
 * triangles). This call replaces only the camera (resolution included) and clears the radiance buffer, as pathtraceInit
 * does (pathtrace.cu:262-263); the binding in INTEGRATION.md uses it whenever pathtraceInit is called again with the
 * arrays it already holds. */
int kdpt_set_camera(kdpt_ctx* ctx, const kdpt_camera* camera);
/* Run all of this context's work on the caller's stream (a cudaStream_t; NULL restores the context's own stream). */
int kdpt_set_stream(kdpt_ctx* ctx, void* stream);
/* Accumulate radiance into a caller-owned DEVICE buffer of W*H*3 floats (e.g. the tensor an NCCL all-reduce runs
 * on); NULL goes back to the context's own buffer. The buffer is not cleared. */
int kdpt_bind_image(kdpt_ctx* ctx, float* d_rgb);
/* cumulative since kdpt_create: ray-bounces traced by the render loop and kernels launched (synchronises) */
int kdpt_counters(kdpt_ctx* ctx, uint64_t* rayBounces, uint64_t* launches);
/* accumulated radiance, W*H*3 floats (scene->state.image, pathtrace.cu:2631) */
int kdpt_get_image(kdpt_ctx* ctx, float* rgb);
/* restore an accumulation buffer saved with kdpt_get_image (checkpoint / resume of a long render: the reference has
 * none — its buffer cannot be reloaded); continue with iter_begin = last rendered iteration + 1 */
int kdpt_set_image(kdpt_ctx* ctx, const float* rgb);
/* device pointer of that buffer (for an NCCL all-reduce issued by the caller) */
int kdpt_image_device(kdpt_ctx* ctx, float** d_rgb, int* numFloats);
/* Sum the radiance buffers of n contexts of ONE process (one per GPU, same resolution) and leave the total in every
 * one of them — the step that follows sample-sharded rendering. Done over NVLink peer memory by the
 * library's own kernels: device i sums slice i of all buffers in context order (so every GPU ends with the same bits),
 * then the slices are gathered. Needs peer access between the devices (KDPT_ERR_STATE otherwise). With one process
 * per GPU use an NCCL all-reduce on kdpt_image_device / kdpt_bind_image instead. Synchronises all contexts. */
int kdpt_allreduce_image(kdpt_ctx** ctxs, int n);

#ifdef __cplusplus
}
#endif
#endif
