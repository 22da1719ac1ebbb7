// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Thin extern "C" shim around the UNMODIFIED reference sources, compiled where
// they lie under /root/reference (see oracle/Makefile, target _ref/libkdref.so).
// Nothing from the reference is copied into this repository: this file only
// #includes the reference translation unit and forwards calls to its functions.
//
// What it exposes (all reference code, executed on the host):
//   * Scene(txt) + Scene::loadObj(obj)            scene.cpp:7-61, 579-901
//   * the flattened NodeBare[] / TriBare[] arrays  scene.cpp:905-968
//   * traverseKDbare / traverseKDbareShortHybrid   pathtrace.cu:881-1235
//   * boxIntersectionTest / sphereIntersectionTest intersections.h:107-203
//   * boxIntersectionTestBox (the KD-viz mode's node-box test)   intersections.h:51-104
//   * scatterRay + makeSeededRandomEngine          interactions.h:195-358, pathtrace.cu:62-66
//   * KDtree(path) -> updateBbox -> split -> writeKDtoFile   KDtree.cpp:37-45,115-135
//   * pathtraceInit / pathtrace / pathtraceFree on a GPU (only usable on a GPU box)
#include <thrust/sort.h>   // the reference forgets this include
#include <cfloat>
#include <cstring>
#include <cstdio>
#include <thread>
#include <chrono>
#include <algorithm>

#include "pathtrace.cu"    // resolved through -I/root/reference/src

static Scene* g_scene = nullptr;

extern "C" {

// Load a scene .txt and (optionally) an OBJ through the reference's own loader.
int kdref_load(const char* scene_txt, const char* obj_path, const char* mtl_dir)
{
    g_scene = new Scene(std::string(scene_txt));
    if (obj_path && obj_path[0])
        g_scene->loadObj(std::string(obj_path), std::string(mtl_dir));
    return 0;
}

int kdref_sizes(int* out /* numNodes,numTris,numGeoms,numMats,numShapes,traceDepth,iterations,resx,resy */)
{
    if (!g_scene) return -1;
    out[0] = g_scene->hasObj ? g_scene->numNodes : 0;
    out[1] = g_scene->hasObj ? g_scene->numTriangles : 0;
    out[2] = (int)g_scene->geoms.size();
    out[3] = (int)g_scene->materials.size();
    out[4] = g_scene->obj_numshapes;
    out[5] = g_scene->state.traceDepth;
    out[6] = (int)g_scene->state.iterations;
    out[7] = g_scene->state.camera.resolution.x;
    out[8] = g_scene->state.camera.resolution.y;
    return 0;
}

int kdref_struct_sizes(int* out)
{
    out[0] = sizeof(KDN::NodeBare); out[1] = sizeof(KDN::TriBare); out[2] = sizeof(Geom);
    out[3] = sizeof(Material); out[4] = sizeof(Camera); out[5] = sizeof(PathSegment);
    out[6] = sizeof(ShadeableIntersection); out[7] = sizeof(Ray);
    return 0;
}

int kdref_copy(void* nodes, void* tris, void* geoms, void* mats, int* offsets, void* cam)
{
    if (!g_scene) return -1;
    if (g_scene->hasObj) {
        if (nodes) memcpy(nodes, g_scene->newNodesBare, sizeof(KDN::NodeBare) * g_scene->numNodes);
        if (tris) memcpy(tris, g_scene->newTrianglesBare, sizeof(KDN::TriBare) * g_scene->numTriangles);
        if (offsets) memcpy(offsets, g_scene->obj_materialOffsets, sizeof(int) * g_scene->obj_numshapes);
    }
    if (geoms) memcpy(geoms, g_scene->geoms.data(), sizeof(Geom) * g_scene->geoms.size());
    if (mats) memcpy(mats, g_scene->materials.data(), sizeof(Material) * g_scene->materials.size());
    if (cam) memcpy(cam, &g_scene->state.camera, sizeof(Camera));
    return 0;
}

// Run the reference's host traversal on caller-supplied flattened arrays.
//   mode 0 = traverseKDbare, 1 = traverseKDbareShortHybrid
//   rays: n x 6 floats (origin, direction);  t_min_in: per-ray incoming t_min or NULL (FLT_MAX)
// Outputs per ray: hit flag, t_min, normal[3], bary[3] (last glm write), objMaterialIdx, hit_geom_index
int kdref_traverse(int mode, const void* nodes_, int numNodes_, const void* tris_,
                   int material_size, const int* offsets,
                   const float* rays, int n, const float* t_min_in,
                   unsigned char* hit, float* t_min_out, float* normal_out, float* bary_out,
                   int* objmat_out, int* geomidx_out)
{
    KDN::NodeBare* nodes = (KDN::NodeBare*)nodes_;
    KDN::TriBare* tris = (KDN::TriBare*)tris_;
    ShadeableIntersection isect;
    for (int i = 0; i < n; i++) {
        PathSegment ps;
        ps.ray.origin = glm::vec3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]);
        ps.ray.direction = glm::vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        ps.ray.isinside = false; ps.ray.sdepth = 0.f;
        ps.color = glm::vec3(1.f); ps.pixelIndex = i; ps.remainingBounces = 1; ps.materialIdHit = -1;
        float t = 0.f, t_min = t_min_in ? t_min_in[i] : FLT_MAX;
        glm::vec3 bary(0.f), hitp(0.f), norm(0.f), ipt(0.f), normal(0.f), tmpi(0.f), tmpn(0.f);
        int objMaterialIdx = -1, hit_geom_index = -1, path_index = 0;
        bool obj_intersect = false;
        isect.t = 0.f;
        if (mode == 0)
            traverseKDbare(nodes, numNodes_, t, ps, tris, bary, objMaterialIdx, material_size, hitp, norm, t_min,
                           hit_geom_index, ipt, normal, tmpi, tmpn, obj_intersect, &isect, (int*)offsets, path_index);
        else
            traverseKDbareShortHybrid(nodes, numNodes_, t, ps, tris, bary, objMaterialIdx, material_size, hitp, norm,
                                      t_min, hit_geom_index, ipt, normal, tmpi, tmpn, obj_intersect, &isect,
                                      (int*)offsets, path_index);
        hit[i] = obj_intersect ? 1 : 0;
        t_min_out[i] = t_min;
        if (normal_out) { normal_out[3 * i] = normal.x; normal_out[3 * i + 1] = normal.y; normal_out[3 * i + 2] = normal.z; }
        if (bary_out) { bary_out[3 * i] = bary.x; bary_out[3 * i + 1] = bary.y; bary_out[3 * i + 2] = bary.z; }
        if (objmat_out) objmat_out[i] = objMaterialIdx;
        if (geomidx_out) geomidx_out[i] = hit_geom_index;
    }
    return 0;
}

// glm::intersectRayTriangle on n (ray, triangle) pairs; tri = 9 floats v0 v1 v2.
int kdref_raytri(const float* rays, const float* tris, int n, unsigned char* hit, float* bary_out)
{
    for (int i = 0; i < n; i++) {
        glm::vec3 o(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        glm::vec3 a(tris[9 * i], tris[9 * i + 1], tris[9 * i + 2]), b(tris[9 * i + 3], tris[9 * i + 4], tris[9 * i + 5]),
            c(tris[9 * i + 6], tris[9 * i + 7], tris[9 * i + 8]);
        glm::vec3 bary(-7.f, -7.f, -7.f);
        hit[i] = glm::intersectRayTriangle(o, d, a, b, c, bary) ? 1 : 0;
        bary_out[3 * i] = bary.x; bary_out[3 * i + 1] = bary.y; bary_out[3 * i + 2] = bary.z;
    }
    return 0;
}

// intersectAABBarrays on n (ray, box) pairs; box = mins[3], maxs[3].
int kdref_aabb(const float* rays, const float* boxes, int n, unsigned char* hit, float* dist_out)
{
    for (int i = 0; i < n; i++) {
        Ray r; r.origin = glm::vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        r.direction = glm::vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        float mins[3] = {boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]};
        float maxs[3] = {boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]};
        float dist = -1.f;
        hit[i] = intersectAABBarrays(r, mins, maxs, dist) ? 1 : 0;
        dist_out[i] = dist;
    }
    return 0;
}

// Analytic geoms of the loaded scene: returns t, point, normal, outside for n rays against geom g.
int kdref_geom_test(const void* geom_, const float* rays, int n, float* t_out, float* pt_out, float* n_out,
                    unsigned char* outside_out)
{
    const Geom& g = *(const Geom*)geom_;
    for (int i = 0; i < n; i++) {
        Ray r; r.origin = glm::vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        r.direction = glm::vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        r.isinside = false; r.sdepth = 0.f;
        glm::vec3 p(0.f), nn(0.f); bool outside = true; float t = -1.f;
        if (g.type == CUBE) t = boxIntersectionTest(g, r, p, nn, outside);
        else if (g.type == SPHERE) t = sphereIntersectionTest(g, r, p, nn, outside);
        t_out[i] = t;
        pt_out[3 * i] = p.x; pt_out[3 * i + 1] = p.y; pt_out[3 * i + 2] = p.z;
        n_out[3 * i] = nn.x; n_out[3 * i + 1] = nn.y; n_out[3 * i + 2] = nn.z;
        outside_out[i] = outside ? 1 : 0;
    }
    return 0;
}

// boxIntersectionTestBox (what pathTraceOneBounceKDbareBoxes, pathtrace.cu:1738-1885, runs on every node box): box i =
// mins/maxs[3i..3i+2] against ray i.
int kdref_box_test_box(const float* mins, const float* maxs, const float* rays, int n, float* t_out, float* pt_out, float* n_out,
                       unsigned char* outside_out)
{
    for (int i = 0; i < n; i++) {
        Ray r; r.origin = glm::vec3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        r.direction = glm::vec3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        r.isinside = false; r.sdepth = 0.f;
        float mn[3] = {mins[3 * i], mins[3 * i + 1], mins[3 * i + 2]}, mx[3] = {maxs[3 * i], maxs[3 * i + 1], maxs[3 * i + 2]};
        glm::vec3 p(0.f), nn(0.f); bool outside = true;
        t_out[i] = boxIntersectionTestBox(mn, mx, r, p, nn, outside);
        pt_out[3 * i] = p.x; pt_out[3 * i + 1] = p.y; pt_out[3 * i + 2] = p.z;
        n_out[3 * i] = nn.x; n_out[3 * i + 1] = nn.y; n_out[3 * i + 2] = nn.z;
        outside_out[i] = outside ? 1 : 0;
    }
    return 0;
}

// makeSeededRandomEngine(iter,index,depth) then k draws of uniform_real_distribution<float>(0,1).
int kdref_rng(int iter, int index, int depth, int k, float* out, unsigned int* seed_out)
{
    thrust::default_random_engine rng = makeSeededRandomEngine(iter, index, depth);
    if (seed_out) *seed_out = (unsigned int)(utilhash((1 << 31) | (depth << 22) | iter) ^ utilhash(index));
    thrust::uniform_real_distribution<float> u01(0, 1);
    for (int i = 0; i < k; i++) out[i] = u01(rng);
    return 0;
}

// scatterRay on n records. in/out ray: origin[3] dir[3] isinside sdepth (8 floats), color[3],
// intersect[3], normal[3]; material = reference Material (56 B); rng seeded (iter, index0+i, depth).
int kdref_scatter(float* ray8, float* color3, const float* isect3, const float* normal3, const void* mat,
                  int iter, int index0, int depth, float softness, int n)
{
    const Material& m = *(const Material*)mat;
    for (int i = 0; i < n; i++) {
        Ray r; r.origin = glm::vec3(ray8[8 * i], ray8[8 * i + 1], ray8[8 * i + 2]);
        r.direction = glm::vec3(ray8[8 * i + 3], ray8[8 * i + 4], ray8[8 * i + 5]);
        r.isinside = ray8[8 * i + 6] != 0.f; r.sdepth = ray8[8 * i + 7];
        glm::vec3 c(color3[3 * i], color3[3 * i + 1], color3[3 * i + 2]);
        glm::vec3 ip(isect3[3 * i], isect3[3 * i + 1], isect3[3 * i + 2]);
        glm::vec3 nn(normal3[3 * i], normal3[3 * i + 1], normal3[3 * i + 2]);
        thrust::default_random_engine rng = makeSeededRandomEngine(iter, index0 + i, depth);
        scatterRay(r, c, ip, nn, m, rng, softness);
        ray8[8 * i] = r.origin.x; ray8[8 * i + 1] = r.origin.y; ray8[8 * i + 2] = r.origin.z;
        ray8[8 * i + 3] = r.direction.x; ray8[8 * i + 4] = r.direction.y; ray8[8 * i + 5] = r.direction.z;
        ray8[8 * i + 6] = r.isinside ? 1.f : 0.f; ray8[8 * i + 7] = r.sdepth;
        color3[3 * i] = c.x; color3[3 * i + 1] = c.y; color3[3 * i + 2] = c.z;
    }
    return 0;
}

// The reference's one golden fixture: triangle soup text -> tree -> bbox dump (KDtree.cpp:37-45, 115-135).
int kdref_soup_tree_dump(const char* soup_path, int maxdepth, const char* out_path)
{
    KDtree* T = new KDtree(soup_path);
    T->rootNode->updateBbox();
    T->split(maxdepth);
    T->writeKDtoFile(T->rootNode, out_path);
    return 0;
}

// ---- CPU baseline: the reference's own host functions driven by a host-side harness ---------------------------
// One iteration of the render loop on the host cores. The per-path work is the reference's code, called unmodified:
// boxIntersectionTest / sphereIntersectionTest (intersections.h), traverseKDbare / traverseKDbareShortHybrid
// (pathtrace.cu:881-1235), makeSeededRandomEngine + scatterRay (interactions.h). The loop around them follows
// pathTraceOneBounceKDbare (1571-1734), shadeMaterial (2304-2369), partialGather, remove_if and the bounce cap of
// pathtrace() (2405-2635); those are __global__ kernels / thrust calls in the reference and cannot run on a host, so
// the harness states their few lines of glue here. paths: W*H PathSegment (generated by the caller), consumed.
// Returns the number of ray-bounces traced; seconds[0] = intersect stage, seconds[1] = everything else.
static void kdref_intersect_range(int lo, int hi, int depth, int iter, PathSegment* paths, Geom* geoms, int geoms_size,
                                  Material* materials, int material_size, ShadeableIntersection* intersections,
                                  float softness, KDN::TriBare* triangles, KDN::NodeBare* nodes, int numNodes,
                                  int* offsets, bool shortstack)
{
    for (int path_index = lo; path_index < hi; path_index++) {
        PathSegment pathSegment = paths[path_index];
        if (paths[path_index].remainingBounces <= 0) continue;
        float t = 0.f, t_min = FLT_MAX;
        glm::vec3 intersect_point, normal, tmp_intersect, tmp_normal, hit, norm, bary;
        int hit_geom_index = -1, objMaterialIdx = -1;
        bool outside = true, obj_intersect = false;
        for (int i = 0; i < geoms_size; i++) {
            Geom& geom = geoms[i];
            if (geom.type == CUBE) t = boxIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, outside);
            else if (geom.type == SPHERE) t = sphereIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, outside);
            if (t > 0.0f && t_min > t) { t_min = t; hit_geom_index = i; intersect_point = tmp_intersect; normal = tmp_normal; }
        }
        if (numNodes > 0) {
            if (!shortstack)
                traverseKDbare(nodes, numNodes, t, pathSegment, triangles, bary, objMaterialIdx, material_size, hit, norm,
                               t_min, hit_geom_index, intersect_point, normal, tmp_intersect, tmp_normal, obj_intersect,
                               intersections, offsets, path_index);
            else
                traverseKDbareShortHybrid(nodes, numNodes, t, pathSegment, triangles, bary, objMaterialIdx, material_size,
                                          hit, norm, t_min, hit_geom_index, intersect_point, normal, tmp_intersect,
                                          tmp_normal, obj_intersect, intersections, offsets, path_index);
        }
        if (hit_geom_index == -1) { intersections[path_index].t = -1.0f; continue; }
        thrust::default_random_engine rng = makeSeededRandomEngine(iter, path_index, depth);
        int mat = obj_intersect ? objMaterialIdx : geoms[hit_geom_index].materialid;
        paths[path_index].materialIdHit = mat;
        scatterRay(paths[path_index].ray, paths[path_index].color, intersect_point, normal, materials[mat], rng, softness);
        intersections[path_index].t = t_min;
        intersections[path_index].materialId = mat;
        intersections[path_index].surfaceNormal = normal;
    }
}

long kdref_cpu_iteration(const void* nodes_, int numNodes, const void* tris_, const void* geoms_, int geoms_size,
                         const void* mats_, int material_size, const int* offsets, void* paths_, int pixelcount,
                         int iter, float softness, int enableSss, int compaction, int shortstack, int sortmode,
                         float* image /* pixelcount*3, += */, int nthreads, double* seconds)
{
    KDN::NodeBare* nodes = (KDN::NodeBare*)nodes_; KDN::TriBare* tris = (KDN::TriBare*)tris_;
    Geom* geoms = (Geom*)geoms_; Material* materials = (Material*)mats_;
    PathSegment* paths = (PathSegment*)paths_;
    std::vector<ShadeableIntersection> isects(pixelcount);
    int num_paths = pixelcount, depth = 0;
    long total = 0;
    double tIsect = 0, tRest = 0;
    if (nthreads < 1) nthreads = 1;
    for (;;) {
        auto c0 = std::chrono::steady_clock::now();
        memset(isects.data(), 0, sizeof(ShadeableIntersection) * pixelcount);
        std::vector<std::thread> th;
        for (int k = 0; k < nthreads; k++) {
            int lo = (int)((long)num_paths * k / nthreads), hi = (int)((long)num_paths * (k + 1) / nthreads);
            th.emplace_back(kdref_intersect_range, lo, hi, depth, iter, paths, geoms, geoms_size, materials, material_size,
                            isects.data(), softness, tris, nodes, numNodes, (int*)offsets, shortstack != 0);
        }
        for (auto& t : th) t.join();
        auto c1 = std::chrono::steady_clock::now();
        for (int i = 0; i < num_paths; i++) total += paths[i].remainingBounces > 0;
        depth++;
        for (int idx = 0; idx < num_paths; idx++) {                      // shadeMaterial
            PathSegment& ps = paths[idx];
            if (ps.remainingBounces <= 0) continue;
            ShadeableIntersection is = isects[idx];
            if (is.t > 0.0f) {
                Material material = materials[is.materialId];
                glm::vec3 materialColor = material.color;
                if (material.emittance > 0.0f) { ps.color *= (materialColor * material.emittance); ps.remainingBounces = 0; }
                else {
                    if (enableSss && (material.transmittance.x > 0.0f || material.transmittance.y > 0.0f || material.transmittance.z > 0.0f)) {
                        float scenescale = 1.0f;
                        float sss = scenescale * ps.ray.sdepth > 1.0 ? 1.0 : ps.ray.sdepth;
                        sss = 1.0f - sss < 0.0 ? 0.0 : sss;
                        sss = glm::pow(sss, 2);
                        ps.color *= (materialColor)* 1.0f + material.hasRefractive * material.specular.color + sss * material.transmittance;
                    }
                    else if (material.hasRefractive > 0.0f) ps.color *= (materialColor)* 1.0f + material.hasRefractive * material.specular.color;
                    else if (material.hasReflective > 0.0f) ps.color *= (materialColor)* 1.0f + material.hasReflective * material.specular.color;
                    else ps.color *= (materialColor) * 1.0f;
                    ps.remainingBounces--;
                }
            } else { ps.color = glm::vec3(0.0f); ps.remainingBounces = 0; }
        }
        if (compaction) {
            for (int i = 0; i < num_paths; i++)
                if (paths[i].remainingBounces == 0) {
                    int px = paths[i].pixelIndex;
                    image[3 * px] += paths[i].color.x; image[3 * px + 1] += paths[i].color.y; image[3 * px + 2] += paths[i].color.z;
                }
            int k = 0;
            for (int i = 0; i < num_paths; i++) if (paths[i].remainingBounces != 0) paths[k++] = paths[i];
            num_paths = k;
        }
        if (sortmode == 2 || (sortmode == 1 && iter == 2))
            std::stable_sort(paths, paths + num_paths, [](const PathSegment& a, const PathSegment& b) { return a.materialIdHit < b.materialIdHit; });
        auto c2 = std::chrono::steady_clock::now();
        tIsect += std::chrono::duration<double>(c1 - c0).count();
        tRest += std::chrono::duration<double>(c2 - c1).count();
        if (num_paths <= 0 || depth > 7) break;
    }
    if (!compaction)
        for (int i = 0; i < num_paths; i++) {
            int idx = paths[i].pixelIndex; PathSegment p = paths[idx];
            image[3 * p.pixelIndex] += p.color.x; image[3 * p.pixelIndex + 1] += p.color.y; image[3 * p.pixelIndex + 2] += p.color.z;
        }
    if (seconds) { seconds[0] = tIsect; seconds[1] = tRest; }
    return total;
}

// ---- GPU-only: the reference's own render loop on whatever CUDA device is present -------------
// Camera as runCuda() recomputes it is supplied by the caller (main.cpp is GL-bound and not compiled).
int kdref_gpu_render(const void* cam_override, int iter_begin, int iter_count, float focal, float dofAngle,
                     int antialias, float softness, int sss, int compaction, int shortstack, int testing,
                     float* image_out /* W*H*3 */)
{
    if (!g_scene) return -1;
    if (cam_override) memcpy(&g_scene->state.camera, cam_override, sizeof(Camera));
    const Camera& cam = g_scene->state.camera;
    int pixelcount = cam.resolution.x * cam.resolution.y;
    g_scene->state.image.resize(pixelcount);
    uchar4* pbo = nullptr;
    if (cudaMalloc(&pbo, pixelcount * sizeof(uchar4)) != cudaSuccess) return -2;
    pathtraceInit(g_scene, true);
    for (int it = iter_begin; it < iter_begin + iter_count; it++)
        pathtrace(pbo, 0, it, focal, dofAngle, false, antialias != 0, softness, sss != 0, testing != 0,
                  compaction != 0, true, false, false, shortstack != 0);
    if (image_out) memcpy(image_out, g_scene->state.image.data(), sizeof(float) * 3 * pixelcount);
    pathtraceFree(g_scene, true);
    cudaFree(pbo);
    return 0;
}

} // extern "C"
