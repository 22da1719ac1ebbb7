// main_headless.cpp — a GL-free stand-in for the reference's main.cpp, used only to RUN the drop-in binding above with the
// reference's own, unmodified Scene / KDtree / KDnode / objmesh / tinyobj classes: load (main.cpp:1031-1033), first-frame
// camera (main.cpp:1053-1073 + runCuda 1111-1129, through libkdpt's host helper), then what runCuda does per frame
// (main.cpp:1134-1166): pathtraceInit once, pathtrace(iteration) per sample with the reference's default toggles
// (main.cpp:35-60), and the accumulated scene->state.image dumped as raw floats. TEST INFRASTRUCTURE.
//     usage: dropin_render SCENE.txt MESH.obj MTL_DIR/ WIDTH HEIGHT SPP OUT.f32 [shortstack=1 [OUT_AFTER_CAMERA_MOVE.f32]]
#include <vector_types.h>   // uchar4 (the reference gets it through its CUDA includes)
#include "pathtrace.h"
#include "scene.h"
#include <kdpt.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

int main(int argc, char** argv)
{
    if (argc < 8) { std::fprintf(stderr, "usage: dropin_render SCENE.txt MESH.obj MTL_DIR/ W H SPP OUT.f32 [shortstack]\n"); return 2; }
    const int W = std::atoi(argv[4]), H = std::atoi(argv[5]), spp = std::atoi(argv[6]);
    const bool shortstack = argc > 8 ? std::atoi(argv[8]) != 0 : true;
    Scene* scene = new Scene(std::string(argv[1]));
    scene->loadObj(std::string(argv[2]), std::string(argv[3]));
    // scene.cpp:236-271 never initialises Material::transmittance of the scene-file materials while scatterRay / shadeMaterial
    // branch on it: zero it, as INTEGRATION.md tells the maintainer to
    const int sceneMaterials = scene->hasObj && scene->obj_numshapes > 0 ? scene->obj_materialOffsets[0] : (int)scene->materials.size();
    for (int m = 0; m < sceneMaterials && m < (int)scene->materials.size(); m++) scene->materials[m].transmittance = glm::vec3(0.0f);
    Camera& cam = scene->state.camera;
    {   // RES override, the way the loader derives pixelLength from it (scene.cpp:216-223)
        const float PI_F = 3.1415926535897932384626422832795028841971f;
        float yscaled = tanf(cam.fov.y * (PI_F / 180));
        float xscaled = (yscaled * W) / H;
        cam.resolution.x = W; cam.resolution.y = H;
        cam.pixelLength = glm::vec2(2 * xscaled / (float)W, 2 * yscaled / (float)H);
    }
    static_assert(sizeof(Camera) == sizeof(kdpt_camera), "Camera ABI");
    kdpt_camera_first_frame(reinterpret_cast<kdpt_camera*>(&cam));
    scene->state.image.assign((size_t)W * H, glm::vec3(0.0f));
    pathtraceInit(scene, true);
    for (int it = 1; it <= spp; it++)
        pathtrace(nullptr, 0, it, /*focal*/ 6.0f, /*dofAngle*/ 0.0f, /*cacherays*/ false, /*antialias*/ true, /*softness*/ 0.0f,
                  /*sss*/ false, /*testing*/ false, /*compaction*/ true, /*enablekd*/ true, /*vizkd*/ false, /*usebbox*/ false, shortstack);
    FILE* f = std::fopen(argv[7], "wb");
    if (!f) return 1;
    std::fwrite(scene->state.image.data(), sizeof(glm::vec3), scene->state.image.size(), f);
    std::fclose(f);
    if (argc > 9) {
        // a camera move as runCuda() handles it (main.cpp:1134-1137): pathtraceFree + pathtraceInit with the same scene,
        // iteration restarts at 1; second image -> argv[9]
        cam.position.x += 0.25f;
        pathtraceFree(scene, true);
        scene->state.image.assign((size_t)W * H, glm::vec3(0.0f));
        pathtraceInit(scene, true);
        for (int it = 1; it <= spp; it++)
            pathtrace(nullptr, 0, it, 6.0f, 0.0f, false, true, 0.0f, false, false, true, true, false, false, shortstack);
        FILE* g = std::fopen(argv[9], "wb");
        if (!g) return 1;
        std::fwrite(scene->state.image.data(), sizeof(glm::vec3), scene->state.image.size(), g);
        std::fclose(g);
    }
    pathtraceFree(scene, true);
    std::printf("{\"nodes\": %d, \"triangles\": %d, \"geoms\": %d, \"materials\": %d}\n", scene->hasObj ? scene->numNodes : 0,
                scene->hasObj ? scene->numTriangles : 0, (int)scene->geoms.size(), (int)scene->materials.size());
    return 0;
}
