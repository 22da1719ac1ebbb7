// kdpt_render — headless driver over the C ABI (include/kdpt.h): what the reference's main.cpp does without the GLFW
// window: `prog SCENE.txt [MESH.obj]` (main.cpp:1013-1038), first-frame camera (main.cpp:1053-1073, 1111-1129),
// ITERATIONS calls of pathtrace() (runCuda, main.cpp:1139-1178) and saveImage() (main.cpp:1087-1108: x mirrored,
// divided by the sample count; image.cpp:22-44: clamp to [0,1] * 255, truncated). Runtime toggles that the reference
// flips with keys (main.cpp:35-60, 1187-1306) are command-line flags here.
#include <kdpt.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

static void usage()
{
    fprintf(stderr,
            "usage: kdpt_render SCENE.txt [MESH.obj] [options]\n"
            "  --spp N            samples (default: ITERATIONS of the scene file)\n"
            "  --res WxH          override RES\n"
            "  --mode M           ref_t2 (default, = SHORTSTACK on), ref_t1, shortstack, fullstack, brute (ENABLEKD off), brute_bbox (+ USEBBOX)\n"
            "  --no-aa  --dof-angle A  --focal F  --softness S  --sss  --no-compaction  --sort 0|1|2  --cache-rays  --testing\n"
            "  --kd-depth D       KD build depth (default 13)   --synth-normals   --device N\n"
            "  --out FILE         .png (8-bit, like the reference's savePNG), .hdr (Radiance RGBE, its saveHDR), .pfm or .ppm\n"
            "  --cache-dir DIR    content-addressed binary cache of the loaded scene + KD tree (skips OBJ parse and build)\n"
            "  --checkpoint FILE  write the accumulation buffer + sample count there (at the end and every --checkpoint-every N)\n"
            "  --resume FILE      continue a render from such a file (same scene and resolution)\n"
            "  --json             print one JSON line with timings\n");
}

// ---- minimal PNG writer (stored deflate blocks; no external library) ------------------------------------------------
static uint32_t crcTable[256];
static void crcInit() { for (uint32_t n = 0; n < 256; n++) { uint32_t c = n; for (int k = 0; k < 8; k++) c = c & 1 ? 0xedb88320u ^ (c >> 1) : c >> 1; crcTable[n] = c; } }
static uint32_t crc(uint32_t c, const uint8_t* p, size_t n) { c = ~c; for (size_t i = 0; i < n; i++) c = crcTable[(c ^ p[i]) & 0xff] ^ (c >> 8); return ~c; }
static void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void chunk(FILE* f, const char* tag, const std::vector<uint8_t>& data)
{
    std::vector<uint8_t> b; be32(b, (uint32_t)data.size());
    fwrite(b.data(), 1, 4, f);
    std::vector<uint8_t> td(tag, tag + 4); td.insert(td.end(), data.begin(), data.end());
    fwrite(td.data(), 1, td.size(), f);
    std::vector<uint8_t> c; be32(c, crc(0, td.data(), td.size()));
    fwrite(c.data(), 1, 4, f);
}
static bool writePng(const char* path, const uint8_t* rgb, int w, int h)
{
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    crcInit();
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    std::vector<uint8_t> ihdr; be32(ihdr, w); be32(ihdr, h); ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(f, "IHDR", ihdr);
    std::vector<uint8_t> raw; raw.reserve((size_t)h * (3 * w + 1));
    for (int y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgb + (size_t)y * 3 * w, rgb + (size_t)(y + 1) * 3 * w); }
    std::vector<uint8_t> z; z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size() || off == 0;) {
        size_t n = raw.size() - off; if (n > 65535) n = 65535;
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        for (size_t i = 0; i < n; i++) { a = (a + raw[off + i]) % 65521; b = (b + a) % 65521; }
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        off += n;
        if (n == 0) break;
    }
    be32(z, (b << 16) | a);
    chunk(f, "IDAT", z);
    chunk(f, "IEND", std::vector<uint8_t>());
    fclose(f);
    return true;
}

// Radiance .hdr (what stbi_write_hdr produces for image::saveHDR, image.cpp:40-44), flat RGBE scanlines
static bool writeHdr(const char* path, const float* rgb, int w, int h)
{
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", h, w);
    std::vector<uint8_t> row((size_t)w * 4);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const float* px = rgb + 3 * ((size_t)y * w + x);
            float m = px[0] > px[1] ? (px[0] > px[2] ? px[0] : px[2]) : (px[1] > px[2] ? px[1] : px[2]);
            uint8_t* o = &row[(size_t)x * 4];
            if (!(m >= 1e-32f)) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
            int e; float n = frexpf(m, &e) * 256.0f / m;
            o[0] = (uint8_t)(px[0] * n); o[1] = (uint8_t)(px[1] * n); o[2] = (uint8_t)(px[2] * n); o[3] = (uint8_t)(e + 128);
        }
        fwrite(row.data(), 1, row.size(), f);
    }
    fclose(f);
    return true;
}

// what a checkpoint belongs to: the scene arrays (sizes, the small arrays whole, the ends of the two big ones), the camera
// and the render parameters. A checkpoint of another scene, resolution, mode or parameter set must not be blended in.
static uint64_t renderKey(const kdpt_scene& sc, const kdpt_params& p)
{
    uint64_t h = 1469598103934665603ull;
    auto add = [&](const void* ptr, size_t n) { const unsigned char* b = (const unsigned char*)ptr; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    auto addEnds = [&](const void* ptr, size_t n) { const size_t k = n < (1u << 16) ? n : (1u << 16); add(ptr, k); add((const char*)ptr + (n - k), k); };
    const int sizes[6] = {sc.numNodes, sc.numTris, sc.numGeoms, sc.numMaterials, sc.numShapes, sc.traceDepth};
    add(sizes, sizeof sizes);
    if (sc.numNodes) addEnds(sc.nodes, sizeof(kdpt_node_bare) * (size_t)sc.numNodes);
    if (sc.numTris) addEnds(sc.tris, sizeof(kdpt_tri_bare) * (size_t)sc.numTris);
    if (sc.numGeoms) add(sc.geoms, sizeof(kdpt_geom) * (size_t)sc.numGeoms);
    if (sc.numMaterials) add(sc.materials, sizeof(kdpt_material) * (size_t)sc.numMaterials);
    add(&sc.camera, sizeof sc.camera);
    kdpt_params q = p; q.testingmode = 0; q.laneQueue = 0;       // tuning / timing switches do not change the image
    add(&q, sizeof q);
    return h;
}

// checkpoint of a long render: "KDPTCKP2", W, H, samples rendered so far, render key, W*H*3 floats (the accumulation buffer)
static bool writeCheckpoint(const std::string& path, const std::vector<float>& img, int w, int h, int samples, uint64_t key)
{
    const std::string tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const int hdr[3] = {w, h, samples};
    bool ok = fwrite("KDPTCKP2", 1, 8, f) == 8 && fwrite(hdr, sizeof(int), 3, f) == 3 && fwrite(&key, sizeof key, 1, f) == 1 &&
              fwrite(img.data(), sizeof(float), img.size(), f) == img.size();
    ok = fclose(f) == 0 && ok;
    return ok && rename(tmp.c_str(), path.c_str()) == 0;
}
static bool readCheckpoint(const std::string& path, std::vector<float>& img, int w, int h, int* samples, uint64_t key)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[8]; int hdr[3]; uint64_t k = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, "KDPTCKP2", 8) && fread(hdr, sizeof(int), 3, f) == 3 && hdr[0] == w && hdr[1] == h && hdr[2] >= 0 &&
              fread(&k, sizeof k, 1, f) == 1 && k == key;
    if (ok) { img.resize((size_t)w * h * 3); ok = fread(img.data(), sizeof(float), img.size(), f) == img.size(); }
    fclose(f);
    if (ok) *samples = hdr[2];
    return ok;
}

static int fail(const char* what) { fprintf(stderr, "kdpt_render: %s: %s\n", what, kdpt_last_error()); return 1; }

int main(int argc, char** argv)
{
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) { usage(); return argc < 2 ? 2 : 0; }
    std::string sceneTxt = argv[1], obj, out = "", cacheDir, ckpt, resume;
    int ckptEvery = 0, cacheHit = 0, done = 0;
    int spp = -1, resW = 0, resH = 0, kdDepth = 13, synth = 0, device = 0, json = 0;
    kdpt_params p; memset(&p, 0, sizeof p);
    p.focalLength = 6.0f; p.dofAngle = 0.0f; p.softness = 0.0f; p.antialias = 1; p.enableSss = 0; p.compaction = 1;
    p.mode = KDPT_REF_T2; p.sortByMaterial = 1;                      // main.cpp:35-60
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
        if (a == "--spp") spp = atoi(next());
        else if (a == "--res") { if (sscanf(next(), "%dx%d", &resW, &resH) != 2) { usage(); return 2; } }
        else if (a == "--mode") {
            std::string m = next();
            if (m == "ref_t2") p.mode = KDPT_REF_T2; else if (m == "ref_t1") p.mode = KDPT_REF_T1;
            else if (m == "shortstack") p.mode = KDPT_SHORTSTACK; else if (m == "fullstack") p.mode = KDPT_FULLSTACK;
            else if (m == "brute") p.mode = KDPT_BRUTE; else if (m == "brute_bbox") p.mode = KDPT_BRUTE_BBOX;
            else if (m == "kdviz") p.mode = KDPT_KDVIZ;
            else { usage(); return 2; }
        }
        else if (a == "--no-aa") p.antialias = 0;
        else if (a == "--cache-rays") p.cacheRays = 1;
        else if (a == "--testing") p.testingmode = 1;
        else if (a == "--dof-angle") p.dofAngle = (float)atof(next());
        else if (a == "--focal") p.focalLength = (float)atof(next());
        else if (a == "--softness") p.softness = (float)atof(next());
        else if (a == "--sss") p.enableSss = 1;
        else if (a == "--no-compaction") p.compaction = 0;
        else if (a == "--sort") p.sortByMaterial = atoi(next());
        else if (a == "--kd-depth") kdDepth = atoi(next());
        else if (a == "--synth-normals") synth = 1;
        else if (a == "--device") device = atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--cache-dir") cacheDir = next();
        else if (a == "--checkpoint") ckpt = next();
        else if (a == "--checkpoint-every") ckptEvery = atoi(next());
        else if (a == "--resume") resume = next();
        else if (a == "--json") json = 1;
        else if (a.size() && a[0] != '-' && obj.empty()) obj = a;
        else { usage(); return 2; }
    }
    kdpt_host_scene* hs = nullptr;
    if (!cacheDir.empty()) {
        if (kdpt_scene_load_cached(sceneTxt.c_str(), obj.empty() ? nullptr : obj.c_str(), nullptr, kdDepth, synth, cacheDir.c_str(), &hs, &cacheHit)) return fail("scene");
    } else if (kdpt_scene_load(sceneTxt.c_str(), obj.empty() ? nullptr : obj.c_str(), nullptr, kdDepth, synth, &hs)) return fail("scene");
    kdpt_scene sc; int iterations = 0;
    kdpt_scene_view(hs, &sc, &iterations);
    if (resW > 0 && resH > 0) {                                       // scene.cpp:216-223 with the new RES
        const float PI = 3.1415926535897932384626422832795028841971f;
        float yscaled = tanf(sc.camera.fov[1] * (PI / 180));
        float xscaled = (yscaled * resW) / resH;
        sc.camera.resolution[0] = resW; sc.camera.resolution[1] = resH;
        sc.camera.pixelLength[0] = 2 * xscaled / (float)resW; sc.camera.pixelLength[1] = 2 * yscaled / (float)resH;
    }
    kdpt_camera_first_frame(&sc.camera);
    if (spp < 0) spp = iterations;
    const int W = sc.camera.resolution[0], H = sc.camera.resolution[1];
    kdpt_ctx* ctx = nullptr;
    if (kdpt_create(&sc, device, &ctx)) return fail("create");
    kdpt_render_stats st; memset(&st, 0, sizeof st);
    const uint64_t key = renderKey(sc, p);
    std::vector<float> img((size_t)W * H * 3);
    if (!resume.empty()) {       // the reference cannot do this: its buffer is lost with the process
        if (!readCheckpoint(resume, img, W, H, &done, key)) {
            fprintf(stderr, "kdpt_render: %s is not a checkpoint of this %dx%d render (scene, camera and parameters must match)\n", resume.c_str(), W, H);
            return 1;
        }
        if (done > spp) { fprintf(stderr, "kdpt_render: %s already holds %d samples, more than --spp %d\n", resume.c_str(), done, spp); return 1; }
        if (kdpt_set_image(ctx, img.data())) return fail("set_image");
    }
    // iterations done+1 .. spp, in slices when checkpoints are wanted (iteration numbers seed the RNG, so a resumed
    // render produces the same image as an uninterrupted one)
    while (done < spp) {
        const int n = (!ckpt.empty() && ckptEvery > 0 && ckptEvery < spp - done) ? ckptEvery : spp - done;
        kdpt_render_stats part; memset(&part, 0, sizeof part);
        if (kdpt_render(ctx, done + 1, n, 1, &p, &part)) return fail("render");
        st.rayBounces += part.rayBounces; st.launches += part.launches; st.totalMs += part.totalMs; st.intersectMs += part.intersectMs;
        done += n;
        if (!ckpt.empty()) {
            if (kdpt_get_image(ctx, img.data())) return fail("get_image");
            if (!writeCheckpoint(ckpt, img, W, H, done, key)) { fprintf(stderr, "kdpt_render: cannot write %s\n", ckpt.c_str()); return 1; }
        }
    }
    if (kdpt_get_image(ctx, img.data())) return fail("get_image");
    if (!out.empty()) {
        const float samples = (float)(done > 0 ? done : 1);       // the samples actually accumulated
        bool pfm = out.size() > 4 && out.substr(out.size() - 4) == ".pfm";
        bool ppm = out.size() > 4 && out.substr(out.size() - 4) == ".ppm";
        bool hdr = out.size() > 4 && out.substr(out.size() - 4) == ".hdr";
        if (hdr) {
            std::vector<float> lin((size_t)W * H * 3);
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++)
                    for (int c = 0; c < 3; c++) lin[3 * ((size_t)y * W + (W - 1 - x)) + c] = img[3 * ((size_t)y * W + x) + c] / samples;
            if (!writeHdr(out.c_str(), lin.data(), W, H)) { fprintf(stderr, "kdpt_render: cannot write %s\n", out.c_str()); return 1; }
        } else if (pfm) {
            FILE* f = fopen(out.c_str(), "wb");
            if (!f) { fprintf(stderr, "kdpt_render: cannot write %s\n", out.c_str()); return 1; }
            fprintf(f, "PF\n%d %d\n-1.0\n", W, H);
            for (int y = H - 1; y >= 0; y--)
                for (int x = 0; x < W; x++) { float px[3]; for (int c = 0; c < 3; c++) px[c] = img[3 * ((size_t)y * W + (W - 1 - x)) + c] / samples; fwrite(px, 4, 3, f); }
            fclose(f);
        } else {
            std::vector<uint8_t> rgb((size_t)W * H * 3);
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++)
                    for (int c = 0; c < 3; c++) {
                        float v = img[3 * ((size_t)y * W + x) + c] / samples;           // main.cpp:1092-1097: setPixel(W-1-x, y, pix/samples)
                        v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
                        rgb[3 * ((size_t)y * W + (W - 1 - x)) + c] = (uint8_t)(v * 255.f);
                    }
            if (ppm) {
                FILE* f = fopen(out.c_str(), "wb");
                if (!f) { fprintf(stderr, "kdpt_render: cannot write %s\n", out.c_str()); return 1; }
                fprintf(f, "P6\n%d %d\n255\n", W, H); fwrite(rgb.data(), 1, rgb.size(), f); fclose(f);
            } else if (!writePng(out.c_str(), rgb.data(), W, H)) { fprintf(stderr, "kdpt_render: cannot write %s\n", out.c_str()); return 1; }
        }
    }
    if (json)
        printf("{\"scene\": \"%s\", \"mesh\": \"%s\", \"resolution\": [%d, %d], \"spp\": %d, \"ray_bounces\": %llu, \"ms\": %.3f, "
               "\"mrays_per_s\": %.2f, \"spp_per_s\": %.2f, \"launches\": %llu, \"cache_hit\": %d, \"intersect_ms_per_iteration\": %.4f}\n",
               sceneTxt.c_str(), obj.c_str(), W, H, spp, (unsigned long long)st.rayBounces, st.totalMs,
               st.totalMs > 0 ? st.rayBounces / (st.totalMs * 1e3) : 0.0, st.totalMs > 0 ? spp / (st.totalMs * 1e-3) : 0.0,
               (unsigned long long)st.launches, cacheHit, spp > 0 ? st.intersectMs / spp : 0.0);
    kdpt_destroy(ctx);
    kdpt_scene_free(hs);
    return 0;
}
