// kdpt_render.cu — the render loop (sm_100a): generate rays -> [intersect+scatter+shade+gather -> compact -> sort]* ->
// accumulate. Replaces pathtrace() and its kernels (src/pathtrace.cu:315-397, 1571-1734, 2304-2399, 2405-2635) and
// the thrust::remove_if / thrust::sort calls (2591-2606).
//
// Design (B200-first, not the reference's):
//  * path state is SoA (3 x float4 + int2 = 56 B/path) instead of the 56-B AoS PathSegment, so every access is a
//    coalesced 128-bit load/store;
//  * one fused persistent kernel per bounce does what the reference spreads over cudaMemset + intersect + shade +
//    partialGather, and leaves a per-packet survivor count for the compaction;
//  * nothing in an iteration needs the host: path counts live in device memory (np[bounce]), kernels are launched
//    with a fixed persistent grid, so an iteration is a fixed launch sequence with no cudaDeviceSynchronize, no
//    per-bounce cudaMemset of 20 B x pixels and no per-iteration D2H image copy (the reference has all three);
//  * stream compaction = warp-ballot prefix sums (packet counts -> one-block scan -> stable scatter), the material
//    sort = hand-written stable LSD radix sort (8-bit digits) on materialIdHit. No thrust.
//  * the per-iteration camera jitter, which the reference recomputes identically in every pixel thread
//    (RNG seeded by utilhash(iter) only, pathtrace.cu:334), is computed once per iteration on the host.
#include "kdpt_internal.h"
#include "kdpt_render.cuh"
#include "kdpt_lean.cuh"
#include <string.h>

namespace kdpt {

constexpr int kThreads = 128;
constexpr int kMaxBounces = 8;          // the reference's hard cap `depth > 7`, pathtrace.cu:2608
constexpr int kMaxSmemGeoms = 32;
constexpr int kSortTile = 512;
#ifndef KDPT_COOP_MIN
#define KDPT_COOP_MIN 8
#endif
constexpr int kBounceCoopMin = KDPT_COOP_MIN;

struct PathSoA {
    float4* od;   // origin.xyz, sdepth
    float4* dr;   // direction.xyz, isinside (0/1 as float)
    float4* cp;   // color.rgb, pixelIndex (int bits)
    int2* bm;     // remainingBounces, materialIdHit
};

// per-iteration constants of generateRayFromCamera (pathtrace.cu:333-389), identical for every pixel
struct IterConst {
    int iter; int antialias;
    float jx, jy, jz;          // normalised jitter vector (scaled by 0.002 in the kernel)
    float qw, qx, qy, qz;      // DOF rotation quaternion
};

#ifndef KDPT_MAX_BATCH
#define KDPT_MAX_BATCH 64
#endif
constexpr int kMaxBatch = KDPT_MAX_BATCH;   // iterations (samples) processed together by every launch
constexpr int kSegShift = 26;               // the lane queue packs (segment, path index) in one word: <= 64 segments, < 2^26 paths each
static_assert(kMaxBatch <= 64, "the lane queue packs the segment in 6 bits");
constexpr uint64_t kQueueBinMinBytes = 100ull << 20;   // leaf-ordered draining of the wavefront queue pays once the tree exceeds the L2 (see launchBounce)
constexpr int kWavefrontMinTris = 0;         // refined leaf entries from which the bounce runs in its wavefront form
constexpr int kTicketPackets = 1;       // 32-path packets a warp takes per ticket (amortises the atomic + segment lookup)

struct LoopState {
    int np[kMaxBatch][kMaxBounces + 2];      // live paths of each batched iteration at the start of each bounce
    unsigned ticket[2 * (kMaxBounces + 1)];  // packet counters: [b] bounce kernel, [9+b] compaction
    unsigned gqCount[kMaxBounces];           // wavefront form: rays queued for the traversal kernel at each bounce ...
    unsigned gqTicket[kMaxBounces];          // ... and how many of them the traversal kernel has handed to its lanes
    unsigned long long bounceRays[kMaxBounces];
    unsigned long long travRays[kMaxBounces];   // of those, rays that entered the KD tree's root box
    // entering fraction per bounce depth of the most recent batch (< 0: not measured yet) and the counter values it
    // was taken at; maintained by scanPacketsKernel
    unsigned long long seenRays[kMaxBounces], seenTrav[kMaxBounces];
    float enterFrac[kMaxBounces];
};

// A batch = B consecutive iterations of this rank, each owning one segment of `cap` path slots (and capPackets packet
// slots). Paths never move between segments: the reference seeds its RNG with the path's index inside its own
// iteration's compacted array (pathtrace.cu:1691), so compaction and sort are per segment.
struct Batch {
    int B, cap, capPackets, iterBase, iterStride, npix;
};
struct IterConsts { IterConst c[kMaxBatch]; };

// flat ticket -> (segment, first packet); pk[] = exclusive prefix of tickets per segment (smem); a ticket is
// `per` consecutive 32-path packets of one segment
__device__ __forceinline__ void buildPacketPrefix(int* pk, const LoopState* st, int B, int depth, int per)
{
    // the per-segment counts are fetched by B threads at once (one L2 round trip instead of B dependent ones: with 64
    // segments the serial form cost every launch ~25 us before its first packet), then summed from shared memory
    for (int s = threadIdx.x; s < B; s += blockDim.x) pk[s + 1] = (((st->np[s][depth] + 31) >> 5) + per - 1) / per;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int s = 0; s < B; s++) { const int c = pk[s + 1]; pk[s] = acc; acc += c; }
        pk[B] = acc;
    }
    __syncthreads();
}
__device__ __forceinline__ int findSegment(const int* pk, int B, int t)
{
    int s = 0;
    while (s + 1 < B && t >= pk[s + 1]) s++;
    return s;
}

// Wavefront form of a bounce (exact modes on a refined tree): the rays that enter the tree are not traversed inside the
// bounce kernel but appended to a queue in global memory — word-major, one column per ray: the kWorkWords words of its
// PathWork record — which the lean traversal kernel (kdpt_lean.cuh: one ray per lane, lanes refilled as rays finish)
// drains into kHitWords words per ray; a third kernel then runs phase C for the queued paths.
constexpr int kHitWords = 7;        // t, z, tri, nx, ny, nz, mtl
struct RayQueue { float* work; float* hit; unsigned cap; };

struct SceneView {
    const kdpt_geom* geoms; int numGeoms;
    const float4* geomBounds;
    const kdpt_material* materials; int numMaterials;
};

struct BounceArgs {
    int iter, depth, stages;    // stages: kdpt_stage bits; 4 = gather finished paths into the image (compaction on)
    float softness; int enableSss;
    float eps;
    int queuePolicy;            // 0 direct, 1 queued, 2 decided by the kernel (see bounceKernel)
};

__host__ __device__ inline unsigned hUtilhash(unsigned a)
{
    a = (a + 0x7ed55d16) + (a << 12);
    a = (a ^ 0xc761c23c) ^ (a >> 19);
    a = (a + 0x165667b1) + (a << 5);
    a = (a + 0xd3a2646c) ^ (a << 9);
    a = (a + 0xfd7046c5) + (a << 3);
    a = (a ^ 0xb55a4f09) ^ (a >> 16);
    return a;
}
struct HostRng {
    unsigned x;
    void seed(unsigned s) { unsigned v = s % 2147483647u; x = v == 0 ? 1u : v; }
    float u01()
    {
        x = (unsigned)(((unsigned long long)x * 48271ull) % 2147483647ull);
        float r = (float)(x - 1u);
        r /= (1.0f + (float)(2147483646u - 1u));
        return (r * (1.0f - 0.0f)) + 0.0f;
    }
};

// host: the iteration-constant part of generateRayFromCamera (compiled with -ffp-contract=off)
static IterConst makeIterConst(int iter, const kdpt_params& p)
{
    IterConst c; memset(&c, 0, sizeof c);
    c.iter = iter; c.antialias = p.antialias;
    HostRng rng; rng.seed(hUtilhash((unsigned)iter));
    if (p.antialias) {
        float a = rng.u01(), b = rng.u01(), cc = rng.u01();
        float s = 1.0f / sqrtf((a * a + b * b) + cc * cc);
        c.jx = a * s; c.jy = b * s; c.jz = cc * s;
    }
    float u = cosf(KDPT_PI * rng.u01());
    float u2 = u * u;
    float s1 = sqrtf(1 - u2);
    float theta = 2 * KDPT_PI * rng.u01();
    float vx = s1 * cosf(theta), vy = s1 * sinf(theta), vz = u;
    float s = 1.0f / sqrtf((vx * vx + vy * vy) + vz * vz);
    vx = vx * s; vy = vy * s; vz = vz * s;
    (void)rng.u01(); (void)rng.u01();
    float randangle = rng.u01() * KDPT_PI * p.dofAngle;
    float sh = sinf(randangle / 2.0f);
    c.qw = cosf(randangle / 2.0f); c.qx = vx * sh; c.qy = vy * sh; c.qz = vz * sh;
    return c;
}

// ---- generateRayFromCamera ------------------------------------------------------------------------------------
// per-pixel part of generateRayFromCamera (pathtrace.cu:320-393) given the iteration's jitter constants
__device__ __forceinline__ void primaryRay(const kdpt_camera& cam, const IterConst& ic, int index, float focal, V3& org, V3& dir)
{
    int W = cam.resolution[0], H = cam.resolution[1];
    int x = index % W, y = index / W;
    V3 view = mk(cam.view[0], cam.view[1], cam.view[2]), right = mk(cam.right[0], cam.right[1], cam.right[2]);
    V3 up = mk(cam.up[0], cam.up[1], cam.up[2]), pos = mk(cam.position[0], cam.position[1], cam.position[2]);
    dir = normalize((view - (right * cam.pixelLength[0]) * ((float)x - (float)W * 0.5f)) -
                    (up * cam.pixelLength[1]) * ((float)y - (float)H * 0.5f));
    if (ic.antialias) {
        dir = dir + mk(ic.jx, ic.jy, ic.jz) * 0.002f;
        dir = normalize(dir);
    }
    V3 qv = mk(ic.qx, ic.qy, ic.qz);
    V3 uv = cross(qv, dir), uuv = cross(qv, uv);
    V3 rot = dir + ((uv * ic.qw) + uuv) * 2.0f;
    org = (pos + dir * focal) - rot * focal;
    dir = normalize(rot);
}

// stand-alone ray generation (kdpt_generate_rays); the render loop generates the rays inside the first bounce kernel
__global__ void __launch_bounds__(256)
generateRaysKernel(kdpt_camera cam, IterConsts ics, Batch bt, int traceDepth, float focal, PathSoA P, LoopState* st)
{
    const int n = bt.npix;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) {
        for (int k = 0; k < 2 * (kMaxBounces + 1); k++) st->ticket[k] = 0;
        for (int k = 0; k < kMaxBounces; k++) { st->gqCount[k] = 0; st->gqTicket[k] = 0; }
    }
    if (g < bt.B) st->np[g][0] = n;
    if (g >= bt.B * n) return;
    const int seg = g / n, index = g - seg * n;
    V3 org, dir;
    primaryRay(cam, ics.c[seg], index, focal, org, dir);
    const size_t slot = (size_t)seg * bt.cap + index;
    P.od[slot] = make_float4(org.x, org.y, org.z, 0.0f);
    P.dr[slot] = make_float4(dir.x, dir.y, dir.z, 0.0f);
    P.cp[slot] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(index));
    P.bm[slot] = make_int2(traceDepth, 0);
}

// start of a batch: every iteration begins with one path per pixel; packet counters to zero
__global__ void beginBatchKernel(LoopState* st, int B, int npix)
{
    const int g = threadIdx.x;
    if (g < 2 * (kMaxBounces + 1)) st->ticket[g] = 0;
    if (g < kMaxBounces) { st->gqCount[g] = 0; st->gqTicket[g] = 0; }
    if (g < B) st->np[g][0] = npix;
}

// what the first bounce needs to make its own primary rays
struct GenArgs { kdpt_camera cam; IterConsts ics; int traceDepth; float focal; };

// ---- one bounce: intersect (+scatter) + shade + gather, persistent warps pulling 32-path packets ------------------
// resident blocks per SM the register allocation aims for: 5 (<= 96 registers) pays on the light full-stack /
// reference-compatible kernels (measured +9 % on the bunny scene); the short-stack variant keeps its four stack
// entries in registers and is better off with 4 (<= 128)
#ifndef KDPT_BOUNCE_MIN_BLOCKS
#define KDPT_BOUNCE_MIN_BLOCKS(MODE) ((MODE) == KDPT_SHORTSTACK ? 4 : 5)
#endif
// Everything phase C needs about one path (also the record of the per-warp deferred-traversal queue).
struct PathWork {
    PathRay r; float4 cp; int2 bm;
    float tMin; int hitGeom; V3 ipoint, normal;
    int seg, i, packet;        // iteration segment, path index inside it, packet index inside the segment
};
constexpr int kWorkWords = 23;
constexpr int kParkWords = 14;      // the part of it the direct (unqueued) variant parks during the traversal
#ifndef KDPT_DENSE_PACKET
#define KDPT_DENSE_PACKET 16
#endif
constexpr int kDensePacket = KDPT_DENSE_PACKET;   // a packet with at least this many lanes inside the root box is traversed at once

// ---- phase C for one path: scatter, shade, gather, write back. Returns whether the path survives. ----
template <int MODE>
__device__ __forceinline__ bool finishPath(PathWork& w, const Hit& H, bool doIsect, const Tree& T, const SceneView& S,
                                           const PathSoA& P, const BounceArgs& A, const Batch& bt, float* __restrict__ accum,
                                           const unsigned* packetAlive, kdpt_intersection* __restrict__ isectOut)
{
    const size_t slot = (size_t)w.seg * bt.cap + w.i;
    int remaining = w.bm.x;
    V3 color = mk(w.cp.x, w.cp.y, w.cp.z);
    float t = 0.0f; int matId = 0;      // dev_intersections is memset to 0 per bounce (pathtrace.cu:2473)
    if (doIsect) {
        bool objHit = H.tri >= 0;
        if (objHit) {
            // the reference's nudged point: o + d*z + n*eps (pathtrace.cu:993-995)
            w.normal = mk(H.nx, H.ny, H.nz);
            w.ipoint = (w.r.o + w.r.d * H.z) + w.normal * A.eps;
            w.tMin = H.t;
            // reference: material of the LAST triangle glm accepted (pathtrace.cu:991); exact modes: of the hit
            int mtl = (MODE == KDPT_REF_T1 || MODE == KDPT_REF_T2) ? H.anyMtl : H.mtl;
            matId = mtl + T.materialSize - 1;
        } else if (w.hitGeom >= 0) matId = S.geoms[w.hitGeom].materialid;
        // multi-shape OBJs index past materials[] in the reference (Appendix A #5); stay in bounds here
        matId = matId < 0 ? 0 : (matId >= S.numMaterials ? S.numMaterials - 1 : matId);
        if (!objHit && w.hitGeom < 0) t = -1.0f;          // pathtrace.cu:1683-1686
        else {
            const kdpt_material m = S.materials[matId];
            const int iter = bt.iterBase + w.seg * bt.iterStride;
            scatterRay(w.r, w.ipoint, w.normal, m, pathSeed(iter, w.i, A.depth), A.softness);
            t = w.tMin;
            w.bm.y = matId;
        }
        if (isectOut) {
            kdpt_intersection is; is.t = t; is.surfaceNormal[0] = t == -1.0f ? 0.f : w.normal.x;
            is.surfaceNormal[1] = t == -1.0f ? 0.f : w.normal.y; is.surfaceNormal[2] = t == -1.0f ? 0.f : w.normal.z;
            is.materialId = t == -1.0f ? 0 : matId;
            isectOut[slot] = is;
        }
    } else if (isectOut) {      // shade-only call: the caller supplies the intersections
        kdpt_intersection is = isectOut[slot];
        t = is.t; matId = is.materialId;
    }
    if (A.stages & KDPT_STAGE_SHADE) {
        if (t > 0.0f) remaining = shadeHit(S.materials[matId], A.enableSss != 0, w.r.sdepth, color, remaining);
        else { color = mk(0, 0, 0); remaining = 0; }
    }
    if ((A.stages & 4) && remaining == 0) {     // partialGather: each pixel's path finishes exactly once per iteration
        float* px = accum + 3 * ((size_t)w.seg * bt.npix + __float_as_int(w.cp.w));
        px[0] = color.x; px[1] = color.y; px[2] = color.z;
    }
    const bool alive = remaining > 0;
    if ((A.stages & 8) && alive && accum) {     // last bounce of the loop: a path still alive is never gathered
        float* px = accum + 3 * ((size_t)w.seg * bt.npix + __float_as_int(w.cp.w));
        px[0] = 0.f; px[1] = 0.f; px[2] = 0.f;
    }
    // with compaction on, a finished path is never looked at again (its radiance went to accum above and the
    // compaction reads the packet's alive mask, not the path): only survivors are written back
    if (alive || !packetAlive) {
        P.od[slot] = make_float4(w.r.o.x, w.r.o.y, w.r.o.z, w.r.sdepth);
        P.dr[slot] = make_float4(w.r.d.x, w.r.d.y, w.r.d.z, w.r.inside ? 1.0f : 0.0f);
        P.cp[slot] = make_float4(color.x, color.y, color.z, w.cp.w);
        w.bm.x = remaining;
        P.bm[slot] = w.bm;
    }
    return alive;
}

// One bounce over all live paths of the batch. QUEUE selects how the lanes of a packet that enter the KD tree are
// traversed:
//  QUEUE: they are parked in a per-warp ring in shared memory (word-major, conflict free) and the warp traverses when 32
//    have accumulated, with every lane busy; all other lanes of the packet finish immediately. When only a few lanes of a
//    packet enter the tree (the mesh covers a small part of the scene), traversing right away would run the whole
//    traversal with 1-3 of 32 lanes active (measured: a third of all issue slots on the cornell scenes). A deferred path
//    reports its survival to the packet's compaction bookkeeping with atomics afterwards. A packet that is mostly inside
//    the box is flushed right away, so that its rays are traversed together instead of being mixed with strangers.
//  !QUEUE: every packet is traversed at once by the lanes that need it; the same array only parks what phase C needs but
//    the traversal does not (column = lane), so that the traversal runs with ~14 fewer live registers. Better when a
//    third or more of the lanes enter the tree anyway (measured: the queue's bookkeeping and the loss of coherence then
//    cost more than the idle lanes). The kernel picks per launch from the entering fraction it measured at this bounce
//    depth in the previous batch (LoopState::enterFrac), see bounceKernel.
template <int MODE, bool GEN, bool QUEUE, bool GQ = false>
__device__ __forceinline__ void
bounceBody(const Tree& T, const SceneView& S, const PathSoA& P, LoopState* st, const BounceArgs& A, const Batch& bt,
           float* __restrict__ accum, int* __restrict__ packetCount, unsigned* __restrict__ packetAlive,
           kdpt_intersection* __restrict__ isectOut, const GenArgs& gen,
           const kdpt_geom* sGeoms, const float4* sBounds, const int* pk, float* sDefRaw, int ng, RayQueue gq = RayQueue{nullptr, nullptr, 0})
{
    constexpr int kSlots = QUEUE ? 64 : 32;
    constexpr int kWords = QUEUE ? kWorkWords : kParkWords;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    float (*def)[kSlots] = reinterpret_cast<float (*)[kSlots]>(sDefRaw + (threadIdx.x >> 5) * (kWords * kSlots));
    unsigned long long liveLocal = 0, travLocal = 0;
    int segHint = 0;
    int qHead = 0, defCount = 0;            // warp-uniform: ring start and fill
    const int totalTickets = pk[bt.B];
    bool finished = false, flushWanted = false;
    for (;;) {
        // One pass of this loop either drains up to 32 queued paths (QUEUE only: phase B + C, one path per lane, every
        // lane busy; when 32 are waiting, after a dense packet, and at the end until the queue is empty) or takes a new
        // packet (phase A; then QUEUE parks the lanes that enter the tree and finishes the others, !QUEUE traverses at
        // once). Both feed the single phase-C call at the bottom.
        PathWork w;
        w.bm = make_int2(0, 0); w.cp = make_float4(0, 0, 0, 0);
        w.r.o = mk(0, 0, 0); w.r.d = mk(0, 0, 0); w.r.inside = false; w.r.sdepth = 0.f;
        w.tMin = FLT_MAX; w.hitGeom = -1; w.ipoint = mk(0, 0, 0); w.normal = mk(0, 0, 0);
        w.seg = 0; w.i = 0; w.packet = 0;
        Hit H;
        bool finishNow = false, doIsect = false;
        const bool drain = QUEUE && !GQ && (defCount >= 32 || (flushWanted && defCount > 0));
        bool alive = false, live = false;
        unsigned defMask = 0;
        if (drain) {
            const int nq = defCount < 32 ? defCount : 32;
            const bool have = (int)lane < nq;
            const int e = (qHead + (int)lane) & (kSlots - 1);
            if (have) {
                w.r.o = mk(def[0][e], def[1][e], def[2][e]); w.r.d = mk(def[3][e], def[4][e], def[5][e]);
                w.tMin = def[6][e];
            }
            initHit(H, w.tMin);
            if (MODE == KDPT_FULLSTACK) traceExactWarp<kMaxDepth, kBounceCoopMin>(T, have, w.r.o, w.r.d, w.tMin, 0.00001f, H);
            else if (MODE == KDPT_SHORTSTACK) traceExactWarp<4, kBounceCoopMin>(T, have, w.r.o, w.r.d, w.tMin, 0.0001f, H);
            else if (MODE == KDPT_REF_T1) traceRefWarp<false>(T, have, w.r.o, w.r.d, w.tMin, H);
            else if (MODE == KDPT_REF_T2) traceRefWarp<true>(T, have, w.r.o, w.r.d, w.tMin, H);
            else if (MODE == KDPT_BRUTE) traceBruteWarp<false>(T, have, w.r.o, w.r.d, w.tMin, 0.0001f, H);
            else traceBruteWarp<true>(T, have, w.r.o, w.r.d, w.tMin, 0.0001f, H);
            if (have) {
                w.r.inside = def[7][e] != 0.0f; w.r.sdepth = def[8][e];
                w.cp = make_float4(def[9][e], def[10][e], def[11][e], def[12][e]);
                w.bm = make_int2(__float_as_int(def[13][e]), __float_as_int(def[14][e]));
                w.hitGeom = __float_as_int(def[15][e]);
                w.ipoint = mk(def[16][e], def[17][e], def[18][e]); w.normal = mk(def[19][e], def[20][e], def[21][e]);
                const int segi = __float_as_int(def[22][e]);
                w.seg = (int)((unsigned)segi >> kSegShift); w.i = segi & ((1 << kSegShift) - 1); w.packet = w.i >> 5;
            }
            finishNow = have; doIsect = true;
            qHead = (qHead + nq) & (kSlots - 1);
            defCount -= nq;
            flushWanted = false;
        } else {
            if (finished) {
                if (defCount == 0) break;
                flushWanted = true;
                continue;
            }
            int ticket = 0;
            if (lane == 0) ticket = (int)atomicAdd(&st->ticket[A.depth], 1u);
            ticket = __shfl_sync(FULL, ticket, 0);
            if (ticket >= totalTickets) { finished = true; continue; }
            // ---- phase A (per lane): load the path, test the analytic geoms ----
            float4 od = make_float4(0, 0, 0, 0), dr = od;
            bool inRange = false, needTrav = false;
            while (segHint + 1 < bt.B && ticket >= pk[segHint + 1]) segHint++;     // a warp's tickets only grow
            const int seg = segHint;
            const int n = st->np[seg][A.depth];
            const int packet = ticket - pk[seg];
            const unsigned i = (unsigned)packet * 32u + lane;                  // path index inside its iteration
            w.seg = seg; w.i = (int)i; w.packet = packet;
            inRange = i < (unsigned)n;
            if (inRange) {
                const size_t slot = (size_t)seg * bt.cap + i;
                if (GEN) {          // first bounce: path i of an iteration is pixel i; nothing to read
                    V3 org, dir;
                    primaryRay(gen.cam, gen.ics.c[seg], (int)i, gen.focal, org, dir);
                    od = make_float4(org.x, org.y, org.z, 0.0f); dr = make_float4(dir.x, dir.y, dir.z, 0.0f);
                    w.cp = make_float4(1.0f, 1.0f, 1.0f, __int_as_float((int)i)); w.bm = make_int2(gen.traceDepth, 0);
                } else {
                    w.bm = P.bm[slot]; od = P.od[slot]; dr = P.dr[slot]; w.cp = P.cp[slot];     // four independent loads, one round trip
                }
                if (w.bm.x > 0) {
                    live = true;
                    w.r.o = mk(od.x, od.y, od.z); w.r.d = mk(dr.x, dr.y, dr.z); w.r.inside = dr.w != 0.0f; w.r.sdepth = od.w;
                }
            }
            doIsect = live && (A.stages & KDPT_STAGE_INTERSECT);
            if (doIsect) {
                const float dlen = length(w.r.d);
                const V3 ginv = mk(1.0f / w.r.d.x, 1.0f / w.r.d.y, 1.0f / w.r.d.z);
                for (int g = 0; g < S.numGeoms; g++) {
                    // world-AABB filter (see kdpt_create): skip geoms whose padded box the ray misses, that lie behind
                    // the origin, or whose box begins beyond the closest hit so far. The reference's result is the
                    // minimum over geoms of a distance that is -1 for exactly those, so nothing changes. NaN -> no skip.
                    const float4 blo = g < ng ? sBounds[2 * g] : __ldg(S.geomBounds + 2 * g);
                    const float4 bhi = g < ng ? sBounds[2 * g + 1] : __ldg(S.geomBounds + 2 * g + 1);
                    const float v1 = (blo.x - w.r.o.x) * ginv.x, v2 = (bhi.x - w.r.o.x) * ginv.x;
                    const float v3 = (blo.y - w.r.o.y) * ginv.y, v4 = (bhi.y - w.r.o.y) * ginv.y;
                    const float v5 = (blo.z - w.r.o.z) * ginv.z, v6 = (bhi.z - w.r.o.z) * ginv.z;
                    const float tn = fmaxf(fmaxf(fminf(v1, v2), fminf(v3, v4)), fminf(v5, v6));
                    const float tf = fminf(fminf(fmaxf(v1, v2), fmaxf(v3, v4)), fmaxf(v5, v6));
                    if (tn > tf || tf < 0.0f) continue;
                    if (w.tMin < FLT_MAX && tn * dlen > w.tMin * 1.0001f + 1e-3f) continue;
                    V3 pt, nr;
#if KDPT_OUTLINE >= 2       // one copy of the test behind a generic pointer instead of a shared- and a global-memory copy
                    const kdpt_geom* gp = g < ng ? &sGeoms[g] : &S.geoms[g];
                    float tg = geomTest(*gp, w.r.o, w.r.d, pt, nr);
#else
                    float tg = g < ng ? geomTest(sGeoms[g], w.r.o, w.r.d, pt, nr) : geomTest(S.geoms[g], w.r.o, w.r.d, pt, nr);
#endif
                    if (tg > 0.0f && w.tMin > tg) { w.tMin = tg; w.hitGeom = g; w.ipoint = pt; w.normal = nr; }
                }
                // does the ray reach the tree at all? (the traversals' own first step: a miss of the root box ends them)
                if (MODE == KDPT_BRUTE) needTrav = T.numTris > 0;        // every triangle for every ray, no box of any kind
                else if (T.numNodes != 0) {
                    float t0, t1;
                    needTrav = slab(w.r.o, ginv, T.rootLo, T.rootHi, t0, t1);
                    // exact modes: also done when the box begins beyond the closest geom hit (the kernels' early-exit rule)
                    if ((MODE == KDPT_FULLSTACK || MODE == KDPT_SHORTSTACK) && needTrav && w.tMin < FLT_MAX && dlen > 0.0f &&
                        t0 > w.tMin / dlen + kExitSlack / dlen) needTrav = false;
                }
            }
            defMask = __ballot_sync(FULL, needTrav);
            initHit(H, w.tMin);
            if (GQ) {
                // wavefront form: lanes whose ray enters the tree append their record to the global queue, the others finish now
                unsigned qb = 0;
                if (defMask && lane == 0) qb = atomicAdd(&st->gqCount[A.depth], (unsigned)__popc(defMask));
                qb = __shfl_sync(FULL, qb, 0);
                if (needTrav) {
                    float* q = gq.work + (qb + (unsigned)__popc(defMask & ((1u << lane) - 1u)));
                    const size_t cs = gq.cap;
                    q[0] = w.r.o.x; q[cs] = w.r.o.y; q[2 * cs] = w.r.o.z;
                    q[3 * cs] = w.r.d.x; q[4 * cs] = w.r.d.y; q[5 * cs] = w.r.d.z; q[6 * cs] = w.tMin;
                    q[7 * cs] = w.r.inside ? 1.0f : 0.0f; q[8 * cs] = w.r.sdepth;
                    q[9 * cs] = w.cp.x; q[10 * cs] = w.cp.y; q[11 * cs] = w.cp.z; q[12 * cs] = w.cp.w;
                    q[13 * cs] = __int_as_float(w.bm.x); q[14 * cs] = __int_as_float(w.bm.y);
                    q[15 * cs] = __int_as_float(w.hitGeom);
                    q[16 * cs] = w.ipoint.x; q[17 * cs] = w.ipoint.y; q[18 * cs] = w.ipoint.z;
                    q[19 * cs] = w.normal.x; q[20 * cs] = w.normal.y; q[21 * cs] = w.normal.z;
                    q[22 * cs] = __int_as_float((int)(((unsigned)w.seg << kSegShift) | (unsigned)w.i));
                }
                finishNow = live && !needTrav;
            } else if (QUEUE) {
                // lanes whose ray enters the tree go to the queue (fewer than 32 are waiting there), the others finish now
                if (needTrav) {
                    const int e = (qHead + defCount + __popc(defMask & ((1u << lane) - 1u))) & (kSlots - 1);
                    def[0][e] = w.r.o.x; def[1][e] = w.r.o.y; def[2][e] = w.r.o.z;
                    def[3][e] = w.r.d.x; def[4][e] = w.r.d.y; def[5][e] = w.r.d.z; def[6][e] = w.tMin;
                    def[7][e] = w.r.inside ? 1.0f : 0.0f; def[8][e] = w.r.sdepth;
                    def[9][e] = w.cp.x; def[10][e] = w.cp.y; def[11][e] = w.cp.z; def[12][e] = w.cp.w;
                    def[13][e] = __int_as_float(w.bm.x); def[14][e] = __int_as_float(w.bm.y);
                    def[15][e] = __int_as_float(w.hitGeom);
                    def[16][e] = w.ipoint.x; def[17][e] = w.ipoint.y; def[18][e] = w.ipoint.z;
                    def[19][e] = w.normal.x; def[20][e] = w.normal.y; def[21][e] = w.normal.z;
                    def[22][e] = __int_as_float((int)(((unsigned)w.seg << kSegShift) | (unsigned)w.i));
                }
                defCount += __popc(defMask);
                // camera rays (first bounce) are coherent: a packet's rays that reach the mesh are traversed together at
                // once; later bounces are incoherent and are pooled unless most of the packet is in the box anyway
                if (__popc(defMask) >= (GEN ? 1 : kDensePacket)) flushWanted = true;
                finishNow = live && !needTrav;
            } else {
                if (defMask) {      // ---- phase B (whole warp): KD traversal + triangle tests for the lanes that need it
                    def[0][lane] = w.cp.x; def[1][lane] = w.cp.y; def[2][lane] = w.cp.z; def[3][lane] = w.cp.w;
                    def[4][lane] = __int_as_float(w.bm.x); def[5][lane] = __int_as_float(w.bm.y); def[6][lane] = w.r.sdepth;
                    def[7][lane] = __int_as_float(w.hitGeom);
                    def[8][lane] = w.ipoint.x; def[9][lane] = w.ipoint.y; def[10][lane] = w.ipoint.z;
                    def[11][lane] = w.normal.x; def[12][lane] = w.normal.y; def[13][lane] = w.normal.z;
                    if (MODE == KDPT_FULLSTACK) traceExactWarp<kMaxDepth, kBounceCoopMin>(T, needTrav, w.r.o, w.r.d, w.tMin, 0.00001f, H);
                    else if (MODE == KDPT_SHORTSTACK) traceExactWarp<4, kBounceCoopMin>(T, needTrav, w.r.o, w.r.d, w.tMin, 0.0001f, H);
                    else if (MODE == KDPT_REF_T1) traceRefWarp<false>(T, needTrav, w.r.o, w.r.d, w.tMin, H);
                    else if (MODE == KDPT_REF_T2) traceRefWarp<true>(T, needTrav, w.r.o, w.r.d, w.tMin, H);
                    else if (MODE == KDPT_BRUTE) traceBruteWarp<false>(T, needTrav, w.r.o, w.r.d, w.tMin, 0.0001f, H);
                    else traceBruteWarp<true>(T, needTrav, w.r.o, w.r.d, w.tMin, 0.0001f, H);
                    w.cp = make_float4(def[0][lane], def[1][lane], def[2][lane], def[3][lane]);
                    w.bm = make_int2(__float_as_int(def[4][lane]), __float_as_int(def[5][lane])); w.r.sdepth = def[6][lane];
                    w.hitGeom = __float_as_int(def[7][lane]);
                    w.ipoint = mk(def[8][lane], def[9][lane], def[10][lane]); w.normal = mk(def[11][lane], def[12][lane], def[13][lane]);
                }
                finishNow = live;
            }
            if (!live && inRange) {
                alive = w.bm.x > 0;
                if (GEN) { const size_t slot = (size_t)w.seg * bt.cap + w.i; P.od[slot] = od; P.dr[slot] = dr; P.cp[slot] = w.cp; P.bm[slot] = w.bm; }     // traceDepth <= 0
            }
        }
        // ---- phase C (per lane): scatter, shade, gather, write back ----
        if (finishNow) alive = finishPath<MODE>(w, H, doIsect, T, S, P, A, bt, accum, packetAlive, isectOut);
        if (drain) {
            // a drained path adds itself to its packet's compaction entry, which was stored when the packet was taken
            // (ordered by the __syncwarp that ended that pass)
            if (alive) {
                if (packetCount) atomicAdd(&packetCount[w.seg * bt.capPackets + w.packet], 1);
                if (packetAlive) atomicOr(&packetAlive[w.seg * bt.capPackets + w.packet], 1u << (w.i & 31));
            }
        } else {
            const unsigned aliveMask = __ballot_sync(FULL, alive);
            const unsigned liveMask = __ballot_sync(FULL, live);
            if (lane == 0) {
                if (packetCount) packetCount[w.seg * bt.capPackets + w.packet] = __popc(aliveMask);
                if (packetAlive) packetAlive[w.seg * bt.capPackets + w.packet] = aliveMask;
                liveLocal += (unsigned)__popc(liveMask);
                travLocal += (unsigned)__popc(defMask);
            }
        }
        __syncwarp();       // queue slots read above may be refilled, and the packet entry is visible to later atomics
    }
    if (lane == 0 && liveLocal && A.depth < kMaxBounces) {
        atomicAdd(&st->bounceRays[A.depth], liveLocal);
        if (travLocal) atomicAdd(&st->travRays[A.depth], travLocal);
    }
}

// entering fraction below which the queued variant is used
#ifndef KDPT_QUEUE_BELOW
#define KDPT_QUEUE_BELOW 0.30f
#endif
template <int MODE, bool GEN>
__global__ void __launch_bounds__(kThreads, KDPT_BOUNCE_MIN_BLOCKS(MODE))
bounceKernel(Tree T, SceneView S, PathSoA P, LoopState* st, BounceArgs A, Batch bt, float* __restrict__ accum,
             int* __restrict__ packetCount, unsigned* __restrict__ packetAlive, kdpt_intersection* __restrict__ isectOut,
             const GenArgs gen)
{
    __shared__ kdpt_geom sGeoms[kMaxSmemGeoms];
    __shared__ float4 sBounds[2 * kMaxSmemGeoms];
    __shared__ int pk[kMaxBatch + 1];
    __shared__ float sDef[(kThreads / 32) * (GEN ? kParkWords * 32 : kWorkWords * 64)];
    const int ng = S.numGeoms < kMaxSmemGeoms ? S.numGeoms : kMaxSmemGeoms;
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(S.geoms);
        unsigned* dst = reinterpret_cast<unsigned*>(sGeoms);
        for (int k = threadIdx.x; k < ng * (int)(sizeof(kdpt_geom) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
        for (int k = threadIdx.x; k < 2 * ng; k += blockDim.x) sBounds[k] = __ldg(S.geomBounds + k);
    }
    buildPacketPrefix(pk, st, bt.B, A.depth, kTicketPackets);
    // Camera rays are coherent and are always traversed directly. Later bounces: the fraction of rays that entered the
    // tree at this depth in the previous batch (stable from batch to batch; written by scanPacketsKernel), or, in a
    // context's first batch, at the previous depth of this batch. Uniform over the grid: nothing writes these during
    // the launch.
    bool queue = false;
    if (!GEN) {
        if (A.queuePolicy == 2) {
            float f = 1.0f;
            if (st->enterFrac[A.depth] >= 0.0f) f = st->enterFrac[A.depth];
            else if (A.depth > 1 && st->enterFrac[A.depth - 1] >= 0.0f) f = st->enterFrac[A.depth - 1];
            queue = f < KDPT_QUEUE_BELOW;
        } else queue = A.queuePolicy == 1;
    }
    if (!GEN && queue) bounceBody<MODE, false, true>(T, S, P, st, A, bt, accum, packetCount, packetAlive, isectOut, gen, sGeoms, sBounds, pk, sDef, ng);
    else bounceBody<MODE, GEN, false>(T, S, P, st, A, bt, accum, packetCount, packetAlive, isectOut, gen, sGeoms, sBounds, pk, sDef, ng);
}

// ---- wavefront form: front kernel (phase A, phase C of the paths that do not enter the tree), traversal, finish ------
template <int MODE, bool GEN>
__global__ void __launch_bounds__(kThreads, 5)
frontKernel(Tree T, SceneView S, PathSoA P, LoopState* st, BounceArgs A, Batch bt, float* __restrict__ accum,
            int* __restrict__ packetCount, unsigned* __restrict__ packetAlive, const GenArgs gen, RayQueue gq)
{
    __shared__ kdpt_geom sGeoms[kMaxSmemGeoms];
    __shared__ float4 sBounds[2 * kMaxSmemGeoms];
    __shared__ int pk[kMaxBatch + 1];
    const int ng = S.numGeoms < kMaxSmemGeoms ? S.numGeoms : kMaxSmemGeoms;
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(S.geoms);
        unsigned* dst = reinterpret_cast<unsigned*>(sGeoms);
        for (int k = threadIdx.x; k < ng * (int)(sizeof(kdpt_geom) / 4); k += blockDim.x) dst[k] = __ldg(src + k);
        for (int k = threadIdx.x; k < 2 * ng; k += blockDim.x) sBounds[k] = __ldg(S.geomBounds + k);
    }
    buildPacketPrefix(pk, st, bt.B, A.depth, kTicketPackets);
    bounceBody<MODE, GEN, true, true>(T, S, P, st, A, bt, accum, packetCount, packetAlive, nullptr, gen, sGeoms, sBounds, pk, nullptr, ng, gq);
}

// leaf-ordered draining of the queue (as kdpt_trace does for big batches): counting sort of the queued rays by
// (reference leaf they start in, direction octant); the traversal kernel then takes its rays through order[]
__global__ void __launch_bounds__(256)
queueBinKeysKernel(Tree T, LoopState* st, int depth, RayQueue gq, unsigned* __restrict__ keys, unsigned* __restrict__ hist)
{
    const unsigned n = st->gqCount[depth];
    const size_t cs = gq.cap;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* q = gq.work + i;
        const unsigned key = rayBinKey(T, mk(q[0], q[cs], q[2 * cs]), mk(q[3 * cs], q[4 * cs], q[5 * cs]));
        keys[i] = key;
        atomicAdd(hist + key, 1u);
    }
}
__global__ void __launch_bounds__(1024)
queueBinScanKernel(unsigned* __restrict__ hist, int bins)
{
    __shared__ unsigned warpSums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < bins; base += 1024) {
        const int k = base + threadIdx.x;
        const unsigned v = k < bins ? hist[k] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = warpSums[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            warpSums[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const unsigned excl = carry + warpSums[threadIdx.x >> 5] + incl - v;
        if (k < bins) hist[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
queueBinScatterKernel(const unsigned* __restrict__ keys, LoopState* st, int depth, unsigned* __restrict__ offsets, unsigned* __restrict__ order)
{
    const unsigned n = st->gqCount[depth];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        order[atomicAdd(offsets + keys[i], 1u)] = i;
}

#ifndef KDPT_LEAN_MIN_BLOCKS
#define KDPT_LEAN_MIN_BLOCKS 8
#endif
// the queued rays through the refined tree: kdpt_lean.cuh, rays and results in the queue's columns
template <bool RESTART, bool WIDE>
__global__ void __launch_bounds__(kLeanThreads, KDPT_LEAN_MIN_BLOCKS)
traceQueueKernel(Tree T, LoopState* st, int depth, RayQueue gq, float eps, const unsigned* __restrict__ order)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    __shared__ uint2 sStackAll[(RESTART ? kLeanRing : kLeanSmemStack) * kLeanThreads];
    __shared__ uint4 sTop[kTopMaxNodes];
    leanStageTop(T, sTop);
    uint2* sStack = sStackAll + threadIdx.x;
    uint2 lStack[RESTART ? 1 : (WIDE ? kWideStack : kMaxDepth) - kLeanSmemStack];
    const unsigned n = st->gqCount[depth];
    const size_t cs = gq.cap;
    LeanState S; Hit H;
    bool alive = false, more = n > 0;
    unsigned slot = 0;
    auto store = [&](unsigned i) {
        float* h = gq.hit + i;
        h[0] = H.t; h[cs] = H.z; h[2 * cs] = __int_as_float(H.tri); h[3 * cs] = H.nx; h[4 * cs] = H.ny; h[5 * cs] = H.nz;
        h[6 * cs] = __int_as_float(H.mtl);
    };
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !alive);
        if (more && __popc(idle) >= KDPT_LEAN_REFILL_MIN) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&st->gqTicket[depth], (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)__popc(idle) >= n) more = false;
            if (!alive) {
                unsigned i = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (i < n) {
                    if (order) i = __ldg(order + i);
                    const float* q = gq.work + i;
                    const V3 o = mk(__ldg(q), __ldg(q + cs), __ldg(q + 2 * cs)), d = mk(__ldg(q + 3 * cs), __ldg(q + 4 * cs), __ldg(q + 5 * cs));
                    slot = i;
                    alive = leanBegin<WIDE>(T, o, d, __ldg(q + 6 * cs), S, H);
                    if (!alive) store(i);
                }
            }
        }
        if (!__any_sync(FULL, alive)) { if (!more) break; continue; }
        const bool was = alive;
        alive = leanRound<RESTART, WIDE>(T, alive, S, H, eps, sStack, lStack, sTop);
        if (was && !alive) store(slot);
    }
}

// phase C of the queued paths
template <int MODE>
__global__ void __launch_bounds__(kThreads, 5)
finishQueueKernel(Tree T, SceneView S, PathSoA P, LoopState* st, BounceArgs A, Batch bt, float* __restrict__ accum,
                  int* __restrict__ packetCount, unsigned* __restrict__ packetAlive, RayQueue gq)
{
    const unsigned n = st->gqCount[A.depth];
    const size_t cs = gq.cap;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const float* q = gq.work + e;
        const float* h = gq.hit + e;
        PathWork w;
        w.r.o = mk(q[0], q[cs], q[2 * cs]); w.r.d = mk(q[3 * cs], q[4 * cs], q[5 * cs]); w.tMin = q[6 * cs];
        w.r.inside = q[7 * cs] != 0.0f; w.r.sdepth = q[8 * cs];
        w.cp = make_float4(q[9 * cs], q[10 * cs], q[11 * cs], q[12 * cs]);
        w.bm = make_int2(__float_as_int(q[13 * cs]), __float_as_int(q[14 * cs]));
        w.hitGeom = __float_as_int(q[15 * cs]);
        w.ipoint = mk(q[16 * cs], q[17 * cs], q[18 * cs]); w.normal = mk(q[19 * cs], q[20 * cs], q[21 * cs]);
        const int segi = __float_as_int(q[22 * cs]);
        w.seg = (int)((unsigned)segi >> kSegShift); w.i = segi & ((1 << kSegShift) - 1); w.packet = w.i >> 5;
        Hit H;
        H.t = h[0]; H.z = h[cs]; H.u = 0.f; H.v = 0.f; H.tri = __float_as_int(h[2 * cs]); H.nx = h[3 * cs]; H.ny = h[4 * cs]; H.nz = h[5 * cs];
        H.mtl = __float_as_int(h[6 * cs]); H.anyMtl = H.mtl;
        const bool alive = finishPath<MODE>(w, H, true, T, S, P, A, bt, accum, packetAlive, nullptr);
        if (alive) {        // the packet's entry was stored by the front kernel; a queued survivor adds itself
            if (packetCount) atomicAdd(&packetCount[w.seg * bt.capPackets + w.packet], 1);
            if (packetAlive) atomicOr(&packetAlive[w.seg * bt.capPackets + w.packet], 1u << (w.i & 31));
        }
    }
}

// ---- stream compaction (replaces thrust::remove_if, pathtrace.cu:2591-2597) ---------------------------------------
// packetCount[k] = survivors in paths [32k, 32k+32). One block turns them into exclusive offsets and publishes the
// next bounce's path count.
__global__ void __launch_bounds__(1024)
scanPacketsKernel(int* __restrict__ packetCountAll, LoopState* st, int depth, int compaction, int capPackets)
{
    __shared__ int warpSums[32];
    __shared__ int carry;
    const int seg = blockIdx.x;                     // one block per batched iteration
    int* packetCount = packetCountAll + (size_t)seg * capPackets;
    const int n = st->np[seg][depth];
    if (seg == 0 && threadIdx.x == 0 && depth < kMaxBounces) {      // entering fraction of the bounce that just ran
        const unsigned long long r = st->bounceRays[depth], t = st->travRays[depth];
        const unsigned long long dr = r - st->seenRays[depth];
        if (dr) st->enterFrac[depth] = (float)(t - st->seenTrav[depth]) / (float)dr;
        st->seenRays[depth] = r; st->seenTrav[depth] = t;
    }
    if (!compaction) { if (threadIdx.x == 0) st->np[seg][depth + 1] = n; return; }
    const int numPackets = (n + 31) >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < numPackets; base += 1024) {
        int k = base + threadIdx.x;
        int v = k < numPackets ? packetCount[k] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warpSums[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            warpSums[threadIdx.x] = wi - w;
        }
        __syncthreads();
        int excl = carry + warpSums[threadIdx.x >> 5] + incl - v;
        if (k < numPackets) packetCount[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) st->np[seg][depth + 1] = carry;
}

__global__ void __launch_bounds__(kThreads)
compactScatterKernel(PathSoA P, PathSoA Q, const int* __restrict__ packetOffset, const unsigned* __restrict__ packetAlive,
                     LoopState* st, int depth, Batch bt)
{
    const unsigned lane = threadIdx.x & 31u;
    const int warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, numWarps = (gridDim.x * blockDim.x) >> 5;
    // one flat packet space over the batched iterations (as in the bounce kernel): a loop over the segments would pay
    // one dependent L2 round trip per segment and warp before doing anything
    __shared__ int pk[kMaxBatch + 1];
    buildPacketPrefix(pk, st, bt.B, depth, 1);
    const int total = pk[bt.B];
    int seg = 0;
    for (int t = warpId; t < total; t += numWarps) {      // uniform work: static striding; a warp's tickets only grow
        while (seg + 1 < bt.B && t >= pk[seg + 1]) seg++;
        const int packet = t - pk[seg];
        const unsigned i = (unsigned)packet * 32u + lane;
        const size_t slot = (size_t)seg * bt.cap + i;
        const unsigned mask = packetAlive[seg * bt.capPackets + packet];   // remainingBounces != 0 (is_zero_bounce, pathtrace.cu:103-110)
        if ((mask >> lane) & 1u) {
            size_t dst = (size_t)seg * bt.cap + packetOffset[seg * bt.capPackets + packet] + __popc(mask & ((1u << lane) - 1u));
            Q.od[dst] = P.od[slot]; Q.dr[dst] = P.dr[slot]; Q.cp[dst] = P.cp[slot]; Q.bm[dst] = P.bm[slot];
        }
    }
}

// ---- stable LSD radix sort by materialIdHit (replaces thrust::sort, pathtrace.cu:2600-2606) -----------------------
__device__ __forceinline__ unsigned sortDigit(int materialIdHit, int shift) { return ((unsigned)(materialIdHit + 1) >> shift) & 255u; }

// All three kernels work per batched iteration (segment `seg` of the path arrays, its own histogram block of
// 256 x maxTiles counters): paths are only ever reordered inside their own iteration.
__global__ void __launch_bounds__(kSortTile)
sortHistKernel(PathSoA P, LoopState* st, int depth, int shift, unsigned* __restrict__ histAll, int maxTiles, Batch bt)
{
    __shared__ unsigned h[256];
    for (int seg = 0; seg < bt.B; seg++) {
        const int n = st->np[seg][depth];
        const int numTiles = (n + kSortTile - 1) / kSortTile;
        unsigned* hist = histAll + (size_t)seg * 256 * maxTiles;
        for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
            if (threadIdx.x < 256) h[threadIdx.x] = 0;
            __syncthreads();
            int i = tile * kSortTile + threadIdx.x;
            if (i < n) atomicAdd(&h[sortDigit(P.bm[(size_t)seg * bt.cap + i].y, shift)], 1u);
            __syncthreads();
            if (threadIdx.x < 256) hist[threadIdx.x * maxTiles + tile] = h[threadIdx.x];
            __syncthreads();
        }
    }
}
// exclusive scan over (digit-major, tile-minor) counts; one block per batched iteration
__global__ void __launch_bounds__(1024)
sortScanKernel(LoopState* st, int depth, unsigned* __restrict__ histAll, int maxTiles)
{
    __shared__ unsigned warpSums[32];
    __shared__ unsigned carry;
    const int seg = blockIdx.x;
    unsigned* hist = histAll + (size_t)seg * 256 * maxTiles;
    const int n = st->np[seg][depth];
    const int numTiles = (n + kSortTile - 1) / kSortTile;
    const int total = 256 * numTiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        int k = base + threadIdx.x;
        int digit = k / (numTiles > 0 ? numTiles : 1), tile = k - digit * numTiles;
        unsigned v = k < total ? hist[digit * maxTiles + tile] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = warpSums[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            warpSums[threadIdx.x] = wi - w;
        }
        __syncthreads();
        unsigned excl = carry + warpSums[threadIdx.x >> 5] + incl - v;
        if (k < total) hist[digit * maxTiles + tile] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kSortTile)
sortScatterKernel(PathSoA P, PathSoA Q, LoopState* st, int depth, int shift, const unsigned* __restrict__ histAll, int maxTiles, Batch bt)
{
    __shared__ unsigned short warpDigitCount[kSortTile / 32][256];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int seg = 0; seg < bt.B; seg++) {
        const int n = st->np[seg][depth];
        const int numTiles = (n + kSortTile - 1) / kSortTile;
        const unsigned* hist = histAll + (size_t)seg * 256 * maxTiles;
        const size_t base = (size_t)seg * bt.cap;
        for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
            for (int k = threadIdx.x; k < (kSortTile / 32) * 256; k += blockDim.x) (&warpDigitCount[0][0])[k] = 0;
            __syncthreads();
            int i = tile * kSortTile + threadIdx.x;
            bool valid = i < n;
            int2 bm = valid ? P.bm[base + i] : make_int2(0, 0);
            unsigned digit = valid ? sortDigit(bm.y, shift) : 0xffffffffu;
            unsigned peers = __match_any_sync(0xffffffffu, digit);
            unsigned rank = __popc(peers & ((1u << lane) - 1u));
            if (valid && rank == 0) warpDigitCount[warp][digit] = (unsigned short)__popc(peers);
            __syncthreads();
            if (valid) {
                unsigned dst = hist[digit * maxTiles + tile] + rank;
                for (unsigned w = 0; w < warp; w++) dst += warpDigitCount[w][digit];
                Q.od[base + dst] = P.od[base + i]; Q.dr[base + dst] = P.dr[base + i]; Q.cp[base + dst] = P.cp[base + i]; Q.bm[base + dst] = bm;
            }
            __syncthreads();
        }
    }
}

// finalGather (compaction off), pathtrace.cu:2373-2383: every path adds its colour
__global__ void finalGatherKernel(PathSoA P, LoopState* st, int depth, Batch bt, float* __restrict__ accum)
{
    for (int seg = 0; seg < bt.B; seg++) {
        const int n = st->np[seg][depth];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            float4 cp = P.cp[(size_t)seg * bt.cap + i];
            float* px = accum + 3 * ((size_t)seg * bt.npix + __float_as_int(cp.w));
            px[0] = cp.x; px[1] = cp.y; px[2] = cp.z;
        }
    }
}

// image += contribution of iteration 0, then 1, ... of the batch: the same summation order as rendering the
// iterations one after the other, so the result does not depend on the batch size
__global__ void accumulateKernel(float* __restrict__ image, const float* __restrict__ accum, int B, int numFloats)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < numFloats; k += gridDim.x * blockDim.x) {
        float v = image[k];
        for (int s = 0; s < B; s++) v += accum[(size_t)s * numFloats + k];
        image[k] = v;
    }
}

// ---- radiance all-reduce over peer memory (kdpt_allreduce_image) ------------------------------------------------
struct PeerImages { float* img[16]; int n; };
// device `self` owns floats [lo, hi): total over all buffers in context order, written into its own buffer
__global__ void __launch_bounds__(256)
peerReduceSliceKernel(PeerImages pi, int self, size_t lo, size_t hi)
{
    // float4 body (the slice bounds are multiples of 4 except possibly the very end of the image)
    const size_t n4 = (hi - lo) / 4;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
        float4 acc = reinterpret_cast<const float4*>(pi.img[0] + lo)[k];
        for (int j = 1; j < pi.n; j++) {
            const float4 v = reinterpret_cast<const float4*>(pi.img[j] + lo)[k];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(pi.img[self] + lo)[k] = acc;
    }
    for (size_t k = lo + 4 * n4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < hi; k += (size_t)gridDim.x * blockDim.x) {
        float acc = pi.img[0][k];
        for (int j = 1; j < pi.n; j++) acc += pi.img[j][k];
        pi.img[self][k] = acc;
    }
}
// device `self` pulls every other device's finished slice
__global__ void __launch_bounds__(256)
peerGatherKernel(PeerImages pi, int self, size_t per, size_t total)
{
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int owner = (int)(k / per) < pi.n ? (int)(k / per) : pi.n - 1;
        if (owner != self) pi.img[self][k] = pi.img[owner][k];
    }
}

// AoS <-> SoA for the stage-level entry points
__global__ void packPathsKernel(const kdpt_path_segment* __restrict__ in, int n, PathSoA P, LoopState* st, int depth)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { st->np[0][depth] = n; st->ticket[depth] = 0; st->gqCount[depth] = 0; st->gqTicket[depth] = 0; }
    if (i >= n) return;
    kdpt_path_segment s = in[i];
    P.od[i] = make_float4(s.origin[0], s.origin[1], s.origin[2], s.sdepth);
    P.dr[i] = make_float4(s.direction[0], s.direction[1], s.direction[2], s.isinside ? 1.0f : 0.0f);
    P.cp[i] = make_float4(s.color[0], s.color[1], s.color[2], __int_as_float(s.pixelIndex));
    P.bm[i] = make_int2(s.remainingBounces, s.materialIdHit);
}
__global__ void unpackPathsKernel(PathSoA P, int n, kdpt_path_segment* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 od = P.od[i], dr = P.dr[i], cp = P.cp[i]; int2 bm = P.bm[i];
    kdpt_path_segment s; memset(&s, 0, sizeof s);
    s.origin[0] = od.x; s.origin[1] = od.y; s.origin[2] = od.z; s.sdepth = od.w;
    s.direction[0] = dr.x; s.direction[1] = dr.y; s.direction[2] = dr.z; s.isinside = dr.w != 0.0f;
    s.color[0] = cp.x; s.color[1] = cp.y; s.color[2] = cp.z; s.pixelIndex = __float_as_int(cp.w);
    s.remainingBounces = bm.x; s.materialIdHit = bm.y;
    out[i] = s;
}

// stage-level compaction entry point: packet alive masks / counts from remainingBounces (inside the render loop the
// bounce kernel leaves them behind itself)
__global__ void __launch_bounds__(kThreads)
aliveMaskKernel(PathSoA P, int n, int* __restrict__ packetCount, unsigned* __restrict__ packetAlive)
{
    const unsigned lane = threadIdx.x & 31u;
    const int warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, numWarps = (gridDim.x * blockDim.x) >> 5;
    for (int packet = warpId; packet * 32 < n; packet += numWarps) {
        const int i = packet * 32 + (int)lane;
        const unsigned mask = __ballot_sync(0xffffffffu, i < n && P.bm[i].x != 0);     // is_zero_bounce, pathtrace.cu:103-110
        if (lane == 0) { packetCount[packet] = __popc(mask); packetAlive[packet] = mask; }
    }
}

// ---- host driver --------------------------------------------------------------------------------------------------
struct RenderBuffers {
    PathSoA A, B;
    LoopState* st;
    int* packetCount; unsigned* packetAlive;
    unsigned* hist; int maxTiles;
    float* image;
    float* accum;
    int cap, capPackets;
    RayQueue gq;        // wavefront form only (work == nullptr otherwise)
};

// n = path slots per iteration, batch = iterations per launch
// The wavefront form pays for rays that enter a refined tree with enough in it that traversal dominates the bounce.
// KDPT_WAVEFRONT=0/1 forces it off/on in development.
static bool useWavefront(const kdpt_ctx* ctx, int mode)
{
    if (mode != KDPT_FULLSTACK && mode != KDPT_SHORTSTACK) return false;
    if (ctx->xNumNodes == 0) return false;
    static const int env = getenv("KDPT_WAVEFRONT") ? atoi(getenv("KDPT_WAVEFRONT")) : -1;
    if (env >= 0) return env != 0;
    return ctx->xtree.numTris >= kWavefrontMinTris;
}

static int ensureRenderBuffers(kdpt_ctx* ctx, int n, int batch, RenderBuffers& rb, bool needAccum, bool wavefront = false)
{
    int rc;
    const size_t per = sizeof(float4) * 3 + sizeof(int2);
    const size_t slots = (size_t)n * batch;
    if ((rc = ctx->pathsA.ensure(per * slots)) || (rc = ctx->pathsB.ensure(per * slots))) return rc;
    auto carve = [&](DeviceBuffer& b, PathSoA& P) {
        char* p = (char*)b.p;
        size_t cap = b.bytes / per;
        P.od = (float4*)p; P.dr = (float4*)(p + sizeof(float4) * cap); P.cp = (float4*)(p + 2 * sizeof(float4) * cap);
        P.bm = (int2*)(p + 3 * sizeof(float4) * cap);
    };
    carve(ctx->pathsA, rb.A); carve(ctx->pathsB, rb.B);
    bool fresh = ctx->scanTmp.p == nullptr;
    if ((rc = ctx->scanTmp.ensure(sizeof(LoopState)))) return rc;
    if (fresh) {
        static const LoopState init = [] { LoopState z; memset(&z, 0, sizeof z); for (float& f : z.enterFrac) f = -1.0f; return z; }();
        KDPT_CUDA(cudaMemcpyAsync(ctx->scanTmp.p, &init, sizeof(LoopState), cudaMemcpyHostToDevice, ctx->stream));
    }
    rb.st = ctx->scanTmp.as<LoopState>();
    rb.cap = n; rb.capPackets = (n + 31) / 32 + 1;
    if ((rc = ctx->keysTmp.ensure(2 * sizeof(int) * (size_t)rb.capPackets * batch))) return rc;
    rb.packetCount = ctx->keysTmp.as<int>();
    rb.packetAlive = (unsigned*)(rb.packetCount + (size_t)rb.capPackets * batch);
    rb.maxTiles = (n + kSortTile - 1) / kSortTile + 1;
    if ((rc = ctx->hist.ensure(sizeof(unsigned) * 256 * (size_t)rb.maxTiles * batch))) return rc;
    rb.hist = ctx->hist.as<unsigned>();
    rb.image = ctx->boundImage ? ctx->boundImage : ctx->image.as<float>();
    rb.gq = RayQueue{nullptr, nullptr, 0};
    if (wavefront) {
        if ((rc = ctx->gqWork.ensure(sizeof(float) * kWorkWords * slots)) || (rc = ctx->gqHit.ensure(sizeof(float) * kHitWords * slots))) return rc;
        rb.gq.work = ctx->gqWork.as<float>(); rb.gq.hit = ctx->gqHit.as<float>();
        rb.gq.cap = (unsigned)(ctx->gqHit.bytes / (sizeof(float) * kHitWords) < ctx->gqWork.bytes / (sizeof(float) * kWorkWords)
                               ? ctx->gqHit.bytes / (sizeof(float) * kHitWords) : ctx->gqWork.bytes / (sizeof(float) * kWorkWords));
    }
    rb.accum = nullptr;
    if (needAccum) {
        if ((rc = ctx->accum.ensure(sizeof(float) * 3 * (size_t)ctx->pixelCount * batch))) return rc;
        rb.accum = ctx->accum.as<float>();
    }
    return KDPT_OK;
}

static int ensureImage(kdpt_ctx* ctx)
{
    if (!ctx->scene.hasCamera) { setError("render: the scene has no camera (resolution 0)"); return KDPT_ERR_STATE; }
    int n = ctx->scene.camera.resolution[0] * ctx->scene.camera.resolution[1];
    if (ctx->image.p && ctx->pixelCount == n) return KDPT_OK;
    int rc = ctx->image.alloc(sizeof(float) * 3 * (size_t)n);
    if (rc) return rc;
    ctx->pixelCount = n;
    KDPT_CUDA(cudaMemsetAsync(ctx->image.p, 0, sizeof(float) * 3 * (size_t)n, ctx->stream));
    return KDPT_OK;
}

static int checkParams(const kdpt_params* p)
{
    if (!p) { setError("render: null params"); return KDPT_ERR_ARG; }
    if (p->mode < 0 || p->mode >= KDPT_MODE_COUNT) { setError("render: unknown traversal mode"); return KDPT_ERR_ARG; }
    if (p->sortByMaterial < 0 || p->sortByMaterial > 2) { setError("render: sortByMaterial must be 0, 1 or 2"); return KDPT_ERR_ARG; }
    if (p->laneQueue < 0 || p->laneQueue > 2) { setError("render: laneQueue must be 0, 1 or 2"); return KDPT_ERR_ARG; }
    return KDPT_OK;
}

// KDPT_KDVIZ (pathTraceOneBounceKDbareBoxes, pathtrace.cu:1738-1885): the reference tests the scene's geoms, then every node
// box with boxIntersectionTestBox — its cube test on matrices made from the box — and gives a box hit the last material. That is
// the ordinary loop over an extended geom list: the scene's geoms followed by one cube geom per node. Built on first use.
static int ensureVizScene(kdpt_ctx* ctx)
{
    const int numNodes = (int)(ctx->nodeBoxes.size() / 6);
    const int total = ctx->scene.numGeoms + numNodes;
    if (ctx->vizCount == total && (ctx->vizGeoms.p || total == 0)) return KDPT_OK;
    std::vector<kdpt_geom> g((size_t)(total > 0 ? total : 1));
    std::vector<float> gb(8 * g.size(), 0.f);
    if (ctx->scene.numGeoms > 0) {
        KDPT_CUDA(cudaMemcpyAsync(g.data(), ctx->scene.geoms.p, sizeof(kdpt_geom) * (size_t)ctx->scene.numGeoms, cudaMemcpyDeviceToHost, ctx->stream));
        KDPT_CUDA(cudaMemcpyAsync(gb.data(), ctx->scene.geomBounds.p, sizeof(float) * 8 * (size_t)ctx->scene.numGeoms, cudaMemcpyDeviceToHost, ctx->stream));
        KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    for (int i = 0; i < numNodes; i++) {
        kdpt_geom& c = g[(size_t)ctx->scene.numGeoms + i];
        nodeBoxGeom(&ctx->nodeBoxes[6 * (size_t)i], &ctx->nodeBoxes[6 * (size_t)i + 3], ctx->scene.numMaterials - 1, &c);
        geomWorldBounds(c.transform, &gb[8 * ((size_t)ctx->scene.numGeoms + i)]);
    }
    int rc;
    if ((rc = ctx->vizGeoms.alloc(sizeof(kdpt_geom) * g.size()))) return rc;
    if ((rc = ctx->vizBounds.alloc(sizeof(float) * gb.size()))) return rc;
    KDPT_CUDA(cudaMemcpyAsync(ctx->vizGeoms.p, g.data(), sizeof(kdpt_geom) * g.size(), cudaMemcpyHostToDevice, ctx->stream));
    KDPT_CUDA(cudaMemcpyAsync(ctx->vizBounds.p, gb.data(), sizeof(float) * gb.size(), cudaMemcpyHostToDevice, ctx->stream));
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->vizCount = total;
    return KDPT_OK;
}

static int launchBounce(kdpt_ctx* ctx, const kdpt_params& p, const RenderBuffers& rb, PathSoA P, const Batch& bt, int depth,
                        int stages, kdpt_intersection* dIsects, uint64_t& launches, const GenArgs* gen)
{
    SceneView S{ctx->scene.geoms.as<kdpt_geom>(), ctx->scene.numGeoms, ctx->scene.geomBounds.as<float4>(), ctx->scene.materials.as<kdpt_material>(), ctx->scene.numMaterials};
    if (p.mode == KDPT_KDVIZ) {
        if (ctx->scene.numMaterials < 1) { setError("render: KDPT_KDVIZ needs at least one material (the boxes take the last one)"); return KDPT_ERR_STATE; }
        int rc = ensureVizScene(ctx);
        if (rc) return rc;
        S.geoms = ctx->vizGeoms.as<kdpt_geom>(); S.numGeoms = ctx->vizCount; S.geomBounds = ctx->vizBounds.as<float4>();
    }
    BounceArgs A; A.iter = bt.iterBase; A.depth = depth; A.stages = stages; A.softness = p.softness; A.enableSss = p.enableSss;
    A.eps = (p.mode == KDPT_FULLSTACK || p.mode == KDPT_REF_T1) ? 0.00001f : 0.0001f;
    int grid = ctx->numSMs * 8;
    cudaStream_t s = ctx->stream;
    int* pc = p.compaction ? rb.packetCount : nullptr;
    unsigned* pa = p.compaction ? rb.packetAlive : nullptr;
    static const GenArgs noGen{};
    // kdpt_params.laneQueue (0 auto, 1 direct, 2 queued); KDPT_QUEUE=0/1/2 (direct/queued/auto) overrides it in development
    static const int envQueue = getenv("KDPT_QUEUE") ? atoi(getenv("KDPT_QUEUE")) : -1;
    A.queuePolicy = envQueue >= 0 ? envQueue : (p.laneQueue == 1 ? 0 : p.laneQueue == 2 ? 1 : 2);
    // the refined view is walked by the lean traversal kernel only; the packet kernels take the tree as the reference built it
    const Tree& TV = (rb.gq.work && !dIsects) ? ctx->xtree : ctx->tree;
    if (rb.gq.work && !dIsects) {
        // wavefront form: front (phase A + the paths that miss the tree) -> lean traversal of the queued rays -> finish
        static const int envBlocks = getenv("KDPT_LEAN_BLOCKS") ? atoi(getenv("KDPT_LEAN_BLOCKS")) : KDPT_LEAN_MIN_BLOCKS;
        const int lgrid = ctx->numSMs * envBlocks;
        // leaf-ordered draining (KDPT_QUEUE_BIN=0/1 forces it off/on in development): only the later bounces are incoherent, and
        // only trees beyond the L2 gain from it (1.6 M dragon +3 %; dragon_5 and the bunny lose 10-12 % to the three extra kernels)
        static const int envBin = getenv("KDPT_QUEUE_BIN") ? atoi(getenv("KDPT_QUEUE_BIN")) : -1;
        const bool wantBin = envBin >= 0 ? envBin != 0 : ctx->deviceBytes >= kQueueBinMinBytes;
        const unsigned* order = nullptr;
        if (wantBin && depth > 0 && TV.numRefNodes > 1) {
            const size_t slots = rb.gq.cap;
            const int bins = TV.numRefNodes * 8;
            int rc;
            if ((rc = ctx->binTmp.ensure(sizeof(unsigned) * (2 * slots + (size_t)bins)))) return rc;
            unsigned* keys = ctx->binTmp.as<unsigned>();
            unsigned* ord = keys + slots;
            unsigned* hist = ord + slots;
            order = ord;
            KDPT_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * (size_t)bins, s));
        }
        auto binQueue = [&]() {
            if (!order) return;
            const size_t slots = rb.gq.cap;
            unsigned* keys = ctx->binTmp.as<unsigned>();
            unsigned* ord = keys + slots;
            unsigned* hist = ord + slots;
            queueBinKeysKernel<<<grid, 256, 0, s>>>(TV, rb.st, depth, rb.gq, keys, hist);
            queueBinScanKernel<<<1, 1024, 0, s>>>(hist, TV.numRefNodes * 8);
            queueBinScatterKernel<<<grid, 256, 0, s>>>(keys, rb.st, depth, hist, ord);
            launches += 3;
        };
        if (p.mode == KDPT_FULLSTACK) {
            if (gen) frontKernel<KDPT_FULLSTACK, true><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, *gen, rb.gq);
            else frontKernel<KDPT_FULLSTACK, false><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, noGen, rb.gq);
            binQueue();
            if (TV.wide) traceQueueKernel<false, true><<<lgrid, kLeanThreads, 0, s>>>(TV, rb.st, depth, rb.gq, A.eps, order);
            else traceQueueKernel<false, false><<<lgrid, kLeanThreads, 0, s>>>(TV, rb.st, depth, rb.gq, A.eps, order);
            finishQueueKernel<KDPT_FULLSTACK><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, rb.gq);
        } else {
            if (gen) frontKernel<KDPT_SHORTSTACK, true><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, *gen, rb.gq);
            else frontKernel<KDPT_SHORTSTACK, false><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, noGen, rb.gq);
            binQueue();
            if (TV.wide) traceQueueKernel<true, true><<<lgrid, kLeanThreads, 0, s>>>(TV, rb.st, depth, rb.gq, A.eps, order);
            else traceQueueKernel<true, false><<<lgrid, kLeanThreads, 0, s>>>(TV, rb.st, depth, rb.gq, A.eps, order);
            finishQueueKernel<KDPT_SHORTSTACK><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, rb.gq);
        }
        launches += 3;
        KDPT_CUDA(cudaGetLastError());
        return KDPT_OK;
    }
#define KDPT_LAUNCH_BOUNCE(M)                                                                                                   \
    if (gen) bounceKernel<M, true><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, dIsects, *gen);   \
    else bounceKernel<M, false><<<grid, kThreads, 0, s>>>(TV, S, P, rb.st, A, bt, rb.accum, pc, pa, dIsects, noGen)
    switch (p.mode) {
    case KDPT_FULLSTACK: KDPT_LAUNCH_BOUNCE(KDPT_FULLSTACK); break;
    case KDPT_SHORTSTACK: KDPT_LAUNCH_BOUNCE(KDPT_SHORTSTACK); break;
    case KDPT_REF_T1: KDPT_LAUNCH_BOUNCE(KDPT_REF_T1); break;
    case KDPT_REF_T2: KDPT_LAUNCH_BOUNCE(KDPT_REF_T2); break;
    case KDPT_BRUTE: KDPT_LAUNCH_BOUNCE(KDPT_BRUTE); break;
    case KDPT_KDVIZ: {      // no triangles at all: the brute-force instantiation over an empty triangle list
        Tree none = ctx->tree; none.numNodes = 0; none.numTris = 0; none.numRefNodes = 0;
        if (gen) bounceKernel<KDPT_BRUTE, true><<<grid, kThreads, 0, s>>>(none, S, P, rb.st, A, bt, rb.accum, pc, pa, dIsects, *gen);
        else bounceKernel<KDPT_BRUTE, false><<<grid, kThreads, 0, s>>>(none, S, P, rb.st, A, bt, rb.accum, pc, pa, dIsects, noGen);
        break;
    }
    default: KDPT_LAUNCH_BOUNCE(KDPT_BRUTE_BBOX); break;
    }
#undef KDPT_LAUNCH_BOUNCE
    launches++;
    KDPT_CUDA(cudaGetLastError());
    return KDPT_OK;
}

} // namespace kdpt

using namespace kdpt;

extern "C" {

int kdpt_clear_image(kdpt_ctx* ctx)
{
    if (!ctx) { setError("kdpt_clear_image: null context"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc = ensureImage(ctx);
    if (rc) return rc;
    KDPT_CUDA(cudaMemsetAsync(ctx->boundImage ? (void*)ctx->boundImage : ctx->image.p, 0, sizeof(float) * 3 * (size_t)ctx->pixelCount, ctx->stream));
    return KDPT_OK;
}

int kdpt_set_camera(kdpt_ctx* ctx, const kdpt_camera* camera)
{
    if (!ctx || !camera) { setError("kdpt_set_camera: null argument"); return KDPT_ERR_ARG; }
    if (camera->resolution[0] <= 0 || camera->resolution[1] <= 0) { setError("kdpt_set_camera: resolution must be positive"); return KDPT_ERR_ARG; }
    const long long n = (long long)camera->resolution[0] * camera->resolution[1];
    if (n >= (1ll << kSegShift)) { setError("kdpt_set_camera: more than 2^26 pixels"); return KDPT_ERR_ARG; }
    if (ctx->boundImage && ctx->scene.hasCamera && (int)n != ctx->pixelCount) {
        setError("kdpt_set_camera: a caller-owned radiance buffer is bound; unbind it (kdpt_bind_image NULL) before changing the resolution");
        return KDPT_ERR_STATE;
    }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));      // launches in flight took the old camera by value; the image may be resized
    ctx->scene.camera = *camera;
    ctx->scene.hasCamera = true;
    return kdpt_clear_image(ctx);                       // pathtraceInit zeroes dev_image (pathtrace.cu:262-263); resizes if needed
}

int kdpt_get_image(kdpt_ctx* ctx, float* rgb)
{
    if (!ctx || !rgb) { setError("kdpt_get_image: null argument"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc = ensureImage(ctx);
    if (rc) return rc;
    KDPT_CUDA(cudaMemcpyAsync(rgb, ctx->boundImage ? (void*)ctx->boundImage : ctx->image.p, sizeof(float) * 3 * (size_t)ctx->pixelCount, cudaMemcpyDeviceToHost, ctx->stream));
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    return KDPT_OK;
}

int kdpt_set_image(kdpt_ctx* ctx, const float* rgb)
{
    if (!ctx || !rgb) { setError("kdpt_set_image: null argument"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc = ensureImage(ctx);
    if (rc) return rc;
    KDPT_CUDA(cudaMemcpyAsync(ctx->boundImage ? (void*)ctx->boundImage : ctx->image.p, rgb, sizeof(float) * 3 * (size_t)ctx->pixelCount, cudaMemcpyHostToDevice, ctx->stream));
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    return KDPT_OK;
}

int kdpt_bind_image(kdpt_ctx* ctx, float* d_rgb)
{
    if (!ctx) { setError("kdpt_bind_image: null context"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc = ensureImage(ctx);
    if (rc) return rc;
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->boundImage = d_rgb;
    return KDPT_OK;
}

int kdpt_counters(kdpt_ctx* ctx, uint64_t* rayBounces, uint64_t* launches)
{
    if (!ctx) { setError("kdpt_counters: null context"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    if (launches) *launches = ctx->launches;
    if (rayBounces) {
        *rayBounces = 0;
        if (ctx->scanTmp.p) {
            unsigned long long c[kMaxBounces];
            KDPT_CUDA(cudaMemcpyAsync(c, ctx->scanTmp.as<LoopState>()->bounceRays, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
            KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int b = 0; b < kMaxBounces; b++) *rayBounces += c[b];
        }
    }
    return KDPT_OK;
}

int kdpt_image_device(kdpt_ctx* ctx, float** d_rgb, int* numFloats)
{
    if (!ctx || !d_rgb) { setError("kdpt_image_device: null argument"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc = ensureImage(ctx);
    if (rc) return rc;
    KDPT_CUDA(cudaStreamSynchronize(ctx->stream));
    *d_rgb = ctx->boundImage ? ctx->boundImage : ctx->image.as<float>();
    if (numFloats) *numFloats = 3 * ctx->pixelCount;
    return KDPT_OK;
}

int kdpt_allreduce_image(kdpt_ctx** ctxs, int n)
{
    if (!ctxs || n < 1 || n > 16) { setError("kdpt_allreduce_image: need 1..16 contexts"); return KDPT_ERR_ARG; }
    for (int i = 0; i < n; i++) if (!ctxs[i]) { setError("kdpt_allreduce_image: null context"); return KDPT_ERR_ARG; }
    PeerImages pi; memset(&pi, 0, sizeof pi); pi.n = n;
    int rc;
    for (int i = 0; i < n; i++) {
        KDPT_CUDA(cudaSetDevice(ctxs[i]->device));
        if ((rc = ensureImage(ctxs[i]))) return rc;
        if (ctxs[i]->pixelCount != ctxs[0]->pixelCount) { setError("kdpt_allreduce_image: resolutions differ"); return KDPT_ERR_ARG; }
        for (int j = 0; j < i; j++) if (ctxs[j]->device == ctxs[i]->device) { setError("kdpt_allreduce_image: two contexts on one device"); return KDPT_ERR_ARG; }
        pi.img[i] = ctxs[i]->boundImage ? ctxs[i]->boundImage : ctxs[i]->image.as<float>();
    }
    if (n == 1) return KDPT_OK;
    for (int i = 0; i < n; i++) {
        KDPT_CUDA(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int can = 0;
            KDPT_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device));
            if (!can) { setError("kdpt_allreduce_image: no peer access between the devices"); return KDPT_ERR_STATE; }
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return cudaFail(e, "cudaDeviceEnablePeerAccess");
            // the contexts' own radiance buffers come from the stream-ordered pool (DeviceBuffer), whose memory is
            // visible to a peer only after the pool says so (cudaDeviceEnablePeerAccess covers cudaMalloc memory)
            cudaMemPool_t pool;
            KDPT_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctxs[j]->device));
            cudaMemAccessDesc desc; memset(&desc, 0, sizeof desc);
            desc.location.type = cudaMemLocationTypeDevice; desc.location.id = ctxs[i]->device;
            desc.flags = cudaMemAccessFlagsProtReadWrite;
            KDPT_CUDA(cudaMemPoolSetAccess(pool, &desc, 1));
        }
    }
    const size_t total = 3 * (size_t)ctxs[0]->pixelCount;
    size_t per = (total + n - 1) / n;
    per = (per + 3) & ~(size_t)3;                  // slices start on float4 boundaries
    // every GPU has finished rendering before anyone reads its buffer
    for (int i = 0; i < n; i++) { KDPT_CUDA(cudaSetDevice(ctxs[i]->device)); KDPT_CUDA(cudaEventRecord(ctxs[i]->ev2, ctxs[i]->stream)); }
    for (int i = 0; i < n; i++) {
        KDPT_CUDA(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; j++) if (j != i) KDPT_CUDA(cudaStreamWaitEvent(ctxs[i]->stream, ctxs[j]->ev2, 0));
        const size_t lo = per * i < total ? per * i : total, hi = per * (i + 1) < total ? per * (i + 1) : total;
        if (hi > lo) peerReduceSliceKernel<<<ctxs[i]->numSMs * 2, 256, 0, ctxs[i]->stream>>>(pi, i, lo, hi);
        KDPT_CUDA(cudaGetLastError());
        KDPT_CUDA(cudaEventRecord(ctxs[i]->ev3, ctxs[i]->stream));
        ctxs[i]->launches++;
    }
    // every slice is final before anyone gathers it
    for (int i = 0; i < n; i++) {
        KDPT_CUDA(cudaSetDevice(ctxs[i]->device));
        for (int j = 0; j < n; j++) if (j != i) KDPT_CUDA(cudaStreamWaitEvent(ctxs[i]->stream, ctxs[j]->ev3, 0));
        peerGatherKernel<<<ctxs[i]->numSMs * 2, 256, 0, ctxs[i]->stream>>>(pi, i, per, total);
        KDPT_CUDA(cudaGetLastError());
        ctxs[i]->launches++;
    }
    // nobody's buffer changes again (next render) before all gathers have read it
    for (int i = 0; i < n; i++) { KDPT_CUDA(cudaSetDevice(ctxs[i]->device)); KDPT_CUDA(cudaStreamSynchronize(ctxs[i]->stream)); }
    return KDPT_OK;
}

int kdpt_render(kdpt_ctx* ctx, int iter_begin, int iter_count, int iter_stride, const kdpt_params* params, kdpt_render_stats* stats)
{
    if (!ctx || iter_count < 0 || iter_stride < 1) { setError("kdpt_render: bad argument"); return KDPT_ERR_ARG; }
    int rc = checkParams(params);
    if (rc) return rc;
    KDPT_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensureImage(ctx))) return rc;
    const kdpt_params& p = *params;
    const int n = ctx->pixelCount;
    if (n >= (1 << kSegShift)) { setError("kdpt_render: more than 2^26 pixels"); return KDPT_ERR_ARG; }
    // iterations per launch: bounded by kMaxBatch and by 16 GB of path state + ray queue (of 180 GB)
    int maxBatch = kMaxBatch;
    while (maxBatch > 1 && (size_t)maxBatch * n * 256 > ((size_t)16 << 30)) maxBatch >>= 1;
    if (iter_count < maxBatch) maxBatch = iter_count > 0 ? iter_count : 1;
    RenderBuffers rb;
    if ((rc = ensureRenderBuffers(ctx, n, maxBatch, rb, true, useWavefront(ctx, p.mode)))) return rc;
    cudaStream_t s = ctx->stream;
    uint64_t launches = 0;
    unsigned long long before[kMaxBounces], after[kMaxBounces];
    if (stats) KDPT_CUDA(cudaMemcpyAsync(before, rb.st->bounceRays, sizeof before, cudaMemcpyDeviceToHost, s));
    if (stats) KDPT_CUDA(cudaEventRecord(ctx->ev0, s));
    const bool timing = stats && p.testingmode;
    size_t evUsed = 0;
    auto nextEvent = [&](cudaEvent_t& e) -> int {
        if (evUsed == ctx->eventPool.size()) { cudaEvent_t ne; KDPT_CUDA(cudaEventCreate(&ne)); ctx->eventPool.push_back(ne); }
        e = ctx->eventPool[evUsed++];
        return KDPT_OK;
    };
    const int grid = ctx->numSMs * 8;
    int sortPasses = 1;
    for (int m = ctx->scene.numMaterials + 1; m > 256; m >>= 8) sortPasses++;
    // remainingBounces starts at traceDepth and drops every bounce, so nothing survives min(8, traceDepth) bounces
    const int maxBounces = ctx->scene.traceDepth < kMaxBounces ? (ctx->scene.traceDepth < 0 ? 0 : ctx->scene.traceDepth) : kMaxBounces;
    for (int k = 0; k < iter_count;) {
        // a batch is sorted as a whole or not at all: the reference sorts only in iteration 2 (pathtrace.cu:2600), so
        // with sortByMaterial == 1 that iteration gets a batch of its own
        auto sorts = [&](int kk) { int it = iter_begin + kk * iter_stride; return p.sortByMaterial == 2 || (p.sortByMaterial == 1 && it == 2); };
        int B = 1;
        const bool sortThis = sorts(k);
        while (B < maxBatch && k + B < iter_count && sorts(k + B) == sortThis && (p.sortByMaterial != 1 || !sortThis)) B++;
        Batch bt; bt.B = B; bt.cap = rb.cap; bt.capPackets = rb.capPackets; bt.iterBase = iter_begin + k * iter_stride;
        bt.iterStride = iter_stride; bt.npix = n;
        GenArgs gen; memset(&gen, 0, sizeof gen);
        gen.cam = ctx->scene.camera; gen.traceDepth = ctx->scene.traceDepth; gen.focal = p.focalLength;
        // rayCaching: every iteration starts from iteration 1's camera rays (pathtrace.cu:2448-2456)
        for (int q = 0; q < B; q++) { gen.ics.c[q] = makeIterConst(p.cacheRays ? 1 : bt.iterBase + q * iter_stride, p); gen.ics.c[q].iter = bt.iterBase + q * iter_stride; }
        if (maxBounces > 0) {
            beginBatchKernel<<<1, 64, 0, s>>>(rb.st, B, n);     // the first bounce kernel generates the camera rays itself
        } else {                                                // DEPTH <= 0: nothing is traced
            generateRaysKernel<<<(B * n + 255) / 256, 256, 0, s>>>(ctx->scene.camera, gen.ics, bt, ctx->scene.traceDepth, p.focalLength, rb.A, rb.st);
            KDPT_CUDA(cudaMemsetAsync(rb.accum, 0, sizeof(float) * 3 * (size_t)n * B, s));
        }
        launches++;
        PathSoA cur = rb.A, other = rb.B;
        for (int depth = 0; depth < maxBounces; depth++) {
            int stages = KDPT_STAGE_INTERSECT | KDPT_STAGE_SHADE | (p.compaction ? 4 : 0) | ((p.compaction && depth == maxBounces - 1) ? 8 : 0);
            cudaEvent_t t0 = nullptr, t1 = nullptr;
            if (timing) { if ((rc = nextEvent(t0)) || (rc = nextEvent(t1))) return rc; KDPT_CUDA(cudaEventRecord(t0, s)); }
            if ((rc = launchBounce(ctx, p, rb, cur, bt, depth, stages, nullptr, launches, depth == 0 ? &gen : nullptr))) return rc;
            if (timing) KDPT_CUDA(cudaEventRecord(t1, s));
            scanPacketsKernel<<<B, 1024, 0, s>>>(rb.packetCount, rb.st, depth, p.compaction, rb.capPackets);
            launches++;
            if (p.compaction) {
                compactScatterKernel<<<grid, kThreads, 0, s>>>(cur, other, rb.packetCount, rb.packetAlive, rb.st, depth, bt);
                launches++;
                PathSoA t = cur; cur = other; other = t;
            }
            if (sortThis) {
                for (int pass = 0; pass < sortPasses; pass++) {
                    sortHistKernel<<<grid, kSortTile, 0, s>>>(cur, rb.st, depth + 1, 8 * pass, rb.hist, rb.maxTiles, bt);
                    sortScanKernel<<<B, 1024, 0, s>>>(rb.st, depth + 1, rb.hist, rb.maxTiles);
                    sortScatterKernel<<<grid, kSortTile, 0, s>>>(cur, other, rb.st, depth + 1, 8 * pass, rb.hist, rb.maxTiles, bt);
                    launches += 3;
                    PathSoA t = cur; cur = other; other = t;
                }
            }
        }
        if (!p.compaction) {
            finalGatherKernel<<<grid, 256, 0, s>>>(cur, rb.st, maxBounces, bt, rb.accum);
            launches++;
        }
        accumulateKernel<<<grid, 256, 0, s>>>(rb.image, rb.accum, B, 3 * n);
        launches++;
        KDPT_CUDA(cudaGetLastError());
        k += B;
    }
    ctx->launches += launches;
    if (stats) {
        KDPT_CUDA(cudaEventRecord(ctx->ev1, s));
        KDPT_CUDA(cudaMemcpyAsync(after, rb.st->bounceRays, sizeof after, cudaMemcpyDeviceToHost, s));
        KDPT_CUDA(cudaStreamSynchronize(s));
        memset(stats, 0, sizeof *stats);
        for (int b = 0; b < kMaxBounces; b++) stats->rayBounces += after[b] - before[b];
        if (getenv("KDPT_DEBUG_STATS")) {
            unsigned long long tr[kMaxBounces];
            KDPT_CUDA(cudaMemcpy(tr, rb.st->travRays, sizeof tr, cudaMemcpyDeviceToHost));
            for (int b = 0; b < kMaxBounces; b++) fprintf(stderr, "kdpt stats: bounce %d rays %llu (cumulative) entering the tree %llu (%.3f)\n", b, after[b], tr[b], after[b] ? (double)tr[b] / after[b] : 0.0);
        }
        stats->launches = launches;
        KDPT_CUDA(cudaEventElapsedTime(&stats->totalMs, ctx->ev0, ctx->ev1));
        stats->intersectMs = 0.0f;
        for (size_t k = 0; k + 1 < evUsed; k += 2) {
            float ms = 0.f;
            KDPT_CUDA(cudaEventElapsedTime(&ms, ctx->eventPool[k], ctx->eventPool[k + 1]));
            stats->intersectMs += ms;
            stats->intersectLaunches++;
        }
    }
    return KDPT_OK;
}

int kdpt_generate_rays(kdpt_ctx* ctx, int iter, const kdpt_params* params, kdpt_path_segment* out)
{
    if (!ctx || !out) { setError("kdpt_generate_rays: null argument"); return KDPT_ERR_ARG; }
    int rc = checkParams(params);
    if (rc) return rc;
    KDPT_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensureImage(ctx))) return rc;
    const int n = ctx->pixelCount;
    RenderBuffers rb;
    if ((rc = ensureRenderBuffers(ctx, n, 1, rb, false))) return rc;
    if ((rc = ctx->isects.ensure(sizeof(kdpt_path_segment) * (size_t)n))) return rc;
    cudaStream_t s = ctx->stream;
    IterConsts ics; ics.c[0] = makeIterConst(iter, *params);
    Batch bt; bt.B = 1; bt.cap = rb.cap; bt.capPackets = rb.capPackets; bt.iterBase = iter; bt.iterStride = 1; bt.npix = n;
    generateRaysKernel<<<(n + 255) / 256, 256, 0, s>>>(ctx->scene.camera, ics, bt, ctx->scene.traceDepth, params->focalLength, rb.A, rb.st);
    unpackPathsKernel<<<(n + 255) / 256, 256, 0, s>>>(rb.A, n, ctx->isects.as<kdpt_path_segment>());
    KDPT_CUDA(cudaGetLastError());
    KDPT_CUDA(cudaMemcpyAsync(out, ctx->isects.p, sizeof(kdpt_path_segment) * (size_t)n, cudaMemcpyDeviceToHost, s));
    KDPT_CUDA(cudaStreamSynchronize(s));
    return KDPT_OK;
}

// thrust::remove_if(paths, is_zero_bounce) (pathtrace.cu:2591-2597) / thrust::sort by materialIdHit (2600-2606) on a
// caller-owned HOST array, through the render loop's own kernels
int kdpt_compact_paths(kdpt_ctx* ctx, kdpt_path_segment* paths, int n, int* n_out)
{
    if (!ctx || n < 0 || (n > 0 && !paths) || !n_out) { setError("kdpt_compact_paths: bad argument"); return KDPT_ERR_ARG; }
    if (n >= (1 << kSegShift)) { setError("kdpt_compact_paths: more than 2^26 paths"); return KDPT_ERR_ARG; }
    *n_out = 0;
    if (n == 0) return KDPT_OK;
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc;
    RenderBuffers rb;
    if ((rc = ensureRenderBuffers(ctx, n, 1, rb, false))) return rc;
    Batch bt; bt.B = 1; bt.cap = rb.cap; bt.capPackets = rb.capPackets; bt.iterBase = 1; bt.iterStride = 1; bt.npix = n;
    const size_t pb = sizeof(kdpt_path_segment) * (size_t)n;
    if ((rc = ctx->isects.ensure(pb))) return rc;
    kdpt_path_segment* dPaths = ctx->isects.as<kdpt_path_segment>();
    cudaStream_t s = ctx->stream;
    const int grid = ctx->numSMs * 8;
    KDPT_CUDA(cudaMemcpyAsync(dPaths, paths, pb, cudaMemcpyHostToDevice, s));
    packPathsKernel<<<(n + 255) / 256, 256, 0, s>>>(dPaths, n, rb.A, rb.st, 0);
    aliveMaskKernel<<<grid, kThreads, 0, s>>>(rb.A, n, rb.packetCount, rb.packetAlive);
    scanPacketsKernel<<<1, 1024, 0, s>>>(rb.packetCount, rb.st, 0, 1, rb.capPackets);
    compactScatterKernel<<<grid, kThreads, 0, s>>>(rb.A, rb.B, rb.packetCount, rb.packetAlive, rb.st, 0, bt);
    int m = 0;
    KDPT_CUDA(cudaMemcpyAsync(&m, &rb.st->np[0][1], sizeof(int), cudaMemcpyDeviceToHost, s));
    KDPT_CUDA(cudaStreamSynchronize(s));
    if (m > 0) {
        unpackPathsKernel<<<(m + 255) / 256, 256, 0, s>>>(rb.B, m, dPaths);
        KDPT_CUDA(cudaMemcpyAsync(paths, dPaths, sizeof(kdpt_path_segment) * (size_t)m, cudaMemcpyDeviceToHost, s));
    }
    KDPT_CUDA(cudaGetLastError());
    KDPT_CUDA(cudaStreamSynchronize(s));
    ctx->launches += 5;
    *n_out = m;
    return KDPT_OK;
}

int kdpt_sort_paths(kdpt_ctx* ctx, kdpt_path_segment* paths, int n)
{
    if (!ctx || n < 0 || (n > 0 && !paths)) { setError("kdpt_sort_paths: bad argument"); return KDPT_ERR_ARG; }
    if (n >= (1 << kSegShift)) { setError("kdpt_sort_paths: more than 2^26 paths"); return KDPT_ERR_ARG; }
    if (n == 0) return KDPT_OK;
    KDPT_CUDA(cudaSetDevice(ctx->device));
    int rc;
    RenderBuffers rb;
    if ((rc = ensureRenderBuffers(ctx, n, 1, rb, false))) return rc;
    Batch bt; bt.B = 1; bt.cap = rb.cap; bt.capPackets = rb.capPackets; bt.iterBase = 1; bt.iterStride = 1; bt.npix = n;
    const size_t pb = sizeof(kdpt_path_segment) * (size_t)n;
    if ((rc = ctx->isects.ensure(pb))) return rc;
    kdpt_path_segment* dPaths = ctx->isects.as<kdpt_path_segment>();
    cudaStream_t s = ctx->stream;
    const int grid = ctx->numSMs * 8;
    KDPT_CUDA(cudaMemcpyAsync(dPaths, paths, pb, cudaMemcpyHostToDevice, s));
    packPathsKernel<<<(n + 255) / 256, 256, 0, s>>>(dPaths, n, rb.A, rb.st, 0);
    // all four key bytes: the caller's materialIdHit values are arbitrary ints here (the render loop knows the material count)
    PathSoA cur = rb.A, other = rb.B;
    for (int pass = 0; pass < 4; pass++) {
        sortHistKernel<<<grid, kSortTile, 0, s>>>(cur, rb.st, 0, 8 * pass, rb.hist, rb.maxTiles, bt);
        sortScanKernel<<<1, 1024, 0, s>>>(rb.st, 0, rb.hist, rb.maxTiles);
        sortScatterKernel<<<grid, kSortTile, 0, s>>>(cur, other, rb.st, 0, 8 * pass, rb.hist, rb.maxTiles, bt);
        PathSoA t = cur; cur = other; other = t;
    }
    unpackPathsKernel<<<(n + 255) / 256, 256, 0, s>>>(cur, n, dPaths);
    KDPT_CUDA(cudaGetLastError());
    KDPT_CUDA(cudaMemcpyAsync(paths, dPaths, pb, cudaMemcpyDeviceToHost, s));
    KDPT_CUDA(cudaStreamSynchronize(s));
    ctx->launches += 14;
    return KDPT_OK;
}

int kdpt_bounce(kdpt_ctx* ctx, int iter, int depth, const kdpt_params* params, kdpt_path_segment* paths, int n,
                kdpt_intersection* isects, int stages)
{
    if (!ctx || n < 0 || (n > 0 && !paths) || depth < 0 || depth >= kMaxBounces || (stages & ~3) || !stages) {
        setError("kdpt_bounce: bad argument");
        return KDPT_ERR_ARG;
    }
    if (stages == KDPT_STAGE_SHADE && !isects) { setError("kdpt_bounce: the shade stage alone needs the intersections"); return KDPT_ERR_ARG; }
    if (n >= (1 << kSegShift)) { setError("kdpt_bounce: more than 2^26 paths"); return KDPT_ERR_ARG; }
    int rc = checkParams(params);
    if (rc) return rc;
    if (n == 0) return KDPT_OK;
    KDPT_CUDA(cudaSetDevice(ctx->device));
    RenderBuffers rb;
    if ((rc = ensureRenderBuffers(ctx, n, 1, rb, false))) return rc;
    Batch bt; bt.B = 1; bt.cap = rb.cap; bt.capPackets = rb.capPackets; bt.iterBase = iter; bt.iterStride = 1; bt.npix = n;
    size_t pb = sizeof(kdpt_path_segment) * (size_t)n, ib = sizeof(kdpt_intersection) * (size_t)n;
    if ((rc = ctx->isects.ensure(pb + ib))) return rc;
    kdpt_path_segment* dPaths = ctx->isects.as<kdpt_path_segment>();
    kdpt_intersection* dIs = (kdpt_intersection*)((char*)ctx->isects.p + pb);
    cudaStream_t s = ctx->stream;
    KDPT_CUDA(cudaMemcpyAsync(dPaths, paths, pb, cudaMemcpyHostToDevice, s));
    if (isects && stages == KDPT_STAGE_SHADE) KDPT_CUDA(cudaMemcpyAsync(dIs, isects, ib, cudaMemcpyHostToDevice, s));
    else KDPT_CUDA(cudaMemsetAsync(dIs, 0, ib, s));
    packPathsKernel<<<(n + 255) / 256, 256, 0, s>>>(dPaths, n, rb.A, rb.st, depth);
    kdpt_params p = *params; p.compaction = 0;
    uint64_t launches = 0;
    if ((rc = launchBounce(ctx, p, rb, rb.A, bt, depth, stages, dIs, launches, nullptr))) return rc;
    unpackPathsKernel<<<(n + 255) / 256, 256, 0, s>>>(rb.A, n, dPaths);
    KDPT_CUDA(cudaGetLastError());
    KDPT_CUDA(cudaMemcpyAsync(paths, dPaths, pb, cudaMemcpyDeviceToHost, s));
    if (isects) KDPT_CUDA(cudaMemcpyAsync(isects, dIs, ib, cudaMemcpyDeviceToHost, s));
    KDPT_CUDA(cudaStreamSynchronize(s));
    return KDPT_OK;
}

} // extern "C"
