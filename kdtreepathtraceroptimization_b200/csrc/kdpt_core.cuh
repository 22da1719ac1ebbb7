// kdpt_core.cuh — device data layout and the per-ray arithmetic of the hot path.
//
// Everything that decides a hit is written as separate IEEE-754 binary32 multiplies and adds in the reference's
// (glm 0.9.6.3) evaluation order; the translation unit is compiled with -fmad=false, default -prec-div/-prec-sqrt
// and no -use_fast_math, so the device results are bit-identical to the reference's host functions:
//   glm::intersectRayTriangle     external/include/glm/gtx/intersect.inl:37-71
//   intersectAABBarrays           src/intersections.h:253-286
//   leaf body (normal, nudge, t)  src/pathtrace.cu:985-1010 / 1133-1161
//   glm dot/cross/normalize       external/include/glm/detail/func_geometric.inl:64-72, 134-142, 158
//
// The functions are KDPT_HD (device-only under nvcc) so that tests/hostsim can compile the very same source with g++ for CPU-only logic
// tests; the product library only ever calls them from __global__ kernels.
#pragma once
#include <stdint.h>
#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define KDPT_HD __device__ __forceinline__
#define KDPT_NOINLINE_FN static __device__ __noinline__
#define KDPT_LDG(p) __ldg(p)
#include <cuda_runtime.h>
#else
#define KDPT_HD inline
#define KDPT_NOINLINE_FN static inline
#define KDPT_LDG(p) (*(p))
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline float __uint_as_float(unsigned i) { float f; __builtin_memcpy(&f, &i, 4); return f; }
static inline float __int_as_float(int i) { float f; __builtin_memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; __builtin_memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; __builtin_memcpy(&i, &f, 4); return i; }
#endif

#ifndef KDPT_CHUNK_PAIR
#define KDPT_CHUNK_PAIR 0     // 1: stream two chunks of a fat leaf per step (more loads in flight, ~10 more registers)
#endif

namespace kdpt {

constexpr int kMaxDepth = 32;           // deepest tree the per-ray stacks accept (reference: split(13) -> 14)
constexpr float kExitSlack = 1e-3f;     // early-exit slack in world units (acceptance is on the nudged distance)
constexpr unsigned kLeafFlag = 3u;
constexpr unsigned kNoChunks = 0xffffffffu;

// ---- flattened tree in HBM -------------------------------------------------------------------------------------
// node (16 B, index == reference node ID, IDs are DFS pre-order):
//   interior: x = split plane (float bits) = nodes[left].maxs[axis]   y = right child   z = left child
//             w = axis (bits 0-1) | (parent + 1) << 2
//   leaf:     x = first leaf triangle            y = triangle count (may be 0)          z = first chunk or ~0u
//             w = 3 | (parent + 1) << 2
// leaf triangles keep the reference's leaf order (TriBare index i), split into a hot and a cold part:
//   triA[i] = (v0.x, v0.y, v0.z, e1.x)  triB[i] = (e1.y, e1.z, e2.x, e2.y)  triC[i] = (e2.z, canonical index)
//   nrmA[i] = (n1.x, n1.y, n1.z, n2.x)  nrmB[i] = (n2.y, n2.z, n3.x, n3.y)  nrmC[i] = (n3.z, mtlIdx)
//   with e1 = v1 - v0, e2 = v2 - v0 (the first two operations of glm::intersectRayTriangle, done once).
// boxes (reference-compatible modes only): boxLo[i] = (mins, 0), boxHi[i] = (maxs, NodeBare.axis as int bits).
//
// Two views of one scene share this struct. The REFERENCE view is the tree exactly as the reference built it (the
// reference-compatible modes depend on its leaves and their order). The REFINED view (exact modes) has every fat
// reference leaf split further (kdpt_flatten.cpp:refineLeaves): nodes / triA / triB / triC / chunk* point at the
// refined arrays, whose leaf entries carry the canonical TriBare index; the cold per-triangle data (normals,
// material) stays in the reference view's arrays and is reached through entryOf[canonical index].
struct Tree {
    const uint4* nodes;
    const float4* boxLo;
    const float4* boxHi;
    const float4* triA;
    const float4* triB;
    const float2* triC;
    const float4* nrmA;
    const float4* nrmB;
    const float2* nrmC;
    const int* triOrig;      // TriBare index of entry i (entries of chunked leaves are Morton-reordered)
    const float4* chunkLo;   // padded AABBs of the 32-entry chunks of chunked leaves (leaf record .z = first chunk)
    const float4* chunkHi;
    const float4* superLo;   // boxes of groups of 32 chunks (leaves with >= 64 chunks); chunkLo[first].w = first group
    const float4* superHi;
    const int* entryOf;      // refined view only (else NULL): TriBare index -> entry of nrmA/nrmB/nrmC (see below)
    const uint4* top;        // refined view only: breadth-first copy of the top levels for shared-memory staging (numTop nodes;
    int numTop;              // children inside the copy are named by position | 0x80000000), see kdpt_lean.cuh
    const uint4* wide;       // refined view only (else NULL): the same tree, two levels per 32-byte record (kdpt_lean.cuh)
    unsigned wideRoot;       // reference of the root in that form
    int numNodes;
    int numTris;
    int numRefNodes;         // nodes[i], i < numRefNodes, are the reference's nodes (both views)
    float rootLo[3], rootHi[3];
    int tagCand[4];     // nodes[0].leftID, nodes[0].rightID, nodes[1].leftID, nodes[1].rightID (T2 quirk)
    int materialSize;   // hst_scene->materials.size()
    int laneCost32;     // cost of one lane-private triangle step relative to a cooperative one, x32 (see traceExactWarp)
};

struct Hit {
    float t;            // nudged distance (reference t_min)
    float z, u, v;
    int tri;            // canonical TriBare index, -1 = none
    float nx, ny, nz;
    int mtl;            // mtlIdx of the accepted triangle
    int anyMtl;         // mtlIdx of the LAST triangle glm accepted (reference objMaterialIdx source), -1 none
};

struct V3 { float x, y, z; };
KDPT_HD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
KDPT_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
KDPT_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
KDPT_HD V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
KDPT_HD V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
KDPT_HD V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
KDPT_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
KDPT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
KDPT_HD V3 cross(V3 x, V3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
KDPT_HD float length(V3 a) { return sqrtf(dot(a, a)); }
KDPT_HD V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
KDPT_HD float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// glm::intersectRayTriangle with e1/e2 precomputed. Returns 0 = rejected before baryPosition.z was written,
// 1 = z written but negative (glm returns false), 2 = hit (glm returns true). u, v, z are only meaningful as
// far as glm wrote them.
KDPT_HD int rayTri(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float& u, float& v, float& z)
{
    V3 p = cross(d, e2);
    float a = dot(e1, p);
    if (a < FLT_EPSILON) return 0;
    float f = 1.0f / a;
    V3 s = o - v0;
    u = f * dot(s, p);
    if (u < 0.0f) return 0;
    if (u > 1.0f) return 0;
    V3 q = cross(s, e1);
    v = f * dot(d, q);
    if (v < 0.0f) return 0;
    if (v + u > 1.0f) return 0;
    z = f * dot(e2, q);
    return z >= 0.0f ? 2 : 1;
}

// The same predicate for the kernels' hot loops: returns exactly what rayTri returns as far as hit / no hit goes, but
// skips the IEEE division for triangles that certainly fail a barycentric test. With f = fl(1/a) > 0:
//   su < -a*2^-120            =>  fl(f*su) < 0 (the product cannot underflow to -0)         => u < 0, reject
//   su >  a*(1 + 2^-20)       =>  fl(f*su) > 1 (three roundings of 2^-24 each cannot undo it) => u > 1, reject
// and likewise for v and for u + v (margin 2^-18). Everything that survives the filter is evaluated with the exact
// sequence above, so accepted hits and their u, v, z are bit-identical. (Only the exact-hit kernels use it: the
// reference-compatible ones also need the u/v/z glm leaves behind on rejected triangles.)
KDPT_HD bool rayTriHit(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float& u, float& v, float& z)
{
    V3 p = cross(d, e2);
    float a = dot(e1, p);
    if (a < FLT_EPSILON) return false;
    V3 s = o - v0;
    float su = dot(s, p);
    const float tiny = a * 7.52316385e-37f;            // a * 2^-120 (exact scaling, a >= 2^-23 so never zero)
    const float aHi = a * 1.00000095367431640625f;    // a * (1 + 2^-20)
    if (su < -tiny || su > aHi) return false;
    V3 q = cross(s, e1);
    float dq = dot(d, q);
    if (dq < -tiny || dq > aHi || su + dq > a * 1.000003814697265625f) return false;   // 1 + 2^-18
    float f = 1.0f / a;
    u = f * su;
    if (u < 0.0f) return false;
    if (u > 1.0f) return false;
    v = f * dq;
    if (v < 0.0f) return false;
    if (v + u > 1.0f) return false;
    z = f * dot(e2, q);
    return z >= 0.0f;
}

// intersectAABBarrays
KDPT_HD bool slab(V3 o, V3 inv, const float* lo, const float* hi, float& dmin, float& dmax)
{
    float v1 = (lo[0] - o.x) * inv.x, v2 = (hi[0] - o.x) * inv.x;
    float v3 = (lo[1] - o.y) * inv.y, v4 = (hi[1] - o.y) * inv.y;
    float v5 = (lo[2] - o.z) * inv.z, v6 = (hi[2] - o.z) * inv.z;
    dmin = fmaxf(fmaxf(fminf(v1, v2), fminf(v3, v4)), fminf(v5, v6));
    dmax = fminf(fminf(fmaxf(v1, v2), fmaxf(v3, v4)), fmaxf(v5, v6));
    return !(dmax < 0.0f) && !(dmin > dmax);
}
KDPT_HD bool refBox(V3 o, V3 inv, float4 lo, float4 hi, float& dist)
{
    float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
    float dmin, dmax;
    // the reference tests dmax < 0 first, then dmin > dmax; both misses report dmax, a hit reports dmin
    bool ok = slab(o, inv, l, h, dmin, dmax);
    dist = ok ? dmin : dmax;
    return ok;
}

struct TriHot { V3 v0, e1, e2; int canon; };
KDPT_HD TriHot loadTri(const Tree& T, int i)
{
    float4 a = KDPT_LDG(T.triA + i), b = KDPT_LDG(T.triB + i);
    float2 c = KDPT_LDG(T.triC + i);
    TriHot t;
    t.v0 = mk(a.x, a.y, a.z); t.e1 = mk(a.w, b.x, b.y); t.e2 = mk(b.z, b.w, c.x);
    t.canon = __float_as_int(c.y);
    return t;
}

// Leaf body after glm said "intersected": interpolated normal, nudged point, distance, acceptance
// (`t > 0 && t_min > t`, pathtrace.cu:1000). Equal distances: the candidate also wins when tieBest >= 0 and
// tieKey < tieBest — the order-independent form of "first in the reference's order wins" (the exact modes key on the
// canonical index, the reference-compatible ones on the TriBare index within the current leaf). Returns accepted.
KDPT_HD bool acceptHit(const Tree& T, int i, int canon, V3 o, V3 d, float u, float v, float z, float eps,
                       int tieKey, int tieBest, Hit& H)
{
    if (T.entryOf) i = KDPT_LDG(T.entryOf + canon);     // refined view: i indexed the refined hot arrays
    float4 a = KDPT_LDG(T.nrmA + i), b = KDPT_LDG(T.nrmB + i);
    float2 c = KDPT_LDG(T.nrmC + i);
    V3 n1 = mk(a.x, a.y, a.z), n2 = mk(a.w, b.x, b.y), n3 = mk(b.z, b.w, c.x);
    int mtl = __float_as_int(c.y);
    V3 hp = o + d * z;
    V3 nrm = normalize(((1 - u - v) * n1 + u * n2) + v * n3);
    hp = hp + nrm * eps;
    float t = length(hp - o);
    bool better = t > 0.0f && (H.t > t || (H.t == t && tieBest >= 0 && tieKey < tieBest));
    if (better) {
        H.t = t; H.z = z; H.u = u; H.v = v; H.tri = canon; H.nx = nrm.x; H.ny = nrm.y; H.nz = nrm.z; H.mtl = mtl;
    }
    return better;
}

// ray (line) against the padded box of chunk c: entry / exit parameters. NaNs (0 * inf) drop out of fminf/fmaxf,
// which only widens the interval.
KDPT_HD void boxInterval(const float4* los, const float4* his, int c, V3 o, V3 inv, float& tn, float& tf)
{
    float4 lo = KDPT_LDG(los + c), hi = KDPT_LDG(his + c);
    float v1 = (lo.x - o.x) * inv.x, v2 = (hi.x - o.x) * inv.x;
    float v3 = (lo.y - o.y) * inv.y, v4 = (hi.y - o.y) * inv.y;
    float v5 = (lo.z - o.z) * inv.z, v6 = (hi.z - o.z) * inv.z;
    tn = fmaxf(fmaxf(fminf(v1, v2), fminf(v3, v4)), fminf(v5, v6));
    tf = fminf(fminf(fmaxf(v1, v2), fmaxf(v3, v4)), fmaxf(v5, v6));
}
// exact modes: a chunk can be skipped when the ray misses its box, leaves it behind the origin, or enters it beyond
// the best hit so far (zlim = best glm t + slack). Written so that any NaN keeps the chunk.
KDPT_HD void chunkInterval(const Tree& T, int c, V3 o, V3 inv, float& tn, float& tf) { boxInterval(T.chunkLo, T.chunkHi, c, o, inv, tn, tf); }
KDPT_HD bool chunkMayHit(const Tree& T, int c, V3 o, V3 inv, float zlim, float slack)
{
    float tn, tf;
    chunkInterval(T, c, o, inv, tn, tf);
    return !(tn > tf) && !(tf < -slack) && !(tn > zlim);
}
KDPT_HD bool superMayHit(const Tree& T, int g, V3 o, V3 inv, float zlim, float slack)
{
    float tn, tf;
    boxInterval(T.superLo, T.superHi, g, o, inv, tn, tf);
    return !(tn > tf) && !(tf < -slack) && !(tn > zlim);
}
KDPT_HD int superFirst(const Tree& T, unsigned chunk0) { return __float_as_int(KDPT_LDG(T.chunkLo + chunk0).w); }

KDPT_HD void initHit(Hit& H, float tMinIn)
{
    H.t = tMinIn; H.z = 0.f; H.u = 0.f; H.v = 0.f; H.tri = -1; H.nx = H.ny = H.nz = 0.f; H.mtl = -1; H.anyMtl = -1;
}

// ---- exact closest hit, one ray per thread --------------------------------------------------------------------
// Near-first descent on the split planes, identical to oracle/kdpt_oracle.c:traverse_stack. STACK = per-ray stack
// capacity: kMaxDepth gives the full-stack variant, a small value the short-stack/restart variant (entries pushed
// when the stack is full push the oldest one out; when the stack runs dry after that, the ray restarts at the
// root with tmin advanced to the exit of the last leaf).
template <int STACK>
struct RayStack {
    unsigned node[STACK]; float tmin[STACK]; float tmax[STACK];
    int n; bool dropped;
    KDPT_HD void clear() { n = 0; dropped = false; }
    KDPT_HD void push(unsigned nd, float a, float b)
    {
        if (STACK >= kMaxDepth) { node[n] = nd; tmin[n] = a; tmax[n] = b; n++; return; }
        if (n == STACK) {
#pragma unroll
            for (int k = 0; k + 1 < STACK; k++) { node[k] = node[k + 1]; tmin[k] = tmin[k + 1]; tmax[k] = tmax[k + 1]; }
            n = STACK - 1; dropped = true;
        }
#pragma unroll
        for (int k = 0; k < STACK; k++) if (k == n) { node[k] = nd; tmin[k] = a; tmax[k] = b; }
        n++;
    }
    KDPT_HD void pop(unsigned& nd, float& a, float& b)
    {
        n--;
        if (STACK >= kMaxDepth) { nd = node[n]; a = tmin[n]; b = tmax[n]; return; }
#pragma unroll
        for (int k = 0; k < STACK; k++) if (k == n) { nd = node[k]; a = tmin[k]; b = tmax[k]; }
    }
};

struct TraceCounters { unsigned nodes, leaves, tris, boxes, hits; };   // boxes: chunk / group boxes tested; hits: triangles accepted by glm

template <int STACK, bool COUNT>
KDPT_HD void traceExact(const Tree& T, V3 o, V3 d, float tMinIn, float eps, Hit& H, TraceCounters* ctr)
{
    initHit(H, tMinIn);
    if (T.numNodes == 0) return;
    float len = length(d);
    if (!(len > 0.0f) || !(len <= FLT_MAX)) return;
    float slack = kExitSlack / len;
    V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float t0, t1;
    if (COUNT) ctr->nodes++;
    if (!slab(o, inv, T.rootLo, T.rootHi, t0, t1)) return;
    if (!(t0 > 0.0f)) t0 = 0.0f;
    float zbound = tMinIn < FLT_MAX ? tMinIn / len : FLT_MAX;
    RayStack<STACK> st; st.clear();
    unsigned cur = 0; float tmin = t0, tmax = t1;
    for (;;) {
        uint4 nd = KDPT_LDG(T.nodes + cur);
        while ((nd.w & 3u) != kLeafFlag) {
            int a = (int)(nd.w & 3u);
            float split = __int_as_float((int)nd.x);
            float oa = comp(o, a), da = comp(d, a), ia = comp(inv, a);
            bool below = oa < split || (oa == split && da <= 0.0f);
            unsigned nearc = below ? nd.z : nd.y, farc = below ? nd.y : nd.z;
            float tp = (split - oa) * ia;
            if (da == 0.0f || tp <= 0.0f || tp >= tmax) cur = nearc;
            else if (tp <= tmin) cur = farc;
            else { st.push(farc, tp, tmax); cur = nearc; tmax = tp; }
            nd = KDPT_LDG(T.nodes + cur);
            if (COUNT) ctr->nodes++;
        }
        int start = (int)nd.x, count = (int)nd.y;
        if (count > 0) {
            if (COUNT) ctr->leaves++;
            const int nch = nd.z != kNoChunks ? (count + 31) >> 5 : 1;
            const int sup0 = nd.z != kNoChunks ? superFirst(T, nd.z) : -1;
            for (int c = 0; c < nch; c++) {
                int b = start, e = start + count;
                if (nd.z != kNoChunks) {
                    float zl = (H.tri >= 0 ? fminf(H.z, zbound) : zbound) + slack;
                    if (COUNT && sup0 >= 0 && (c & 31) == 0) ctr->boxes++;
                    if (sup0 >= 0 && (c & 31) == 0 && !superMayHit(T, sup0 + (c >> 5), o, inv, zl, slack)) { c += 31; continue; }
                    if (COUNT) ctr->boxes++;
                    if (!chunkMayHit(T, (int)nd.z + c, o, inv, zl, slack)) continue;
                    b = start + 32 * c; e = b + 32 < start + count ? b + 32 : start + count;
                }
                if (COUNT) ctr->tris += (unsigned)(e - b);
                for (int i = b; i < e; i++) {
                    TriHot tr = loadTri(T, i);
                    float u, v, z;
                    if (rayTriHit(o, d, tr.v0, tr.e1, tr.e2, u, v, z)) {
                        if (COUNT) ctr->hits++;
                        acceptHit(T, i, tr.canon, o, d, u, v, z, eps, tr.canon, H.tri, H);
                    }
                }
            }
        }
        float zb = H.tri >= 0 ? fminf(H.z, zbound) : zbound;
        bool found = false;
        while (st.n > 0) {
            unsigned n2; float a2, b2;
            st.pop(n2, a2, b2);
            if (a2 > zb + slack) continue;
            cur = n2; tmin = a2; tmax = b2; found = true;
            break;
        }
        if (found) continue;
        if (STACK < kMaxDepth && st.dropped) {
            // kd-restart: everything up to the exit of this leaf is done
            if (tmax >= t1 || tmax > zb + slack) break;
            tmin = tmax; tmax = t1; cur = 0; st.clear();
            continue;
        }
        break;
    }
}

#if defined(__CUDACC__)
// best candidate of the warp (min nudged t, ties to the smaller canonical index) back to the owner lane `src`; L = every
// lane's local best, started from the owner's (bestT, bestTri)
__device__ __forceinline__ void warpBestToOwner(int src, float bestT, int bestTri, const Hit& L, Hit& H)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const bool improved = L.t != bestT || L.tri != bestTri;
    if (!__ballot_sync(FULL, improved)) return;
    float kt = improved ? L.t : FLT_MAX; int kc = improved ? L.tri : 0x7fffffff;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        float t2 = __shfl_xor_sync(FULL, kt, off); int c2 = __shfl_xor_sync(FULL, kc, off);
        if (t2 < kt || (t2 == kt && c2 < kc)) { kt = t2; kc = c2; }
    }
    const int win = __ffs(__ballot_sync(FULL, improved && L.t == kt && L.tri == kc)) - 1;
    Hit W;
    W.t = kt; W.tri = kc;
    W.z = __shfl_sync(FULL, L.z, win); W.u = __shfl_sync(FULL, L.u, win); W.v = __shfl_sync(FULL, L.v, win);
    W.nx = __shfl_sync(FULL, L.nx, win); W.ny = __shfl_sync(FULL, L.ny, win); W.nz = __shfl_sync(FULL, L.nz, win);
    W.mtl = __shfl_sync(FULL, L.mtl, win); W.anyMtl = W.mtl;
    if ((int)lane == src) H = W;
}

// ---- brute force (KDPT_BRUTE / KDPT_BRUTE_BBOX): the reference's KD-less modes, pathTraceOneBounce (pathtrace.cu:402-628)
// Every TriBare entry is tested for every ray — 32 entries per step by the warp, one owner ray at a time — with the nudge
// (1e-4) and the acceptance rule of that kernel; BBOX first rejects rays that miss the mesh's bounding box (the
// reference's per-shape AABB reject, :497-511; here the KD root box, which bounds every triangle). By construction the
// result is what KDPT_SHORTSTACK returns; the modes exist to reproduce the reference's brute force / bounding box /
// kd-tree comparison (README, presentation/benchmarks.py) on this hardware.
template <bool BBOX>
__device__ __forceinline__ void traceBruteWarp(const Tree& T, bool active, V3 o, V3 d, float tMinIn, float eps, Hit& H)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    initHit(H, tMinIn);
    bool run = active && T.numTris > 0;
    if (BBOX && run) {
        const V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float t0, t1;
        if (!slab(o, inv, T.rootLo, T.rootHi, t0, t1)) run = false;
    }
    unsigned todo = __ballot_sync(FULL, run);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const V3 O = mk(__shfl_sync(FULL, o.x, src), __shfl_sync(FULL, o.y, src), __shfl_sync(FULL, o.z, src));
        const V3 D = mk(__shfl_sync(FULL, d.x, src), __shfl_sync(FULL, d.y, src), __shfl_sync(FULL, d.z, src));
        const float bestT = __shfl_sync(FULL, H.t, src);
        const int bestTri = __shfl_sync(FULL, H.tri, src);
        Hit L; L.t = bestT; L.tri = bestTri; L.z = L.u = L.v = 0.f; L.nx = L.ny = L.nz = 0.f; L.mtl = -1; L.anyMtl = -1;
        for (int k = (int)lane; k < T.numTris; k += 32) {
            const TriHot tr = loadTri(T, k);
            float u, v, z;
            if (rayTriHit(O, D, tr.v0, tr.e1, tr.e2, u, v, z)) acceptHit(T, k, tr.canon, O, D, u, v, z, eps, tr.canon, L.tri, L);
        }
        warpBestToOwner(src, bestT, bestTri, L, H);
    }
}
#endif

#if defined(__CUDACC__)
// ---- exact closest hit, warp-cooperative leaves ---------------------------------------------------------------
// Same traversal as traceExact (so the same leaves and triangles are tested and the same hit is returned), but the
// reference tree's fat leaves (hundreds of triangles: the builder stops at depth 14) are not walked by one lane while
// 31 idle: every lane descends its own ray to its next leaf, then the warp takes the pending leaves with at least
// COOP_MIN triangles (when the cost model below says so) one at a time — the owner's ray is broadcast with shuffles, the 32 lanes test 32 consecutive
// triangles per step (three fully coalesced loads), and the best candidate (min nudged t, ties to the smaller
// canonical index: the order-independent form of the reference's strict `t_min > t`) goes back to the owner.
// Small leaves stay lane-private. Must be called by all 32 lanes of a warp (lanes without a ray pass active=false).
template <int STACK, int COOP_MIN>
__device__ __forceinline__ void traceExactWarp(const Tree& T, bool active, V3 o, V3 d, float tMinIn, float eps, Hit& H)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    initHit(H, tMinIn);
    bool run = active && T.numNodes != 0;
    float len = 0.f, slack = 0.f, t0 = 0.f, t1 = 0.f, zbound = FLT_MAX;
    V3 inv = mk(0, 0, 0);
    if (run) {
        len = length(d);
        if (!(len > 0.0f) || !(len <= FLT_MAX)) run = false;
    }
    if (run) {
        slack = kExitSlack / len;
        inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        if (!slab(o, inv, T.rootLo, T.rootHi, t0, t1)) run = false;
        else {
            if (!(t0 > 0.0f)) t0 = 0.0f;
            zbound = tMinIn < FLT_MAX ? tMinIn / len : FLT_MAX;
        }
    }
    RayStack<STACK> st; st.clear();
    unsigned cur = 0; float tmin = t0, tmax = t1;

    // one owner's leaf range [cstart, cstart+ccount) tested by the whole warp; L = lane-local best candidate
    auto testRange = [&](V3 O, V3 D, int cstart, int kbeg, int kend, Hit& L) {
        int k = kbeg + (int)lane;
        for (; k + 32 < kend; k += 64) {      // two triangles per lane and step: six independent loads in flight
            TriHot ta = loadTri(T, cstart + k), tb = loadTri(T, cstart + k + 32);
            float u, v, z;
            if (rayTriHit(O, D, ta.v0, ta.e1, ta.e2, u, v, z)) acceptHit(T, cstart + k, ta.canon, O, D, u, v, z, eps, ta.canon, L.tri, L);
            if (rayTriHit(O, D, tb.v0, tb.e1, tb.e2, u, v, z)) acceptHit(T, cstart + k + 32, tb.canon, O, D, u, v, z, eps, tb.canon, L.tri, L);
        }
        if (k < kend) {
            TriHot tr = loadTri(T, cstart + k);
            float u, v, z;
            if (rayTriHit(O, D, tr.v0, tr.e1, tr.e2, u, v, z)) acceptHit(T, cstart + k, tr.canon, O, D, u, v, z, eps, tr.canon, L.tri, L);
        }
    };
    auto reduceToOwner = [&](int src, float bestT, int bestTri, const Hit& L) { warpBestToOwner(src, bestT, bestTri, L, H); };

    while (__any_sync(FULL, run)) {
        int start = 0, count = 0; unsigned chunk0 = kNoChunks;
        if (run) {
            uint4 nd = __ldg(T.nodes + cur);
            while ((nd.w & 3u) != kLeafFlag) {
                int a = (int)(nd.w & 3u);
                float split = __int_as_float((int)nd.x);
                float oa = comp(o, a), da = comp(d, a), ia = comp(inv, a);
                bool below = oa < split || (oa == split && da <= 0.0f);
                unsigned nearc = below ? nd.z : nd.y, farc = below ? nd.y : nd.z;
                float tp = (split - oa) * ia;
                if (da == 0.0f || tp <= 0.0f || tp >= tmax) cur = nearc;
                else if (tp <= tmin) cur = farc;
                else { st.push(farc, tp, tmax); cur = nearc; tmax = tp; }
                nd = __ldg(T.nodes + cur);
            }
            start = (int)nd.x; count = (int)nd.y; chunk0 = nd.z;
        }
        bool pendingLeaf = run && count > 0;

        // ---- chunked (fat) leaves: always cooperative. The lanes first test the owner's ray against 32 chunk boxes
        // at a time, then the warp streams only the chunks the ray can touch. ----
        unsigned chunked = __ballot_sync(FULL, pendingLeaf && chunk0 != kNoChunks);
        if (pendingLeaf && chunk0 != kNoChunks) pendingLeaf = false;
        while (chunked) {
            const int src = __ffs(chunked) - 1;
            chunked &= chunked - 1;
            V3 O = mk(__shfl_sync(FULL, o.x, src), __shfl_sync(FULL, o.y, src), __shfl_sync(FULL, o.z, src));
            V3 D = mk(__shfl_sync(FULL, d.x, src), __shfl_sync(FULL, d.y, src), __shfl_sync(FULL, d.z, src));
            V3 INV = mk(__shfl_sync(FULL, inv.x, src), __shfl_sync(FULL, inv.y, src), __shfl_sync(FULL, inv.z, src));
            const int cstart = __shfl_sync(FULL, start, src), ccount = __shfl_sync(FULL, count, src);
            const int cchunk0 = (int)__shfl_sync(FULL, chunk0, src);
            const float bestT = __shfl_sync(FULL, H.t, src);
            const int bestTri = __shfl_sync(FULL, H.tri, src);
            const float cslack = __shfl_sync(FULL, slack, src);
            float zlim = __shfl_sync(FULL, (H.tri >= 0 ? fminf(H.z, zbound) : zbound) + slack, src);
            Hit L; L.t = bestT; L.tri = bestTri; L.z = FLT_MAX; L.u = L.v = 0.f; L.nx = L.ny = L.nz = 0.f; L.mtl = -1; L.anyMtl = -1;
            const int nch = (ccount + 31) >> 5;
            // 32 chunk boxes per step; the chunks the ray can touch are then streamed one per step
            auto chunkGroup = [&](int cb) {
                const int c = cb + (int)lane;
                unsigned hitMask = __ballot_sync(FULL, c < nch && chunkMayHit(T, cchunk0 + c, O, INV, zlim, cslack));
                if (!hitMask) return;
#if KDPT_CHUNK_PAIR
                while (hitMask) {       // two chunks per step: six independent loads in flight
                    const int j = __ffs(hitMask) - 1;
                    hitMask &= hitMask - 1;
                    const int j2 = __ffs(hitMask) - 1;          // -1: none left
                    hitMask &= hitMask - 1;
                    const int k = (cb + j) * 32 + (int)lane, k2 = (cb + j2) * 32 + (int)lane;
                    const bool has = k < ccount, has2 = j2 >= 0 && k2 < ccount;
                    TriHot tr, tr2;
                    if (has) tr = loadTri(T, cstart + k);
                    if (has2) tr2 = loadTri(T, cstart + k2);
                    float u, v, z;
                    if (has && rayTriHit(O, D, tr.v0, tr.e1, tr.e2, u, v, z)) acceptHit(T, cstart + k, tr.canon, O, D, u, v, z, eps, tr.canon, L.tri, L);
                    if (has2 && rayTriHit(O, D, tr2.v0, tr2.e1, tr2.e2, u, v, z)) acceptHit(T, cstart + k2, tr2.canon, O, D, u, v, z, eps, tr2.canon, L.tri, L);
                }
#else
                while (hitMask) {
                    const int j = __ffs(hitMask) - 1;
                    hitMask &= hitMask - 1;
                    const int k = (cb + j) * 32 + (int)lane;
                    if (k < ccount) {
                        TriHot tr = loadTri(T, cstart + k);
                        float u, v, z;
                        if (rayTriHit(O, D, tr.v0, tr.e1, tr.e2, u, v, z)) acceptHit(T, cstart + k, tr.canon, O, D, u, v, z, eps, tr.canon, L.tri, L);
                    }
                }
#endif
                // tighten the cull distance with whatever any lane has found so far (only worth it on long leaves)
                if (nch > 32) {
                    float lz = (L.t != bestT || L.tri != bestTri) ? L.z : FLT_MAX;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) lz = fminf(lz, __shfl_xor_sync(FULL, lz, off));
                    if (lz < FLT_MAX) zlim = fminf(zlim, lz + cslack);
                }
            };
            const int sup0 = superFirst(T, (unsigned)cchunk0);
            if (sup0 < 0) {
                for (int cb = 0; cb < nch; cb += 32) chunkGroup(cb);
            } else {
                // very fat leaf: 32 group boxes (32 x 32 x 32 triangles) per step first
                const int nsup = (nch + 31) >> 5;
                for (int sb = 0; sb < nsup; sb += 32) {
                    const int g = sb + (int)lane;
                    unsigned gm = __ballot_sync(FULL, g < nsup && superMayHit(T, sup0 + g, O, INV, zlim, cslack));
                    while (gm) {
                        const int j = __ffs(gm) - 1;
                        gm &= gm - 1;
                        chunkGroup((sb + j) * 32);
                    }
                }
            }
            reduceToOwner(src, bestT, bestTri, L);
        }

        // ---- small leaves: per-round cost model. Testing a leaf of c triangles cooperatively costs about
        // 1.5 + c/32 steps (broadcast + reduction + ceil(c/32) tests); the lane-private pass that follows costs max(c)
        // over the lanes left to it, times a = laneCost32/32 (a lane-private step gathers from 32 different leaves:
        // ~1 when the triangles sit in L1, larger when every lane pulls its own lines from L2). Taking the group of
        // near-largest leaves (within 3/4 of the largest, g lanes) out of the private pass pays when
        // g * (1.5 + m1/32) < a * (m1 - m2), m2 being the largest leaf that would remain. ----
        for (;;) {
            const int m1 = __reduce_max_sync(FULL, pendingLeaf ? count : 0);
            if (m1 < COOP_MIN) break;
            const bool inGroup = pendingLeaf && count * 4 >= m1 * 3;
            unsigned big = __ballot_sync(FULL, inGroup);
            const int m2 = __reduce_max_sync(FULL, (pendingLeaf && !inGroup) ? count : 0);
            if (__popc(big) * (48 + m1) >= T.laneCost32 * (m1 - m2)) break;
            if (inGroup) pendingLeaf = false;
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                V3 O = mk(__shfl_sync(FULL, o.x, src), __shfl_sync(FULL, o.y, src), __shfl_sync(FULL, o.z, src));
                V3 D = mk(__shfl_sync(FULL, d.x, src), __shfl_sync(FULL, d.y, src), __shfl_sync(FULL, d.z, src));
                const int cstart = __shfl_sync(FULL, start, src), ccount = __shfl_sync(FULL, count, src);
                const float bestT = __shfl_sync(FULL, H.t, src);
                const int bestTri = __shfl_sync(FULL, H.tri, src);
                Hit L; L.t = bestT; L.tri = bestTri; L.z = L.u = L.v = 0.f; L.nx = L.ny = L.nz = 0.f; L.mtl = -1; L.anyMtl = -1;
                testRange(O, D, cstart, 0, ccount, L);
                reduceToOwner(src, bestT, bestTri, L);
            }
        }
        if (run) {
            if (pendingLeaf) {
                for (int i = start; i < start + count; i++) {
                    TriHot tr = loadTri(T, i);
                    float u, v, z;
                    if (rayTriHit(o, d, tr.v0, tr.e1, tr.e2, u, v, z)) acceptHit(T, i, tr.canon, o, d, u, v, z, eps, tr.canon, H.tri, H);
                }
            }
            float zb = H.tri >= 0 ? fminf(H.z, zbound) : zbound;
            bool found = false;
            while (st.n > 0) {
                unsigned n2; float a2, b2;
                st.pop(n2, a2, b2);
                if (a2 > zb + slack) continue;
                cur = n2; tmin = a2; tmax = b2; found = true;
                break;
            }
            if (!found) {
                if (STACK < kMaxDepth && st.dropped && !(tmax >= t1 || tmax > zb + slack)) {
                    tmin = tmax; tmax = t1; cur = 0; st.clear();        // kd-restart
                } else run = false;
            }
        }
    }
}
#endif // __CUDACC__

// ---- reference-compatible traversal (KDPT_REF_T1 / KDPT_REF_T2) --------------------------------------------------
// The reference's parent-pointer state machine visits nodes exactly like this depth-first walk (proved equal on
// every ray by tests/test_oracle_cpu.py on the CPU restatement, oracle/kdpt_oracle.c:traverse_dfs):
//   visit(N): box test; prune when missed or dist > bary.z, where bary.z is whatever glm wrote LAST (a farther
//   hit, or a negative t); leaf: test every triangle; interior: first child, re-check the prune, second child.
//   T2 picks the first child by the sign of dir[axis] and, on every triangle hit, tags a child of node 0 (or of
//   node 1 when the leaf's parent is itself tagged) as visited — pathtrace.cu:1136/1199.
template <bool T2>
KDPT_HD void traceRef(const Tree& T, V3 o, V3 d, float tMinIn, Hit& H, float* lastBary)
{
    initHit(H, tMinIn);
    float bu = 0.f, bv = 0.f, bz = FLT_MAX;
    const float eps = T2 ? 0.0001f : 0.00001f;
    if (T.numNodes != 0) {
        V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        unsigned tagMask = 0;
        auto tagged = [&](int id) -> bool {
            if (!T2 || id < 0) return false;
            return ((tagMask & 1u) && id == T.tagCand[0]) || ((tagMask & 2u) && id == T.tagCand[1]) ||
                   ((tagMask & 4u) && id == T.tagCand[2]) || ((tagMask & 8u) && id == T.tagCand[3]);
        };
        int stNode[kMaxDepth]; int stSecond[kMaxDepth]; float stDist[kMaxDepth];
        int sp = 0;
        int cur = 0;
        for (;;) {
            uint4 nd = KDPT_LDG(T.nodes + cur);
            float dist;
            bool hitGeom = refBox(o, inv, KDPT_LDG(T.boxLo + cur), KDPT_LDG(T.boxHi + cur), dist);
            bool descend = hitGeom && !(dist > bz);
            if (descend) {
                int nodeAxis = __float_as_int(KDPT_LDG(T.boxHi + cur).w); // NodeBare.axis (level % 3)
                bool leftFirst = !(T2 && !(comp(d, nodeAxis) > 0.0f));
                if ((nd.w & 3u) == kLeafFlag) {
                    int start = (int)nd.x, count = (int)nd.y;
                    int parent = (int)(nd.w >> 2) - 1;
                    // The reference walks the leaf in TriBare order and three things depend on that order: bary.z
                    // is what the LAST triangle that got through the barycentric tests wrote (pathtrace.cu:951 prunes
                    // on it), objMaterialIdx comes from the LAST glm hit (:991), and among equal nudged distances the
                    // FIRST hit stays (strict `t_min > t`, :1000). All three are recovered from the entries' TriBare
                    // indices (triOrig), so the entries may be visited in any order and whole chunks whose box the
                    // ray's LINE misses can be skipped: a triangle writes bary.z only if the line pierces it (glm
                    // writes z before rejecting t < 0, so no culling on the sign or size of t here).
                    int zOrig = -1, mtlOrig = -1, accOrig = -1, nHits = 0;
                    float zNew = bz;
                    const int nch = nd.z != kNoChunks ? (count + 31) >> 5 : 1;
                    const int sup0 = nd.z != kNoChunks ? superFirst(T, nd.z) : -1;
                    for (int c = 0; c < nch; c++) {
                        int b = start, e = start + count;
                        if (nd.z != kNoChunks) {
                            float tn, tf;
                            if (sup0 >= 0 && (c & 31) == 0) {
                                boxInterval(T.superLo, T.superHi, sup0 + (c >> 5), o, inv, tn, tf);
                                if (tn > tf) { c += 31; continue; }
                            }
                            chunkInterval(T, (int)nd.z + c, o, inv, tn, tf);
                            if (tn > tf) continue;
                            b = start + 32 * c; e = b + 32 < start + count ? b + 32 : start + count;
                        }
                        for (int i = b; i < e; i++) {
                            TriHot tr = loadTri(T, i);
                            float u, v, z;
                            int r = rayTri(o, d, tr.v0, tr.e1, tr.e2, u, v, z);
                            if (r == 0) continue;
                            const int orig = KDPT_LDG(T.triOrig + i);
                            if (orig > zOrig) { zOrig = orig; zNew = z; bu = u; bv = v; }
                            if (r == 2) {
                                nHits++;
                                if (acceptHit(T, i, tr.canon, o, d, u, v, z, eps, orig, accOrig, H)) accOrig = orig;
                                if (orig > mtlOrig) { mtlOrig = orig; H.anyMtl = __float_as_int(KDPT_LDG(T.nrmC + i).y); }
                            }
                        }
                    }
                    bz = zNew;
                    if (T2) {
                        // "skip other side" (:1136/:1199) runs once per glm hit; its effect saturates after two
                        for (int h = 0; h < (nHits < 2 ? nHits : 2); h++) {
                            int k = tagged(parent) ? 1 : 0;
                            if (k < T.numNodes) {
                                int bit = k * 2 + (leftFirst ? 1 : 0);
                                if (T.tagCand[bit] >= 0) tagMask |= 1u << bit;
                            }
                        }
                    }
                } else {
                    int first = leftFirst ? (int)nd.z : (int)nd.y, second = leftFirst ? (int)nd.y : (int)nd.z;
                    bool fOk = !tagged(first), sOk = !tagged(second);
                    if (fOk) { stNode[sp] = cur; stSecond[sp] = sOk ? second : -1; stDist[sp] = dist; sp++; cur = first; continue; }
                    if (sOk) { stNode[sp] = cur; stSecond[sp] = -1; stDist[sp] = dist; sp++; cur = second; continue; }
                }
            }
            int next = -1;
            while (sp > 0) {
                int n = stNode[sp - 1], sec = stSecond[sp - 1]; float pd = stDist[sp - 1];
                if (tagged(n)) { sp--; continue; }
                if (pd > bz) { sp--; continue; }
                if (sec != -1 && !tagged(sec)) { stSecond[sp - 1] = -1; next = sec; break; }
                sp--;
            }
            if (next == -1) break;
            cur = next;
        }
    }
    // bary.x/.y of the triangle that wrote bary.z last (the reference's own x/y may come from a later triangle that
    // failed the second test; nothing ever reads them)
    if (lastBary) { lastBary[0] = bu; lastBary[1] = bv; lastBary[2] = bz; }
}

#if defined(__CUDACC__)
// ---- reference-compatible traversal, warp-cooperative leaves ------------------------------------------------------
// Same walk and same results as traceRef (bit for bit), but a lane that arrives at a leaf of at least kRefCoopMin
// entries does not loop over it alone: the warp takes such leaves one owner at a time, 32 entries per step (and 32 chunk
// boxes per step on chunked leaves, line test only — see traceRef), and the order-dependent leaf results are rebuilt by
// reductions keyed on the entries' TriBare indices: last barycentric-test survivor (max index) -> bary.z, last glm hit
// (max index) -> material, best nudged t with the smaller index winning ties inside the leaf -> accepted hit, and the
// number of glm hits -> T2's tag updates. Must be called by all 32 lanes (lanes without a ray pass active = false).
constexpr int kRefCoopMin = 24;
template <bool T2>
__device__ __forceinline__ void traceRefWarp(const Tree& T, bool active, V3 o, V3 d, float tMinIn, Hit& H)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    initHit(H, tMinIn);
    float bz = FLT_MAX;
    const float eps = T2 ? 0.0001f : 0.00001f;
    bool run = active && T.numNodes != 0;
    const V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    unsigned tagMask = 0;
    auto tagged = [&](int id) -> bool {
        if (!T2 || id < 0) return false;
        return ((tagMask & 1u) && id == T.tagCand[0]) || ((tagMask & 2u) && id == T.tagCand[1]) ||
               ((tagMask & 4u) && id == T.tagCand[2]) || ((tagMask & 8u) && id == T.tagCand[3]);
    };
    int stNode[kMaxDepth]; int stSecond[kMaxDepth]; float stDist[kMaxDepth];
    int sp = 0, cur = 0;
    // the leaf the lane stands at (pending) and what its T2 tag update needs
    int lstart = 0, lcount = 0, lparent = -1; unsigned lchunk0 = kNoChunks; bool lleftFirst = true;
    // pop: the next node of the walk, -1 when the ray is finished
    auto popNext = [&]() -> int {
        while (sp > 0) {
            const int n = stNode[sp - 1], sec = stSecond[sp - 1]; const float pd = stDist[sp - 1];
            if (tagged(n)) { sp--; continue; }
            if (pd > bz) { sp--; continue; }
            if (sec != -1 && !tagged(sec)) { stSecond[sp - 1] = -1; return sec; }
            sp--;
        }
        return -1;
    };
    // what a leaf did to the ray's order-dependent state, then on with the walk
    auto afterLeaf = [&](int zOrig, float zNew, int nHits) {
        if (zOrig >= 0) bz = zNew;
        if (T2) {
            for (int h = 0; h < (nHits < 2 ? nHits : 2); h++) {      // "skip other side" (:1136/:1199), saturates after two
                const int k = tagged(lparent) ? 1 : 0;
                if (k < T.numNodes) {
                    const int bit = k * 2 + (lleftFirst ? 1 : 0);
                    if (T.tagCand[bit] >= 0) tagMask |= 1u << bit;
                }
            }
        }
        const int nx = popNext();
        if (nx == -1) run = false; else cur = nx;
    };
    while (__any_sync(FULL, run)) {
        // ---- phase 1 (per lane): walk until the lane stands at a leaf it has to test, or the ray is finished ----
        bool pending = false;
        while (run && !pending) {
            const uint4 nd = __ldg(T.nodes + cur);
            float dist;
            const float4 bhi = __ldg(T.boxHi + cur);
            const bool hitGeom = refBox(o, inv, __ldg(T.boxLo + cur), bhi, dist);
            if (hitGeom && !(dist > bz)) {
                const int nodeAxis = __float_as_int(bhi.w);
                const bool leftFirst = !(T2 && !(comp(d, nodeAxis) > 0.0f));
                if ((nd.w & 3u) == kLeafFlag) {
                    lstart = (int)nd.x; lcount = (int)nd.y; lchunk0 = nd.z; lparent = (int)(nd.w >> 2) - 1; lleftFirst = leftFirst;
                    pending = true;
                    break;
                }
                const int first = leftFirst ? (int)nd.z : (int)nd.y, second = leftFirst ? (int)nd.y : (int)nd.z;
                const bool fOk = !tagged(first), sOk = !tagged(second);
                if (fOk) { stNode[sp] = cur; stSecond[sp] = sOk ? second : -1; stDist[sp] = dist; sp++; cur = first; continue; }
                if (sOk) { stNode[sp] = cur; stSecond[sp] = -1; stDist[sp] = dist; sp++; cur = second; continue; }
            }
            const int nx = popNext();
            if (nx == -1) run = false; else cur = nx;
        }
        // ---- phase 2 (whole warp): the big pending leaves, one owner at a time ----
        unsigned big = __ballot_sync(FULL, pending && lcount >= kRefCoopMin);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const V3 O = mk(__shfl_sync(FULL, o.x, src), __shfl_sync(FULL, o.y, src), __shfl_sync(FULL, o.z, src));
            const V3 D = mk(__shfl_sync(FULL, d.x, src), __shfl_sync(FULL, d.y, src), __shfl_sync(FULL, d.z, src));
            const V3 INV = mk(__shfl_sync(FULL, inv.x, src), __shfl_sync(FULL, inv.y, src), __shfl_sync(FULL, inv.z, src));
            const int cstart = __shfl_sync(FULL, lstart, src), ccount = __shfl_sync(FULL, lcount, src);
            const unsigned cchunk0 = __shfl_sync(FULL, lchunk0, src);
            const float bestT = __shfl_sync(FULL, H.t, src);
            // lane-local results over the entries this lane tests
            int zOrig = -1; float zz = 0.f, zu = 0.f, zv = 0.f;
            int mOrig = -1, mMtl = -1, nHits = 0, accOrig = -1;
            Hit L; L.t = bestT; L.tri = -1; L.z = L.u = L.v = 0.f; L.nx = L.ny = L.nz = 0.f; L.mtl = -1; L.anyMtl = -1;
            auto testEntry = [&](int i) {
                const TriHot tr = loadTri(T, i);
                float u, v, z;
                const int r = rayTri(O, D, tr.v0, tr.e1, tr.e2, u, v, z);
                if (r == 0) return;
                const int orig = __ldg(T.triOrig + i);
                if (orig > zOrig) { zOrig = orig; zz = z; zu = u; zv = v; }
                if (r == 2) {
                    nHits++;
                    if (acceptHit(T, i, tr.canon, O, D, u, v, z, eps, orig, accOrig, L)) accOrig = orig;
                    if (orig > mOrig) { mOrig = orig; mMtl = __float_as_int(__ldg(T.nrmC + i).y); }
                }
            };
            if (cchunk0 == kNoChunks) {
                for (int k = (int)lane; k < ccount; k += 32) testEntry(cstart + k);
            } else {
                const int nch = (ccount + 31) >> 5;
                auto chunkGroup = [&](int cb) {
                    const int c = cb + (int)lane;
                    bool may = false;
                    if (c < nch) { float tn, tf; chunkInterval(T, (int)cchunk0 + c, O, INV, tn, tf); may = !(tn > tf); }
                    unsigned hitMask = __ballot_sync(FULL, may);
                    while (hitMask) {
                        const int j = __ffs(hitMask) - 1;
                        hitMask &= hitMask - 1;
                        const int k = (cb + j) * 32 + (int)lane;
                        if (k < ccount) testEntry(cstart + k);
                    }
                };
                const int sup0 = superFirst(T, cchunk0);
                if (sup0 < 0) {
                    for (int cb = 0; cb < nch; cb += 32) chunkGroup(cb);
                } else {
                    const int nsup = (nch + 31) >> 5;
                    for (int sb = 0; sb < nsup; sb += 32) {
                        const int g = sb + (int)lane;
                        bool may = false;
                        if (g < nsup) { float tn, tf; boxInterval(T.superLo, T.superHi, sup0 + g, O, INV, tn, tf); may = !(tn > tf); }
                        unsigned gm = __ballot_sync(FULL, may);
                        while (gm) {
                            const int j = __ffs(gm) - 1;
                            gm &= gm - 1;
                            chunkGroup((sb + j) * 32);
                        }
                    }
                }
            }
            // reductions: last bary writer, last glm hit, accepted hit, hit count
            const int zWin = __reduce_max_sync(FULL, zOrig);
            const int mWin = __reduce_max_sync(FULL, mOrig);
            const int hitsAll = __reduce_add_sync(FULL, nHits);
            float rz = 0.f; int rm = -1;
            if (zWin >= 0) {
                const int wl = __ffs(__ballot_sync(FULL, zOrig == zWin)) - 1;
                rz = __shfl_sync(FULL, zz, wl);
            }
            if (mWin >= 0) {
                const int wl = __ffs(__ballot_sync(FULL, mOrig == mWin)) - 1;
                rm = __shfl_sync(FULL, mMtl, wl);
            }
            // accepted: min t, ties to the smaller TriBare index
            float kt = accOrig >= 0 ? L.t : FLT_MAX; int ko = accOrig >= 0 ? accOrig : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float t2 = __shfl_xor_sync(FULL, kt, off); const int o2 = __shfl_xor_sync(FULL, ko, off);
                if (t2 < kt || (t2 == kt && o2 < ko)) { kt = t2; ko = o2; }
            }
            const bool anyAcc = ko != 0x7fffffff;
            Hit W = L;
            if (anyAcc) {
                const int wl = __ffs(__ballot_sync(FULL, accOrig >= 0 && L.t == kt && accOrig == ko)) - 1;
                W.t = kt;
                W.z = __shfl_sync(FULL, L.z, wl); W.u = __shfl_sync(FULL, L.u, wl); W.v = __shfl_sync(FULL, L.v, wl);
                W.tri = __shfl_sync(FULL, L.tri, wl);
                W.nx = __shfl_sync(FULL, L.nx, wl); W.ny = __shfl_sync(FULL, L.ny, wl); W.nz = __shfl_sync(FULL, L.nz, wl);
                W.mtl = __shfl_sync(FULL, L.mtl, wl);
            }
            if ((int)lane == src) {
                if (anyAcc) { H.t = W.t; H.z = W.z; H.u = W.u; H.v = W.v; H.tri = W.tri; H.nx = W.nx; H.ny = W.ny; H.nz = W.nz; H.mtl = W.mtl; }
                if (mWin >= 0) H.anyMtl = rm;
                pending = false;
                afterLeaf(zWin, rz, hitsAll);
            }
        }
        // ---- phase 3 (per lane): small leaves stay lane-private ----
        if (pending) {
            int zOrig = -1, mtlOrig = -1, accOrig = -1, nHits = 0;
            float zNew = bz;
            for (int i = lstart; i < lstart + lcount; i++) {
                const TriHot tr = loadTri(T, i);
                float u, v, z;
                const int r = rayTri(o, d, tr.v0, tr.e1, tr.e2, u, v, z);
                if (r == 0) continue;
                const int orig = __ldg(T.triOrig + i);
                if (orig > zOrig) { zOrig = orig; zNew = z; }
                if (r == 2) {
                    nHits++;
                    if (acceptHit(T, i, tr.canon, o, d, u, v, z, eps, orig, accOrig, H)) accOrig = orig;
                    if (orig > mtlOrig) { mtlOrig = orig; H.anyMtl = __float_as_int(__ldg(T.nrmC + i).y); }
                }
            }
            afterLeaf(zOrig, zNew, nHits);
        }
    }
}
#endif // __CUDACC__

} // namespace kdpt
