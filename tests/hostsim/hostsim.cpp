// TEST INFRASTRUCTURE ONLY — compiles the product's per-ray device functions (csrc/kdpt_core.cuh) and host flatten
// (csrc/kdpt_flatten.cpp) with g++ so the traversal LOGIC can be checked against the oracle on machines without a
// GPU (pytest -m "not gpu"). It is never loaded by the package, bench.py or the C ABI; the product has no CPU path.
#include "../../kdtreepathtraceroptimization_b200/csrc/kdpt_core.cuh"
#include "../../kdtreepathtraceroptimization_b200/csrc/kdpt_lean.cuh"
#include "../../kdtreepathtraceroptimization_b200/csrc/kdpt_flatten.h"
#include <string>
#include <string.h>
#include <vector>

using namespace kdpt;

// the lean kernels' per-lane loop (kdpt_lean.cuh) for one ray: the same leanBegin / leanRound the kernels call, with the
// "shared-memory" part of the stack in a host array of the same stride. steps = rounds (leaves visited).
template <bool RESTART, bool WIDE>
static void leanTraceOne(const Tree& X, V3 o, V3 d, float tin, float eps, Hit& H, unsigned long long* rounds)
{
    static thread_local std::vector<uint2> smem((size_t)(kLeanSmemStack > kLeanRing ? kLeanSmemStack : kLeanRing) * kLeanThreads);
    uint2 lStack[kWideStack];
    LeanState S;
    bool alive = leanBegin<WIDE>(X, o, d, tin, S, H);
    while (alive) { alive = leanRound<RESTART, WIDE>(X, alive, S, H, eps, smem.data() + 5, lStack, nullptr); if (rounds) ++*rounds; }
}

// refinement of the exact modes' tree: -1 = the product's default, 0 = off (walk the reference tree), n = leaf size
static int g_refine = -1;
extern "C" void hostsim_set_refine(int leafMax) { g_refine = leafMax; }

extern "C" int hostsim_trace(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris,
                             int materialSize, const kdpt_ray* rays, const float* tMinIn, int n, kdpt_hit* hits,
                             int mode, unsigned long long* counters /* nodes, leaves, tris, boxes, hits (5) or NULL */,
                             float* lastBary /* n*3 or NULL, ref modes */)
{
    FlatTree F; std::string err;
    FlattenOpts fo;
    if (g_refine >= 0) fo.refineLeafMax = g_refine;
    if (mode == 202 || mode == 203) fo.wideNodes = 1;
    if (!flatten(nodes, numNodes, tris, numTris, F, err, fo)) { fprintf(stderr, "%s\n", err.c_str()); return -1; }
    Tree T;
    T.nodes = (const uint4*)F.nodes.data();
    T.boxLo = (const float4*)F.boxLo.data(); T.boxHi = (const float4*)F.boxHi.data();
    T.triA = (const float4*)F.triA.data(); T.triB = (const float4*)F.triB.data(); T.triC = (const float2*)F.triC.data();
    T.nrmA = (const float4*)F.nrmA.data(); T.nrmB = (const float4*)F.nrmB.data(); T.nrmC = (const float2*)F.nrmC.data();
    T.triOrig = F.triOrig.data(); T.chunkLo = (const float4*)F.chunkLo.data(); T.chunkHi = (const float4*)F.chunkHi.data();
    T.superLo = (const float4*)F.superLo.data(); T.superHi = (const float4*)F.superHi.data();
    T.entryOf = nullptr; T.top = nullptr; T.numTop = 0; T.wide = nullptr; T.wideRoot = 0;
    T.numNodes = F.numNodes; T.numTris = F.numTris; T.numRefNodes = F.numNodes;
    for (int a = 0; a < 3; a++) { T.rootLo[a] = F.rootLo[a]; T.rootHi[a] = F.rootHi[a]; }
    for (int k = 0; k < 4; k++) T.tagCand[k] = F.tagCand[k];
    T.materialSize = materialSize; T.laneCost32 = 32;
    Tree X = T;         // the refined view the exact modes walk (kdpt_api.cu:kdpt_create_opts does the same on the device)
    if (!F.xnodes.empty()) {
        X.nodes = (const uint4*)F.xnodes.data();
        X.triA = (const float4*)F.xtriA.data(); X.triB = (const float4*)F.xtriB.data(); X.triC = (const float2*)F.xtriC.data();
        X.chunkLo = (const float4*)F.xchunkLo.data(); X.chunkHi = (const float4*)F.xchunkHi.data();
        X.superLo = (const float4*)F.xsuperLo.data(); X.superHi = (const float4*)F.xsuperHi.data();
        X.entryOf = F.entryOf.data();
        X.numNodes = (int)F.xnodes.size(); X.numTris = F.xNumTris;
        if (!F.xwide.empty()) { X.wide = (const uint4*)F.xwide.data(); X.wideRoot = F.xwideRoot; }
    }
    if (mode >= 200 && mode < 204 && F.xnodes.empty()) return -3;       // the lean loop walks the refined view only
    if ((mode == 202 || mode == 203) && !X.wide) return -4;
    TraceCounters tc = {0, 0, 0, 0, 0};
    unsigned long long cn = 0, cl = 0, ct = 0, cb = 0, ch = 0;
    for (int i = 0; i < n; i++) {
        V3 o = mk(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
        V3 d = mk(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        float tin = tMinIn ? tMinIn[i] : FLT_MAX;
        Hit H;
        tc.nodes = tc.leaves = tc.tris = tc.boxes = tc.hits = 0;
        switch (mode) {
        case KDPT_FULLSTACK: traceExact<kMaxDepth, true>(X, o, d, tin, 0.00001f, H, &tc); break;
        case KDPT_SHORTSTACK: traceExact<4, true>(X, o, d, tin, 0.0001f, H, &tc); break;
        case 100: traceExact<1, true>(X, o, d, tin, 0.0001f, H, &tc); break; // restart stress: 1-entry stack
        // the lean kernels' loop: 200 / 201 = full stack / ring + restart on 16-byte nodes, 202 / 203 = the same on two-level records
        case 200: leanTraceOne<false, false>(X, o, d, tin, 0.00001f, H, &cl); break;
        case 201: leanTraceOne<true, false>(X, o, d, tin, 0.0001f, H, &cl); break;
        case 202: leanTraceOne<false, true>(X, o, d, tin, 0.00001f, H, &cl); break;
        case 203: leanTraceOne<true, true>(X, o, d, tin, 0.0001f, H, &cl); break;
        case KDPT_REF_T1: traceRef<false>(T, o, d, tin, H, lastBary ? lastBary + 3 * i : nullptr); break;
        case KDPT_REF_T2: traceRef<true>(T, o, d, tin, H, lastBary ? lastBary + 3 * i : nullptr); break;
        default: return -2;
        }
        cn += tc.nodes; cl += tc.leaves; ct += tc.tris; cb += tc.boxes; ch += tc.hits;
        hits[i].t = H.t; hits[i].z = H.z; hits[i].u = H.u; hits[i].v = H.v; hits[i].tri = H.tri;
        hits[i].normal[0] = H.nx; hits[i].normal[1] = H.ny; hits[i].normal[2] = H.nz;
    }
    if (counters) { counters[0] = cn; counters[1] = cl; counters[2] = ct; counters[3] = cb; counters[4] = ch; }
    return 0;
}

// shape of the refined tree: nodes, leaf entries, max depth, max leaf
extern "C" int hostsim_refined_info(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris, int leafMax, long long* info)
{
    FlatTree F; std::string err;
    FlattenOpts fo; if (leafMax >= 0) fo.refineLeafMax = leafMax;
    if (!flatten(nodes, numNodes, tris, numTris, F, err, fo)) { fprintf(stderr, "%s\n", err.c_str()); return -1; }
    info[0] = (long long)F.xnodes.size(); info[1] = F.xNumTris; info[2] = F.xMaxDepth; info[3] = F.xMaxLeaf;
    return 0;
}

// the two-level records against the 16-byte nodes they were made from: the same tree — planes, axes, leaves (inline range or record
// index) — reached through the records; returns the number of leaves reached, or -(1 + the first node that differs)
static long wideCheck(const FlatTree& F, uint32_t ref, uint32_t id)
{
    const HostNode& n = F.xnodes[id];
    if (ref & kWideLeaf) {
        if ((n.w & 3u) != 3u) return -(long)(1 + id);
        const unsigned cnt = (ref >> 27) & 15u, idx = ref & kWideIndexMask;
        if (cnt == 15u) return idx == id ? 1 : -(long)(1 + id);
        if (n.z != kNoChunks || n.y != cnt || (cnt && n.x != idx)) return -(long)(1 + id);
        return 1;
    }
    if ((n.w & 3u) == 3u || 2 * (size_t)ref + 1 >= F.xwide.size()) return -(long)(1 + id);
    const HostNode &A = F.xwide[2 * (size_t)ref], &B = F.xwide[2 * (size_t)ref + 1];
    if (A.x != n.x || (A.w & 3u) != (n.w & 3u)) return -(long)(1 + id);
    long leaves = 0;
    const uint32_t kid[2] = {n.z, n.y}, slot[2][2] = {{B.x, B.y}, {B.z, B.w}}, split[2] = {A.y, A.z};
    for (int k = 0; k < 2; k++) {
        const HostNode& c = F.xnodes[kid[k]];
        const unsigned ax = (A.w >> (2 + 2 * k)) & 3u;
        if (ax != (c.w & 3u)) return -(long)(1 + kid[k]);
        long r;
        if (ax == 3u) r = wideCheck(F, slot[k][0], kid[k]);
        else {
            if (split[k] != c.x) return -(long)(1 + kid[k]);
            r = wideCheck(F, slot[k][0], c.z);
            if (r < 0) return r;
            leaves += r;
            r = wideCheck(F, slot[k][1], c.y);
        }
        if (r < 0) return r;
        leaves += r;
    }
    return leaves;
}
extern "C" long hostsim_wide_check(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris, int leafMax, long long* info)
{
    FlatTree F; std::string err;
    FlattenOpts fo; fo.wideNodes = 1; if (leafMax >= 0) fo.refineLeafMax = leafMax;
    if (!flatten(nodes, numNodes, tris, numTris, F, err, fo)) { fprintf(stderr, "%s\n", err.c_str()); return -1; }
    if (F.xwide.empty()) return -2;
    long leavesInNodes = 0;
    for (const HostNode& h : F.xnodes) if ((h.w & 3u) == 3u) leavesInNodes++;
    if (info) { info[0] = (long long)F.xnodes.size(); info[1] = (long long)F.xwide.size() / 2; info[2] = leavesInNodes; }
    return wideCheck(F, F.xwideRoot, 0);
}

// canonical index of every TriBare entry as the product's flatten computes it (what kdpt_hit.tri reports)
extern "C" int hostsim_canonical(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris, int* canonOut)
{
    FlatTree F; std::string err;
    if (!flatten(nodes, numNodes, tris, numTris, F, err)) { fprintf(stderr, "%s\n", err.c_str()); return -1; }
    for (int i = 0; i < numTris; i++) { int c; memcpy(&c, &F.triC[i].y, 4); canonOut[F.triOrig[i]] = c; }
    return 0;
}

#ifdef HOSTSIM_COUNT
// development: counters of the lean loop (rounds, rounds on an empty leaf, triangles, node steps); reset = 1 zeroes them
extern "C" void hostsim_lean_counts(unsigned long long* c, int reset)
{
    if (reset) { g_leanRounds = g_leanEmpty = g_leanTris = g_leanSteps = 0; return; }
    c[0] = g_leanRounds; c[1] = g_leanEmpty; c[2] = g_leanTris; c[3] = g_leanSteps;
}
#endif
