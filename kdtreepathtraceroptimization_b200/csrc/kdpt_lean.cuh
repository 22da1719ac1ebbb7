// kdpt_lean.cuh — exact closest hit on the REFINED tree: one ray per lane, lanes refilled as their rays finish.
//
// The refined view (kdpt_flatten.cpp:refineLeaves) has small leaves, so there is nothing for a warp to share inside a
// leaf: every lane walks its own ray — the same near-first traversal, early-exit rule and triangle arithmetic as
// traceExact / traceExactWarp in kdpt_core.cuh, hence the same hits, bit for bit — and what decides the speed is how
// many lanes have a ray at any time. A packet-at-a-time kernel idles the lanes of short rays until the longest ray of
// the packet is done (the soups: 2 to 100 leaves per ray). Here a lane whose ray has finished takes the next ray of
// the batch at the top of the loop, so a warp stays full until the batch runs dry.
//
// Per-ray stack: a pending far child needs only (node, exit distance): its entry distance is the exit distance of
// whatever was traversed before it (the intervals of a kd traversal tile the ray), and the entries lie along the ray
// in stack order — so the first popped entry that begins beyond the best hit ends the ray. The first
// kLeanSmemStack entries per lane live in shared memory ([depth][thread]: conflict-free whatever the depths), deeper
// ones in local memory. RESTART (the short-stack mode): shared-memory ring only; a push onto a full ring drops the
// OLDEST (farthest) entry, and a ray that runs out of entries after a drop restarts at the root from where it is.
#pragma once
#include "kdpt_core.cuh"

namespace kdpt {

#ifndef KDPT_LEAN_THREADS
#define KDPT_LEAN_THREADS 128
#endif
#ifndef KDPT_LEAN_SMEM_STACK
#define KDPT_LEAN_SMEM_STACK 12
#endif
#ifndef KDPT_LEAN_REFILL_MIN
#define KDPT_LEAN_REFILL_MIN 8      // idle lanes needed before the warp fetches new rays (while rays are left): the fetch
                                    // code runs with only the idle lanes active (1 / 4 / 8: 1623 / 1695 / 1722 Mrays/s on dragon1m)
#endif
constexpr int kLeanThreads = KDPT_LEAN_THREADS;
constexpr int kLeanSmemStack = KDPT_LEAN_SMEM_STACK;
constexpr int kLeanRing = 8;        // RESTART: entries in the ring
#ifndef KDPT_TOP_MAX_LEVELS
#define KDPT_TOP_MAX_LEVELS 0       // compile-time bound of the shared-memory copy of the tree's top levels (0: not compiled in)
#endif
constexpr int kTopMaxNodes = KDPT_TOP_MAX_LEVELS > 0 ? (1 << KDPT_TOP_MAX_LEVELS) - 1 : 1;
constexpr unsigned kTopFlag = 0x80000000u;

// the block's copy of the top levels (Tree::top); call from every thread of the block before the first round
__device__ __forceinline__ void leanStageTop(const Tree& T, uint4* sTop)
{
#if KDPT_TOP_MAX_LEVELS > 0
    const int m = T.numTop < kTopMaxNodes ? T.numTop : 0;       // a copy that does not fit is not used at all
    for (int k = threadIdx.x; k < m; k += blockDim.x) sTop[k] = __ldg(T.top + k);
    __syncthreads();
#endif
}
__device__ __forceinline__ unsigned leanRoot(const Tree& T)
{
#if KDPT_TOP_MAX_LEVELS > 0
    return (T.numTop > 0 && T.numTop < kTopMaxNodes) ? kTopFlag : 0u;
#else
    return 0u;
#endif
}
__device__ __forceinline__ uint4 leanNode(const Tree& T, const uint4* sTop, unsigned cur)
{
#if KDPT_TOP_MAX_LEVELS > 0
    if (cur & kTopFlag) return sTop[cur & ~kTopFlag];
#endif
    return __ldg(T.nodes + cur);
}

struct HitCand { float t, nx, ny, nz; int mtl; };
// interpolated normal, nudged point and its distance (pathtrace.cu:993-998) of a triangle glm accepted: the rare part of
// the leaf loop, kept out of line so that the loop itself stays small
static __device__ __noinline__ HitCand nudgedHit(const float4* __restrict__ nrmA, const float4* __restrict__ nrmB, const float2* __restrict__ nrmC,
                                          int i, float ox, float oy, float oz, float dx, float dy, float dz, float u, float v, float z, float eps)
{
    const float4 a = __ldg(nrmA + i), b = __ldg(nrmB + i);
    const float2 c = __ldg(nrmC + i);
    const V3 o = mk(ox, oy, oz), d = mk(dx, dy, dz);
    const V3 n1 = mk(a.x, a.y, a.z), n2 = mk(a.w, b.x, b.y), n3 = mk(b.z, b.w, c.x);
    V3 hp = o + d * z;
    const V3 nrm = normalize(((1 - u - v) * n1 + u * n2) + v * n3);
    hp = hp + nrm * eps;
    HitCand r;
    r.t = length(hp - o); r.nx = nrm.x; r.ny = nrm.y; r.nz = nrm.z; r.mtl = __float_as_int(c.y);
    return r;
}

// Per-lane traversal state between two rounds of the lean loop.
struct LeanState {
    V3 o, d, inv;
    float tmin, tmax, t1, slack, zbound;
    unsigned cur;
    int sp, ringBase;
    int li, le;         // leaf cursor: li < le = the lane is inside a leaf (next entry li), else at node `cur`
    bool dropped;
};

// ray set-up shared by the kernels: false = the ray never enters the tree (H stays the "no hit" record)
__device__ __forceinline__ bool leanBegin(const Tree& T, V3 o, V3 d, float tMinIn, LeanState& S, Hit& H)
{
    initHit(H, tMinIn);
    S.o = o; S.d = d; S.sp = 0; S.ringBase = 0; S.dropped = false; S.cur = leanRoot(T); S.li = 0; S.le = 0;
    if (T.numNodes == 0) return false;
    const float len = length(d);
    if (!(len > 0.0f) || !(len <= FLT_MAX)) return false;
    S.slack = kExitSlack / len;
    S.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float t0, t1;
    if (!slab(o, S.inv, T.rootLo, T.rootHi, t0, t1)) return false;
    if (!(t0 > 0.0f)) t0 = 0.0f;
    S.zbound = tMinIn < FLT_MAX ? tMinIn / len : FLT_MAX;
    S.tmin = t0; S.tmax = t1; S.t1 = t1;
    return true;
}

// the leaf at hand is done: take the next pending subtree — it begins where this leaf's cell ends. false = ray finished.
template <bool RESTART>
__device__ __forceinline__ bool leanPop(const Tree& T, LeanState& S, const Hit& H, const uint2* sStack, const uint2* lStack)
{
    const float tmax = S.tmax;
    const float zb = (H.tri >= 0 ? fminf(H.z, S.zbound) : S.zbound) + S.slack;
    if (S.sp > 0) {
        if (tmax > zb) return false;            // ... beyond the best hit: so do all the others (they lie farther along the ray)
        S.sp--;
        uint2 e;
        if (RESTART) e = sStack[((S.ringBase + S.sp) & (kLeanRing - 1)) * kLeanThreads];
        else e = S.sp < kLeanSmemStack ? sStack[S.sp * kLeanThreads] : lStack[S.sp - kLeanSmemStack];
        S.cur = e.x; S.tmin = tmax; S.tmax = __uint_as_float(e.y);
        return true;
    }
    if (RESTART && S.dropped && !(tmax >= S.t1 || tmax > zb)) {      // kd-restart: everything up to the exit of this leaf is done
        S.cur = leanRoot(T); S.tmin = tmax; S.tmax = S.t1; S.ringBase = 0; S.dropped = false;
        return true;
    }
    return false;
}

// glm accepted entry i (canonical index canon): nudged distance, then the reference's acceptance rule
__device__ __forceinline__ void leanAccept(const Tree& T, const LeanState& S, Hit& H, float eps, int i, int canon, float u, float v, float z)
{
    const int ni = T.entryOf ? __ldg(T.entryOf + canon) : i;
    const HitCand hc = nudgedHit(T.nrmA, T.nrmB, T.nrmC, ni, S.o.x, S.o.y, S.o.z, S.d.x, S.d.y, S.d.z, u, v, z, eps);
    if (hc.t > 0.0f && (H.t > hc.t || (H.t == hc.t && H.tri >= 0 && canon < H.tri))) {
        H.t = hc.t; H.z = z; H.u = u; H.v = v; H.tri = canon; H.nx = hc.nx; H.ny = hc.ny; H.nz = hc.nz; H.mtl = hc.mtl;
    }
}
__device__ __forceinline__ void leanTestTri(const Tree& T, const LeanState& S, Hit& H, float eps, int i)
{
    const TriHot tr = loadTri(T, i);
    float u, v, z;
    if (rayTriHit(S.o, S.d, tr.v0, tr.e1, tr.e2, u, v, z)) leanAccept(T, S, H, eps, i, tr.canon, u, v, z);
}

// a leaf the refinement could not split (>= 128 coincident triangles): all of it at once, through its chunk boxes.
// Rare; kept out of line, with everything passed by value so that the callers' ray state stays in registers.
struct LeanLeafRefs {
    const float4 *chunkLo, *chunkHi, *superLo, *superHi, *triA, *triB, *nrmA, *nrmB;
    const float2 *triC, *nrmC;
    const int* entryOf;
};
static __device__ __noinline__ Hit leanChunkedLeaf(LeanLeafRefs r, V3 o, V3 d, V3 inv, float zbound, float slack, Hit H, float eps,
                                                   int start, int count, unsigned chunk0)
{
    Tree T;
    T.chunkLo = r.chunkLo; T.chunkHi = r.chunkHi; T.superLo = r.superLo; T.superHi = r.superHi;
    T.triA = r.triA; T.triB = r.triB; T.triC = r.triC; T.nrmA = r.nrmA; T.nrmB = r.nrmB; T.nrmC = r.nrmC; T.entryOf = r.entryOf;
    LeanState S;
    S.o = o; S.d = d; S.inv = inv; S.zbound = zbound; S.slack = slack;
    const int nch = (count + 31) >> 5;
    const int sup0 = superFirst(T, chunk0);
    for (int c = 0; c < nch; c++) {
        const float zl = (H.tri >= 0 ? fminf(H.z, zbound) : zbound) + slack;
        if (sup0 >= 0 && (c & 31) == 0 && !superMayHit(T, sup0 + (c >> 5), o, inv, zl, slack)) { c += 31; continue; }
        if (!chunkMayHit(T, (int)chunk0 + c, o, inv, zl, slack)) continue;
        const int b = start + 32 * c, e = b + 32 < start + count ? b + 32 : start + count;
        for (int i = b; i < e; i++) leanTestTri(T, S, H, eps, i);
    }
    return H;
}

#ifndef KDPT_LEAN_NODE_WEIGHT
#define KDPT_LEAN_NODE_WEIGHT 1     // a round takes the leaf step when leafLanes * NODE_WEIGHT > nodeLanes * LEAF_WEIGHT
#endif
#ifndef KDPT_LEAN_LEAF_WEIGHT
#define KDPT_LEAN_LEAF_WEIGHT 1
#endif

// One round of the warp (all 32 lanes call it; `alive` = the lane has a ray). Rays differ wildly in how many nodes
// they cross between leaves and in how many triangles their leaves hold, so a "descend to the next leaf, then test the
// leaf" step makes every lane wait for the longest descent and the fullest leaf of the warp (measured: 6 of 32 lanes
// active per instruction). Here a round is ONE unit of work — one node step or one triangle test — and the warp takes
// the kind more of its lanes are waiting for; the others keep their state and wait for a round of their kind.
// Returns whether the lane still has a ray.
template <bool RESTART>
__device__ __forceinline__ bool leanRoundUnit(const Tree& T, bool alive, LeanState& S, Hit& H, float eps, uint2* sStack, uint2* lStack,
                                          const uint4* sTop)
{
    const bool inLeaf = alive && S.li < S.le;
    const unsigned leafMask = __ballot_sync(0xffffffffu, inLeaf);
    const unsigned nodeMask = __ballot_sync(0xffffffffu, alive && !inLeaf);
    if (__popc(leafMask) * KDPT_LEAN_NODE_WEIGHT > __popc(nodeMask) * KDPT_LEAN_LEAF_WEIGHT) {
        if (inLeaf) {
            leanTestTri(T, S, H, eps, S.li);
            if (++S.li == S.le) alive = leanPop<RESTART>(T, S, H, sStack, lStack);
        }
        return alive;
    }
    if (alive && !inLeaf) {
        const uint4 nd = leanNode(T, sTop, S.cur);
        if ((nd.w & 3u) != kLeafFlag) {
            const int a = (int)(nd.w & 3u);
            const float split = __int_as_float((int)nd.x);
            const float oa = comp(S.o, a), da = comp(S.d, a), ia = comp(S.inv, a);
            const bool below = oa < split || (oa == split && da <= 0.0f);
            const unsigned nearc = below ? nd.z : nd.y, farc = below ? nd.y : nd.z;
            const float tp = (split - oa) * ia;
            if (da == 0.0f || tp <= 0.0f || tp >= S.tmax) S.cur = nearc;
            else if (tp <= S.tmin) S.cur = farc;
            else {
                const uint2 e = make_uint2(farc, __float_as_uint(S.tmax));
                if (RESTART) {
                    if (S.sp == kLeanRing) { S.ringBase = (S.ringBase + 1) & (kLeanRing - 1); S.sp--; S.dropped = true; }
                    sStack[((S.ringBase + S.sp) & (kLeanRing - 1)) * kLeanThreads] = e;
                } else {
                    if (S.sp < kLeanSmemStack) sStack[S.sp * kLeanThreads] = e;
                    else lStack[S.sp - kLeanSmemStack] = e;
                }
                S.sp++;
                S.cur = nearc; S.tmax = tp;
            }
        } else {
            const int start = (int)nd.x, count = (int)nd.y;
            if (count > 0 && nd.z == kNoChunks) { S.li = start; S.le = start + count; }
            else {
                if (count > 0) {
                    const LeanLeafRefs r = {T.chunkLo, T.chunkHi, T.superLo, T.superHi, T.triA, T.triB, T.nrmA, T.nrmB, T.triC, T.nrmC, T.entryOf};
                    H = leanChunkedLeaf(r, S.o, S.d, S.inv, S.zbound, S.slack, H, eps, start, count, nd.z);
                }
                alive = leanPop<RESTART>(T, S, H, sStack, lStack);
            }
        }
    }
    return alive;
}

#ifndef KDPT_LEAN_TRI_UNROLL
#define KDPT_LEAN_TRI_UNROLL 1      // triangles of a leaf tested per loop trip (their loads are issued together)
#endif
// "while-while" round: the lane descends to its next leaf, tests all of it, pops. The lanes of a warp wait for the longest
// descent and the fullest leaf among them (measured on dragon_5: 6 of 32 lanes active per instruction), but a lane's own
// instruction stream is minimal; this is the form that measured fastest (profiles/r2_lean_variants.txt).
template <bool RESTART>
__device__ __forceinline__ bool leanRoundLeaf(const Tree& T, bool alive, LeanState& S, Hit& H, float eps, uint2* sStack, uint2* lStack,
                                              const uint4* sTop)
{
    if (!alive) return false;
    unsigned cur = S.cur; float tmin = S.tmin, tmax = S.tmax;
    uint4 nd = leanNode(T, sTop, cur);
    while ((nd.w & 3u) != kLeafFlag) {
        const int a = (int)(nd.w & 3u);
        const float split = __int_as_float((int)nd.x);
        const float oa = comp(S.o, a), da = comp(S.d, a), ia = comp(S.inv, a);
        const bool below = oa < split || (oa == split && da <= 0.0f);
        const unsigned nearc = below ? nd.z : nd.y, farc = below ? nd.y : nd.z;
        const float tp = (split - oa) * ia;
        if (da == 0.0f || tp <= 0.0f || tp >= tmax) cur = nearc;
        else if (tp <= tmin) cur = farc;
        else {
            const uint2 e = make_uint2(farc, __float_as_uint(tmax));
            if (RESTART) {
                if (S.sp == kLeanRing) { S.ringBase = (S.ringBase + 1) & (kLeanRing - 1); S.sp--; S.dropped = true; }
                sStack[((S.ringBase + S.sp) & (kLeanRing - 1)) * kLeanThreads] = e;
            } else {
                if (S.sp < kLeanSmemStack) sStack[S.sp * kLeanThreads] = e;
                else lStack[S.sp - kLeanSmemStack] = e;
            }
            S.sp++;
            cur = nearc; tmax = tp;
        }
        nd = leanNode(T, sTop, cur);
    }
    S.tmax = tmax;
    const int start = (int)nd.x, count = (int)nd.y;
    if (count > 0) {
        if (nd.z == kNoChunks) {
            int i = start;
            const int e = start + count;
#if KDPT_LEAN_TRI_UNROLL >= 2
            for (; i + 1 < e; i += 2) {
                const TriHot ta = loadTri(T, i), tb = loadTri(T, i + 1);
                float u, v, z;
                if (rayTriHit(S.o, S.d, ta.v0, ta.e1, ta.e2, u, v, z)) leanAccept(T, S, H, eps, i, ta.canon, u, v, z);
                if (rayTriHit(S.o, S.d, tb.v0, tb.e1, tb.e2, u, v, z)) leanAccept(T, S, H, eps, i + 1, tb.canon, u, v, z);
            }
#endif
            for (; i < e; i++) leanTestTri(T, S, H, eps, i);
        } else {
            const LeanLeafRefs r = {T.chunkLo, T.chunkHi, T.superLo, T.superHi, T.triA, T.triB, T.nrmA, T.nrmB, T.triC, T.nrmC, T.entryOf};
            H = leanChunkedLeaf(r, S.o, S.d, S.inv, S.zbound, S.slack, H, eps, start, count, nd.z);
        }
    }
    return leanPop<RESTART>(T, S, H, sStack, lStack);
}

#ifndef KDPT_LEAN_NODE_CAP
#define KDPT_LEAN_NODE_CAP 8
#endif
#ifndef KDPT_LEAN_TRI_CAP
#define KDPT_LEAN_TRI_CAP 8
#endif
// "capped while-while" round: as leanRoundLeaf, but a lane takes at most NODE_CAP node steps and then at most TRI_CAP
// triangle tests per round and keeps its place (node, or cursor inside the leaf) for the next round. The first descent of a
// ray crosses ~20 levels, later ones 2-5, and leaves hold 0 to dozens of entries: uncapped, every round lasts as long as
// the warp's longest descent plus its fullest leaf while the average lane needs a third of either.
template <bool RESTART>
__device__ __forceinline__ bool leanRoundCapped(const Tree& T, bool alive, LeanState& S, Hit& H, float eps, uint2* sStack, uint2* lStack,
                                                const uint4* sTop)
{
    if (!alive) return false;
    if (S.li >= S.le) {
        unsigned cur = S.cur; float tmin = S.tmin, tmax = S.tmax;
        uint4 nd = leanNode(T, sTop, cur);
        int steps = KDPT_LEAN_NODE_CAP;
        while ((nd.w & 3u) != kLeafFlag) {
            if (steps-- == 0) { S.cur = cur; S.tmin = tmin; S.tmax = tmax; return true; }      // on with this descent next round
            const int a = (int)(nd.w & 3u);
            const float split = __int_as_float((int)nd.x);
            const float oa = comp(S.o, a), da = comp(S.d, a), ia = comp(S.inv, a);
            const bool below = oa < split || (oa == split && da <= 0.0f);
            const unsigned nearc = below ? nd.z : nd.y, farc = below ? nd.y : nd.z;
            const float tp = (split - oa) * ia;
            if (da == 0.0f || tp <= 0.0f || tp >= tmax) cur = nearc;
            else if (tp <= tmin) cur = farc;
            else {
                const uint2 e = make_uint2(farc, __float_as_uint(tmax));
                if (RESTART) {
                    if (S.sp == kLeanRing) { S.ringBase = (S.ringBase + 1) & (kLeanRing - 1); S.sp--; S.dropped = true; }
                    sStack[((S.ringBase + S.sp) & (kLeanRing - 1)) * kLeanThreads] = e;
                } else {
                    if (S.sp < kLeanSmemStack) sStack[S.sp * kLeanThreads] = e;
                    else lStack[S.sp - kLeanSmemStack] = e;
                }
                S.sp++;
                cur = nearc; tmax = tp;
            }
            nd = leanNode(T, sTop, cur);
        }
        S.cur = cur; S.tmin = tmin; S.tmax = tmax;
        const int start = (int)nd.x, count = (int)nd.y;
        if (count > 0 && nd.z != kNoChunks) {
            const LeanLeafRefs r = {T.chunkLo, T.chunkHi, T.superLo, T.superHi, T.triA, T.triB, T.nrmA, T.nrmB, T.triC, T.nrmC, T.entryOf};
            H = leanChunkedLeaf(r, S.o, S.d, S.inv, S.zbound, S.slack, H, eps, start, count, nd.z);
            return leanPop<RESTART>(T, S, H, sStack, lStack);
        }
        if (count <= 0) return leanPop<RESTART>(T, S, H, sStack, lStack);
        S.li = start; S.le = start + count;
    }
    const int e = S.le < S.li + KDPT_LEAN_TRI_CAP ? S.le : S.li + KDPT_LEAN_TRI_CAP;
    for (int i = S.li; i < e; i++) leanTestTri(T, S, H, eps, i);
    S.li = e;
    if (e < S.le) return true;              // the rest of this leaf next round
    return leanPop<RESTART>(T, S, H, sStack, lStack);
}

#ifndef KDPT_LEAN_V
#define KDPT_LEAN_V 1       // 1: leanRoundLeaf (whole leaf per round), 2: leanRoundUnit (one node step or one triangle per round),
#endif                      // 3: leanRoundCapped
template <bool RESTART>
__device__ __forceinline__ bool leanRound(const Tree& T, bool alive, LeanState& S, Hit& H, float eps, uint2* sStack, uint2* lStack,
                                          const uint4* sTop)
{
#if KDPT_LEAN_V == 2
    return leanRoundUnit<RESTART>(T, alive, S, H, eps, sStack, lStack, sTop);
#elif KDPT_LEAN_V == 3
    return leanRoundCapped<RESTART>(T, alive, S, H, eps, sStack, lStack, sTop);
#else
    return leanRoundLeaf<RESTART>(T, alive, S, H, eps, sStack, lStack, sTop);
#endif
}

} // namespace kdpt
