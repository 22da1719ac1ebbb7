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
//
// Like kdpt_core.cuh this header also compiles with g++ (KDPT_HD): tests/hostsim runs leanBegin / leanRound ray by ray on the
// CPU against the brute-force oracle, for both node forms and both stack forms.
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
// Shared-memory staging of the tree's top levels (north star / SURVEY §7 step 3): measured slower at 8, 10 and 11 levels
// (-5 % / -16 % / -32 % on dragon1m, profiles/r2_lean_variants.txt — the top levels already hit L1 and the copy costs
// residency), so it is not compiled in by default; -DKDPT_TOP_MAX_LEVELS=n with KDPT_TOP_LEVELS=n reproduces the sweep.
#ifndef KDPT_TOP_MAX_LEVELS
#define KDPT_TOP_MAX_LEVELS 0       // compile-time bound of the shared-memory copy of the tree's top levels (0: not compiled in)
#endif
// Two levels per record (Tree::wide, built by kdpt_flatten.cpp:buildWide): a node with both children inlined,
//   wide[2k]   = (split plane of the node, of its left child, of its right child, axis | left axis << 2 | right axis << 4)
//   wide[2k+1] = (left child's left, left child's right, right child's left, right child's right)
// where a child axis of 3 means "this child is a leaf" and its first slot names it. A reference is either a record index or
//   kWideLeaf | count << 27 | first entry    (count <= 14: the leaf's range inline — no record to fetch, an empty leaf costs nothing)
//   kWideLeaf | 15 << 27 | xnodes index      (chunked or bigger leaves: their 16-byte leaf record)
// One fetch decides two levels, so a descent is half as many dependent loads (dragon_5: 29 -> 12 per ray, leaf records included);
// the far subtrees of both levels are cut by their planes at once and go to the stack as up to three (reference, exit
// distance) entries, farthest first. Measured (profiles/r2_call20_empty_leaf_policies.txt, r2_call23_axis_select.txt): +6-10 % on
// the 10 M soup, +3 % on the 1.6 M dragon (with compSel below; +-1 % before it), -1 to -5 % on trees that live in L1/L2
// (dragon_5, bunny) where the extra plane arithmetic costs more than the saved loads — so kdpt_create picks this form for trees
// of 256 MB and more (kWideMinBytes).
constexpr unsigned kWideLeaf = 0x80000000u;
constexpr unsigned kWideIndexMask = (1u << 27) - 1u;
constexpr int kWideStack = 48;      // 3 entries per record level, at most 16 record levels (buildWide refuses deeper trees)
constexpr int kTopMaxNodes = KDPT_TOP_MAX_LEVELS > 0 ? (1 << KDPT_TOP_MAX_LEVELS) - 1 : 1;
constexpr unsigned kTopFlag = 0x80000000u;

// the block's copy of the top levels (Tree::top); call from every thread of the block before the first round
KDPT_HD void leanStageTop(const Tree& T, uint4* sTop)
{
#if KDPT_TOP_MAX_LEVELS > 0 && defined(__CUDACC__)
    const int m = T.numTop < kTopMaxNodes ? T.numTop : 0;       // a copy that does not fit is not used at all
    for (int k = threadIdx.x; k < m; k += blockDim.x) sTop[k] = __ldg(T.top + k);
    __syncthreads();
#endif
}
KDPT_HD unsigned leanRoot(const Tree& T)
{
#if KDPT_TOP_MAX_LEVELS > 0
    return (T.numTop > 0 && T.numTop < kTopMaxNodes) ? kTopFlag : 0u;
#else
    return 0u;
#endif
}
KDPT_HD uint4 leanNode(const Tree& T, const uint4* sTop, unsigned cur)
{
#if KDPT_TOP_MAX_LEVELS > 0
    if (cur & kTopFlag) return sTop[cur & ~kTopFlag];
#endif
    return KDPT_LDG(T.nodes + cur);
}

// Bin key of a ray for leaf-ordered tracing: (the reference leaf in which the ray is when it is first inside the root
// box, direction octant). Node IDs are DFS pre-order, so consecutive keys are neighbouring cells.
KDPT_HD unsigned rayBinKey(const Tree& T, V3 o, V3 d)
{
    const V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float t0, t1;
    unsigned node = 0;
    if (T.numNodes > 0 && slab(o, inv, T.rootLo, T.rootHi, t0, t1)) {
        // the point where the ray is when it is first inside the root box, then plain point location
        const float te = t0 > 0.0f ? t0 : 0.0f;
        const V3 p = mk(o.x + d.x * te, o.y + d.y * te, o.z + d.z * te);
        // (down to the reference's leaf: in the refined view the nodes below it have indices >= numRefNodes)
        for (int level = 0; level < kMaxDepth + 1; level++) {
            const uint4 nd = KDPT_LDG(T.nodes + node);
            const unsigned axis = nd.w & 3u;
            if (axis == kLeafFlag) break;
            const float split = __uint_as_float(nd.x);
            const float pc = comp(p, (int)axis), dc = comp(d, (int)axis);
            const bool left = pc < split || (pc == split && dc < 0.0f);
            const unsigned child = left ? nd.z : nd.y;
            if (child >= (unsigned)T.numRefNodes) break;
            node = child;
        }
    }
    return node * 8u + ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u));
}

// component k of a vector for the descent: two selects. Written as PTX because the compiler turns comp()'s nested conditional for
// the three vectors of a plane test into a branch on the axis — three-way divergence in the middle of every node step.
KDPT_HD float compSel(const V3& a, int k)
{
#if defined(__CUDA_ARCH__) && !defined(KDPT_LEAN_BRANCHY_AXIS)
    float r;
    asm("{ .reg .pred p1, p2;\n\t"
        "setp.eq.s32 p1, %4, 1;\n\t"
        "setp.eq.s32 p2, %4, 2;\n\t"
        "selp.f32 %0, %2, %1, p1;\n\t"
        "selp.f32 %0, %3, %0, p2; }" : "=&f"(r) : "f"(a.x), "f"(a.y), "f"(a.z), "r"(k));
    return r;
#else
    return comp(a, k);
#endif
}

struct HitCand { float t, nx, ny, nz; int mtl; };
// interpolated normal, nudged point and its distance (pathtrace.cu:993-998) of a triangle glm accepted: the rare part of
// the leaf loop, kept out of line so that the loop itself stays small
KDPT_NOINLINE_FN HitCand nudgedHit(const float4* __restrict__ nrmA, const float4* __restrict__ nrmB, const float2* __restrict__ nrmC,
                                          int i, float ox, float oy, float oz, float dx, float dy, float dz, float u, float v, float z, float eps)
{
    const float4 a = KDPT_LDG(nrmA + i), b = KDPT_LDG(nrmB + i);
    const float2 c = KDPT_LDG(nrmC + i);
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
    bool dropped;
};

// ray set-up shared by the kernels: false = the ray never enters the tree (H stays the "no hit" record)
template <bool WIDE = false>
KDPT_HD bool leanBegin(const Tree& T, V3 o, V3 d, float tMinIn, LeanState& S, Hit& H)
{
    initHit(H, tMinIn);
    S.o = o; S.d = d; S.sp = 0; S.ringBase = 0; S.dropped = false; S.cur = WIDE ? T.wideRoot : leanRoot(T);
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
template <bool RESTART, bool WIDE = false>
KDPT_HD bool leanPop(const Tree& T, LeanState& S, const Hit& H, const uint2* sStack, const uint2* lStack)
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
        S.cur = WIDE ? T.wideRoot : leanRoot(T); S.tmin = tmax; S.tmax = S.t1; S.ringBase = 0; S.dropped = false;
        return true;
    }
    return false;
}

// glm accepted entry i (canonical index canon): nudged distance, then the reference's acceptance rule
KDPT_HD void leanAccept(const Tree& T, const LeanState& S, Hit& H, float eps, int i, int canon, float u, float v, float z)
{
    const int ni = T.entryOf ? KDPT_LDG(T.entryOf + canon) : i;
    const HitCand hc = nudgedHit(T.nrmA, T.nrmB, T.nrmC, ni, S.o.x, S.o.y, S.o.z, S.d.x, S.d.y, S.d.z, u, v, z, eps);
    if (hc.t > 0.0f && (H.t > hc.t || (H.t == hc.t && H.tri >= 0 && canon < H.tri))) {
        H.t = hc.t; H.z = z; H.u = u; H.v = v; H.tri = canon; H.nx = hc.nx; H.ny = hc.ny; H.nz = hc.nz; H.mtl = hc.mtl;
    }
}
KDPT_HD void leanTestTri(const Tree& T, const LeanState& S, Hit& H, float eps, int i)
{
    const TriHot tr = loadTri(T, i);
    float u, v, z;
    if (!rayTriHit(S.o, S.d, tr.v0, tr.e1, tr.e2, u, v, z)) return;
    // a hit more than the exit slack (1e-3 world units) behind the best one cannot win on the nudged distance, which
    // differs from |d| z by at most the nudge (1e-4): skip its normal fetch, normalisation and square roots
    if (H.tri >= 0 && z > H.z + S.slack) return;
    leanAccept(T, S, H, eps, i, tr.canon, u, v, z);
}

// a leaf the refinement could not split (>= 128 coincident triangles): all of it at once, through its chunk boxes.
// Rare; kept out of line, with everything passed by value so that the callers' ray state stays in registers.
struct LeanLeafRefs {
    const float4 *chunkLo, *chunkHi, *superLo, *superHi, *triA, *triB, *nrmA, *nrmB;
    const float2 *triC, *nrmC;
    const int* entryOf;
};
KDPT_NOINLINE_FN Hit leanChunkedLeaf(LeanLeafRefs r, V3 o, V3 d, V3 inv, float zbound, float slack, Hit H, float eps,
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

// one plane of a record against the interval [tmin, tmax] of the ray: which of the two sides first, where it ends, and
// whether the other side follows (from there to tmax) — the near / far / both decision of the binary descent
#ifdef KDPT_LEAN_COUNT
static unsigned long long g_leanRounds, g_leanEmpty, g_leanTris, g_leanSteps;
#endif
struct PlaneCut { bool lowFirst, both; float tp; };
KDPT_HD PlaneCut leanCut(const LeanState& S, int a, float split, float tmin, float tmax)
{
    const float oa = compSel(S.o, a), da = compSel(S.d, a), ia = compSel(S.inv, a);
    const bool below = oa < split || (oa == split && da <= 0.0f);
    const float tp = (split - oa) * ia;
    const bool nearOnly = da == 0.0f || tp <= 0.0f || tp >= tmax;
    const bool farOnly = !nearOnly && tp <= tmin;
    PlaneCut c;
    c.lowFirst = below != farOnly; c.both = !nearOnly && !farOnly; c.tp = tp;
    return c;
}

template <bool RESTART, bool WIDE>
KDPT_HD void leanPush(LeanState& S, uint2* sStack, uint2* lStack, unsigned node, float texit)
{
    const uint2 e = make_uint2(node, __float_as_uint(texit));
    if (RESTART) {
        if (S.sp == kLeanRing) { S.ringBase = (S.ringBase + 1) & (kLeanRing - 1); S.sp--; S.dropped = true; }
        sStack[((S.ringBase + S.sp) & (kLeanRing - 1)) * kLeanThreads] = e;
    } else {
        if (S.sp < kLeanSmemStack) sStack[S.sp * kLeanThreads] = e;
        else lStack[S.sp - kLeanSmemStack] = e;
    }
    S.sp++;
}

// one step of the descent on a two-level record: cur becomes the first of the (up to four) grandchildren the ray visits, the
// others go to the stack farthest first
template <bool RESTART>
KDPT_HD void leanWideStep(const Tree& T, LeanState& S, uint2* sStack, uint2* lStack, unsigned& cur, float tmin, float& tmax)
{
    const uint4 A = KDPT_LDG(T.wide + 2 * (size_t)cur), B = KDPT_LDG(T.wide + 2 * (size_t)cur + 1);
    const PlaneCut c0 = leanCut(S, (int)(A.w & 3u), __uint_as_float(A.x), tmin, tmax);
    // the child the ray is in first, over [tmin, e1]; the other one, if the ray reaches it, over [tp, tmax]
    const float e1 = c0.both ? c0.tp : tmax;
    const unsigned ax1 = c0.lowFirst ? (A.w >> 2) & 3u : (A.w >> 4) & 3u, ax2 = c0.lowFirst ? (A.w >> 4) & 3u : (A.w >> 2) & 3u;
    const unsigned l1 = c0.lowFirst ? B.x : B.z, r1 = c0.lowFirst ? B.y : B.w;
    const unsigned l2 = c0.lowFirst ? B.z : B.x, r2 = c0.lowFirst ? B.w : B.y;
    const float s1 = __uint_as_float(c0.lowFirst ? A.y : A.z), s2 = __uint_as_float(c0.lowFirst ? A.z : A.y);
    if (c0.both) {
        if (ax2 == 3u) leanPush<RESTART, true>(S, sStack, lStack, l2, tmax);
        else {
            const PlaneCut c2 = leanCut(S, (int)ax2, s2, c0.tp, tmax);
            if (c2.both) leanPush<RESTART, true>(S, sStack, lStack, c2.lowFirst ? r2 : l2, tmax);
            leanPush<RESTART, true>(S, sStack, lStack, c2.lowFirst ? l2 : r2, c2.both ? c2.tp : tmax);
        }
    }
    if (ax1 == 3u) { cur = l1; tmax = e1; }
    else {
        const PlaneCut c1 = leanCut(S, (int)ax1, s1, tmin, e1);
        if (c1.both) leanPush<RESTART, true>(S, sStack, lStack, c1.lowFirst ? r1 : l1, e1);
        cur = c1.lowFirst ? l1 : r1; tmax = c1.both ? c1.tp : e1;
    }
}
// a leaf reference of that form: its range, inline or from its 16-byte record
KDPT_HD void leanWideLeaf(const Tree& T, unsigned cur, int& start, int& count, unsigned& chunk0)
{
    if (((cur >> 27) & 15u) == 15u) {
        const uint4 nd = KDPT_LDG(T.nodes + (cur & kWideIndexMask));
        start = (int)nd.x; count = (int)nd.y; chunk0 = nd.z;
    } else { start = (int)(cur & kWideIndexMask); count = (int)((cur >> 27) & 15u); chunk0 = kNoChunks; }
}
// one step of the descent on a 16-byte interior node
template <bool RESTART>
KDPT_HD void leanStep(const uint4 nd, LeanState& S, uint2* sStack, uint2* lStack, unsigned& cur, float tmin, float& tmax)
{
    const int a = (int)(nd.w & 3u);
    const float split = __int_as_float((int)nd.x);
    const float oa = compSel(S.o, a), da = compSel(S.d, a), ia = compSel(S.inv, a);
    const bool below = oa < split || (oa == split && da <= 0.0f);
    const unsigned nearc = below ? nd.z : nd.y, farc = below ? nd.y : nd.z;
    const float tp = (split - oa) * ia;
    if (da == 0.0f || tp <= 0.0f || tp >= tmax) cur = nearc;
    else if (tp <= tmin) cur = farc;
    else { leanPush<RESTART, false>(S, sStack, lStack, farc, tmax); cur = nearc; tmax = tp; }
}

// One round of the warp's loop for a lane that has a ray ("while-while"): descend to the next leaf, test all of it, pop.
// The lanes of a warp wait for the longest descent and the fullest leaf among them (measured on dragon_5: 6 of 32 lanes
// active per instruction), but a lane's own instruction stream is minimal. Two forms that keep more lanes busy were built
// and measured slower — one node step OR one triangle per round with the warp voting on the kind (10.7 lanes per
// instruction, 2.8x the thread instructions), and rounds capped at 8 node steps + 8 triangles (no gain) — see
// profiles/r2_lean_variants.txt (code: git history, commit "lean v3 (capped rounds) experiment"). A third, with the
// triangle tests postponed into a per-warp shared-memory ring and run 32 wide, lost too (commit "queued triangle tests
// experiment"): the lanes that walk on before their tests are done prune with a stale best hit; and the same dealing
// without speculation (commit "dealt leaf tests experiment") costs in scans and shared-memory traffic what it returns.
//
// Empty leaves (the surface-area heuristic cuts empty space off: two leaves of three a ray enters on dragon_5 are empty) end a
// lane's round like any other leaf. Popping them inside the descent — nested in a pop loop, or as ONE loop whose trips are a
// node step or a pop, with or without a cap on the trips — cuts the rounds per ray from 5.8 to 1.9 and was measured 8-20 %
// SLOWER on every workload (profiles/r2_call19_*.txt, r2_call20_*.txt; code: commit "empty-leaf policies of the lean loop"):
// the rounds get as long as the longest chain of descents among the lanes, finished lanes wait that much longer for their
// refill, and the pop code joins the divergent part of the loop.
template <bool RESTART, bool WIDE = false>
KDPT_HD bool leanRoundLeaf(const Tree& T, bool alive, LeanState& S, Hit& H, float eps, uint2* sStack, uint2* lStack,
                           const uint4* sTop)
{
    if (!alive) return false;
    unsigned cur = S.cur; float tmin = S.tmin, tmax = S.tmax;
    int start, count; unsigned chunk0 = kNoChunks;
    if (WIDE) {
        while (!(cur & kWideLeaf)) {
#ifdef KDPT_LEAN_COUNT
            g_leanSteps++;
#endif
            leanWideStep<RESTART>(T, S, sStack, lStack, cur, tmin, tmax);
        }
        leanWideLeaf(T, cur, start, count, chunk0);
    } else {
        uint4 nd = leanNode(T, sTop, cur);
        while ((nd.w & 3u) != kLeafFlag) {
#ifdef KDPT_LEAN_COUNT
            g_leanSteps++;
#endif
            leanStep<RESTART>(nd, S, sStack, lStack, cur, tmin, tmax);
            nd = leanNode(T, sTop, cur);
        }
        start = (int)nd.x; count = (int)nd.y; chunk0 = nd.z;
    }
    S.tmax = tmax;
#ifdef KDPT_LEAN_COUNT
    g_leanRounds++; g_leanTris += count; if (count == 0) g_leanEmpty++;
#endif
    if (count > 0) {
        if (chunk0 == kNoChunks) {
            for (int i = start; i < start + count; i++) leanTestTri(T, S, H, eps, i);
        } else {
            const LeanLeafRefs r = {T.chunkLo, T.chunkHi, T.superLo, T.superHi, T.triA, T.triB, T.nrmA, T.nrmB, T.triC, T.nrmC, T.entryOf};
            H = leanChunkedLeaf(r, S.o, S.d, S.inv, S.zbound, S.slack, H, eps, start, count, chunk0);
        }
    }
    return leanPop<RESTART, WIDE>(T, S, H, sStack, lStack);
}

template <bool RESTART, bool WIDE = false>
KDPT_HD bool leanRound(const Tree& T, bool alive, LeanState& S, Hit& H, float eps, uint2* sStack, uint2* lStack,
                       const uint4* sTop)
{
    return leanRoundLeaf<RESTART, WIDE>(T, alive, S, H, eps, sStack, lStack, sTop);
}

} // namespace kdpt
