// kdpt_flatten.cpp — see kdpt_flatten.h.
#include "kdpt_flatten.h"
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <math.h>
#include <thread>
#include <unordered_map>

namespace kdpt {

static const int kChunkMinLeaf = 128;     // leaves at least this large get 32-triangle chunks with bounding boxes
static const int kSuperMinChunks = 64;    // leaves with at least this many chunks (2048 triangles) get a second level

uint64_t FlatTree::bytes() const
{
    return nodes.size() * sizeof(HostNode) + (boxLo.size() + boxHi.size()) * sizeof(HostF4) +
           (triA.size() + triB.size() + nrmA.size() + nrmB.size()) * sizeof(HostF4) +
           (triC.size() + nrmC.size()) * sizeof(HostF2) + triOrig.size() * sizeof(int) +
           (chunkLo.size() + chunkHi.size() + superLo.size() + superHi.size()) * sizeof(HostF4) +
           (xnodes.size() + xwide.size()) * sizeof(HostNode) + (xtriA.size() + xtriB.size()) * sizeof(HostF4) + xtriC.size() * sizeof(HostF2) +
           (xchunkLo.size() + xchunkHi.size() + xsuperLo.size() + xsuperHi.size()) * sizeof(HostF4) + entryOf.size() * sizeof(int);
}

namespace {
struct TriKey {
    uint32_t w[19];
    bool operator==(const TriKey& o) const { return memcmp(w, o.w, sizeof w) == 0; }
};
struct TriKeyHash {
    size_t operator()(const TriKey& k) const
    {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 19; i++) { h ^= k.w[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
inline float asFloat(int i) { float f; memcpy(&f, &i, 4); return f; }
inline uint32_t asBits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

template <class F> void parallelFor(int n, int nthreads, F f)
{
    if (nthreads <= 1 || n < (1 << 14)) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(f, (int)((long long)n * t / nthreads), (int)((long long)n * (t + 1) / nthreads));
    for (auto& t : th) t.join();
}
template <class F> void parallelTasks(size_t n, int nthreads, F f)
{
    if (nthreads <= 1 || n < 2) { for (size_t q = 0; q < n; q++) f(q); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&]() { for (size_t q = next.fetch_add(1); q < n; q = next.fetch_add(1)) f(q); });
    for (auto& t : th) t.join();
}
inline float vertOf(const kdpt_tri_bare& t, int k, int a) { return (&t.x1)[3 * a + k]; }     // x1 x2 x3 y1 y2 y3 z1 z2 z3

// hot triangle records of the entries order[i]: (v0, e1, e2, canonical index)
void fillHot(const std::vector<int>& order, const kdpt_tri_bare* tris, const std::vector<int>& canon, int nthreads,
             BulkVector<HostF4>& A, BulkVector<HostF4>& B, BulkVector<HostF2>& C)
{
    parallelFor((int)order.size(), nthreads, [&](int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            const int src = order[i];
            const kdpt_tri_bare& t = tris[src];
            // e1 = v1 - v0, e2 = v2 - v0: binary32 subtractions, the same the reference performs per test
            float e1x = t.x2 - t.x1, e1y = t.y2 - t.y1, e1z = t.z2 - t.z1;
            float e2x = t.x3 - t.x1, e2y = t.y3 - t.y1, e2z = t.z3 - t.z1;
            A[i] = HostF4{t.x1, t.y1, t.z1, e1x};
            B[i] = HostF4{e1y, e1z, e2x, e2y};
            C[i] = HostF2{e2z, asFloat(canon[src])};
        }
    });
}

// Leaf-local culling structure. The reference builder stops at depth 14 or as soon as a split keeps every
// triangle on one side, so leaves hold hundreds to tens of thousands of triangles (986 on dragon_5, 62 080 on the
// 1.6 M dragon). Leaves with at least kChunkMinLeaf entries are reordered along a Morton curve of the triangle
// centroids and cut into chunks of 32 consecutive entries with a padded AABB each; the kernels test a ray against
// the chunk boxes first and only stream the chunks it can touch. A triangle is never skipped if its exact test
// could pass: the computed hit point lies on the triangle up to rounding (<< pad), hence inside its chunk box.
// The order inside a leaf is free: the exact modes break ties on the canonical index and the reference-compatible
// modes reconstruct the reference's sequential semantics from triOrig (see traceRef).
// `order[e]` = TriBare index of leaf entry e (reordered in place inside the chunked leaves); leaf records get
// .z = first chunk or ~0u.
void buildChunks(std::vector<HostNode>& nodesOut, std::vector<int>& order, const kdpt_tri_bare* tris, float pad, int nthreads,
                 std::vector<HostF4>& chunkLo, std::vector<HostF4>& chunkHi, std::vector<HostF4>& superLo, std::vector<HostF4>& superHi)
{
    const int numNodes = (int)nodesOut.size();
    // pass 1: which leaves are chunked and where their chunks / groups go
    std::vector<int> chunkedLeaves;
    std::vector<size_t> superOfLeaf;
    {
        size_t nChunks = 0, nSupers = 0;
        for (int nId = 0; nId < numNodes; nId++) {
            HostNode& h = nodesOut[nId];
            if ((h.w & 3u) != 3u) continue;
            h.z = 0xffffffffu;
            const int count = (int)h.y;
            if (count < kChunkMinLeaf) continue;
            h.z = (uint32_t)nChunks;
            chunkedLeaves.push_back(nId);
            superOfLeaf.push_back(nSupers);
            const size_t nch = ((size_t)count + 31) / 32;
            nChunks += nch;
            if ((int)nch >= kSuperMinChunks) nSupers += (nch + 31) / 32;
        }
        chunkLo.assign(nChunks, HostF4{0, 0, 0, 0}); chunkHi.assign(nChunks, HostF4{0, 0, 0, 0});
        superLo.assign(nSupers, HostF4{0, 0, 0, 0}); superHi.assign(nSupers, HostF4{0, 0, 0, 0});
    }
    // pass 2: leaves are independent -> spread over the host cores
    auto doLeaf = [&](size_t q) {
        const HostNode& h = nodesOut[chunkedLeaves[q]];
        const int start = (int)h.x, count = (int)h.y;
        const size_t firstChunk = h.z;
        // Morton order of centroids inside the leaf's own centroid bounds
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        std::vector<float> cen(3 * (size_t)count);
        for (int j = 0; j < count; j++) {
            const kdpt_tri_bare& t = tris[order[start + j]];
            for (int a = 0; a < 3; a++) {
                float c = (vertOf(t, 0, a) + vertOf(t, 1, a) + vertOf(t, 2, a)) * (1.0f / 3.0f);
                if (!(c == c)) c = 0.f;
                cen[3 * (size_t)j + a] = c; lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c);
            }
        }
        std::vector<std::pair<uint32_t, int>> keys((size_t)count);
        for (int j = 0; j < count; j++) {
            uint32_t code = 0;
            uint32_t qv[3];
            for (int a = 0; a < 3; a++) {
                float ext = hi[a] - lo[a];
                float f = ext > 0 ? (cen[3 * (size_t)j + a] - lo[a]) / ext : 0.f;
                int v = (int)(f * 1023.0f); qv[a] = (uint32_t)(v < 0 ? 0 : (v > 1023 ? 1023 : v));
            }
            for (int b = 9; b >= 0; b--) for (int a = 0; a < 3; a++) code = (code << 1) | ((qv[a] >> b) & 1u);
            keys[j] = std::make_pair(code, order[start + j]);
        }
        std::sort(keys.begin(), keys.end());
        for (int j = 0; j < count; j++) order[start + j] = keys[j].second;
        size_t nch = 0;
        for (int c0 = 0; c0 < count; c0 += 32, nch++) {
            float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int j = c0; j < count && j < c0 + 32; j++)
                for (int k = 0; k < 3; k++)
                    for (int a = 0; a < 3; a++) { float x = vertOf(tris[order[start + j]], k, a); blo[a] = fminf(blo[a], x); bhi[a] = fmaxf(bhi[a], x); }
            chunkLo[firstChunk + nch] = HostF4{blo[0] - pad, blo[1] - pad, blo[2] - pad, asFloat(-1)};
            chunkHi[firstChunk + nch] = HostF4{bhi[0] + pad, bhi[1] + pad, bhi[2] + pad, 0.f};
        }
        // second level for very fat leaves: one box per 32 consecutive chunks (1024 triangles); the first chunk's
        // lo.w carries the index of the leaf's first group box (-1: none)
        if ((int)nch >= kSuperMinChunks) {
            size_t sp = superOfLeaf[q];
            chunkLo[firstChunk].w = asFloat((int)sp);
            for (size_t g = 0; g < nch; g += 32, sp++) {
                float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (size_t c = g; c < nch && c < g + 32; c++) {
                    const HostF4& l = chunkLo[firstChunk + c]; const HostF4& u = chunkHi[firstChunk + c];
                    blo[0] = fminf(blo[0], l.x); blo[1] = fminf(blo[1], l.y); blo[2] = fminf(blo[2], l.z);
                    bhi[0] = fmaxf(bhi[0], u.x); bhi[1] = fmaxf(bhi[1], u.y); bhi[2] = fmaxf(bhi[2], u.z);
                }
                superLo[sp] = HostF4{blo[0], blo[1], blo[2], 0.f};
                superHi[sp] = HostF4{bhi[0], bhi[1], bhi[2], 0.f};
            }
        }
    };
    parallelTasks(chunkedLeaves.size(), nthreads, doLeaf);
}

// ---- leaf refinement (exact modes) ---------------------------------------------------------------------------------
// The reference builder is a spatial-median split that stops at depth 14 or whenever one side keeps every triangle
// (KDnode.cpp:164-191), so its leaves are fat. The exact modes must return the brute-force closest hit, not walk the
// reference's leaves in the reference's order, so for them every fat leaf is split further, inside its own cell, by a
// binned surface-area-heuristic KD builder. Inclusion uses the reference's own rule with a scene-scaled slack: a
// triangle goes to every side its AABB padded by `pad` reaches (KDnode.cpp:179-186 uses 1e-4), so a computed hit point
// — on the triangle up to rounding << pad — always lies in a leaf that holds the triangle; the traversal's arguments
// (near-first order, early exit with kExitSlack) are those that already hold on the reference tree.
struct SubTree {
    std::vector<HostNode> nodes;    // local indices; nodes[0] is the reference leaf itself
    std::vector<int> entries;       // TriBare indices (unique inside the reference leaf)
    int maxDepth = 0, maxLeaf = 0;
};
struct RefineItem { float lo[3], hi[3]; int tri; };

void refineRec(SubTree& st, int self, std::vector<RefineItem>& items, const float* cLo, const float* cHi, int depthLeft, int depth,
               const FlattenOpts& opts)
{
    const int n = (int)items.size();
    auto makeLeaf = [&]() {
        HostNode h; h.x = (uint32_t)st.entries.size(); h.y = (uint32_t)n; h.z = 0xffffffffu; h.w = 3u;
        for (const RefineItem& it : items) st.entries.push_back(it.tri);
        st.nodes[self] = h;
        if (n > st.maxLeaf) st.maxLeaf = n;
        if (depth > st.maxDepth) st.maxDepth = depth;
    };
    if (n <= opts.refineLeafMax || depthLeft <= 0) { makeLeaf(); return; }
    const double ext[3] = {(double)cHi[0] - cLo[0], (double)cHi[1] - cLo[1], (double)cHi[2] - cLo[2]};
    const double saCell = 2.0 * (ext[0] * ext[1] + ext[1] * ext[2] + ext[2] * ext[0]);
    const int K = 16;
    double bestCost = (double)n; int bestAxis = -1; float bestPlane = 0.f;
    for (int a = 0; a < 3; a++) {
        float mn = INFINITY, mx = -INFINITY;
        for (const RefineItem& it : items) { mn = fminf(mn, it.lo[a]); mx = fmaxf(mx, it.hi[a]); }
        const float tlo = fmaxf(cLo[a], mn), thi = fminf(cHi[a], mx);
        if (!(thi > tlo)) continue;
        int cntL[K + 2] = {0}, cntH[K + 2] = {0};
        const float scale = (float)K / (thi - tlo);
        for (const RefineItem& it : items) {
            // planes p_k = tlo + (thi - tlo) k / K: the item is on the left of every plane above its lo, on the right
            // of every plane up to its hi (counts drive the heuristic only; the partition below is exact)
            int kl = it.lo[a] < tlo ? 0 : (int)((it.lo[a] - tlo) * scale) + 1;
            int kh = it.hi[a] >= thi ? K : (int)((it.hi[a] - tlo) * scale);
            cntL[kl > K + 1 ? K + 1 : kl]++;
            cntH[kh < 0 ? 0 : kh]++;
        }
        const int o1 = (a + 1) % 3, o2 = (a + 2) % 3;
        const double cross = ext[o1] * ext[o2], perim = ext[o1] + ext[o2];
        int nl = 0, nr = n;
        for (int k = 0; k <= K; k++) {
            nl += cntL[k];
            if (k > 0) nr -= cntH[k - 1];
            float plane = tlo + (thi - tlo) * ((float)k / (float)K);
            int nrk = nr;
            if (k == K) { plane = nextafterf(thi, INFINITY); nrk = 0; }      // cut the empty space above the triangles
            if (!(plane > cLo[a]) || !(plane < cHi[a])) continue;
            const double wl = (double)plane - cLo[a], wr = (double)cHi[a] - plane;
            const double saL = 2.0 * (cross + wl * perim), saR = 2.0 * (cross + wr * perim);
            double cost = opts.traversalCost + (saCell > 0 ? (saL * nl + saR * nrk) / saCell : (double)(nl + nrk));
            if (nl == 0 || nrk == 0) cost *= 0.8;                            // an empty child ends rays early
            if (cost < bestCost) { bestCost = cost; bestAxis = a; bestPlane = plane; }
        }
    }
    if (bestAxis < 0) { makeLeaf(); return; }
    std::vector<RefineItem> left, right;
    for (const RefineItem& it : items) {
        if (it.lo[bestAxis] < bestPlane) left.push_back(it);
        if (it.hi[bestAxis] >= bestPlane) right.push_back(it);
    }
    if (((int)left.size() == n && (int)right.size() == n) || left.size() + right.size() > (size_t)n * 2) { makeLeaf(); return; }
    { std::vector<RefineItem>().swap(items); }
    float lHi[3] = {cHi[0], cHi[1], cHi[2]}, rLo[3] = {cLo[0], cLo[1], cLo[2]};
    lHi[bestAxis] = bestPlane; rLo[bestAxis] = bestPlane;
    // children in DFS pre-order, like the reference's IDs: the near part of a subtree sits next to its parent
    const int l = (int)st.nodes.size();
    st.nodes.push_back(HostNode{0, 0, 0, 0});
    refineRec(st, l, left, cLo, lHi, depthLeft - 1, depth + 1, opts);
    const int r = (int)st.nodes.size();
    st.nodes.push_back(HostNode{0, 0, 0, 0});
    refineRec(st, r, right, rLo, cHi, depthLeft - 1, depth + 1, opts);
    HostNode h; h.x = asBits(bestPlane); h.y = (uint32_t)r; h.z = (uint32_t)l; h.w = (uint32_t)bestAxis;
    st.nodes[self] = h;
}

// out.xnodes + the entry list of the refined tree (xorder[e] = TriBare index)
void refineLeaves(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, const std::vector<int>& canon,
                  const std::vector<int>& depth, float pad, const FlattenOpts& opts, int nthreads, FlatTree& out, std::vector<int>& xorder)
{
    std::vector<int> leaves;
    for (int i = 0; i < numNodes; i++) if ((out.nodes[i].w & 3u) == 3u) leaves.push_back(i);
    // big leaves first, so that the longest task does not start last
    std::sort(leaves.begin(), leaves.end(), [&](int a, int b) { return out.nodes[a].y != out.nodes[b].y ? out.nodes[a].y > out.nodes[b].y : a < b; });
    std::vector<SubTree> subs(leaves.size());
    parallelTasks(leaves.size(), nthreads, [&](size_t q) {
        const int nId = leaves[q];
        const int start = (int)out.nodes[nId].x, count = (int)out.nodes[nId].y;
        SubTree& st = subs[q];
        // unique triangles of the leaf (identical 76-byte entries test identically; the smallest index is reported)
        std::vector<int> ids((size_t)count);
        for (int j = 0; j < count; j++) ids[j] = canon[start + j];
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        std::vector<RefineItem> items(ids.size());
        for (size_t j = 0; j < ids.size(); j++) {
            const kdpt_tri_bare& t = tris[ids[j]];
            RefineItem& it = items[j];
            it.tri = ids[j];
            for (int a = 0; a < 3; a++) {
                const float v0 = vertOf(t, 0, a), v1 = vertOf(t, 1, a), v2 = vertOf(t, 2, a);
                float mn = fminf(v0, fminf(v1, v2)), mx = fmaxf(v0, fmaxf(v1, v2));
                if (!(mn == mn) || !(mx == mx) || !(v0 == v0) || !(v1 == v1) || !(v2 == v2)) { mn = -INFINITY; mx = INFINITY; }   // NaN vertex: everywhere
                it.lo[a] = mn - pad; it.hi[a] = mx + pad;
            }
        }
        st.nodes.push_back(HostNode{0, 0, 0, 0});
        refineRec(st, 0, items, nodes[nId].mins, nodes[nId].maxs, 31 - depth[nId], depth[nId], opts);
    });
    // global numbering: xnodes[i] = reference node i for i < numNodes (a refined leaf becomes the root of its subtree),
    // the new nodes of every subtree follow, subtree by subtree in node-ID order
    std::vector<size_t> subOf((size_t)numNodes, (size_t)-1);
    for (size_t q = 0; q < leaves.size(); q++) subOf[leaves[q]] = q;
    std::vector<size_t> nodeBase(leaves.size()), entryBase(leaves.size());
    size_t nn = (size_t)numNodes, ne = 0;
    for (int i = 0; i < numNodes; i++) {
        const size_t q = subOf[i];
        if (q == (size_t)-1) continue;
        nodeBase[q] = nn; entryBase[q] = ne;
        nn += subs[q].nodes.size() - 1; ne += subs[q].entries.size();
        if (subs[q].maxDepth > out.xMaxDepth) out.xMaxDepth = subs[q].maxDepth;
        if (subs[q].maxLeaf > out.xMaxLeaf) out.xMaxLeaf = subs[q].maxLeaf;
    }
    out.xnodes.assign(nn, HostNode{0, 0, 0, 0});
    xorder.assign(ne, 0);
    for (int i = 0; i < numNodes; i++) if (subOf[i] == (size_t)-1) out.xnodes[i] = out.nodes[i];
    parallelTasks(leaves.size(), nthreads, [&](size_t q) {
        const SubTree& st = subs[q];
        const int nId = leaves[q];
        auto global = [&](uint32_t local) { return local == 0 ? (uint32_t)nId : (uint32_t)(nodeBase[q] + local - 1); };
        const uint32_t rootParent = out.nodes[nId].w & ~3u;
        for (size_t k = 0; k < st.nodes.size(); k++) {
            HostNode h = st.nodes[k];
            const uint32_t self = global((uint32_t)k);
            if ((h.w & 3u) == 3u) h.x += (uint32_t)entryBase[q];
            else {
                h.y = global(h.y); h.z = global(h.z);
                out.xnodes[h.y].w |= (self + 1) << 2; out.xnodes[h.z].w |= (self + 1) << 2;     // children come later in k
            }
            HostNode& dst = out.xnodes[self];
            dst.x = h.x; dst.y = h.y; dst.z = h.z; dst.w = (dst.w & ~3u) | (h.w & 3u);
            if (k == 0) dst.w |= rootParent;
        }
        for (size_t e = 0; e < st.entries.size(); e++) xorder[entryBase[q] + e] = st.entries[e];
    });
}
} // namespace

// The refined tree, two levels per record (kdpt_lean.cuh: kWide*). A record is made for the root and for every interior
// grandchild of a record's node, numbered in depth-first order (left subtrees first, like the nodes themselves); a leaf is
// named by its parent: (first entry, count) inline when count <= 14 and the leaf has no chunk boxes, else the index of its
// 16-byte leaf record in xnodes. Left empty if an index does not fit the 27 bits of a reference or the tree is too deep for
// the kernels' stack (3 entries per record level).
static void buildWide(FlatTree& out)
{
    const std::vector<HostNode>& N = out.xnodes;
    out.xwide.clear(); out.xwideRoot = 0;
    if (N.empty() || N.size() >= (1u << 27) || (size_t)out.xNumTris >= (1u << 27)) return;
    auto isLeaf = [&](uint32_t id) { return (N[id].w & 3u) == 3u; };
    auto leafRef = [&](uint32_t id) -> uint32_t {
        const HostNode& h = N[id];
        if (h.z == 0xffffffffu && h.y <= 14u) return 0x80000000u | (h.y << 27) | (h.y ? h.x : 0u);
        return 0x80000000u | (15u << 27) | id;
    };
    if (isLeaf(0)) { out.xwideRoot = leafRef(0); out.xwide.assign(2, HostNode{0, 0, 0, 0}); return; }
    // pass 1: number the records depth-first; pass 2: fill them
    std::vector<uint32_t> recOf(N.size(), 0xffffffffu), order;
    std::vector<uint32_t> stack{0u};
    std::vector<int> level(N.size(), 0);
    while (!stack.empty()) {
        const uint32_t id = stack.back(); stack.pop_back();
        if (level[id] >= 16) return;                        // kernels' stack: 3 entries per record level, kWideStack = 48
        recOf[id] = (uint32_t)order.size(); order.push_back(id);
        const uint32_t kids[2] = {N[id].z, N[id].y};        // left, right
        uint32_t grand[4]; int ng = 0;
        for (int k = 0; k < 2; k++)
            if (!isLeaf(kids[k])) { grand[ng++] = N[kids[k]].z; grand[ng++] = N[kids[k]].y; }
        for (int k = ng - 1; k >= 0; k--) if (!isLeaf(grand[k])) { level[grand[k]] = level[id] + 1; stack.push_back(grand[k]); }
    }
    auto ref = [&](uint32_t id) { return isLeaf(id) ? leafRef(id) : recOf[id]; };
    out.xwide.resize(2 * order.size());
    for (size_t r = 0; r < order.size(); r++) {
        const HostNode& n = N[order[r]];
        HostNode a{n.x, 0u, 0u, n.w & 3u}, b{0u, 0u, 0u, 0u};
        const uint32_t kids[2] = {n.z, n.y};
        for (int k = 0; k < 2; k++) {
            const HostNode& c = N[kids[k]];
            uint32_t axis = c.w & 3u, c0, c1 = 0u;
            if (axis == 3u) c0 = leafRef(kids[k]);
            else { c0 = ref(c.z); c1 = ref(c.y); (k == 0 ? a.y : a.z) = c.x; }
            a.w |= axis << (2 + 2 * k);
            if (k == 0) { b.x = c0; b.y = c1; } else { b.z = c0; b.w = c1; }
        }
        out.xwide[2 * r] = a; out.xwide[2 * r + 1] = b;
    }
    out.xwideRoot = 0;
}

FlattenOpts::FlattenOpts()
{
    if (const char* e = getenv("KDPT_REFINE_LEAF")) refineLeafMax = atoi(e);
    if (const char* e = getenv("KDPT_REFINE_CT")) traversalCost = (float)atof(e);
    if (const char* e = getenv("KDPT_TOP_LEVELS")) topLevels = atoi(e);
    if (const char* e = getenv("KDPT_WIDE")) wideNodes = atoi(e);
}

bool flatten(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris, FlatTree& out,
             std::string& err, const FlattenOpts& opts)
{
    static_assert(sizeof(kdpt_node_bare) == 64, "NodeBare ABI");
    static_assert(sizeof(kdpt_tri_bare) == 76, "TriBare ABI");
    out = FlatTree();
    if (numNodes < 0 || numTris < 0 || (numNodes > 0 && !nodes) || (numTris > 0 && !tris)) {
        err = "flatten: null array or negative count";
        return false;
    }
    out.numNodes = numNodes; out.numTris = numTris;
    if (numNodes == 0) return true;
    // the reference looks for the first node with parentID == -1 (pathtrace.cu:909-916); its flatten always puts
    // the root at index 0 with index == ID (scene.cpp:889-895)
    if (nodes[0].parentID != -1) { err = "flatten: nodes[0] is not the root"; return false; }
    out.nodes.resize(numNodes); out.boxLo.resize(numNodes); out.boxHi.resize(numNodes);
    std::vector<int> depth(numNodes, -1);
    depth[0] = 0;
    for (int i = 0; i < numNodes; i++) {
        const kdpt_node_bare& n = nodes[i];
        if (n.ID != i) { err = "flatten: node index != ID at " + std::to_string(i); return false; }
        bool leaf = n.leftID == -1 && n.rightID == -1;
        if (!leaf && (n.leftID <= i || n.rightID <= i || n.leftID >= numNodes || n.rightID >= numNodes || n.leftID == n.rightID)) {
            err = "flatten: node " + std::to_string(i) + " needs two different children with larger IDs";
            return false;
        }
        // the links must agree in both directions: a node listed as a child by two parents (a DAG) would make paths
        // through child links longer than the parent-link depth checked below, and the per-ray stacks are sized by it
        if (!leaf && (nodes[n.leftID].parentID != i || nodes[n.rightID].parentID != i)) {
            err = "flatten: a child of node " + std::to_string(i) + " names another parent";
            return false;
        }
        if (i > 0 && (n.parentID < 0 || n.parentID >= i)) { err = "flatten: bad parentID at " + std::to_string(i); return false; }
        if (i > 0) {
            const kdpt_node_bare& p = nodes[n.parentID];
            if (p.leftID != i && p.rightID != i) { err = "flatten: parent/child mismatch at " + std::to_string(i); return false; }
            depth[i] = depth[n.parentID] + 1;
            if (depth[i] > out.maxDepth) out.maxDepth = depth[i];
        }
        if ((unsigned)n.axis > 2u) { err = "flatten: axis out of range at " + std::to_string(i); return false; }
        HostNode h;
        uint32_t parentBits = (uint32_t)(n.parentID + 1) << 2;
        if (leaf) {
            int start = n.triIdStart, count = n.triIdSize;
            if (count <= 0) { start = 0; count = 0; }
            if (start < 0 || (int64_t)start + count > numTris) { err = "flatten: leaf range outside TriBare[] at " + std::to_string(i); return false; }
            h.x = (uint32_t)start; h.y = (uint32_t)count; h.z = 0; h.w = 3u | parentBits;
            if (count > out.maxLeaf) out.maxLeaf = count;
        } else {
            // split plane = the boundary shared by the two children = left child's max on this node's axis
            // (== parent bbox centre, KDnode.cpp:211; NodeBare.splitPos is NOT the plane, Appendix A #14)
            float split = nodes[n.leftID].maxs[n.axis];
            if (asBits(split) != asBits(nodes[n.rightID].mins[n.axis])) {
                err = "flatten: children of node " + std::to_string(i) + " do not share a split plane";
                return false;
            }
            h.x = asBits(split); h.y = (uint32_t)n.rightID; h.z = (uint32_t)n.leftID; h.w = (uint32_t)n.axis | parentBits;
        }
        out.nodes[i] = h;
        out.boxLo[i] = HostF4{n.mins[0], n.mins[1], n.mins[2], 0.f};
        out.boxHi[i] = HostF4{n.maxs[0], n.maxs[1], n.maxs[2], asFloat(n.axis)};
    }
    if (out.maxDepth + 1 > 32) { err = "flatten: tree deeper than 32 levels"; return false; }
    {
        // leaf ranges must not overlap (the leaf-local reordering below works on them in parallel, and a triangle
        // belongs to the leaf the reference put it in)
        std::vector<std::pair<int, int>> ranges;
        for (int i = 0; i < numNodes; i++)
            if ((out.nodes[i].w & 3u) == 3u && out.nodes[i].y > 0) ranges.emplace_back((int)out.nodes[i].x, (int)out.nodes[i].y);
        std::sort(ranges.begin(), ranges.end());
        for (size_t k = 1; k < ranges.size(); k++)
            if (ranges[k - 1].first + ranges[k - 1].second > ranges[k].first) { err = "flatten: leaf triangle ranges overlap"; return false; }
    }
    for (int a = 0; a < 3; a++) { out.rootLo[a] = nodes[0].mins[a]; out.rootHi[a] = nodes[0].maxs[a]; }
    out.tagCand[0] = nodes[0].leftID; out.tagCand[1] = nodes[0].rightID;
    if (numNodes > 1) { out.tagCand[2] = nodes[1].leftID; out.tagCand[3] = nodes[1].rightID; }

    // canonical index = first TriBare entry holding identical 76 bytes (what kdpt_hit.tri reports)
    std::vector<int> canon((size_t)numTris);
    unsigned hwc = std::thread::hardware_concurrency();
    const int nthreads = numTris < (1 << 16) ? 1 : (int)(hwc == 0 ? 4 : (hwc > 32 ? 32 : hwc));
    {
        // Sort-based, cache-friendly and parallel (a hash map of 76-byte keys costs a few cache misses per entry — it
        // was most of pathtraceInit on the multi-million-triangle scenes): entries are keyed by a 64-bit hash of their
        // 76 bytes, partitioned by the top hash bits, each part sorted by (hash, index); inside a run of equal hashes
        // the smallest index whose bytes match is the canonical one (runs are almost always one distinct triangle;
        // true hash collisions are resolved by comparing against the run's earlier distinct heads).
        struct HashIdx { uint64_t h; int i; };
        BulkVector<HashIdx> keys((size_t)numTris);
        const int parts = nthreads == 1 ? 1 : nthreads * 8;
        std::vector<size_t> partBegin((size_t)parts + 1, 0);
        auto partOf = [&](uint64_t h) { return (int)(((h >> 32) * (uint64_t)parts) >> 32); };
        auto canonPart = [&](size_t lo, size_t hi) {
            std::sort(keys.begin() + lo, keys.begin() + hi, [](const HashIdx& x, const HashIdx& y) { return x.h != y.h ? x.h < y.h : x.i < y.i; });
            size_t r = lo;
            while (r < hi) {
                size_t e = r + 1;
                while (e < hi && keys[e].h == keys[r].h) e++;
                // run [r, e): heads = distinct contents seen so far in this run, in index order
                int headsInline[4]; int numHeads = 0; std::vector<int> headsMore;
                for (size_t k = r; k < e; k++) {
                    const int i = keys[k].i;
                    int found = -1;
                    for (int q = 0; q < numHeads && found < 0; q++) {
                        const int hd = q < 4 ? headsInline[q] : headsMore[q - 4];
                        if (memcmp(&tris[hd], &tris[i], sizeof(kdpt_tri_bare)) == 0) found = hd;
                    }
                    if (found < 0) { found = i; if (numHeads < 4) headsInline[numHeads] = i; else headsMore.push_back(i); numHeads++; }
                    canon[i] = found;
                }
                r = e;
            }
        };
        if (nthreads == 1) {
            TriKeyHash H;
            for (int i = 0; i < numTris; i++) { TriKey k; memcpy(k.w, &tris[i], sizeof k.w); keys[i] = HashIdx{(uint64_t)H(k), i}; }
            canonPart(0, (size_t)numTris);
        } else {
            // pass 1: hash + per-thread part counts; pass 2: stable scatter into parts; pass 3: sort + resolve parts
            BulkVector<uint64_t> hashes((size_t)numTris);
            std::vector<std::vector<size_t>> counts((size_t)nthreads, std::vector<size_t>((size_t)parts, 0));
            auto range = [&](int t, int& lo, int& hi) { lo = (int)((long long)numTris * t / nthreads); hi = (int)((long long)numTris * (t + 1) / nthreads); };
            std::vector<std::thread> th;
            for (int t = 0; t < nthreads; t++)
                th.emplace_back([&, t]() {
                    int lo, hi; range(t, lo, hi);
                    TriKeyHash H;
                    for (int i = lo; i < hi; i++) { TriKey k; memcpy(k.w, &tris[i], sizeof k.w); hashes[i] = (uint64_t)H(k); counts[t][partOf(hashes[i])]++; }
                });
            for (auto& t : th) t.join();
            th.clear();
            std::vector<std::vector<size_t>> offs((size_t)nthreads, std::vector<size_t>((size_t)parts, 0));
            size_t acc = 0;
            for (int p = 0; p < parts; p++) {
                partBegin[p] = acc;
                for (int t = 0; t < nthreads; t++) { offs[t][p] = acc; acc += counts[t][p]; }
            }
            partBegin[parts] = acc;
            for (int t = 0; t < nthreads; t++)
                th.emplace_back([&, t]() {
                    int lo, hi; range(t, lo, hi);
                    for (int i = lo; i < hi; i++) keys[offs[t][partOf(hashes[i])]++] = HashIdx{hashes[i], i};
                });
            for (auto& t : th) t.join();
            th.clear();
            std::atomic<int> nextPart(0);
            for (int t = 0; t < nthreads; t++)
                th.emplace_back([&]() { for (int p = nextPart.fetch_add(1); p < parts; p = nextPart.fetch_add(1)) canonPart(partBegin[p], partBegin[p + 1]); });
            for (auto& t : th) t.join();
        }
    }
    float sceneMax = 0.f;
    for (int a = 0; a < 3; a++) sceneMax = fmaxf(sceneMax, fmaxf(fabsf(out.rootLo[a]), fabsf(out.rootHi[a])));
    const float pad = fmaxf(1e-4f, 2e-5f * sceneMax);

    // reference tree: entry i holds TriBare order[i] (identity outside the chunked leaves)
    std::vector<int> order((size_t)numTris);
    for (int i = 0; i < numTris; i++) order[i] = i;
    buildChunks(out.nodes, order, tris, pad, nthreads, out.chunkLo, out.chunkHi, out.superLo, out.superHi);
    out.numChunks = (int)out.chunkLo.size();
    out.triA.resize(numTris); out.triB.resize(numTris); out.triC.resize(numTris);
    out.nrmA.resize(numTris); out.nrmB.resize(numTris); out.nrmC.resize(numTris);
    out.triOrig.resize(numTris);
    fillHot(order, tris, canon, nthreads, out.triA, out.triB, out.triC);
    parallelFor(numTris, nthreads, [&](int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            const kdpt_tri_bare& t = tris[order[i]];
            out.nrmA[i] = HostF4{t.nx1, t.ny1, t.nz1, t.nx2};
            out.nrmB[i] = HostF4{t.ny2, t.nz2, t.nx3, t.ny3};
            out.nrmC[i] = HostF2{t.nz3, asFloat(t.mtlIdx)};
            out.triOrig[i] = order[i];
        }
    });

    // refined tree for the exact modes
    if (opts.refineLeafMax > 0) {
        std::vector<int> xorder;
        refineLeaves(nodes, numNodes, tris, canon, depth, pad, opts, nthreads, out, xorder);
        buildChunks(out.xnodes, xorder, tris, pad, nthreads, out.xchunkLo, out.xchunkHi, out.xsuperLo, out.xsuperHi);
        out.xNumTris = (int)xorder.size();
        out.xtriA.resize(xorder.size()); out.xtriB.resize(xorder.size()); out.xtriC.resize(xorder.size());
        fillHot(xorder, tris, canon, nthreads, out.xtriA, out.xtriB, out.xtriC);
        out.entryOf.resize(numTris);
        parallelFor(numTris, nthreads, [&](int lo, int hi) { for (int i = lo; i < hi; i++) out.entryOf[order[i]] = i; });
        if (opts.wideNodes > 0 || (opts.wideNodes < 0 && out.bytes() >= kWideMinBytes)) buildWide(out);
        if (opts.topLevels > 0) {
            // breadth-first copy of the first topLevels levels (at most 2^levels - 1 nodes)
            const int levels = opts.topLevels > 11 ? 11 : opts.topLevels;
            std::vector<uint32_t> ids{0u}; std::vector<int> lvl{0};
            for (size_t q = 0; q < ids.size(); q++) {
                const HostNode& h = out.xnodes[ids[q]];
                if ((h.w & 3u) == 3u || lvl[q] + 1 >= levels) continue;
                ids.push_back(h.z); lvl.push_back(lvl[q] + 1);
                ids.push_back(h.y); lvl.push_back(lvl[q] + 1);
            }
            std::unordered_map<uint32_t, uint32_t> pos;
            for (size_t q = 0; q < ids.size(); q++) pos[ids[q]] = (uint32_t)q;
            out.xtop.resize(ids.size());
            for (size_t q = 0; q < ids.size(); q++) {
                HostNode h = out.xnodes[ids[q]];
                if ((h.w & 3u) != 3u) {
                    auto a = pos.find(h.y), b = pos.find(h.z);
                    if (a != pos.end()) h.y = a->second | 0x80000000u;
                    if (b != pos.end()) h.z = b->second | 0x80000000u;
                }
                out.xtop[q] = h;
            }
        }
    }
    return true;
}

} // namespace kdpt
