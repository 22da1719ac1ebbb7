// kdpt_flatten.cpp — see kdpt_flatten.h.
#include "kdpt_flatten.h"
#include <string.h>
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
           (chunkLo.size() + chunkHi.size() + superLo.size() + superHi.size()) * sizeof(HostF4);
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
} // namespace

bool flatten(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris, FlatTree& out,
             std::string& err)
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
        if (!leaf && (n.leftID <= i || n.rightID <= i || n.leftID >= numNodes || n.rightID >= numNodes)) {
            err = "flatten: node " + std::to_string(i) + " needs two children with larger IDs";
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
    // Leaf-local culling structure. The reference builder stops at depth 14 or as soon as a split keeps every
    // triangle on one side, so leaves hold hundreds to tens of thousands of triangles (986 on dragon_5, 62 080 on the
    // 1.6 M dragon). Leaves with at least kChunkMinLeaf entries are reordered along a Morton curve of the triangle
    // centroids and cut into chunks of 32 consecutive entries with a padded AABB each; the kernels test a ray against
    // the chunk boxes first and only stream the chunks it can touch. A triangle is never skipped if its exact test
    // could pass: the computed hit point lies on the triangle up to rounding (<< pad), hence inside its chunk box.
    // The order inside a leaf is free: the exact modes break ties on the canonical index and the reference-compatible
    // modes reconstruct the reference's sequential semantics from triOrig (see traceRef).
    std::vector<int> order((size_t)numTris);
    for (int i = 0; i < numTris; i++) order[i] = i;
    float sceneMax = 0.f;
    for (int a = 0; a < 3; a++) sceneMax = fmaxf(sceneMax, fmaxf(fabsf(out.rootLo[a]), fabsf(out.rootHi[a])));
    const float pad = fmaxf(1e-4f, 2e-5f * sceneMax);
    auto vert = [&](int i, int k, int a) -> float {
        const kdpt_tri_bare& t = tris[i];
        const float* x = &t.x1;          // x1 x2 x3 y1 y2 y3 z1 z2 z3
        return x[3 * a + k];
    };
    // pass 1: which leaves are chunked and where their chunks / groups go
    std::vector<int> chunkedLeaves;
    {
        size_t nChunks = 0, nSupers = 0;
        for (int nId = 0; nId < numNodes; nId++) {
            HostNode& h = out.nodes[nId];
            if ((h.w & 3u) != 3u) continue;
            h.z = 0xffffffffu;
            const int count = (int)h.y;
            if (count < kChunkMinLeaf) continue;
            h.z = (uint32_t)nChunks;
            chunkedLeaves.push_back(nId);
            const size_t nch = ((size_t)count + 31) / 32;
            nChunks += nch;
            if ((int)nch >= kSuperMinChunks) nSupers += (nch + 31) / 32;
        }
        out.chunkLo.resize(nChunks); out.chunkHi.resize(nChunks);
        out.superLo.resize(nSupers); out.superHi.resize(nSupers);
    }
    std::vector<size_t> superOfLeaf(chunkedLeaves.size());
    {
        size_t acc = 0;
        for (size_t q = 0; q < chunkedLeaves.size(); q++) {
            const size_t nch = ((size_t)out.nodes[chunkedLeaves[q]].y + 31) / 32;
            superOfLeaf[q] = acc;
            if ((int)nch >= kSuperMinChunks) acc += (nch + 31) / 32;
        }
    }
    // pass 2: leaves are independent -> spread over the host cores
    auto doLeaf = [&](size_t q) {
        const HostNode& h = out.nodes[chunkedLeaves[q]];
        const int start = (int)h.x, count = (int)h.y;
        const size_t firstChunk = h.z;
        // Morton order of centroids inside the leaf's own centroid bounds
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        std::vector<float> cen(3 * (size_t)count);
        for (int j = 0; j < count; j++)
            for (int a = 0; a < 3; a++) {
                float c = (vert(start + j, 0, a) + vert(start + j, 1, a) + vert(start + j, 2, a)) * (1.0f / 3.0f);
                if (!(c == c)) c = 0.f;
                cen[3 * (size_t)j + a] = c; lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c);
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
            keys[j] = std::make_pair(code, start + j);
        }
        std::sort(keys.begin(), keys.end());
        for (int j = 0; j < count; j++) order[start + j] = keys[j].second;
        size_t nch = 0;
        for (int c0 = 0; c0 < count; c0 += 32, nch++) {
            float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int j = c0; j < count && j < c0 + 32; j++)
                for (int k = 0; k < 3; k++)
                    for (int a = 0; a < 3; a++) { float x = vert(order[start + j], k, a); blo[a] = fminf(blo[a], x); bhi[a] = fmaxf(bhi[a], x); }
            out.chunkLo[firstChunk + nch] = HostF4{blo[0] - pad, blo[1] - pad, blo[2] - pad, asFloat(-1)};
            out.chunkHi[firstChunk + nch] = HostF4{bhi[0] + pad, bhi[1] + pad, bhi[2] + pad, 0.f};
        }
        // second level for very fat leaves: one box per 32 consecutive chunks (1024 triangles); the first chunk's
        // lo.w carries the index of the leaf's first group box (-1: none)
        if ((int)nch >= kSuperMinChunks) {
            size_t sp = superOfLeaf[q];
            out.chunkLo[firstChunk].w = asFloat((int)sp);
            for (size_t g = 0; g < nch; g += 32, sp++) {
                float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (size_t c = g; c < nch && c < g + 32; c++) {
                    const HostF4& l = out.chunkLo[firstChunk + c]; const HostF4& u = out.chunkHi[firstChunk + c];
                    blo[0] = fminf(blo[0], l.x); blo[1] = fminf(blo[1], l.y); blo[2] = fminf(blo[2], l.z);
                    bhi[0] = fmaxf(bhi[0], u.x); bhi[1] = fmaxf(bhi[1], u.y); bhi[2] = fmaxf(bhi[2], u.z);
                }
                out.superLo[sp] = HostF4{blo[0], blo[1], blo[2], 0.f};
                out.superHi[sp] = HostF4{bhi[0], bhi[1], bhi[2], 0.f};
            }
        }
    };
    if (nthreads == 1 || chunkedLeaves.size() < 2) { for (size_t q = 0; q < chunkedLeaves.size(); q++) doLeaf(q); }
    else {
        std::atomic<size_t> nextLeaf(0);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&]() { for (size_t q = nextLeaf.fetch_add(1); q < chunkedLeaves.size(); q = nextLeaf.fetch_add(1)) doLeaf(q); });
        for (auto& t : th) t.join();
    }
    out.numChunks = (int)out.chunkLo.size();

    out.triA.resize(numTris); out.triB.resize(numTris); out.triC.resize(numTris);
    out.nrmA.resize(numTris); out.nrmB.resize(numTris); out.nrmC.resize(numTris);
    out.triOrig.resize(numTris);
    auto fill = [&](int lo, int hi) {
        for (int i = lo; i < hi; i++) {
            const int src = order[i];
            const kdpt_tri_bare& t = tris[src];
            // e1 = v1 - v0, e2 = v2 - v0: binary32 subtractions, the same the reference performs per test
            float e1x = t.x2 - t.x1, e1y = t.y2 - t.y1, e1z = t.z2 - t.z1;
            float e2x = t.x3 - t.x1, e2y = t.y3 - t.y1, e2z = t.z3 - t.z1;
            out.triA[i] = HostF4{t.x1, t.y1, t.z1, e1x};
            out.triB[i] = HostF4{e1y, e1z, e2x, e2y};
            out.triC[i] = HostF2{e2z, asFloat(canon[src])};
            out.nrmA[i] = HostF4{t.nx1, t.ny1, t.nz1, t.nx2};
            out.nrmB[i] = HostF4{t.ny2, t.nz2, t.nx3, t.ny3};
            out.nrmC[i] = HostF2{t.nz3, asFloat(t.mtlIdx)};
            out.triOrig[i] = src;
        }
    };
    if (nthreads == 1) fill(0, numTris);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(fill, (int)((long long)numTris * t / nthreads), (int)((long long)numTris * (t + 1) / nthreads));
        for (auto& t : th) t.join();
    }
    return true;
}

} // namespace kdpt
