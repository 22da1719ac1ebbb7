// kdpt_flatten.cpp — see kdpt_flatten.h.
#include "kdpt_flatten.h"
#include <string.h>
#include <unordered_map>

namespace kdpt {

uint64_t FlatTree::bytes() const
{
    return nodes.size() * sizeof(HostNode) + (boxLo.size() + boxHi.size()) * sizeof(HostF4) +
           (triA.size() + triB.size() + nrmA.size() + nrmB.size()) * sizeof(HostF4) +
           (triC.size() + nrmC.size()) * sizeof(HostF2);
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

    out.triA.resize(numTris); out.triB.resize(numTris); out.triC.resize(numTris);
    out.nrmA.resize(numTris); out.nrmB.resize(numTris); out.nrmC.resize(numTris);
    std::unordered_map<TriKey, int, TriKeyHash> first;
    first.reserve((size_t)numTris);
    for (int i = 0; i < numTris; i++) {
        const kdpt_tri_bare& t = tris[i];
        TriKey k; memcpy(k.w, &t, sizeof k.w);
        auto it = first.find(k);
        int canon = i;
        if (it == first.end()) first.emplace(k, i); else canon = it->second;
        // e1 = v1 - v0, e2 = v2 - v0: binary32 subtractions, the same the reference performs per test
        float e1x = t.x2 - t.x1, e1y = t.y2 - t.y1, e1z = t.z2 - t.z1;
        float e2x = t.x3 - t.x1, e2y = t.y3 - t.y1, e2z = t.z3 - t.z1;
        out.triA[i] = HostF4{t.x1, t.y1, t.z1, e1x};
        out.triB[i] = HostF4{e1y, e1z, e2x, e2y};
        out.triC[i] = HostF2{e2z, asFloat(canon)};
        out.nrmA[i] = HostF4{t.nx1, t.ny1, t.nz1, t.nx2};
        out.nrmB[i] = HostF4{t.ny2, t.nz2, t.nx3, t.ny3};
        out.nrmC[i] = HostF2{t.nz3, asFloat(t.mtlIdx)};
    }
    return true;
}

} // namespace kdpt
