// kdpt_flatten.h — host-side repack of the reference's NodeBare[] / TriBare[] (scene.cpp:905-968) into the
// device layout described in kdpt_core.cuh. Pure host C++, no CUDA.
#pragma once
#include <stdint.h>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>
#include "../../include/kdpt.h"
#include "kdpt_bulk.h"

namespace kdpt {

struct HostNode { uint32_t x, y, z, w; };
struct HostF4 { float x, y, z, w; };
struct HostF2 { float x, y; };

struct FlatTree {
    std::vector<HostNode> nodes;
    std::vector<HostF4> boxLo, boxHi;
    BulkVector<HostF4> triA, triB, nrmA, nrmB;
    BulkVector<HostF2> triC, nrmC;
    BulkVector<int> triOrig;                    // TriBare index of every (possibly reordered) leaf entry
    std::vector<HostF4> chunkLo, chunkHi;       // padded AABB of every 32-triangle chunk of the chunked leaves
    std::vector<HostF4> superLo, superHi;       // AABB of every group of 32 chunks of the very fat leaves
    int numChunks = 0;
    // Refined tree (exact modes only): the reference tree with every leaf of more than FlattenOpts::refineLeafMax
    // triangles split further by a surface-area-heuristic KD builder (same 16-byte node record; xnodes[i] mirrors
    // nodes[i] for i < numNodes, the new nodes follow). Leaf entries are (unique) TriBare indices; normals are found
    // through entryOf. Empty when refinement is off: the exact modes then walk the reference tree itself.
    std::vector<HostNode> xnodes;
    BulkVector<HostF4> xtriA, xtriB;
    BulkVector<HostF2> xtriC;
    std::vector<HostF4> xchunkLo, xchunkHi, xsuperLo, xsuperHi;
    // the top levels of the refined tree in breadth-first order, for the kernels' shared-memory copy: same record, but a
    // child that is itself in this array is named by (its position | 0x80000000), any other child by its xnodes index
    std::vector<HostNode> xtop;
    // the refined tree with two levels per record (FlattenOpts::wideNodes, layout and reference encoding: kdpt_lean.cuh): record k =
    // xwide[2k], xwide[2k+1]; xwideRoot = reference of the root. Empty when the tree does not fit the encoding.
    std::vector<HostNode> xwide;
    uint32_t xwideRoot = 0;
    BulkVector<int> entryOf;                    // TriBare index -> entry of triA/.../nrmC holding it (inverse of triOrig)
    int xNumTris = 0, xMaxDepth = 0, xMaxLeaf = 0;
    float rootLo[3] = {0, 0, 0}, rootHi[3] = {0, 0, 0};
    int tagCand[4] = {-1, -1, -1, -1};
    int numNodes = 0, numTris = 0, maxDepth = 0, maxLeaf = 0;
    uint64_t bytes() const;
};

constexpr uint64_t kWideMinBytes = 256ull << 20;     // two-level records pay once the tree is beyond the L2 (kdpt_lean.cuh)

struct FlattenOpts {
    int refineLeafMax = 8;      // leaves with more triangles than this are split further for the exact modes; 0 = off
    float traversalCost = 2.0f; // SAH: cost of one node step relative to one triangle test
    int wideNodes = -1;         // also emit the refined tree as two-level records with the small leaves named inline by their parents
                                // (what the exact modes' kernel then walks): 1 yes, 0 no, -1 when the device arrays reach kWideMinBytes
    int topLevels = 0;          // levels of the refined tree copied breadth-first for shared-memory staging (0 = none, <= 11)
    FlattenOpts();              // defaults, overridable in development by KDPT_REFINE_LEAF / KDPT_REFINE_CT
};

// Validates the reference invariants the kernels rely on (index == ID, children in range, both-or-no children,
// leaf ranges inside [0, numTris), depth <= kMaxDepth) and fills `out`. Returns false with `err` set otherwise.
bool flatten(const kdpt_node_bare* nodes, int numNodes, const kdpt_tri_bare* tris, int numTris, FlatTree& out,
             std::string& err, const FlattenOpts& opts = FlattenOpts());

} // namespace kdpt
