// kdpt_build.cu — the reference's KD-tree build on the GPU: KDtree(triangles) -> updateBbox ->
// split(depth) -> flatten (KDnode.cpp:112-249, scene.cpp:866-968), producing the SAME NodeBare[] / TriBare[] bytes as
// the recursive host build (kdpt_scene.cpp:buildTree, itself byte-identical to the reference loader).
//
// Why it can be level-synchronous: KDnode::split is a spatial-median build whose per-triangle decision is a pure
// predicate of (triangle bounds along depth % 3, centre of the node's box):
//     left  <=> (double)tri.min[a] <  (double)centre[a] + 0.0001            KDnode.cpp:190-206
//     right <=> (double)tri.max[a] >= (double)centre[a] - 0.0001
// a node is a leaf when it has <= 2 triangles, when depth > maxdepth, or when one side keeps every triangle; children
// keep their parent's triangle ORDER. So one level of the whole tree is one segmented stable partition of a flat index
// array, and node IDs (DFS pre-order, left subtree first) and leaf ranges follow from the per-node counts.
//
// Per level (all nodes of that depth at once, m = triangles referenced by the level):
//   classifyKernel   8 consecutive entries per thread: gather the (min,max) pair of the level's axis, evaluate the two
//                    predicates in binary64 exactly as the reference does, store 2 flag bits per entry, add the
//                    per-node (left,right) counts (warp-aggregated atomics; node slots are monotone along the array)
//                    and the per-tile totals
//   scanTilesKernel  exclusive scan of the tile totals (one block)
//   [host]           reads the per-node counts (8 B per node, <= 16 384 nodes), applies the reference's leaf rules and
//                    lays out the next level: child segments in node order, left list then right list
//   scatterKernel    re-scans the flags of its tile, ranks every entry inside its node's left / right list and writes
//                    the triangle id into the next level's array (or, for a node that became a leaf, into the leaf pool)
// Afterwards the host numbers the (<= 32 767) nodes depth-first with the reference's box arithmetic, and two kernels emit
// the leaf triangle lists in DFS order and the 76-byte TriBare records. Bytes moved per level: m x (4 id + 4 slot + 8
// bounds + 1 flag) read, m' x 8 written — HBM-bound integer work, ~1.4 GB per level for the 10 M-triangle soup.
//
// The root box reduction keeps the host loop's tie rule (the earliest triangle wins among equal-comparing bounds, which
// matters only for +0 / -0) by reducing (ordered value, index) keys.
#include "kdpt_internal.h"
#include <string.h>
#include <algorithm>

namespace kdpt {

namespace {

constexpr int kItems = 8;                 // consecutive entries per thread
constexpr int kThreads = 256;
constexpr int kTile = kItems * kThreads;  // entries per block

typedef unsigned long long u64;

struct LevelNodeIn {       // per node of the level being split
    float centre;          // centre[depth % 3] of the node's box
    int active;            // 0: already a leaf (<= 2 triangles or depth > maxdepth), its entries are not classified
};

struct LevelNodeOut {      // what the host decided for the node
    unsigned startL, startR;   // global exclusive scan of the left / right flags at the node's first entry
    int dstL, dstR;            // first entry of the children's lists in the next level's array (split nodes)
    int slotL, slotR;          // children's slots in the next level
    int poolBase;              // >= 0: the node is a leaf, its list goes to pool[poolBase ...]
    int segStart;              // the node's first entry in this level's array
};

__device__ __forceinline__ unsigned orderedBits(float f)
{
    if (f == 0.0f) f = 0.0f;                               // +0 and -0 compare equal in the host loop
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// KDN::Triangle::computeBounds (KDnode.h:216-225) per triangle + the keys of the root box reduction (KDnode.cpp:112-137)
__global__ void boundsKernel(const float* __restrict__ tri18, int n, float2* __restrict__ bx, float2* __restrict__ by,
                             float2* __restrict__ bz, u64* __restrict__ keys /* [0..2] min, [3..5] max */)
{
    __shared__ u64 sMin[3], sMax[3];
    if (threadIdx.x < 3) { sMin[threadIdx.x] = ~0ull; sMax[threadIdx.x] = 0ull; }
    __syncthreads();
    u64 kmin[3] = {~0ull, ~0ull, ~0ull}, kmax[3] = {0ull, 0ull, 0ull};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* t = tri18 + 18 * (size_t)i;
        float2 b[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float c1 = t[a], c2 = t[3 + a], c3 = t[6 + a];
            b[a].x = c1 < c2 ? (c1 < c3 ? c1 : c3) : (c2 < c3 ? c2 : c3);
            b[a].y = c1 > c2 ? (c1 > c3 ? c1 : c3) : (c2 > c3 ? c2 : c3);
            u64 lo = ((u64)orderedBits(b[a].x) << 32) | (unsigned)i;
            u64 hi = ((u64)orderedBits(b[a].y) << 32) | (0xffffffffu - (unsigned)i);
            kmin[a] = lo < kmin[a] ? lo : kmin[a];
            kmax[a] = hi > kmax[a] ? hi : kmax[a];
        }
        bx[i] = b[0]; by[i] = b[1]; bz[i] = b[2];
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) {
            u64 x = __shfl_xor_sync(0xffffffffu, kmin[a], o), y = __shfl_xor_sync(0xffffffffu, kmax[a], o);
            kmin[a] = x < kmin[a] ? x : kmin[a];
            kmax[a] = y > kmax[a] ? y : kmax[a];
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&sMin[a], kmin[a]); atomicMax(&sMax[a], kmax[a]); }
    }
    __syncthreads();
    if (threadIdx.x < 3) { atomicMin(&keys[threadIdx.x], sMin[threadIdx.x]); atomicMax(&keys[3 + threadIdx.x], sMax[threadIdx.x]); }
}

__global__ void iotaKernel(int* __restrict__ cur, int* __restrict__ slot, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { cur[i] = i; slot[i] = 0; }
}

// flags: bit 0 = goes left, bit 1 = goes right. counts packed as (right << 32) | left.
__global__ void __launch_bounds__(kThreads)
classifyKernel(const int* __restrict__ cur, const int* __restrict__ slot, const float2* __restrict__ bnd,
               const LevelNodeIn* __restrict__ nodesIn, int m, unsigned char* __restrict__ flags,
               u64* __restrict__ nodeCount, u64* __restrict__ tileSum)
{
    const int base = blockIdx.x * kTile + threadIdx.x * kItems;
    u64 total = 0, run = 0;
    int runSlot = -1, firstSlot = -1;
    bool uniform = true;
    u64 packedFlags = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const int i = base + k;
        if (i < m) {
            const int s = slot[i];
            const LevelNodeIn nd = nodesIn[s];
            u64 v = 0;
            if (nd.active) {
                const float2 b = __ldg(&bnd[cur[i]]);
                const double c = (double)nd.centre;
                const bool fl = (double)b.x < c + 0.0001;
                const bool fr = (double)b.y >= c - 0.0001;
                v = (u64)(fl ? 1u : 0u) | ((u64)(fr ? 1u : 0u) << 32);
                packedFlags |= (u64)((fl ? 1u : 0u) | (fr ? 2u : 0u)) << (8 * k);
            }
            if (s != runSlot) {
                if (runSlot >= 0) { if (run) atomicAdd(&nodeCount[runSlot], run); uniform = false; }
                else firstSlot = s;
                runSlot = s; run = 0;
            }
            run += v; total += v;
        }
    }
    if (base < m) *reinterpret_cast<u64*>(flags + base) = packedFlags;       // flags[] is padded to a whole tile
    // per-node counts: one atomic per warp when the whole warp sits inside one node (the common case near the root)
    const int lane0Slot = __shfl_sync(0xffffffffu, firstSlot, 0);
    const bool warpUniform = __all_sync(0xffffffffu, uniform && firstSlot == lane0Slot && firstSlot >= 0);
    if (warpUniform) {
        u64 w = run;
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&nodeCount[lane0Slot], w);
    } else if (runSlot >= 0 && run) {
        atomicAdd(&nodeCount[runSlot], run);
    }
    // tile total
    __shared__ u64 warpTot[kThreads / 32];
    u64 t = total;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) warpTot[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 s = 0;
        for (int w = 0; w < kThreads / 32; w++) s += warpTot[w];
        tileSum[blockIdx.x] = s;
    }
}

// exclusive scan of the tile totals, in place (one block of 1024 threads; <= a few thousand tiles)
__global__ void __launch_bounds__(1024) scanTilesKernel(u64* __restrict__ tileSum, int numTiles)
{
    __shared__ u64 warpTot[32];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < numTiles; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const u64 v = i < numTiles ? tileSum[i] : 0;
        u64 inc = v;
        for (int o = 1; o < 32; o <<= 1) { u64 x = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += x; }
        if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            u64 w = warpTot[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) { u64 x = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += x; }
            warpTot[threadIdx.x] = wi - w;                      // exclusive warp offsets
        }
        __syncthreads();
        const u64 c = carry;
        const u64 excl = c + warpTot[threadIdx.x >> 5] + inc - v;
        if (i < numTiles) tileSum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
scatterKernel(const int* __restrict__ cur, const int* __restrict__ slot, const unsigned char* __restrict__ flags,
              const u64* __restrict__ tileBase, const LevelNodeOut* __restrict__ nodesOut, int m,
              int* __restrict__ next, int* __restrict__ nextSlot, int* __restrict__ pool)
{
    const int base = blockIdx.x * kTile + threadIdx.x * kItems;
    const u64 packedFlags = base < m ? *reinterpret_cast<const u64*>(flags + base) : 0ull;
    u64 mine = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const unsigned f = (unsigned)(packedFlags >> (8 * k)) & 3u;
        mine += (u64)(f & 1u) | ((u64)(f >> 1) << 32);
    }
    // block-wide exclusive scan of `mine`
    __shared__ u64 warpTot[kThreads / 32];
    u64 inc = mine;
    for (int o = 1; o < 32; o <<= 1) { u64 x = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += x; }
    if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = inc;
    __syncthreads();
    u64 before = tileBase[blockIdx.x];
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += warpTot[w];
    u64 g = before + inc - mine;                                 // global exclusive (left | right << 32) at `base`
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        const int i = base + k;
        if (i < m) {
            const unsigned f = (unsigned)(packedFlags >> (8 * k)) & 3u;
            const int s = slot[i];
            const LevelNodeOut nd = nodesOut[s];
            const int t = cur[i];
            if (nd.poolBase >= 0) {
                pool[nd.poolBase + (i - nd.segStart)] = t;
            } else {
                if (f & 1u) { const int p = nd.dstL + (int)((unsigned)g - nd.startL); next[p] = t; nextSlot[p] = nd.slotL; }
                if (f & 2u) { const int p = nd.dstR + (int)((unsigned)(g >> 32) - nd.startR); next[p] = t; nextSlot[p] = nd.slotR; }
            }
            g += (u64)(f & 1u) | ((u64)(f >> 1) << 32);
        }
    }
}

struct LeafCopy { int poolOff, dstOff, num; };

// leaf lists from the pool (level order) into TriBare order (DFS order of the leaves): one block per leaf, strided
__global__ void leafGatherKernel(const int* __restrict__ pool, const LeafCopy* __restrict__ leaves, int numLeaves,
                                 int* __restrict__ leafTri)
{
    for (int l = blockIdx.x; l < numLeaves; l += gridDim.x) {
        const LeafCopy lc = leaves[l];
        for (int k = threadIdx.x; k < lc.num; k += blockDim.x) leafTri[lc.dstOff + k] = pool[lc.poolOff + k];
    }
}

// scene.cpp:905-968: one 76-byte component-major TriBare per leaf entry; one thread per 4-byte word (coalesced stores)
__global__ void triBareKernel(const float* __restrict__ tri18, const int* __restrict__ mtl, const int* __restrict__ leafTri,
                              size_t numWords, unsigned* __restrict__ out)
{
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < numWords; w += (size_t)gridDim.x * blockDim.x) {
        const size_t e = w / 19;
        const int k = (int)(w - e * 19);
        const int t = leafTri[e];
        unsigned v;
        if (k == 18) v = mtl ? (unsigned)mtl[t] : 0u;
        else {
            const int grp = k / 9, r = k - grp * 9, comp = r / 3, vert = r - comp * 3;      // x1 x2 x3 y1 y2 y3 z1 z2 z3
            v = __float_as_uint(tri18[18 * (size_t)t + grp * 9 + vert * 3 + comp]);
        }
        out[w] = v;
    }
}

struct HostBuildNode {
    float mins[3], maxs[3], centre[3];
    int axis = 0; float splitPos = 0.f;
    int parent = -1, left = -1, right = -1, depth = 0;
    int num = 0;
    int poolOff = -1;        // leaf: its list in the pool
    int id = -1;             // DFS pre-order
};

struct Scope {               // frees every device buffer of the build on every exit path
    std::vector<DeviceBuffer*> bufs;
    ~Scope() { for (auto* b : bufs) b->release(); }
};

} // namespace

int buildTreeDevice(const float* tri18, const int* mtl, int n, int maxdepth, int device,
                    std::vector<kdpt_node_bare>& outNodes, BulkVector<kdpt_tri_bare>& outTris, BulkVector<int>& outSource,
                    float* deviceMs)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); setError("kdpt_scene_from_triangles_gpu: no CUDA device (there is no CPU fallback behind this entry point)"); return KDPT_ERR_CUDA; }
    if (device < 0 || device >= count) { setError("kdpt_scene_from_triangles_gpu: bad device ordinal"); return KDPT_ERR_ARG; }
    KDPT_CUDA(cudaSetDevice(device));
    cudaStream_t st;
    KDPT_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{st};
    cudaEvent_t e0, e1;
    KDPT_CUDA(cudaEventCreate(&e0)); KDPT_CUDA(cudaEventCreate(&e1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } eg{e0, e1};

    std::vector<HostBuildNode> nodes(1);
    HostBuildNode& root = nodes[0];
    memset(root.mins, 0, sizeof root.mins); memset(root.maxs, 0, sizeof root.maxs); memset(root.centre, 0, sizeof root.centre);
    root.num = n;

    DeviceBuffer dTri, dMtl, dBx, dBy, dBz, dKeys, dCur, dSlot, dNext, dNextSlot, dFlags, dTileSum, dNodeCount, dNodesIn,
        dNodesOut, dPool, dLeafTri, dLeaves, dOut;
    Scope scope;
    scope.bufs = {&dTri, &dMtl, &dBx, &dBy, &dBz, &dKeys, &dCur, &dSlot, &dNext, &dNextSlot, &dFlags, &dTileSum, &dNodeCount,
                  &dNodesIn, &dNodesOut, &dPool, &dLeafTri, &dLeaves, &dOut};
    int rc;
    std::vector<int> poolLevelBase;          // pool offset of every leaf is global; the pool grows level by level
    size_t poolSize = 0;
    const float pad = 0.001f;

    if (n > 0) {
        if ((rc = dTri.alloc(18 * sizeof(float) * (size_t)n))) return rc;
        KDPT_CUDA(cudaMemcpyAsync(dTri.p, tri18, 18 * sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
        if (mtl) {
            if ((rc = dMtl.alloc(sizeof(int) * (size_t)n))) return rc;
            KDPT_CUDA(cudaMemcpyAsync(dMtl.p, mtl, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
        }
        KDPT_CUDA(cudaEventRecord(e0, st));
        if ((rc = dBx.alloc(sizeof(float2) * (size_t)n)) || (rc = dBy.alloc(sizeof(float2) * (size_t)n)) ||
            (rc = dBz.alloc(sizeof(float2) * (size_t)n)) || (rc = dKeys.alloc(6 * sizeof(u64)))) return rc;
        u64 keysInit[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull}, keys[6];
        KDPT_CUDA(cudaMemcpyAsync(dKeys.p, keysInit, sizeof keysInit, cudaMemcpyHostToDevice, st));
        const int grid = std::min((n + 255) / 256, 148 * 8);
        boundsKernel<<<grid, 256, 0, st>>>(dTri.as<float>(), n, dBx.as<float2>(), dBy.as<float2>(), dBz.as<float2>(), dKeys.as<u64>());
        KDPT_CUDA(cudaGetLastError());
        KDPT_CUDA(cudaMemcpyAsync(keys, dKeys.p, sizeof keys, cudaMemcpyDeviceToHost, st));
        KDPT_CUDA(cudaStreamSynchronize(st));
        // the winning triangles' bounds, recomputed on the host from the caller's array (same expressions as the kernel)
        for (int a = 0; a < 3; a++) {
            const unsigned iMin = (unsigned)(keys[a] & 0xffffffffu), iMax = 0xffffffffu - (unsigned)(keys[3 + a] & 0xffffffffu);
            const float* t = tri18 + 18 * (size_t)iMin;
            float c1 = t[a], c2 = t[3 + a], c3 = t[6 + a];
            root.mins[a] = c1 < c2 ? (c1 < c3 ? c1 : c3) : (c2 < c3 ? c2 : c3);
            t = tri18 + 18 * (size_t)iMax;
            c1 = t[a]; c2 = t[3 + a]; c3 = t[6 + a];
            root.maxs[a] = c1 > c2 ? (c1 > c3 ? c1 : c3) : (c2 > c3 ? c2 : c3);
            root.centre[a] = (float)((double)(root.mins[a] + root.maxs[a]) / 2.0);     // computed BEFORE the pad (KDnode.cpp:112-149)
        }
    }
    for (int a = 0; a < 3; a++) { root.mins[a] -= pad; root.maxs[a] += pad; }

    // ---- level loop ---------------------------------------------------------------------------------------------
    std::vector<int> level = {0};            // node indices (into `nodes`) of the current level, in array order
    size_t m = (size_t)n;
    if (n > 0) {
        if ((rc = dCur.alloc(sizeof(int) * m)) || (rc = dSlot.alloc(sizeof(int) * m))) return rc;
        iotaKernel<<<std::min((n + 255) / 256, 148 * 8), 256, 0, st>>>(dCur.as<int>(), dSlot.as<int>(), n);
        KDPT_CUDA(cudaGetLastError());
    }
    std::vector<LevelNodeIn> in;
    std::vector<LevelNodeOut> out;
    std::vector<u64> counts;
    for (int depth = 0; m > 0 && !level.empty(); depth++) {
        const int axis = depth % 3;
        const int numLevelNodes = (int)level.size();
        const int numTiles = (int)((m + kTile - 1) / kTile);
        in.resize(numLevelNodes);
        bool anyActive = false;
        for (int s = 0; s < numLevelNodes; s++) {
            const HostBuildNode& nd = nodes[level[s]];
            in[s].centre = nd.centre[axis];
            in[s].active = (nd.num > 2 && depth <= maxdepth) ? 1 : 0;       // KDnode.cpp:151-160
            anyActive |= in[s].active != 0;
        }
        if ((rc = dNodesIn.ensure(sizeof(LevelNodeIn) * numLevelNodes)) || (rc = dNodesOut.ensure(sizeof(LevelNodeOut) * numLevelNodes)) ||
            (rc = dNodeCount.ensure(sizeof(u64) * numLevelNodes)) || (rc = dTileSum.ensure(sizeof(u64) * numTiles)) ||
            (rc = dFlags.ensure((size_t)numTiles * kTile))) return rc;
        KDPT_CUDA(cudaMemcpyAsync(dNodesIn.p, in.data(), sizeof(LevelNodeIn) * numLevelNodes, cudaMemcpyHostToDevice, st));
        KDPT_CUDA(cudaMemsetAsync(dNodeCount.p, 0, sizeof(u64) * numLevelNodes, st));
        const float2* bnd = axis == 0 ? dBx.as<float2>() : axis == 1 ? dBy.as<float2>() : dBz.as<float2>();
        classifyKernel<<<numTiles, kThreads, 0, st>>>(dCur.as<int>(), dSlot.as<int>(), bnd, dNodesIn.as<LevelNodeIn>(), (int)m,
                                                      dFlags.as<unsigned char>(), dNodeCount.as<u64>(), dTileSum.as<u64>());
        KDPT_CUDA(cudaGetLastError());
        scanTilesKernel<<<1, 1024, 0, st>>>(dTileSum.as<u64>(), numTiles);
        KDPT_CUDA(cudaGetLastError());
        counts.resize(numLevelNodes);
        KDPT_CUDA(cudaMemcpyAsync(counts.data(), dNodeCount.p, sizeof(u64) * numLevelNodes, cudaMemcpyDeviceToHost, st));
        KDPT_CUDA(cudaStreamSynchronize(st));
        (void)anyActive;

        // the reference's decisions (KDnode.cpp:208-249) and the layout of the next level
        out.resize(numLevelNodes);
        std::vector<int> nextLevel;
        size_t nextM = 0, seg = 0;
        unsigned accL = 0, accR = 0;
        for (int s = 0; s < numLevelNodes; s++) {
            const int ni = level[s];
            const int num = nodes[ni].num;
            const unsigned nl = (unsigned)(counts[s] & 0xffffffffu), nr = (unsigned)(counts[s] >> 32);
            LevelNodeOut& o = out[s];
            o.startL = accL; o.startR = accR; o.segStart = (int)seg;
            o.dstL = o.dstR = o.slotL = o.slotR = -1; o.poolBase = -1;
            const bool split = in[s].active && !((int)nl == num || (int)nr == num);
            if (!split) {
                if (poolSize + (size_t)num > 0x7fffffffull) { setError("kdpt_scene_from_triangles_gpu: more than 2^31 leaf entries"); return KDPT_ERR_ARG; }
                o.poolBase = (int)poolSize; nodes[ni].poolOff = (int)poolSize; poolSize += (size_t)num;
            } else {
                if (nextM + nl + nr > 0x7fffffffull) { setError("kdpt_scene_from_triangles_gpu: more than 2^31 entries in one level"); return KDPT_ERR_ARG; }
                const float cl = nodes[ni].centre[axis];
                for (int side = 0; side < 2; side++) {
                    HostBuildNode ch;
                    const HostBuildNode& p = nodes[ni];
                    memcpy(ch.mins, p.mins, sizeof ch.mins); memcpy(ch.maxs, p.maxs, sizeof ch.maxs);
                    ch.axis = (axis + 1) % 3; ch.parent = ni; ch.depth = depth + 1;
                    if (side == 0) { ch.maxs[axis] = cl; ch.splitPos = p.maxs[axis]; ch.num = (int)nl; }
                    else { ch.mins[axis] = cl; ch.splitPos = p.mins[axis]; ch.num = (int)nr; }
                    for (int a = 0; a < 3; a++) ch.centre[a] = (float)((double)(ch.mins[a] + ch.maxs[a]) / 2.0);
                    const int ci = (int)nodes.size();
                    nodes.push_back(ch);
                    if (side == 0) { nodes[ni].left = ci; o.dstL = (int)nextM; o.slotL = (int)nextLevel.size(); nextM += nl; }
                    else { nodes[ni].right = ci; o.dstR = (int)nextM; o.slotR = (int)nextLevel.size(); nextM += nr; }
                    nextLevel.push_back(ci);
                }
            }
            accL += nl; accR += nr; seg += (size_t)num;
        }
        if (seg != m) { setError("kdpt_scene_from_triangles_gpu: internal error (level size mismatch)"); return KDPT_ERR_STATE; }
        KDPT_CUDA(cudaMemcpyAsync(dNodesOut.p, out.data(), sizeof(LevelNodeOut) * numLevelNodes, cudaMemcpyHostToDevice, st));
        if (poolSize > 0 && dPool.bytes < poolSize * sizeof(int)) {
            // grow the pool, keeping what earlier levels wrote
            DeviceBuffer bigger;
            if ((rc = bigger.alloc(std::max(poolSize * sizeof(int) * 2, (size_t)n * sizeof(int) * 2)))) return rc;
            if (dPool.p) KDPT_CUDA(cudaMemcpyAsync(bigger.p, dPool.p, dPool.bytes, cudaMemcpyDeviceToDevice, st));
            KDPT_CUDA(cudaStreamSynchronize(st));
            dPool.release();
            dPool = bigger;
        }
        if (nextM > 0 && ((rc = dNext.alloc(sizeof(int) * nextM)) || (rc = dNextSlot.alloc(sizeof(int) * nextM)))) return rc;
        scatterKernel<<<numTiles, kThreads, 0, st>>>(dCur.as<int>(), dSlot.as<int>(), dFlags.as<unsigned char>(), dTileSum.as<u64>(),
                                                     dNodesOut.as<LevelNodeOut>(), (int)m, dNext.as<int>(), dNextSlot.as<int>(),
                                                     dPool.as<int>());
        KDPT_CUDA(cudaGetLastError());
        KDPT_CUDA(cudaStreamSynchronize(st));
        std::swap(dCur, dNext); std::swap(dSlot, dNextSlot);
        dNext.release(); dNextSlot.release();
        level.swap(nextLevel);
        m = nextM;
    }

    // ---- DFS numbering (the recursion order of KDnode::split: node, left subtree, right subtree) -----------------
    const int numNodes = (int)nodes.size();
    outNodes.assign(numNodes, kdpt_node_bare{});
    std::vector<LeafCopy> leaves;
    {
        std::vector<int> stack = {0};
        int nextId = 0; size_t leafOff = 0;
        while (!stack.empty()) {
            const int ni = stack.back(); stack.pop_back();
            HostBuildNode& nd = nodes[ni];
            nd.id = nextId++;
            if (nd.left >= 0) { stack.push_back(nd.right); stack.push_back(nd.left); }
            else if (nd.num > 0) { leaves.push_back({nd.poolOff, (int)leafOff, nd.num}); leafOff += (size_t)nd.num; }
        }
        if (leafOff != poolSize) { setError("kdpt_scene_from_triangles_gpu: internal error (leaf size mismatch)"); return KDPT_ERR_STATE; }
        size_t li = 0;
        // second walk in ID order to fill the records (leaves were pushed in ID order)
        std::vector<int> byId(numNodes);
        for (int i = 0; i < numNodes; i++) byId[nodes[i].id] = i;
        for (int id = 0; id < numNodes; id++) {
            const HostBuildNode& nd = nodes[byId[id]];
            kdpt_node_bare& o = outNodes[id];
            memset(&o, 0, sizeof o);
            o.axis = nd.axis; o.splitPos = nd.splitPos;
            memcpy(o.mins, nd.mins, sizeof o.mins); memcpy(o.maxs, nd.maxs, sizeof o.maxs);
            o.ID = id; o.parentID = nd.parent >= 0 ? nodes[nd.parent].id : -1;
            o.leftID = nd.left >= 0 ? nodes[nd.left].id : -1; o.rightID = nd.right >= 0 ? nodes[nd.right].id : -1;
            o.triIdStart = o.triIdSize = -1;
            if (nd.left < 0 && nd.num > 0) { o.triIdStart = leaves[li].dstOff; o.triIdSize = nd.num; li++; }
        }
    }

    // ---- leaf lists in DFS order + TriBare records -----------------------------------------------------------------
    if (poolSize > 0) {
        if ((rc = dLeafTri.alloc(sizeof(int) * poolSize)) || (rc = dLeaves.alloc(sizeof(LeafCopy) * leaves.size())) ||
            (rc = dOut.alloc(sizeof(kdpt_tri_bare) * poolSize))) return rc;
        KDPT_CUDA(cudaMemcpyAsync(dLeaves.p, leaves.data(), sizeof(LeafCopy) * leaves.size(), cudaMemcpyHostToDevice, st));
        leafGatherKernel<<<std::min((int)leaves.size(), 148 * 16), 256, 0, st>>>(dPool.as<int>(), dLeaves.as<LeafCopy>(), (int)leaves.size(),
                                                                               dLeafTri.as<int>());
        KDPT_CUDA(cudaGetLastError());
        const size_t words = poolSize * 19;
        triBareKernel<<<(int)std::min<size_t>((words + 255) / 256, 148 * 32), 256, 0, st>>>(dTri.as<float>(), mtl ? dMtl.as<int>() : nullptr,
                                                                                          dLeafTri.as<int>(), words, dOut.as<unsigned>());
        KDPT_CUDA(cudaGetLastError());
        KDPT_CUDA(cudaEventRecord(e1, st));
        outSource.resize(poolSize);          // the host arrays are allocated (and page-faulted) while the kernels run
        outTris.resize(poolSize);
        KDPT_CUDA(cudaMemcpyAsync(outSource.data(), dLeafTri.p, sizeof(int) * poolSize, cudaMemcpyDeviceToHost, st));
        KDPT_CUDA(cudaMemcpyAsync(outTris.data(), dOut.p, sizeof(kdpt_tri_bare) * poolSize, cudaMemcpyDeviceToHost, st));
        KDPT_CUDA(cudaStreamSynchronize(st));
        if (deviceMs) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *deviceMs = ms; }
    } else {
        outSource.clear(); outTris.clear();
        if (deviceMs) *deviceMs = 0.f;
    }
    return KDPT_OK;
}

} // namespace kdpt
