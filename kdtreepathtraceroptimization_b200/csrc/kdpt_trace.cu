// kdpt_trace.cu — traversal + intersection kernels (sm_100a), the kdpt_trace / kdpt_trace_device entry points.
//
// Replaces the per-ray body of pathTraceOneBounceKDbare (src/pathtrace.cu:1651-1681): traverseKDbare (881-1020) and
// traverseKDbareShortHybrid (1023-1235). Compiled with -fmad=false: see kdpt_core.cuh.
#include "kdpt_internal.h"
#include "kdpt_lean.cuh"
#include <stdlib.h>

namespace kdpt {

constexpr int kTraceThreads = 128;
#ifndef KDPT_COOP_MIN
#define KDPT_COOP_MIN 8
#endif
constexpr int kCoopMin = KDPT_COOP_MIN;     // leaves with at least this many triangles may be tested by the whole warp
#ifndef KDPT_BLOCK_RANGE
#define KDPT_BLOCK_RANGE 16
#endif
constexpr int kBlockRange = KDPT_BLOCK_RANGE;   // packets per block-level ticket of a binned batch (see traceKernel)
constexpr int kBinMinRays = 1 << 16;     // batches at least this big are traced in leaf order (see binRaysKernel)
constexpr uint64_t kBinMinTreeBytes = 100ull << 20;

// Persistent threads: a fixed grid (a multiple of the SM count) whose warps pull 32-ray packets from a global
// counter until the batch is drained, so a warp stuck in a fat leaf does not hold back the rest of its block's
// share of the batch.
#ifndef KDPT_TRACE_MIN_BLOCKS
#define KDPT_TRACE_MIN_BLOCKS 7    // <= 72 registers: 28 warps per SM (measured best of 4..8 on dragon_5 and the 1 M / 10 M soups)
#endif
template <int MODE>
__global__ void __launch_bounds__(kTraceThreads, KDPT_TRACE_MIN_BLOCKS)
traceKernel(Tree T, const kdpt_ray* __restrict__ rays, const float* __restrict__ tMinIn, int n,
            kdpt_hit* __restrict__ hits, unsigned* __restrict__ counter, const unsigned* __restrict__ order, int blockTickets)
{
    const unsigned lane = threadIdx.x & 31u;
    __shared__ unsigned sBase, sNext;
    unsigned rangeBase = 0, rangeLeft = 0;      // binned batch: the block's current range of packets
    for (;;) {
        unsigned base = 0;
        if (blockTickets) {
            // binned batch: the block takes a range of kBlockRange neighbouring packets (the same leaf or the next one), so
            // that the leaf's triangles are fetched into this SM's L1 once, and its warps deal the packets of the range
            // among themselves as they become free. The block only meets at a barrier when a range is used up (a warp
            // stuck in a long leaf holds the others back for at most the tail of one range; with one packet per warp
            // and barrier, ncu showed 3.5 warps per issue slot waiting at the barrier on the 10 M soup).
            unsigned k = kBlockRange;
            if (rangeLeft) {
                if (lane == 0) k = atomicAdd(&sNext, 1u);
                k = __shfl_sync(0xffffffffu, k, 0);
            }
            if (k >= (unsigned)kBlockRange || rangeBase + k * 32u >= (unsigned)n) {
                __syncthreads();            // every warp is done with the old range (nobody still reads sBase / sNext)
                if (threadIdx.x == 0) { sBase = atomicAdd(counter, (unsigned)kBlockRange * 32u); sNext = 0; }
                __syncthreads();
                rangeBase = sBase;
                if (rangeBase >= (unsigned)n) break;
                rangeLeft = 1;
                continue;
            }
            base = rangeBase + k * 32u;
        } else {
            if (lane == 0) base = atomicAdd(counter, 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= (unsigned)n) break;
        }
        unsigned i = base + lane;
        const bool active = i < (unsigned)n;
        if (active && order) i = __ldg(order + i);      // binned batch: packet = 32 rays that start in the same leaf
        V3 o = mk(0, 0, 0), d = mk(0, 0, 0); float tin = FLT_MAX;
        if (active) {
            const float* r = reinterpret_cast<const float*>(rays + i);
            o = mk(__ldg(r + 0), __ldg(r + 1), __ldg(r + 2));
            d = mk(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
            tin = tMinIn ? __ldg(tMinIn + i) : FLT_MAX;
        }
        Hit H;
        if (MODE == KDPT_FULLSTACK) traceExactWarp<kMaxDepth, kCoopMin>(T, active, o, d, tin, 0.00001f, H);
        else if (MODE == KDPT_SHORTSTACK) traceExactWarp<4, kCoopMin>(T, active, o, d, tin, 0.0001f, H);
        else if (MODE == KDPT_REF_T1) traceRefWarp<false>(T, active, o, d, tin, H);
        else if (MODE == KDPT_REF_T2) traceRefWarp<true>(T, active, o, d, tin, H);
        else if (MODE == KDPT_BRUTE) traceBruteWarp<false>(T, active, o, d, tin, 0.0001f, H);
        else traceBruteWarp<true>(T, active, o, d, tin, 0.0001f, H);
        if (active) {
            float4* out = reinterpret_cast<float4*>(hits + i);
            out[0] = make_float4(H.t, H.z, H.u, H.v);
            out[1] = make_float4(__int_as_float(H.tri), H.nx, H.ny, H.nz);
        }
    }
}

// ---- the refined tree: one ray per lane, lanes refilled as their rays finish (kdpt_lean.cuh) ----------------------------
#ifndef KDPT_LEAN_MIN_BLOCKS
#define KDPT_LEAN_MIN_BLOCKS 8
#endif
template <bool RESTART, bool WIDE>
__global__ void __launch_bounds__(kLeanThreads, KDPT_LEAN_MIN_BLOCKS)
traceLeanKernel(Tree T, const kdpt_ray* __restrict__ rays, const float* __restrict__ tMinIn, int n, kdpt_hit* __restrict__ hits,
                unsigned* __restrict__ counter, const unsigned* __restrict__ order, float eps)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    __shared__ uint2 sStackAll[(RESTART ? kLeanRing : kLeanSmemStack) * kLeanThreads];
    __shared__ uint4 sTop[kTopMaxNodes];
    leanStageTop(T, sTop);
    uint2* sStack = sStackAll + threadIdx.x;
    uint2 lStack[RESTART ? 1 : (WIDE ? kWideStack : kMaxDepth) - kLeanSmemStack];
    LeanState S; Hit H;
    bool alive = false, more = true;
    unsigned ray = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !alive);
        if (more && __popc(idle) >= KDPT_LEAN_REFILL_MIN) {
            // idle lanes take the next rays of the batch (binned batches: neighbours in leaf order)
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)__popc(idle) >= (unsigned)n) more = false;
            if (!alive) {
                unsigned i = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (i < (unsigned)n) {
                    if (order) i = __ldg(order + i);
                    const float* r = reinterpret_cast<const float*>(rays + i);
                    const V3 o = mk(__ldg(r + 0), __ldg(r + 1), __ldg(r + 2)), d = mk(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
                    const float tin = tMinIn ? __ldg(tMinIn + i) : FLT_MAX;
                    ray = i;
                    alive = leanBegin<WIDE>(T, o, d, tin, S, H);
                    if (!alive) {       // never enters the tree
                        float4* out = reinterpret_cast<float4*>(hits + i);
                        out[0] = make_float4(H.t, H.z, H.u, H.v);
                        out[1] = make_float4(__int_as_float(H.tri), H.nx, H.ny, H.nz);
                    }
                }
            }
        }
        if (!__any_sync(FULL, alive)) { if (!more) break; continue; }
        const bool was = alive;
        alive = leanRound<RESTART, WIDE>(T, alive, S, H, eps, sStack, lStack, sTop);
        if (was && !alive) {
            float4* out = reinterpret_cast<float4*>(hits + ray);
            out[0] = make_float4(H.t, H.z, H.u, H.v);
            out[1] = make_float4(__int_as_float(H.tri), H.nx, H.ny, H.nz);
        }
    }
}

// ---- ray binning ------------------------------------------------------------------------------------------------
// The rays of a batch arrive in the caller's order. When the tree is bigger than the caches, 32 unrelated rays per
// packet mean that every packet drags its own leaves through HBM (measured on the 10 M-triangle soup: 58 KB of DRAM
// traffic per ray, 64 % of the HBM peak, before binning). The batch is therefore traced in the order of a counting sort
// by (leaf the ray starts in, direction octant): the packets of a leaf run back to back on neighbouring warps and find
// its triangles in L1/L2, and the lanes of a packet walk the same nodes. Node IDs are DFS pre-order, so consecutive
// bins are neighbouring cells. Results are written back by original index: the caller sees nothing of this.
__global__ void __launch_bounds__(256)
binRaysKernel(Tree T, const kdpt_ray* __restrict__ rays, int n, unsigned* __restrict__ keys, unsigned* __restrict__ hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = reinterpret_cast<const float*>(rays + i);
    const V3 o = mk(__ldg(r + 0), __ldg(r + 1), __ldg(r + 2)), d = mk(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5));
    const unsigned key = rayBinKey(T, o, d);
    keys[i] = key;
    atomicAdd(hist + key, 1u);
}

// exclusive scan of the bin counts, one block
__global__ void __launch_bounds__(1024)
binScanKernel(unsigned* __restrict__ hist, int bins)
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

// order[] = ray indices bin by bin (the order inside a bin is whatever the atomics give; it does not reach the results)
__global__ void __launch_bounds__(256)
binScatterKernel(const unsigned* __restrict__ keys, int n, unsigned* __restrict__ offsets, unsigned* __restrict__ order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[atomicAdd(offsets + keys[i], 1u)] = (unsigned)i;
}


int launchTrace(kdpt_ctx* ctx, const kdpt_ray* dRays, const float* dTmin, int n, kdpt_hit* dHits, int mode,
                cudaStream_t stream)
{
    if (n <= 0) return KDPT_OK;
    unsigned* counter = ctx->counter.as<unsigned>();
    KDPT_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
    // bin the batch by starting leaf when it is big enough to fill the machine several times over (below that the
    // three extra launches cost more than they save); KDPT_BIN_RAYS=0/1 forces it off/on in development
    static const int envBin = getenv("KDPT_BIN_RAYS") ? atoi(getenv("KDPT_BIN_RAYS")) : -1;
    const bool exact = mode == KDPT_FULLSTACK || mode == KDPT_SHORTSTACK;
    const Tree& T = ctx->tree;      // bin keys are reference leaves: the same in both views
    // ... and when the tree is beyond the L2: on cache-resident trees the lean kernel does not need the coherence and the sort
    // costs a quarter of the batch (dragon_5: 2043 Mrays/s unsorted, 1519 sorted; 10 M soup: 116 unsorted, 198 sorted)
    const bool bigTree = ctx->deviceBytes >= kBinMinTreeBytes || !(exact && ctx->xNumNodes > 0);
    const bool bin = T.numRefNodes > 1 && (envBin >= 0 ? envBin != 0 : (n >= kBinMinRays && bigTree));
    const unsigned* order = nullptr;
    if (bin) {
        const int bins = T.numRefNodes * 8;
        int rc;
        if ((rc = ctx->binTmp.ensure(sizeof(unsigned) * (2 * (size_t)n + (size_t)bins)))) return rc;
        unsigned* keys = ctx->binTmp.as<unsigned>();
        unsigned* ord = keys + n;
        unsigned* hist = ord + n;
        KDPT_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * (size_t)bins, stream));
        binRaysKernel<<<(n + 255) / 256, 256, 0, stream>>>(T, dRays, n, keys, hist);
        binScanKernel<<<1, 1024, 0, stream>>>(hist, bins);
        binScatterKernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, hist, ord);
        ctx->launches += 3;
        order = ord;
    }
    static const int envBlock = getenv("KDPT_BLOCK_TICKETS") ? atoi(getenv("KDPT_BLOCK_TICKETS")) : 1;
    const int blockTickets = order && envBlock;
    int blocksPerSM = 8;
    int grid = ctx->numSMs * blocksPerSM;
    int need = (n + kTraceThreads - 1) / kTraceThreads;
    if (grid > need) grid = need;
    // exact modes on a refined tree: the lean kernel (KDPT_LEAN=0 keeps the packet kernel, development switch)
    static const int envLean = getenv("KDPT_LEAN") ? atoi(getenv("KDPT_LEAN")) : 1;
    if (exact && envLean && ctx->xNumNodes > 0) {
        static const int envBlocks = getenv("KDPT_LEAN_BLOCKS") ? atoi(getenv("KDPT_LEAN_BLOCKS")) : KDPT_LEAN_MIN_BLOCKS;
        int lgrid = ctx->numSMs * envBlocks;
        const int lneed = (n + kLeanThreads - 1) / kLeanThreads;
        if (lgrid > lneed) lgrid = lneed;
        const bool wide = ctx->xtree.wide != nullptr;      // the tree as two-level records (kdpt_lean.cuh)
        if (mode == KDPT_FULLSTACK) {
            if (wide) traceLeanKernel<false, true><<<lgrid, kLeanThreads, 0, stream>>>(ctx->xtree, dRays, dTmin, n, dHits, counter, order, 0.00001f);
            else traceLeanKernel<false, false><<<lgrid, kLeanThreads, 0, stream>>>(ctx->xtree, dRays, dTmin, n, dHits, counter, order, 0.00001f);
        } else {
            if (wide) traceLeanKernel<true, true><<<lgrid, kLeanThreads, 0, stream>>>(ctx->xtree, dRays, dTmin, n, dHits, counter, order, 0.0001f);
            else traceLeanKernel<true, false><<<lgrid, kLeanThreads, 0, stream>>>(ctx->xtree, dRays, dTmin, n, dHits, counter, order, 0.0001f);
        }
        KDPT_CUDA(cudaGetLastError());
        return KDPT_OK;
    }
    switch (mode) {
    case KDPT_FULLSTACK: traceKernel<KDPT_FULLSTACK><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter, order, blockTickets); break;
    case KDPT_SHORTSTACK: traceKernel<KDPT_SHORTSTACK><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter, order, blockTickets); break;
    case KDPT_REF_T1: traceKernel<KDPT_REF_T1><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter, order, blockTickets); break;
    case KDPT_REF_T2: traceKernel<KDPT_REF_T2><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter, order, blockTickets); break;
    case KDPT_BRUTE: traceKernel<KDPT_BRUTE><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter, order, blockTickets); break;
    case KDPT_BRUTE_BBOX: traceKernel<KDPT_BRUTE_BBOX><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter, order, blockTickets); break;
    default: setError("kdpt_trace: unknown mode"); return KDPT_ERR_ARG;
    }
    KDPT_CUDA(cudaGetLastError());
    return KDPT_OK;
}

} // namespace kdpt
