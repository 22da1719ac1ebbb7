// kdpt_trace.cu — traversal + intersection kernels (sm_100a), the kdpt_trace / kdpt_trace_device entry points.
//
// Replaces the per-ray body of pathTraceOneBounceKDbare (src/pathtrace.cu:1651-1681): traverseKDbare (881-1020) and
// traverseKDbareShortHybrid (1023-1235). Compiled with -fmad=false: see kdpt_core.cuh.
#include "kdpt_internal.h"

namespace kdpt {

constexpr int kTraceThreads = 128;
#ifndef KDPT_COOP_MIN
#define KDPT_COOP_MIN 8
#endif
constexpr int kCoopMin = KDPT_COOP_MIN;   // leaves with at least this many triangles are tested by the whole warp

// Persistent threads: a fixed grid (a multiple of the SM count) whose warps pull 32-ray packets from a global
// counter until the batch is drained, so a warp stuck in a fat leaf does not hold back the rest of its block's
// share of the batch.
#ifndef KDPT_TRACE_MIN_BLOCKS
#define KDPT_TRACE_MIN_BLOCKS 7    // <= 72 registers: 28 warps per SM (measured best of 4..8 on dragon_5 and the 1 M / 10 M soups)
#endif
template <int MODE>
__global__ void __launch_bounds__(kTraceThreads, KDPT_TRACE_MIN_BLOCKS)
traceKernel(Tree T, const kdpt_ray* __restrict__ rays, const float* __restrict__ tMinIn, int n,
            kdpt_hit* __restrict__ hits, unsigned* __restrict__ counter)
{
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned)n) break;
        unsigned i = base + lane;
        const bool active = i < (unsigned)n;
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
        else if (active) {
            if (MODE == KDPT_REF_T1) traceRef<false>(T, o, d, tin, H, nullptr);
            else traceRef<true>(T, o, d, tin, H, nullptr);
        }
        if (active) {
            float4* out = reinterpret_cast<float4*>(hits + i);
            out[0] = make_float4(H.t, H.z, H.u, H.v);
            out[1] = make_float4(__int_as_float(H.tri), H.nx, H.ny, H.nz);
        }
    }
}

int launchTrace(kdpt_ctx* ctx, const kdpt_ray* dRays, const float* dTmin, int n, kdpt_hit* dHits, int mode,
                cudaStream_t stream)
{
    if (n <= 0) return KDPT_OK;
    unsigned* counter = ctx->counter.as<unsigned>();
    KDPT_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
    int blocksPerSM = 8;
    int grid = ctx->numSMs * blocksPerSM;
    int need = (n + kTraceThreads - 1) / kTraceThreads;
    if (grid > need) grid = need;
    switch (mode) {
    case KDPT_FULLSTACK: traceKernel<KDPT_FULLSTACK><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter); break;
    case KDPT_SHORTSTACK: traceKernel<KDPT_SHORTSTACK><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter); break;
    case KDPT_REF_T1: traceKernel<KDPT_REF_T1><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter); break;
    case KDPT_REF_T2: traceKernel<KDPT_REF_T2><<<grid, kTraceThreads, 0, stream>>>(ctx->tree, dRays, dTmin, n, dHits, counter); break;
    default: setError("kdpt_trace: unknown mode"); return KDPT_ERR_ARG;
    }
    KDPT_CUDA(cudaGetLastError());
    return KDPT_OK;
}

} // namespace kdpt
