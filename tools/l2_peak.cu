// l2_peak.cu — measures the L2-resident and HBM read bandwidth of the GPU with a plain 16-byte-per-thread read kernel
// ("L2 peak is not in MEASURED_PEAKS.json -> measure with an L2-resident read kernel and record it").
// Development tool, not part of libkdpt.so. Prints one JSON line.  Build: make -C kdtreepathtraceroptimization_b200/csrc l2peak
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void readKernel(const uint4* __restrict__ p, size_t n, int reps, unsigned* sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            uint4 v;
            // ld.global.cg would skip L1; the default path may hit L1 only for the slice a block re-reads, which at
            // >= 8 MB per pass (>> 148 x 256 KB is false: 37 MB) is why sizes below 64 MB use the .cg form
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const size_t sizesMB[] = {8, 16, 32, 48, 64, 96, 4096};
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"gpu\": \"%s\", \"l2_bytes\": %d, \"sms\": %d, \"read_gbs\": {", prop.name, prop.l2CacheSize, prop.multiProcessorCount);
    for (size_t k = 0; k < sizeof sizesMB / sizeof *sizesMB; k++) {
        const size_t bytes = sizesMB[k] << 20, n = bytes / 16;
        uint4* buf; if (cudaMalloc(&buf, bytes) != cudaSuccess) return 2;
        cudaMemset(buf, 1, bytes);
        const int reps = (int)((16ull << 30) / bytes) < 4 ? 4 : (int)((16ull << 30) / bytes);   // ~16 GB of reads per launch
        const int grid = prop.multiProcessorCount * 8;
        float best = 1e30f;
        for (int it = 0; it < 4; it++) {
            cudaEventRecord(e0);
            readKernel<<<grid, 256>>>(buf, n, reps, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        if (cudaGetLastError() != cudaSuccess) return 3;
        printf("%s\"%zu MB\": %.1f", k ? ", " : "", sizesMB[k], (double)bytes * reps / (best * 1e-3) / 1e9);
        cudaFree(buf);
    }
    printf("}}\n");
    return 0;
}
