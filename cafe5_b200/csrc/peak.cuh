// peak.cuh -- FP64 pipe microbenchmarks used as the roofline denominator for the pruning kernel.
// MEASURED_PEAKS.json carries HBM and bf16 figures only; the pruning contraction is bound by the FP64
// pipe, so bench.py measures that pipe on the same GPU, in the same process, right before the timed run.
#pragma once
#include <cuda_runtime.h>

namespace cafe {

// Register-resident DFMA chains: 16 independent accumulators per thread, no memory traffic.
__global__ void __launch_bounds__(256)
dfma_peak_kernel(double* out, int iters, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;   // never true; keeps the chain alive
}

// mma.sync.m8n8k4 FP64 (DMMA): 8 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256)
dmma_peak_kernel(double* out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// Larger DMMA shapes (sm_90+ PTX): m16n8k8 (A 4 regs, B 2, C 4) and m16n8k16 (A 8, B 4, C 4); 4 independent tiles per warp.
__global__ void __launch_bounds__(256)
dmma_peak_kernel_k8(double* out, int iters, double a, double b)
{
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1.0; c[i][3] = 2.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4}, {%5,%5}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256)
dmma_peak_kernel_k16(double* out, int iters, double a, double b)
{
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1.0; c[i][3] = 2.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

}  // namespace cafe
