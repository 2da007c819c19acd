// bulk_probe.cu -- development microbenchmark (not part of the product): latency / throughput of cp.async.bulk global->shared
// from an L2-resident buffer, one CTA per SM, depth D copies in flight, copy size B bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(const double* src, size_t src_doubles, int bytes, int depth, int iters, long long* out)
{
    extern __shared__ __align__(128) double sm[];
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(bar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t stride = bytes / 8;
    size_t off = ((size_t)blockIdx.x * 7919 * stride) % (src_doubles - stride * 4);
    off = off / 16 * 16;
    long long t0 = clock64();
    for (int i = 0; i < depth - 1 && i < iters; ++i) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar + i)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(s32(sm + (size_t)i * stride)), "l"(src + off), "r"(bytes), "r"(s32(bar + i)) : "memory");
        off = (off + 148 * stride * 3) % (src_doubles - stride * 4);
    }
    for (int i = 0; i < iters; ++i) {
        const int nx = i + depth - 1;
        if (nx < iters) {
            const int s = nx % depth;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar + s)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(s32(sm + (size_t)s * stride)), "l"(src + off), "r"(bytes), "r"(s32(bar + s)) : "memory");
            off = (off + 148 * stride * 3) % (src_doubles - stride * 4);
        }
        const int s = i % depth;
        const unsigned parity = (i / depth) & 1;
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar + s)), "r"(parity) : "memory");
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
}

int main()
{
    const size_t n = (size_t)100 * 1024 * 1024 / 8;   // 100 MB: the matrix arena size, L2 resident
    double* src; long long* out;
    cudaMalloc(&src, n * 8); cudaMemset(src, 0, n * 8); cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int grid : {1, 148})
        for (int bytes : {1408, 5760, 11520})
            for (int depth : {1, 2, 3, 4, 8}) {
                if ((size_t)bytes * depth > 190 * 1024) continue;
                for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, (size_t)bytes * depth>>>(src, n, bytes, depth + 1 - 1 ? depth : 1, iters, out);
                cudaDeviceSynchronize();
                long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
                printf("grid %3d bytes %5d depth %d : %.0f clk per copy, %.2f B/clk/SM  (%s)\n", grid, bytes, depth, avg / iters, bytes / (avg / iters), cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
