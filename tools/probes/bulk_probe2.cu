// bulk_probe2.cu -- development microbenchmark: where do the ~560 cycles per cp.async.bulk go?  (issue cost vs completion latency vs try_wait)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint64_t* bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool test(uint64_t* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(128, 1) probe(const double* src, int bytes, int iters, long long* out, int mode)
{
    extern __shared__ __align__(128) double sm[];
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    long long t_issue = 0, t_done = 0, t_wait_done = 0;
    size_t off = (size_t)blockIdx.x * 65536;
    for (int i = 0; i < iters; ++i) {
        long long t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(s32(sm)), "l"(src + off), "r"(bytes), "r"(s32(bar)) : "memory");
        long long t1 = clock64();
        if (mode == 0) wait(bar, i & 1);                       // try_wait spin
        else while (!test(bar, i & 1)) { }                      // test_wait spin (non-blocking poll)
        long long t2 = clock64();
        wait(bar, i & 1);                                       // wait on an already-complete phase
        long long t3 = clock64();
        t_issue += t1 - t0; t_done += t2 - t1; t_wait_done += t3 - t2;
        off = (off + 148 * 65536 + 1440) % (12 * 1024 * 1024);
        off = off / 16 * 16;
    }
    out[blockIdx.x * 3 + 0] = t_issue; out[blockIdx.x * 3 + 1] = t_done; out[blockIdx.x * 3 + 2] = t_wait_done;
}
int main()
{
    const size_t n = (size_t)100 * 1024 * 1024 / 8;
    double* src; long long* out;
    cudaMalloc(&src, n * 8); cudaMemset(src, 0, n * 8); cudaMalloc(&out, 148 * 3 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int mode : {0, 1})
        for (int grid : {1, 148})
            for (int bytes : {16, 1408, 11520}) {
                for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, 16384>>>(src, bytes, iters, out, mode);
                cudaDeviceSynchronize();
                long long h[148 * 3]; cudaMemcpy(h, out, grid * 3 * 8, cudaMemcpyDeviceToHost);
                double a = 0, b = 0, c = 0; for (int i = 0; i < grid; ++i) { a += h[i * 3]; b += h[i * 3 + 1]; c += h[i * 3 + 2]; }
                printf("%s grid %3d bytes %5d : issue %.0f clk, completion after issue %.0f clk, wait on complete phase %.0f clk (%s)\n", mode ? "test_wait" : "try_wait ",
                       grid, bytes, a / grid / iters, b / grid / iters, c / grid / iters, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
