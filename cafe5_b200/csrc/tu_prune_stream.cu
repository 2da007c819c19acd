// tu_prune_stream.cu -- instantiations of prune_dmma_kernel<TMW, TNW> (DMMA, both operands streamed).
#include "launchers.h"
#include "prune_dmma.cuh"

namespace cafe {
namespace {
template <int TMW, int TNW>
cudaError_t go(int grid, int n_stages, cudaStream_t stream, const PruneParams& p)
{
    const size_t smem = DmmaCfg<TMW, TNW>::smem_bytes(n_stages);
    cudaError_t e = allow_max_smem(prune_dmma_kernel<TMW, TNW>);
    if (e != cudaSuccess) return e;
    prune_dmma_kernel<TMW, TNW><<<grid, PRUNE_THREADS, smem, stream>>>(p, n_stages);
    return cudaGetLastError();
}
template <int TNW>
cudaError_t by_tm(int TM, int grid, int n_stages, cudaStream_t stream, const PruneParams& p)
{
    switch (TM) {
    case 8: return go<8, TNW>(grid, n_stages, stream, p);
    case 9: return go<9, TNW>(grid, n_stages, stream, p);
    case 10: return go<10, TNW>(grid, n_stages, stream, p);
    case 11: return go<11, TNW>(grid, n_stages, stream, p);
    case 12: return go<12, TNW>(grid, n_stages, stream, p);
    default: return go<13, TNW>(grid, n_stages, stream, p);
    }
}
}  // namespace

cudaError_t launch_prune_stream(int TM, int TNW, int grid, int n_stages, cudaStream_t stream, const PruneParams& p)
{
    switch (TNW) {
    case 4: return by_tm<4>(TM, grid, n_stages, stream, p);
    case 2: return by_tm<2>(TM, grid, n_stages, stream, p);
    default: return by_tm<1>(TM, grid, n_stages, stream, p);
    }
}
}  // namespace cafe
