// tu_prune_dfma.cu -- instantiations of prune_kernel<TM, TN> (DFMA register tiles).
#include "launchers.h"

namespace cafe {
namespace {
template <int TM, int TN>
cudaError_t go(int grid, int S, cudaStream_t stream, const PruneParams& p)
{
    const size_t smem = PruneCfg<TM, TN>::smem_bytes(S);
    cudaError_t e = allow_max_smem(prune_kernel<TM, TN>);
    if (e != cudaSuccess) return e;
    prune_kernel<TM, TN><<<grid, PRUNE_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}
template <int TN>
cudaError_t by_tm(int TM, int grid, int S, cudaStream_t stream, const PruneParams& p)
{
    switch (TM) {
    case 8: return go<8, TN>(grid, S, stream, p);
    case 9: return go<9, TN>(grid, S, stream, p);
    case 10: return go<10, TN>(grid, S, stream, p);
    case 11: return go<11, TN>(grid, S, stream, p);
    case 12: return go<12, TN>(grid, S, stream, p);
    default: return go<13, TN>(grid, S, stream, p);
    }
}
}  // namespace

cudaError_t launch_prune_dfma(int TM, int TN, int grid, int S, cudaStream_t stream, const PruneParams& p)
{
    switch (TN) {
    case 4: return by_tm<4>(TM, grid, S, stream, p);
    case 2: return by_tm<2>(TM, grid, S, stream, p);
    default: return by_tm<1>(TM, grid, S, stream, p);
    }
}
}  // namespace cafe
