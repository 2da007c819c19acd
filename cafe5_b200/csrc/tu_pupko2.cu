// tu_pupko2.cu -- instantiations of pupko2_kernel<TM, TN, THREADS, JOBS> and the traceback kernel (pupko2.cuh).
#define CAFE_PUPKO2_LAUNCH_IMPL
#include "launchers.h"
#include "pupko2.cuh"

namespace cafe {
namespace {
template <int TM, int TN, bool JOBS>
cudaError_t go(int grid, cudaStream_t stream, const Pupko2Params& p)
{
    constexpr int T = TN >= 2 ? 512 : 256;
    const size_t smem = pupko2_smem_bytes(TM, TN);
    cudaError_t e = allow_max_smem(pupko2_kernel<TM, TN, T, JOBS>);
    if (e != cudaSuccess) return e;
    pupko2_kernel<TM, TN, T, JOBS><<<grid, T, smem, stream>>>(p);
    return cudaGetLastError();
}
template <int TN, bool JOBS>
cudaError_t by_tm(int TM, int grid, cudaStream_t stream, const Pupko2Params& p)
{
    switch (TM) {
    case 8: return go<8, TN, JOBS>(grid, stream, p);
    case 9: return go<9, TN, JOBS>(grid, stream, p);
    case 10: return go<10, TN, JOBS>(grid, stream, p);
    case 11: return go<11, TN, JOBS>(grid, stream, p);
    case 12: return go<12, TN, JOBS>(grid, stream, p);
    default: return go<13, TN, JOBS>(grid, stream, p);
    }
}
template <bool JOBS>
cudaError_t by_tn(int TM, int TN, int grid, cudaStream_t stream, const Pupko2Params& p)
{
    switch (TN) {
    case 4: return by_tm<4, JOBS>(TM, grid, stream, p);
    case 2: return by_tm<2, JOBS>(TM, grid, stream, p);
    default: return by_tm<1, JOBS>(TM, grid, stream, p);
    }
}
}  // namespace

cudaError_t launch_pupko2(int TM, int TN, int grid, cudaStream_t stream, const Pupko2Params& p, bool jobs)
{
    return jobs ? by_tn<true>(TM, TN, grid, stream, p) : by_tn<false>(TM, TN, grid, stream, p);
}

cudaError_t launch_pupko_traceback(cudaStream_t stream, const void* ctab, const int64_t* c_off, const int64_t* c_cols, const int32_t* parent,
                                   const int32_t* leaf_col, const int32_t* tab_of, const int32_t* tab_ids, const int32_t* root_state,
                                   int64_t U, int64_t U_stride, int n_nodes, int K, int SP, int32_t* states)
{
    const int64_t n = U * K;
    pupko_traceback_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ctab, c_off, c_cols, parent, leaf_col, tab_of, tab_ids, root_state,
                                                                             U, U_stride, n_nodes, K, SP, states);
    return cudaGetLastError();
}
}  // namespace cafe
