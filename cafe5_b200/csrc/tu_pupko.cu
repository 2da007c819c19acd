// tu_pupko.cu -- instantiations of pupko_kernel<TM, TN>.
#define CAFE_PUPKO_LAUNCH_IMPL
#include "launchers.h"

namespace cafe {
cudaError_t launch_pupko(int TM, int TN, int grid, int S, cudaStream_t stream, const PupkoParams& p, int threads)
{
    return launch_pupko_impl(TM, TN, grid, S, stream, p, threads);
}
}  // namespace cafe
