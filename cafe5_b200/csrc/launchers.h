// launchers.h -- host-callable launchers of the templated kernels.  Each family of instantiations lives in its own
// translation unit (tu_*.cu) so the library builds in parallel; cafe_b200.cu holds the likelihood-path C ABI and the
// small non-template kernels, context.cuh the context they share with abi_analysis.cu.
#pragma once
#include "kernels.cuh"
#include "pupko.cuh"
#include "pupko2.cuh"

namespace cafe {

// Opt a kernel into the device's whole opt-in shared-memory range.  Always the SAME value for a given kernel: contexts on several
// host threads (device shards, buckets) launch the same instantiation with different dynamic sizes, and a per-launch value would
// let one thread lower the limit between another thread's attribute call and its launch.
template <typename Kernel>
inline cudaError_t allow_max_smem(Kernel kernel)
{
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
}

// DFMA register-tile kernel (kernels.cuh).  TM in 8..13, TN in {1,2,4}.
cudaError_t launch_prune_dfma(int TM, int TN, int grid, int S, cudaStream_t stream, const PruneParams& p);
// DMMA kernel streaming both operands (prune_dmma.cuh).  TNW in {1,2,4}.
cudaError_t launch_prune_stream(int TM, int TNW, int grid, int n_stages, cudaStream_t stream, const PruneParams& p);
// DMMA kernel with the child vector resident in shared memory (prune_resident.cuh).
//   wn4: one 256-thread CTA per SM, BN = 32*TNW, BK = 8;  wn2: two 128-thread CTAs per SM, BN = 16*TNW, BK = 4.
cudaError_t launch_prune_resident_wn4(int TM, int TNW, int grid, int N, int n_stages, cudaStream_t stream, const PruneParams& p, const InlineSchedule& sched);
cudaError_t launch_prune_resident_wn2(int TM, int TNW, int grid, int N, int n_stages, cudaStream_t stream, const PruneParams& p, const InlineSchedule& sched);
// table build of the wn2 geometry: factor tables of the nodes listed as jobs in sched (subtree-pattern reuse)
cudaError_t launch_prune_resident_wn2_tables(int TM, int TNW, int grid, int N, int n_stages, cudaStream_t stream, const PruneParams& p, const InlineSchedule& sched);
cudaError_t launch_prune_resident_wn2probe(int TM, int TNW, int grid, int N, int n_stages, cudaStream_t stream, const PruneParams& p, const InlineSchedule& sched);
// Pupko reconstruction (pupko.cuh).  threads = 512 (default geometry when TN >= 2) or 256.
cudaError_t launch_pupko(int TM, int TN, int grid, int S, cudaStream_t stream, const PupkoParams& p, int threads);

// Pupko, second design (pupko2.cuh): argmax tables + chained values + subtree-pattern tables; jobs = a table launch
cudaError_t launch_pupko2(int TM, int TN, int grid, cudaStream_t stream, const Pupko2Params& p, bool jobs);
cudaError_t launch_pupko_traceback(cudaStream_t stream, const void* ctab, const int64_t* c_off, const int64_t* c_cols, const int32_t* parent,
                                   const int32_t* leaf_col, const int32_t* tab_of, const int32_t* tab_ids, const int32_t* root_state,
                                   int64_t U, int64_t U_stride, int n_nodes, int K, int SP, int32_t* states);

}  // namespace cafe
