// viterbi.cuh -- per-branch change probabilities of the reconstruction report (compute_viterbi_sum,
// src/gene_family_reconstructor.cpp:388-429): for a branch whose reconstructed parent and child sizes are (p, c), the probability
// mass of all child sizes m < max_family_size that are less likely than c given p, plus half the mass of those exactly as likely.
// One thread per (family, node); the matrix column P(p -> .) is a strided walk through the transposed matrix.  The root has no
// branch and unselected families (p-value above the threshold, src/execute.cpp:178-184) are skipped: both are reported as -1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cafe {

__global__ void __launch_bounds__(256)
viterbi_sum_kernel(const double* __restrict__ arena, const int32_t* __restrict__ mat_of, const int32_t* __restrict__ parent,
                   const int32_t* __restrict__ states, const uint8_t* __restrict__ selected, int64_t F, int n_nodes, int LD,
                   int max_family_size, double* __restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= F * n_nodes) return;
    const int64_t f = idx / n_nodes;
    const int node = (int)(idx % n_nodes);
    if (parent[node] < 0 || (selected != nullptr && !selected[f])) { out[idx] = -1.0; return; }
    const int ps = states[f * n_nodes + parent[node]], cs = states[idx];
    const double* __restrict__ col = arena + (size_t)mat_of[node] * LD * LD + ps;      // P(ps -> m) = PT[m][ps]
    const double calculated = col[(size_t)cs * LD];
    double result = 0.0;
    for (int m = 0; m < max_family_size; ++m) {
        const double p = col[(size_t)m * LD];
        if (p == calculated) result = __dadd_rn(result, p / 2.0);
        else if (p < calculated) result = __dadd_rn(result, p);
    }
    out[idx] = result;
}

}  // namespace cafe
