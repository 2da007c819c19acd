// simulate.cuh -- family simulator on the device.
// Reference behaviour: simulator::create_trial (src/simulator.cpp:29-58): the root size of family f is given (the reference reads
// its vectorised root distribution at index f), every other node draws its size from row `parent size` of its branch's transition
// matrix restricted to sizes < max_sim (set_weighted_random_family_size, src/probability.cpp:449-476; matrix::select_random_y,
// src/matrix_cache.cpp:60-66: a discrete distribution over the un-normalised row), a lost family stays lost, and a family that
// does not exist at the root (src/gene_family.cpp:62-91) is redrawn, up to max_redraws times (50 in the reference's simulator, 0 for the conditional
// distributions of the p-value path, whose create_family, src/probability.cpp:355-375, never redraws).  Under the gamma model each family first
// picks one rate category with the category probabilities (gamma_model::get_simulation_lambda, src/gamma_core.cpp:91-95).
// The random stream is counter-based (Philox4x32-10, keyed by the seed, counter = family index), so a run is reproducible for a
// given seed whatever the launch geometry; it is NOT the reference's std::mt19937 stream: parity is distributional.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cafe {

struct Philox {
    uint32_t ctr[4], key[2], out[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t stream)
    {
        ctr[0] = 0; ctr[1] = 0; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        have = 0;
    }
    __device__ void refill()
    {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
        if (++ctr[0] == 0) ++ctr[1];
        have = 4;
    }
    __device__ uint32_t next()
    {
        if (have == 0) refill();
        return out[--have];
    }
    __device__ double uniform()   // [0, 1) with 53 random bits
    {
        const uint64_t hi = next(), lo = next();
        return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
    }
};

// cdf[(mat * max_sim + c) * N + s] = sum_{c' <= c} P(s -> c'), summed left to right as a discrete distribution accumulates it
__global__ void __launch_bounds__(256)
sim_cdf_kernel(const double* __restrict__ arena, int LD, int N, int max_sim, double* __restrict__ cdf)
{
    const double* PT = arena + (size_t)blockIdx.x * LD * LD;
    double* out = cdf + (size_t)blockIdx.x * max_sim * N;
    for (int s = threadIdx.x; s < N; s += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < max_sim; ++c) {
            acc += PT[(size_t)c * LD + s];
            out[(size_t)c * N + s] = acc;
        }
    }
}

struct SimParams {
    const int32_t* parent;       // [n_nodes]
    const int32_t* leaf_col;     // [n_nodes]
    const int32_t* mat_of;       // [K][n_nodes]
    const double* cdf;           // [n_mats][max_sim][N]
    const double* cat_probs;     // [K]
    const int32_t* root_sizes;   // [F]
    int32_t* sizes;              // [n_nodes][F] scratch / output
    uint8_t* has;                // [n_nodes][F] scratch: the subtree holds a leaf with a positive count
    int32_t* counts;             // [F][n_species]
    int32_t* categories;         // [F]
    unsigned long long* exhausted;
    const double* em;            // [em_rows][3] error model of the context, or nullptr: leaf counts are perturbed like adjust_for_error_model
    int32_t em_rows;
    int64_t F;
    uint64_t seed;
    int32_t n_nodes, n_species, K, N, max_sim, max_attempts;
};

__global__ void __launch_bounds__(256)
simulate_kernel(const SimParams p)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.F) return;
    Philox rng(p.seed, (uint64_t)f);
    int cat = 0;
    if (p.K > 1) {                                    // discrete_distribution over the category probabilities
        double tot = 0.0;
        for (int k = 0; k < p.K; ++k) tot += p.cat_probs[k];
        const double u = rng.uniform() * tot;
        double acc = 0.0;
        cat = p.K - 1;
        for (int k = 0; k < p.K; ++k) { acc += p.cat_probs[k]; if (u < acc) { cat = k; break; } }
    }
    p.categories[f] = cat;
    const int32_t* mat_of = p.mat_of + (size_t)cat * p.n_nodes;
    const int root = p.n_nodes - 1;
    bool ok = false;
    for (int attempt = 0; attempt < p.max_attempts && !ok; ++attempt) {
        p.sizes[(size_t)root * p.F + f] = p.root_sizes[f];
        for (int i = root - 1; i >= 0; --i) {         // parents have larger indices: top-down
            const int ps = p.sizes[(size_t)p.parent[i] * p.F + f];
            int c = 0;
            if (ps > 0) {                             // a lost family stays lost (probability.cpp:459)
                const double* col = p.cdf + (size_t)mat_of[i] * p.max_sim * p.N + ps;
                const double u = rng.uniform() * col[(size_t)(p.max_sim - 1) * p.N];
                int lo = 0, hi = p.max_sim - 1;       // smallest c with cdf[c] > u
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (col[(size_t)mid * p.N] > u) hi = mid; else lo = mid + 1;
                }
                c = lo;
            }
            if (p.em != nullptr && p.leaf_col[i] >= 0) {
                // adjust_for_error_model (src/probability.cpp:478-499): one uniform draw moves the simulated leaf count down / up with
                // the model's probabilities for that size (the reference throws beyond the model's last size; rows are clamped here)
                const double* pr = p.em + (size_t)min(c, p.em_rows - 1) * 3;
                const double rnd = rng.uniform();
                if (rnd < pr[0]) c = max(c - 1, 0);
                else if (rnd > (1 - pr[2])) c = c + 1;
            }
            p.sizes[(size_t)i * p.F + f] = c;
        }
        // exists_at_root: every child of the root has a descendant leaf with a positive count (gene_family.cpp:62-91)
        for (int i = 0; i <= root; ++i) p.has[(size_t)i * p.F + f] = 0;
        ok = true;
        for (int i = 0; i < root; ++i) {              // children have smaller indices: bottom-up
            uint8_t h = p.has[(size_t)i * p.F + f];
            if (p.leaf_col[i] >= 0) h = p.sizes[(size_t)i * p.F + f] > 0;
            if (p.parent[i] == root) { if (!h) ok = false; }
            else if (h) p.has[(size_t)p.parent[i] * p.F + f] = 1;
        }
    }
    if (!ok) atomicAdd(p.exhausted, 1ull);
    for (int i = 0; i <= root; ++i)
        if (p.leaf_col[i] >= 0) p.counts[(size_t)f * p.n_species + p.leaf_col[i]] = p.sizes[(size_t)i * p.F + f];
}

}  // namespace cafe
