// pupko.cuh -- Pupko joint ancestral reconstruction on the device (sm_100a).
// Reference: src/gene_family_reconstructor.cpp:30-190 (up-pass with argmax table C, root pick with the
// prior, traceback), src/gamma_core.cpp:271-288,351-357 (per-category passes, weighted average, round).
// Same persistent-CTA tiling as prune_kernel: a CTA owns BN unique families of one category and walks
// the whole tree; the (x, +) contraction is replaced by (x, max), and the argmax the traceback needs is
// recomputed for the one parent state it is asked for, j scanned ASCENDING so that the reference's
// strict '>' tie-break (first maximum wins) is reproduced exactly.
#pragma once
#include "kernels.cuh"

namespace cafe {

struct PupkoParams {
    const Step* steps;
    const StepChild* children;
    const int32_t* mat_of;      // [K][n_nodes]
    const double* arena;
    const int32_t* counts_t;
    const double* prior_d;      // [>= root_len]
    double* scratch;            // [grid][n_slots][slot_stride]  L vectors
    double* mstore;             // [grid][n_steps][m_stride]     M_v = prod_children L_child of every node of the tile in flight
    int32_t* states;            // [K][U_stride][n_nodes] reconstructed states of internal nodes
    int64_t U, U_stride, slot_stride, m_stride;
    int32_t n_steps, n_nodes, n_slots;
    int32_t LD, S, R, N, K, root_len;
    int32_t n_col_tiles, n_mtiles;
};

__host__ __device__ constexpr int pupko_tmp(int tm) { return ((tm + 1) & ~1) % 4 == 0 ? ((tm + 1) & ~1) + 2 : ((tm + 1) & ~1); }

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

// The reference's up-pass keeps, for every node v and every parent state i, L_v[i] = max_j M_v[j] P_v(i -> j) AND the first
// maximiser C_v[i] (reconstruct_internal_node, :78-114); the traceback then reads ONE entry C_v[state of the parent] per node
// (:173-188).  Carrying the argmax through the max-product costs a third select per (i, j) pair on the half-rate ALU pipe and a
// register per accumulator.  Here the up-pass keeps only the values (DMUL, DSETP, 2 FSEL per pair) and parks M_v in global
// memory; the traceback recomputes the one row it needs, argmax_j M_v[j] P_v(parent state -> j), with the same products
// (same operands, same rounding) scanned for the FIRST maximum - identical states, 1/S of the up-pass work.
//
// Geometry: 16 thread-rows x NTN = THREADS / 16 thread-columns; a thread owns TM rows (i * 16 + tm) and CT = BN / NTN columns.
//   THREADS = 512 (BN >= 32): a warp is one thread-row, so the matrix operand is a broadcast load (the stage is laid out
//                 [kk][tm][TM] so that it is TM/2 16-byte loads); 4 warps per sub-partition at <= 128 registers.
//   THREADS = 256: BN = 16 (few families), or CAFE_B200_PUPKO_THREADS=256.
template <int TM, int TN, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
pupko_kernel(const PupkoParams p)
{
    constexpr int BM = 16 * TM, BN = 16 * TN, BK = PRUNE_BK, STAGES = PRUNE_STAGES;
    constexpr int NTN = THREADS / 16, CT = BN / NTN;
    constexpr int BNP = BN + 2;                  // row stride of Ms: the leaf gather writes a column (32 rows of one column per warp)
    // thread-row block of the matrix stage: even (16-byte loads) and an odd multiple of 2, so that the 8-byte cp.async writes of 16
    // consecutive rows (stride TMP doubles) spread over the banks instead of folding onto four of them
    constexpr int TMP = pupko_tmp(TM);
    constexpr int AST = 16 * TMP;                // doubles per k-row of a stage
    constexpr int SUBS = THREADS / BN;           // traceback: lanes per column
    static_assert(CT * NTN == BN && CT >= 1, "column tile does not divide");
    static_assert(SUBS >= 1 && SUBS <= 32 && (SUBS & (SUBS - 1)) == 0, "traceback lanes per column");
    extern __shared__ __align__(16) double smem[];
    const int kpad = (p.S + BK - 1) / BK * BK;
    double* Ms = smem;                                   // [kpad][BNP]  prod_children L_child[j]
    double* As = smem + (size_t)kpad * BNP;              // [STAGES][BK][16][TMP]
    int32_t* st_s = reinterpret_cast<int32_t*>(As + (size_t)STAGES * BK * AST);   // [n_steps][BN] traceback states

    const int tid = threadIdx.x;
    const int tn = tid % NTN, tm = tid / NTN;
    const int n_tiles = p.K * p.n_col_tiles;
    double* const my_scratch = p.scratch + (size_t)blockIdx.x * p.n_slots * p.slot_stride;
    double* const my_m = p.mstore + (size_t)blockIdx.x * p.n_steps * p.m_stride;
    const int n_chunks = kpad / BK;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int k = tile / p.n_col_tiles;
        const int64_t col0 = (int64_t)(tile % p.n_col_tiles) * BN;
        const int32_t* mat_of = p.mat_of + (size_t)k * p.n_nodes;

        for (int st = 0; st < p.n_steps; ++st) {
            const Step sp = p.steps[st];
            // ---- M[j][col] = prod over children (descendant order) of L_child[j][col] ----
            __syncthreads();
            for (int ci = 0; ci < sp.n_children; ++ci) {
                const StepChild ch = p.children[sp.child_begin + ci];
                if (ch.leaf_row >= 0) {
                    // reconstruct_leaf_node (:30-46): L[j] = P(j -> obs) = PT[obs][j]; one warp per column, j coalesced
                    const double* __restrict__ PT = p.arena + (size_t)mat_of[ch.node] * p.LD * p.LD;
                    for (int c = tid >> 5; c < BN; c += THREADS / 32) {
                        int64_t u = col0 + c;
                        if (u >= p.U) u = p.U - 1;
                        const int obs = p.counts_t[(size_t)ch.leaf_row * p.U_stride + u];
                        const double* __restrict__ r = PT + (size_t)obs * p.LD;
                        for (int j = tid & 31; j < kpad; j += 32) {
                            const double l = j < p.S ? __ldg(r + j) : 0.0;
                            Ms[(size_t)j * BNP + c] = ci == 0 ? l : __dmul_rn(Ms[(size_t)j * BNP + c], l);
                        }
                    }
                } else {
                    const double* __restrict__ src = my_scratch + (size_t)ch.slot * p.slot_stride;
                    for (int idx = tid; idx < kpad * BN; idx += THREADS) {
                        const int j = idx / BN, c = idx % BN;
                        const double l = j < p.S ? src[idx] : 0.0;
                        Ms[(size_t)j * BNP + c] = ci == 0 ? l : __dmul_rn(Ms[(size_t)j * BNP + c], l);
                    }
                }
                __syncthreads();
            }

            if (sp.is_root) {
                // reconstruct_root_node (:48-76): argmax_{1 <= j < root_len} M[j] * prior(j), strict '>' from -1
                if (tid < BN) {
                    double best = -1.0;
                    int arg = 0;
                    for (int j = 1; j < p.root_len; ++j) {
                        const double val = __dmul_rn(Ms[(size_t)j * BNP + tid], p.prior_d[j]);
                        if (val > best) { best = val; arg = j; }
                    }
                    st_s[st * BN + tid] = arg;
                }
                continue;
            }

            // park M_v for the traceback
            {
                double* __restrict__ mg = my_m + (size_t)st * p.m_stride;
                for (int idx = tid; idx < kpad * BN / 2; idx += THREADS) {
                    const int j = idx / (BN / 2), c = (idx % (BN / 2)) * 2;
                    *reinterpret_cast<double2*>(mg + (size_t)j * BN + c) = *reinterpret_cast<const double2*>(Ms + (size_t)j * BNP + c);
                }
            }

            // reconstruct_internal_node (:78-114), values only: L[i] = max_j M[j] * P(i -> j)
            const double* __restrict__ PT = p.arena + (size_t)mat_of[sp.node] * p.LD * p.LD;
            double* const out_slot = my_scratch + (size_t)sp.out_slot * p.slot_stride;
            for (int mt = 0; mt < p.n_mtiles; ++mt) {
                const int m0 = mt * BM;
                double best[TM][CT];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < CT; ++j) best[i][j] = -1.0;
                if (mt > 0) __syncthreads();
                auto load_chunk = [&](int chunk) {
                    if (chunk < n_chunks) {
                        double* dst = As + (size_t)(chunk % STAGES) * BK * AST;
                        const double* __restrict__ g = PT + (size_t)chunk * BK * p.LD + m0;
                        for (int idx = tid; idx < BK * BM; idx += THREADS) {
                            const int kk = idx / BM, mm = idx % BM;
                            cp_async8(dst + kk * AST + (mm % 16) * TMP + mm / 16, g + (size_t)kk * p.LD + mm);
                        }
                    }
                    cp_async_commit();
                };
#pragma unroll
                for (int s = 0; s < STAGES - 1; ++s) load_chunk(s);
                for (int chunk = 0; chunk < n_chunks; ++chunk) {
                    cp_async_wait<STAGES - 2>();
                    __syncthreads();
                    load_chunk(chunk + STAGES - 1);
                    const double* a_s = As + (size_t)(chunk % STAGES) * BK * AST + tm * TMP;
                    const double* b_s = Ms + (size_t)chunk * BK * BNP;
#pragma unroll
                    for (int kk = 0; kk < BK; ++kk) {
                        double a[(TM + 1) & ~1], b[CT];
#pragma unroll
                        for (int i = 0; i < ((TM + 1) & ~1); i += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(a_s + kk * AST + i);
                            a[i] = v.x; a[i + 1] = v.y;
                        }
                        if (CT == 2) {
                            const double2 v = *reinterpret_cast<const double2*>(b_s + kk * BNP + 2 * tn);
                            b[0] = v.x; b[CT - 1] = v.y;
                        } else {
#pragma unroll
                            for (int j = 0; j < CT; ++j) b[j] = b_s[kk * BNP + col_of<CT>(tn, j)];
                        }
#pragma unroll
                        for (int i = 0; i < TM; ++i)
#pragma unroll
                            for (int j = 0; j < CT; ++j) {
                                const double val = __dmul_rn(b[j], a[i]);   // value * matrix->get(i, j)
                                best[i][j] = val > best[i][j] ? val : best[i][j];
                            }
                    }
                }
                cp_async_wait<0>();
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < CT; ++j)
                        out_slot[(size_t)(m0 + i * 16 + tm) * BN + col_of<CT>(tn, j)] = best[i][j];
            }
        }
        // ---- traceback (:173-188): parents were scheduled after their children, so walk the steps backwards.  SUBS lanes of one
        // warp share a column: each scans j = sub, sub + SUBS, ... for its first maximum of M_v[j] * P_v(parent state -> j); the
        // butterfly keeps the larger value and, on a tie, the smaller j - the first maximum of the reference's ascending scan.
        __syncthreads();
        {
            const int c = tid / SUBS, sub = tid % SUBS;
            const int64_t u = col0 + c;
            for (int st = p.n_steps - 1; st >= 0; --st) {
                const Step sp = p.steps[st];
                int s;
                if (sp.is_root) s = st_s[st * BN + c];
                else {
                    const int ps = st_s[sp.parent_step * BN + c];
                    const double* __restrict__ PTc = p.arena + (size_t)mat_of[sp.node] * p.LD * p.LD + ps;
                    const double* __restrict__ mg = my_m + (size_t)st * p.m_stride + c;
                    double best = -1.0;
                    int arg = 0;
                    constexpr int TB = 8;                    // loads in flight per lane
                    for (int j0 = sub; j0 < p.S; j0 += SUBS * TB) {
                        double mv[TB], pv[TB];
#pragma unroll
                        for (int t = 0; t < TB; ++t) {
                            const int j = j0 + t * SUBS;
                            mv[t] = j < p.S ? mg[(size_t)j * BN] : 0.0;
                            pv[t] = j < p.S ? __ldg(PTc + (size_t)j * p.LD) : 0.0;
                        }
#pragma unroll
                        for (int t = 0; t < TB; ++t) {
                            const double val = __dmul_rn(mv[t], pv[t]);
                            if (j0 + t * SUBS < p.S && val > best) { best = val; arg = j0 + t * SUBS; }
                        }
                    }
#pragma unroll
                    for (int off = SUBS / 2; off > 0; off >>= 1) {
                        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                        const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
                        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                    }
                    s = arg;
                    if (sub == 0) st_s[st * BN + c] = s;
                    __syncwarp();
                }
                if (sub == 0 && u < p.U) p.states[((size_t)k * p.U_stride + u) * p.n_nodes + sp.node] = s;
            }
        }
        __syncthreads();
    }
}

// dynamic shared memory of pupko_kernel<TM, TN, *>: M_v tile + matrix stages + the traceback states of every step
inline size_t pupko_smem_bytes(int TM, int TN, int S, int n_steps)
{
    const int kpad = (S + PRUNE_BK - 1) / PRUNE_BK * PRUNE_BK;
    const int tmp = pupko_tmp(TM);
    return sizeof(double) * ((size_t)kpad * (16 * TN + 2) + (size_t)PRUNE_STAGES * PRUNE_BK * 16 * tmp) + sizeof(int32_t) * (size_t)n_steps * 16 * TN;
}

#ifdef CAFE_PUPKO_LAUNCH_IMPL   // tu_pupko.cu only
template <int TM, int TN>
inline cudaError_t launch_pupko_t(int grid, int S, cudaStream_t stream, const PupkoParams& p, int threads)
{
    const size_t smem = pupko_smem_bytes(TM, TN, S, p.n_steps);
    cudaError_t e;
    if (TN >= 2 && threads == 512) {
        constexpr int T = TN >= 2 ? 512 : 256;      // (TN = 1 never instantiates the 512-thread geometry)
        if ((e = allow_max_smem(pupko_kernel<TM, TN, T>)) != cudaSuccess) return e;
        pupko_kernel<TM, TN, T><<<grid, T, smem, stream>>>(p);
    } else {
        if ((e = allow_max_smem(pupko_kernel<TM, TN, 256>)) != cudaSuccess) return e;
        pupko_kernel<TM, TN, 256><<<grid, 256, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

template <int TN>
inline cudaError_t launch_pupko_tn(int TM, int grid, int S, cudaStream_t stream, const PupkoParams& p, int threads)
{
    switch (TM) {
    case 8: return launch_pupko_t<8, TN>(grid, S, stream, p, threads);
    case 9: return launch_pupko_t<9, TN>(grid, S, stream, p, threads);
    case 10: return launch_pupko_t<10, TN>(grid, S, stream, p, threads);
    case 11: return launch_pupko_t<11, TN>(grid, S, stream, p, threads);
    case 12: return launch_pupko_t<12, TN>(grid, S, stream, p, threads);
    default: return launch_pupko_t<13, TN>(grid, S, stream, p, threads);
    }
}

inline cudaError_t launch_pupko_impl(int TM, int TN, int grid, int S, cudaStream_t stream, const PupkoParams& p, int threads)
{
    switch (TN) {
    case 4: return launch_pupko_tn<4>(TM, grid, S, stream, p, threads);
    case 2: return launch_pupko_tn<2>(TM, grid, S, stream, p, threads);
    default: return launch_pupko_tn<1>(TM, grid, S, stream, p, threads);
    }
}

#endif  // CAFE_PUPKO_LAUNCH_IMPL

#ifdef CAFE_KERNELS_IMPL
// Expand unique -> family, fill leaves with observed counts, average the categories
// (get_weighted_averages, gamma_core.cpp:271-288: val = sum_k p_k * state_k from 0.0, then round, :356).
__global__ void __launch_bounds__(256)
expand_states_kernel(const int32_t* __restrict__ st_u, const int64_t* __restrict__ f2u, const int32_t* __restrict__ counts_t,
                     const int32_t* __restrict__ leaf_row, const double* __restrict__ cat_probs,
                     int64_t F, int64_t U_stride, int n_nodes, int K, int gamma,
                     int32_t* __restrict__ cat_states, int32_t* __restrict__ states, double* __restrict__ averaged)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= F * n_nodes) return;
    const int64_t f = idx / n_nodes;
    const int i = (int)(idx % n_nodes);
    const int64_t u = f2u[f];
    if (leaf_row[i] >= 0) {
        const int32_t v = counts_t[(size_t)leaf_row[i] * U_stride + u];
        states[idx] = v;
        if (averaged) averaged[idx] = (double)v;
        if (cat_states) for (int k = 0; k < K; ++k) cat_states[((size_t)f * K + k) * n_nodes + i] = v;
        return;
    }
    double val = 0.0;
    int32_t last = 0;
    for (int k = 0; k < K; ++k) {
        last = st_u[((size_t)k * U_stride + u) * n_nodes + i];
        if (cat_states) cat_states[((size_t)f * K + k) * n_nodes + i] = last;
        val = __dadd_rn(val, __dmul_rn(cat_probs[k], (double)last));
    }
    if (!gamma) { states[idx] = last; if (averaged) averaged[idx] = (double)last; }
    else { states[idx] = (int32_t)round(val); if (averaged) averaged[idx] = val; }
}

#endif  // CAFE_KERNELS_IMPL

}  // namespace cafe
