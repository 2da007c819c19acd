// pupko2.cuh -- Pupko joint ancestral reconstruction, second design (sm_100a).
// Reference: src/gene_family_reconstructor.cpp:30-190 (up-pass L_v[i] = max_j M_v[j] P_v(i -> j) with the first maximiser C_v[i],
// root pick with the prior, traceback state[child] = C_child[state[parent]]), src/gamma_core.cpp:271-288,351-357.
//
// What changed against pupko.cuh (round 1: 238.9 ms and 109 GB of DRAM traffic per launch on the bench shard, because every node's
// product vector M_v and value vector L_v of every family was parked in HBM for a traceback that reads one entry per node):
//   * the up-pass keeps what the traceback reads, the argmax table C_v (one byte per parent state, two when S > 256), not M_v;
//   * L_v travels like the factors of the pruning kernel: it stays in registers when the parent is the next step of the post-order
//     schedule (the common case) and is multiplied into the parent's product tile in shared memory in place; only the first-visited
//     child of a node with two internal children goes through a global slot;
//   * subtree-pattern reuse: L_v and C_v of a node whose subtree shows few distinct patterns of leaf counts are computed once per
//     PATTERN (table jobs, level by level, the plan the likelihood path uses) and gathered by the families that share it;
//   * the traceback is a separate kernel: one thread per (family, category) walks the tree top-down reading one C entry per node.
// States are identical to the reference's: the same products in the same order, the first maximum of an ascending scan.
#pragma once
#include "kernels.cuh"
#include "pupko.cuh"

namespace cafe {

struct Pupko2Params {
    const Step* steps;          // main pass: the (reduced) post-order schedule; table launch: one step per job
    const StepChild* children;
    const int32_t* jobs;        // table launch: [n_jobs][JOB_WORDS] {first tile, column tiles, patterns, pattern stride, id offset lo, hi, C offset lo, hi}
    const int32_t* mat_of;      // [K][n_nodes]
    const double* arena;
    const int32_t* ids;         // main pass: count table (+ pattern-id rows of the table nodes it gathers); table launch: the job id tables
    const double* prior_d;
    double* scratch;            // [grid][n_fslots][BM * BN] L vectors that cannot stay in registers
    double* tables;             // L tables of the table nodes: rows of LD doubles, row (rows_before * K + k * D + pattern)
    void* ctab;                 // argmax tables, uint16: node v at c_off[v], element ((k * cols_v + col) * SP + i)
    const int64_t* c_off;       // [n_nodes] offset of every internal node's argmax table (elements)
    const int64_t* c_cols;      // [n_nodes] columns of that table per category (patterns of a table node, U otherwise)
    int32_t* root_state;        // [K][U_stride]
    int64_t U, U_stride, slot_stride;
    int32_t n_steps, n_nodes, n_fslots, n_jobs, n_job_tiles;
    int32_t LD, S, SP, R, N, K, root_len;
    int32_t n_col_tiles;
};

// One persistent CTA takes (category k, tile of BN columns) and walks the schedule (main pass) or its job's single step (table launch).
// Thread (tm, tn): rows i*16 + tm (i < TM), columns col_of<CT>(tn, j).  Argmax entries are two bytes (state spaces beyond 256).
template <int TM, int TN, int THREADS, bool JOBS>
__global__ void __launch_bounds__(THREADS, 1)
pupko2_kernel(const Pupko2Params p)
{
    constexpr int BM = 16 * TM, BN = 16 * TN, BK = PRUNE_BK, STAGES = PRUNE_STAGES;
    constexpr int NTN = THREADS / 16, CT = BN / NTN;
    constexpr int BNP = BN + 2;
    constexpr int TMP = pupko_tmp(TM);
    constexpr int AST = 16 * TMP;
    constexpr int LST = BM + 1;                  // column stride of the L staging tile (odd: column-wise writes spread over the banks)
    static_assert(CT * NTN == BN && CT >= 1, "column tile does not divide");
    using ctype = uint16_t;
    extern __shared__ __align__(16) double smem[];
    const int kpad = (p.S + BK - 1) / BK * BK;
    double* Ms = smem;                                   // [kpad][BNP]  prod_children L_child[j]; after the scan: L staging [BN][LST]
    double* As = smem + (size_t)BM * BNP;                // [STAGES][BK][16][TMP] matrix chunks; after the scan: argmax staging [BN][BM]

    const int tid = threadIdx.x;
    const int tn = tid % NTN, tm = tid / NTN;
    const int n_tiles = JOBS ? p.n_job_tiles : p.K * p.n_col_tiles;
    double* const my_slots = p.scratch + (size_t)blockIdx.x * p.n_fslots * p.slot_stride;
    const int n_chunks = kpad / BK;
    ctype* const ctab = reinterpret_cast<ctype*>(p.ctab);
    int job = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int k;
        int64_t col0, U, U_stride;
        const int32_t* __restrict__ ids;
        if (JOBS) {
            while (job + 1 < p.n_jobs && tile >= p.jobs[(job + 1) * JOB_WORDS]) ++job;
            const int32_t* jw = p.jobs + job * JOB_WORDS;
            const int local = tile - jw[0];
            k = local / jw[1];
            col0 = (int64_t)(local % jw[1]) * BN;
            U = jw[2];
            U_stride = jw[3];
            ids = p.ids + (((int64_t)jw[5] << 32) | (uint32_t)jw[4]);
        } else {
            k = tile / p.n_col_tiles;
            col0 = (int64_t)(tile % p.n_col_tiles) * BN;
            U = p.U;
            U_stride = p.U_stride;
            ids = p.ids;
        }
        const int32_t* mat_of = p.mat_of + (size_t)k * p.n_nodes;
        double best[TM][CT];

        for (int st = JOBS ? job : 0; st < (JOBS ? job + 1 : p.n_steps); ++st) {
            const Step sp = p.steps[st];
            // ---- M[j][col] = prod over children of L_child[j][col], in the schedule's child order (:99-109); the carried child's
            // values are still in this thread's registers and go in last, in place (two-child products only: commutative) ----
            __syncthreads();
            bool first = true;
            for (int ci = 0; ci < sp.n_children; ++ci) {
                const StepChild ch = p.children[sp.child_begin + ci];
                if (ch.kind == 1) continue;
                if (ch.kind == 0 || ch.kind == 3) {
                    // reconstruct_leaf_node (:30-46): L[j] = P(j -> obs) = PT[obs][j]; a table node: row `id` of its L table.
                    // One warp per column, j coalesced.
                    const double* __restrict__ PT = ch.kind == 3 ? p.tables + ((size_t)ch.slot * p.K + (size_t)k * ch.f_slot) * p.LD
                                                                 : p.arena + (size_t)mat_of[ch.node] * p.LD * p.LD;
                    for (int c = tid >> 5; c < BN; c += THREADS / 32) {
                        int64_t u = col0 + c;
                        if (u >= U) u = U - 1;
                        const int obs = ids[(size_t)ch.leaf_row * U_stride + u];
                        const double* __restrict__ r = PT + (size_t)obs * p.LD;
                        for (int j = tid & 31; j < kpad; j += 32) {
                            const double l = j < p.S ? __ldg(r + j) : 0.0;
                            Ms[(size_t)j * BNP + c] = first ? l : __dmul_rn(Ms[(size_t)j * BNP + c], l);
                        }
                    }
                } else {
                    const double* __restrict__ src = my_slots + (size_t)ch.f_slot * p.slot_stride;
                    for (int idx = tid; idx < kpad * BN; idx += THREADS) {
                        const int j = idx / BN, c = idx % BN;
                        const double l = j < p.S ? src[idx] : 0.0;
                        Ms[(size_t)j * BNP + c] = first ? l : __dmul_rn(Ms[(size_t)j * BNP + c], l);
                    }
                }
                first = false;
                __syncthreads();
            }
            if (sp.carry_in) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < CT; ++j) {
                        const int row = i * 16 + tm;
                        double* m = Ms + (size_t)row * BNP + col_of<CT>(tn, j);
                        const double l = row < p.S ? best[i][j] : 0.0;
                        if (row < kpad) *m = first ? l : __dmul_rn(*m, l);
                    }
                __syncthreads();
            }

            if (sp.is_root) {
                // reconstruct_root_node (:48-76): argmax_{1 <= j < root_len} M[j] * prior(j), strict '>' from -1
                if (tid < BN) {
                    double bst = -1.0;
                    int arg = 0;
                    for (int j = 1; j < p.root_len; ++j) {
                        const double val = __dmul_rn(Ms[(size_t)j * BNP + tid], p.prior_d[j]);
                        if (val > bst) { bst = val; arg = j; }
                    }
                    const int64_t u = col0 + tid;
                    if (u < U) p.root_state[(size_t)k * p.U_stride + u] = arg;
                }
                continue;
            }

            // reconstruct_internal_node (:78-114): L[i] = max_j M[j] * P(i -> j), C[i] = the first maximiser (ascending j, strict '>')
            const double* __restrict__ PT = p.arena + (size_t)mat_of[sp.node] * p.LD * p.LD;
            int arg[TM][CT];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j) { best[i][j] = -1.0; arg[i][j] = 0; }
            auto load_chunk = [&](int chunk) {
                if (chunk < n_chunks) {
                    double* dst = As + (size_t)(chunk % STAGES) * BK * AST;
                    const double* __restrict__ g = PT + (size_t)chunk * BK * p.LD;
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
                    const int jj = chunk * BK + kk;
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < CT; ++j) {
                            const double val = __dmul_rn(b[j], a[i]);   // value * matrix->get(i, j)
                            const bool gt = val > best[i][j];
                            best[i][j] = gt ? val : best[i][j];
                            arg[i][j] = gt ? jj : arg[i][j];
                        }
                }
            }
            cp_async_wait<0>();
            __syncthreads();                 // every warp is done with Ms and the stages: both become staging tiles

            // ---- C_v: staged column-major in shared memory, written as contiguous rows of the node's argmax table ----
            ctype* Cs = reinterpret_cast<ctype*>(As);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < CT; ++j) Cs[(size_t)col_of<CT>(tn, j) * BM + i * 16 + tm] = (ctype)arg[i][j];
            const bool to_table = JOBS && sp.dst_kind == 3;
            if (to_table) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < CT; ++j) Ms[(size_t)col_of<CT>(tn, j) * LST + i * 16 + tm] = best[i][j];
            }
            __syncthreads();
            {
                const int64_t cols = p.c_cols[sp.node];
                ctype* __restrict__ cdst = ctab + p.c_off[sp.node] + (size_t)k * cols * p.SP;
                for (int c = tid >> 5; c < BN; c += THREADS / 32) {
                    const int64_t col = col0 + c;
                    if (col >= U) continue;
                    for (int i = tid & 31; i < p.S; i += 32) cdst[(size_t)col * p.SP + i] = Cs[(size_t)c * BM + i];
                }
            }
            // ---- L_v: table rows (table job), a global slot, or it simply stays in `best` for the parent (next step) ----
            if (to_table) {
                double* __restrict__ tb = p.tables + ((size_t)sp.f_slot * p.K + (size_t)k * U) * p.LD;
                for (int c = tid >> 5; c < BN; c += THREADS / 32) {
                    const int64_t col = col0 + c;
                    if (col >= U) continue;
                    for (int i = tid & 31; i < BM; i += 32) tb[(size_t)col * p.LD + i] = Ms[(size_t)c * LST + i];
                }
            } else if (sp.dst_kind == 1) {
                double* const out_slot = my_slots + (size_t)sp.f_slot * p.slot_stride;
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < CT; ++j) out_slot[(size_t)(i * 16 + tm) * BN + col_of<CT>(tn, j)] = best[i][j];
            }
        }
        __syncthreads();
    }
}

// Traceback (:173-188): one thread per (unique family, category) walks the nodes from the root down (a parent's index is larger than
// its children's) and reads one argmax entry per node.  tab_of[v] >= 0: v is a table node and its column is the family's pattern id
// tab_ids[tab_of[v]][u]; otherwise its column is u.
#ifdef CAFE_PUPKO2_LAUNCH_IMPL
__global__ void __launch_bounds__(256)
pupko_traceback_kernel(const void* __restrict__ ctab_v, const int64_t* __restrict__ c_off, const int64_t* __restrict__ c_cols,
                       const int32_t* __restrict__ parent, const int32_t* __restrict__ leaf_col, const int32_t* __restrict__ tab_of,
                       const int32_t* __restrict__ tab_ids, const int32_t* __restrict__ root_state, int64_t U, int64_t U_stride,
                       int n_nodes, int K, int SP, int32_t* __restrict__ states)
{
    const uint16_t* __restrict__ ctab = reinterpret_cast<const uint16_t*>(ctab_v);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= U * K) return;
    const int k = (int)(idx / U);
    const int64_t u = idx % U;
    int32_t* __restrict__ out = states + ((size_t)k * U_stride + u) * n_nodes;
    const int root = n_nodes - 1;
    out[root] = root_state[(size_t)k * U_stride + u];
    for (int v = root - 1; v >= 0; --v) {
        if (leaf_col[v] >= 0) continue;
        const int sp = out[parent[v]];
        const int64_t col = tab_of[v] >= 0 ? tab_ids[(size_t)tab_of[v] * U_stride + u] : u;
        out[v] = (int32_t)ctab[c_off[v] + ((size_t)k * c_cols[v] + col) * SP + sp];
    }
}

#endif  // CAFE_PUPKO2_LAUNCH_IMPL

inline size_t pupko2_smem_bytes(int TM, int TN)
{
    const int tmp = pupko_tmp(TM);
    return sizeof(double) * ((size_t)16 * TM * (16 * TN + 2) + (size_t)PRUNE_STAGES * PRUNE_BK * 16 * tmp);
}

}  // namespace cafe
