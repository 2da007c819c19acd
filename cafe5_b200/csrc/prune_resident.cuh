// prune_resident.cuh -- pruning kernel with the child vector RESIDENT in shared memory (the default path).
//
// Felsenstein pruning rewritten around FACTORS (reference src/probability.cpp:206-233: a node's vector is the
// elementwise product over children c of result_c = P_c . v_c).  A persistent CTA owns BN unique families of one
// gamma category and walks the post-order schedule.  For every non-root internal node v:
//   1. V_v = product of its children's factors, formed in the DMMA accumulator registers:
//        leaf child      -> row gather of the transposed matrix (or the 3-term error-model combination),
//        chain child     -> its factor is ALREADY in the registers (it was the previous step's contraction output),
//        any other child -> its factor is read back from a global "factor slot".
//   2. V_v is written to a shared-memory tile Vres[S][BN] (never to global memory).
//   3. W_v = P_v (BM x S) . Vres, FP64 tensor cores (mma.sync.m8n8k4.f64): B fragments come straight from Vres, the
//      matrix streams through a 3-stage cp.async.bulk + mbarrier pipeline that runs AHEAD ACROSS contractions
//      (the next node's matrix is known from the schedule), so the tensor pipe never waits for a pipeline fill.
//   4. W_v stays in registers when the parent is the next step (post-order makes that the common case), else it
//      is stored once to a global factor slot.
// The root multiplies its children's factors the same way and runs the fused epilogue (prior weighting, max over
// root sizes, failure flag) out of shared memory.  Global traffic per tile: the count columns, the matrices (L2
// resident, shared by all CTAs working on the same category) and one store + one load of a factor for each
// first-visited child of a node with two internal children.
//
// A two-child product is commutative bit for bit, so the order "chain factor first" changes nothing; nodes with
// three or more children keep the reference's descendant order and take every internal child's factor from global
// slots.  Requires the whole state space in one row pass (N <= 208) and Vres + stages within 227 KB; otherwise the
// host falls back to prune_dmma_kernel (prune_dmma.cuh), which streams both operands.
#pragma once
#include "prune_dmma.cuh"

namespace cafe {

constexpr int RS_BK = 8;
constexpr int RS_STAGES = 3;

template <int TMW, int TNW>
struct ResidentCfg {
    static constexpr int BM = 16 * TMW;
    static constexpr int BN = 32 * TNW;
    static constexpr int BMP = BM + 4;
    static constexpr int BNP = BN + 4;
    static constexpr int STAGE_DOUBLES = RS_BK * BMP;
    static constexpr int TAIL_DOUBLES = 2 * PRUNE_THREADS + 2 * RS_STAGES + 2;
    static size_t smem_bytes(int N)   // Vres holds child states [0, S) for the contraction and root rows [1, R] for the epilogue
    {
        const int vrows = (N + RS_BK - 1) / RS_BK * RS_BK;
        return sizeof(double) * ((size_t)vrows * BNP + (size_t)RS_STAGES * STAGE_DOUBLES + TAIL_DOUBLES);
    }
};

template <int TMW, int TNW>
__global__ void __launch_bounds__(PRUNE_THREADS, 1)
prune_resident_kernel(const PruneParams p)
{
    using Cfg = ResidentCfg<TMW, TNW>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BMP = Cfg::BMP, BNP = Cfg::BNP, BK = RS_BK, NS = RS_STAGES;
    extern __shared__ __align__(128) double smem_rs[];
    const int kpad = (p.S + BK - 1) / BK * BK;
    const int n_chunks = kpad / BK;
    const int vrows = (p.N + BK - 1) / BK * BK;
    double* const Vres = smem_rs;                                     // [vrows][BNP]
    double* const stages = Vres + (size_t)vrows * BNP;                // [NS][BK][BMP]
    double* const red = stages + (size_t)NS * Cfg::STAGE_DOUBLES;     // [2][PRUNE_THREADS]
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(red + 2 * PRUNE_THREADS);
    uint64_t* const empty_bar = full_bar + NS;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
    const int g = lane >> 2, q = lane & 3;            // DMMA fragment coordinates
    const int row_base = wm * 8 * TMW + g;            // + i*8
    const int col_base = wn * 8 * TNW + 2 * q;        // + j*8 + e
    const int n_tiles = p.K * p.n_col_tiles;
    double* const my_slots = p.scratch + (size_t)blockIdx.x * p.n_fslots * p.slot_stride;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, PRUNE_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---- producer state (warp 0): one continuous stream of matrix chunks over (tile, contraction, chunk) ----
    unsigned gp = 0;
    int p_tile = blockIdx.x, p_g = 0, p_chunk = 0;
    auto produce_one = [&]() {
        if (p.n_gemm == 0 || p_tile >= n_tiles) return;
        const unsigned stage = gp % NS;
        mbar_wait(empty_bar + stage, ((gp / NS) & 1u) ^ 1u);
        if (lane < BK) {
            const int kcat = p_tile / p.n_col_tiles;
            const int node = p.gemm_nodes[p_g];
            const double* PT = p.arena + (size_t)p.mat_of[(size_t)kcat * p.n_nodes + node] * p.LD * p.LD;
            if (lane == 0) mbar_expect_tx(full_bar + stage, (unsigned)(BK * BM * sizeof(double)));
            __syncwarp(0x000000ffu);
            bulk_g2s(stages + (size_t)stage * Cfg::STAGE_DOUBLES + lane * BMP,
                     PT + (size_t)(p_chunk * BK + lane) * p.LD, BM * sizeof(double), full_bar + stage);
        }
        __syncwarp();
        ++gp;
        if (++p_chunk == n_chunks) {
            p_chunk = 0;
            if (++p_g == p.n_gemm) { p_g = 0; p_tile += gridDim.x; }
        }
    };
    if (warp == 0)
        for (int s = 0; s < NS - 1; ++s) produce_one();
    unsigned gc = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int k = tile / p.n_col_tiles;
        const int64_t col0 = (int64_t)(tile % p.n_col_tiles) * BN;
        const int32_t* mat_of = p.mat_of + (size_t)k * p.n_nodes;
        double acc[TMW][TNW][2];

        for (int st = 0; st < p.n_steps; ++st) {
            const Step sp = p.steps[st];
            bool has_acc = sp.carry_in != 0;
            // ---- 1. V_v = product of the children's factors (probability.cpp:215-217, 229-231) ----
            for (int ci = 0; ci < sp.n_children; ++ci) {
                const StepChild ch = p.children[sp.child_begin + ci];
                if (ch.kind == 1) continue;                       // carried: already in acc
                if (ch.kind == 0) {
                    // leaf: factor[s] = sum_d em[obs][d] * P(s -> obs-1+d)   (probability.cpp:187-202)
                    const double* __restrict__ PT = p.arena + (size_t)mat_of[ch.node] * p.LD * p.LD;
#pragma unroll
                    for (int j = 0; j < TNW; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            int64_t u = col0 + col_base + j * 8 + e;
                            if (u >= p.U) u = p.U - 1;            // padding columns replay the last family; never written out
                            const int obs = p.counts_t[(size_t)ch.leaf_row * p.U_stride + u];
                            if (p.em == nullptr) {
                                const double* __restrict__ r = PT + (size_t)obs * p.LD + row_base;
#pragma unroll
                                for (int i = 0; i < TMW; ++i) {
                                    const double v = __ldg(r + i * 8);
                                    acc[i][j][e] = has_acc ? __dmul_rn(acc[i][j][e], v) : v;
                                }
                            } else {
                                const int er = obs < p.em_rows ? obs : p.em_rows - 1;
                                double pe[3];
                                const double* r[3];
#pragma unroll
                                for (int d = 0; d < 3; ++d) {
                                    const int idx = obs - 1 + d;
                                    const bool ok = idx >= 0 && idx < p.S;
                                    pe[d] = ok ? __ldg(p.em + er * 3 + d) : 0.0;
                                    r[d] = PT + (size_t)(ok ? idx : obs) * p.LD + row_base;
                                }
#pragma unroll
                                for (int i = 0; i < TMW; ++i) {
                                    double f = __dmul_rn(__ldg(r[0] + i * 8), pe[0]);       // c ascending, separately rounded
                                    f = __dadd_rn(f, __dmul_rn(__ldg(r[1] + i * 8), pe[1]));
                                    f = __dadd_rn(f, __dmul_rn(__ldg(r[2] + i * 8), pe[2]));
                                    acc[i][j][e] = has_acc ? __dmul_rn(acc[i][j][e], f) : f;
                                }
                            }
                        }
                } else {
                    // factor of an earlier sibling subtree, parked in a global slot
                    const double* __restrict__ fs = my_slots + (size_t)ch.f_slot * p.slot_stride;
#pragma unroll
                    for (int i = 0; i < TMW; ++i)
#pragma unroll
                        for (int j = 0; j < TNW; ++j) {
                            const double2 f = *reinterpret_cast<const double2*>(fs + (size_t)(row_base + i * 8) * BN + col_base + j * 8);
                            acc[i][j][0] = has_acc ? __dmul_rn(acc[i][j][0], f.x) : f.x;
                            acc[i][j][1] = has_acc ? __dmul_rn(acc[i][j][1], f.y) : f.y;
                        }
                }
                has_acc = true;
            }

            // ---- 2. V_v -> shared memory (states >= S do not exist: zero rows, the matrix rows there may be real) ----
            __syncthreads();          // every warp finished reading Vres in the previous contraction / epilogue
#pragma unroll
            for (int i = 0; i < TMW; ++i) {
                const int row = row_base + i * 8;
                if (row < vrows) {
                    const bool live = sp.is_root || row < p.S;
#pragma unroll
                    for (int j = 0; j < TNW; ++j)
                        *reinterpret_cast<double2*>(Vres + (size_t)row * BNP + col_base + j * 8) =
                            live ? make_double2(acc[i][j][0], acc[i][j][1]) : make_double2(0.0, 0.0);
                }
            }
            __syncthreads();

            if (!sp.is_root) {
                // ---- 3. W_v = P_v . V_v on the FP64 tensor cores (matrix_cache.cpp:49-56) ----
#pragma unroll
                for (int i = 0; i < TMW; ++i)
#pragma unroll
                    for (int j = 0; j < TNW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
                for (int chunk = 0; chunk < n_chunks; ++chunk) {
                    const unsigned stage = gc % NS;
                    mbar_wait(full_bar + stage, (gc / NS) & 1u);
                    const double* As = stages + (size_t)stage * Cfg::STAGE_DOUBLES;
                    const double* Bs = Vres + (size_t)chunk * BK * BNP;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 4; ++k4) {
                        double a[TMW], b[TNW];
                        const double* ap = As + (k4 * 4 + q) * BMP + row_base;
                        const double* bp = Bs + (k4 * 4 + q) * BNP + wn * 8 * TNW + g;
#pragma unroll
                        for (int i = 0; i < TMW; ++i) a[i] = ap[i * 8];
#pragma unroll
                        for (int j = 0; j < TNW; ++j) b[j] = bp[j * 8];
#pragma unroll
                        for (int i = 0; i < TMW; ++i)
#pragma unroll
                            for (int j = 0; j < TNW; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar + stage);
                    ++gc;
                    if (warp == 0) produce_one();
                }
                // ---- 4. the factor stays in registers for the parent, or is parked once in a global slot ----
                if (sp.dst_kind == 1) {
                    double* fs = my_slots + (size_t)sp.f_slot * p.slot_stride;
#pragma unroll
                    for (int i = 0; i < TMW; ++i)
#pragma unroll
                        for (int j = 0; j < TNW; ++j)
                            *reinterpret_cast<double2*>(fs + (size_t)(row_base + i * 8) * BN + col_base + j * 8) =
                                make_double2(acc[i][j][0], acc[i][j][1]);
                }
            } else {
                // ---- root epilogue from shared memory: index j <-> root size j+1 (core.cpp:141), weighted by prior(j)
                constexpr int PARTS = PRUNE_THREADS / BN;
                const int c = tid % BN, part = tid / BN;
                const int64_t u = col0 + c;
                const double* root = Vres + c;
                double best;
                int any = 0;
                if (p.mode == MODE_BASE) {
                    best = -INFINITY;           // max_j log L_j + log prior_j   (base_model.cpp:82-91)
                    for (int j = part; j < p.R; j += PARTS) {
                        const double v = __dadd_rn(log(root[(size_t)(j + 1) * BNP]), p.logprior[j]);
                        if (v > best) best = v;
                    }
                } else if (p.mode == MODE_GAMMA) {
                    best = 0.0;                 // max_j L_j * prior_j ; failure iff sum_j L_j == 0   (gamma_core.cpp:151-160)
                    bool first = true;
                    for (int j = part; j < p.R; j += PARTS) {
                        const double L = root[(size_t)(j + 1) * BNP];
                        any |= (L != 0.0);
                        const double v = __dmul_rn(L, p.prior_d[j]);
                        if (first || v > best) { best = v; first = false; }
                    }
                } else {
                    best = 0.0;
                    if (u < p.U && k == 0)
                        for (int j = part; j < p.R; j += PARTS) p.out_roots[(size_t)u * p.R + j] = root[(size_t)(j + 1) * BNP];
                }
                red[part * BN + c] = best;
                red[(PARTS + part) * BN + c] = (double)any;
                __syncthreads();
                if (part == 0 && u < p.U && p.mode != MODE_ROOTS) {
                    double bb = red[c];
                    int aa = red[PARTS * BN + c] != 0.0;
                    for (int qq = 1; qq < PARTS; ++qq) {
                        const double v = red[qq * BN + c];
                        if (v > bb) bb = v;
                        aa |= red[(PARTS + qq) * BN + c] != 0.0;
                    }
                    p.out_best[(size_t)k * p.U_stride + u] = bb;
                    if (p.mode == MODE_GAMMA) p.out_ok[(size_t)k * p.U_stride + u] = (uint8_t)aa;
                }
            }
        }
        __syncthreads();   // factor slots of this tile are dead; Vres / red are reused by the next tile
    }
}

}  // namespace cafe
