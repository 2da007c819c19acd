// prune_dmma.cuh -- pruning kernel, FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) variant.
//
// Same algorithm, schedule, scratch-slot protocol and fused root epilogue as prune_kernel (kernels.cuh);
// what changes is the contraction P(BM x S) . V(S x BN) of every internal branch:
//   * accumulators are 8x8 DMMA tiles (2 doubles per thread per tile) instead of per-thread DFMA tiles, so one
//     A fragment load feeds 8 columns and one B fragment load feeds 8 rows: 4x fewer shared-memory wavefronts
//     per FMA than the DFMA kernel, whose LSU pipe was the co-limiter (profiles/r01_prune_dfma_ncu.txt);
//   * BOTH operands are streamed through one bulk-async-copy pipeline (cp.async.bulk + mbarrier full/empty pairs:
//     A = 16 rows of the transposed matrix, B = 16 rows of the child's vector tile, one 1-1.4 KB row copy per lane of
//     warp 0), so no child vector stays resident in shared memory, the column tile can be 128 wide (half the matrix
//     traffic per FMA), no thread spends issue slots on address arithmetic and there is no CTA-wide barrier inside
//     the contraction.
// 8 warps as 2 (rows) x 4 (columns); warp tile = (8*TMW) x (8*TNW); CTA tile BM = 16*TMW rows x BN = 32*TNW columns.
#pragma once
#include "kernels.cuh"

namespace cafe {

constexpr int DM_BK = 16;

template <int TMW, int TNW>
struct DmmaCfg {
    static constexpr int BM = 16 * TMW;
    static constexpr int BN = 32 * TNW;
    static constexpr int BMP = BM + 4;   // row strides = 32 bytes mod 128: fragment loads are bank-conflict free
    static constexpr int BNP = BN + 4;
    static constexpr int STAGE_DOUBLES = DM_BK * (BMP + BNP);
    static constexpr int MAX_STAGES = 4;
    static constexpr int TAIL_DOUBLES = 2 * PRUNE_THREADS + 2 * MAX_STAGES;   // epilogue reduction + mbarriers
    static size_t smem_bytes(int n_stages) { return sizeof(double) * ((size_t)n_stages * STAGE_DOUBLES + TAIL_DOUBLES); }
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one contiguous row, global -> shared, completion counted in bytes on an mbarrier (async proxy)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Leaf child of a pruning step, in DMMA-fragment layout (shared by prune_dmma_kernel and prune_resident_kernel):
//   factor[s] = sum_d em[obs][d] * P(s -> obs-1+d)   (probability.cpp:187-202; no error model: the single column P(s -> obs))
// multiplied into (or, for the first child, copied to) the accumulators.  rows: PT + first_row + i*8 for the thread's TMW row
// tiles; columns: col0 + col_base + j*8 + e.  Gathers of the TRANSPOSED matrix: row `obs` is contiguous in s.
// ids_row: the observed counts of this leaf for every column (U columns; padding columns replay the last one).  The same gather
// with em == nullptr serves a subtree-pattern table child: PT = the child's factor table, ids_row = its pattern id per column.
template <int TMW, int TNW>
__device__ __forceinline__ void leaf_factor_into(double (&acc)[TMW][TNW][2], bool has_acc, const PruneParams& p,
                                                 const double* __restrict__ PT, const int32_t* __restrict__ ids_row, int64_t U,
                                                 const double* __restrict__ em, int first_row, int64_t col0, int col_base)
{
#pragma unroll
    for (int j = 0; j < TNW; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int64_t u = col0 + col_base + j * 8 + e;
            if (u >= U) u = U - 1;                        // padding columns replay the last family; never written out
            const int obs = ids_row[u];
            if (em == nullptr) {
                const double* __restrict__ r = PT + (size_t)obs * p.LD + first_row;
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
                    pe[d] = ok ? __ldg(em + er * 3 + d) : 0.0;
                    r[d] = PT + (size_t)(ok ? idx : obs) * p.LD + first_row;
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
}

template <int TMW, int TNW>
__global__ void __launch_bounds__(PRUNE_THREADS, 1)
prune_dmma_kernel(const PruneParams p, const int n_stages)
{
    using Cfg = DmmaCfg<TMW, TNW>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BMP = Cfg::BMP, BNP = Cfg::BNP, BK = DM_BK;
    extern __shared__ __align__(128) double smem_dm[];
    double* const smem = smem_dm;
    double* const red = smem + (size_t)n_stages * Cfg::STAGE_DOUBLES;          // [2][PRUNE_THREADS] epilogue reduction
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(red + 2 * PRUNE_THREADS);
    uint64_t* const empty_bar = full_bar + Cfg::MAX_STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, PRUNE_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned gc = 0;        // chunks consumed since kernel start (same in every thread): stage = gc % n_stages
    unsigned gp = 0;        // chunks produced since kernel start (meaningful in warp 0)
    const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
    const int g = lane >> 2, q = lane & 3;            // DMMA fragment coordinates
    const int row_base = wm * 8 * TMW + g;            // + i*8
    const int col_base = wn * 8 * TNW + 2 * q;        // + j*8 + e
    const int n_tiles = p.K * p.n_col_tiles;
    double* const my_scratch = p.scratch + (size_t)blockIdx.x * p.n_slots * p.slot_stride;
    const int kpad = (p.S + BK - 1) / BK * BK;
    const int n_chunks = kpad / BK;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int k = tile / p.n_col_tiles;
        const int64_t col0 = (int64_t)(tile % p.n_col_tiles) * BN;
        const int32_t* mat_of = p.mat_of + (size_t)k * p.n_nodes;

        for (int st = 0; st < p.n_steps; ++st) {
            const Step sp = p.steps[st];
            double* const out_slot = my_scratch + (size_t)sp.out_slot * p.slot_stride;
            for (int mt = 0; mt < p.n_mtiles; ++mt) {
                const int m0 = mt * BM;
                double acc[TMW][TNW][2];
                bool has_acc = false;
                for (int ci = 0; ci < sp.n_children; ++ci) {
                    const StepChild ch = p.children[sp.child_begin + ci];
                    const double* __restrict__ PT = p.arena + (size_t)mat_of[ch.node] * p.LD * p.LD;
                    if (ch.leaf_row >= 0) {
                        leaf_factor_into<TMW, TNW>(acc, has_acc, p, PT, p.counts_t + (size_t)ch.leaf_row * p.U_stride, p.U, p.em, m0 + row_base,
                                                   col0, col_base);
                    } else {
                        // ---- internal child: acc = P[m0.., 0..S) . V_child   (matrix_cache.cpp:49-56)
                        if (has_acc) {   // park the running product in the output slot while the tiles accumulate
#pragma unroll
                            for (int i = 0; i < TMW; ++i)
#pragma unroll
                                for (int j = 0; j < TNW; ++j)
                                    *reinterpret_cast<double2*>(out_slot + (size_t)(m0 + row_base + i * 8) * BN + col_base + j * 8) =
                                        make_double2(acc[i][j][0], acc[i][j][1]);
                        }
                        const double* __restrict__ src = my_scratch + (size_t)ch.slot * p.slot_stride;
                        // producer: lane l < 16 copies A row l of the chunk, lane 16 + l copies B row l
                        auto produce = [&](int chunk) {
                            const unsigned stage = gp % (unsigned)n_stages;
                            mbar_wait(empty_bar + stage, ((gp / (unsigned)n_stages) & 1u) ^ 1u);   // all 8 warps released it
                            double* As = smem + (size_t)stage * Cfg::STAGE_DOUBLES;
                            double* Bs = As + BK * BMP;
                            const int k0 = chunk * BK;
                            if (lane == 0) mbar_expect_tx(full_bar + stage, (unsigned)(BK * (BM + BN) * sizeof(double)));
                            __syncwarp();
                            if (lane < BK) {
                                bulk_g2s(As + lane * BMP, PT + (size_t)(k0 + lane) * p.LD + m0, BM * sizeof(double), full_bar + stage);
                            } else {
                                const int kk = lane - BK;
                                // child states >= S do not exist: feed zeros (matrix rows there are real when N > S)
                                const double* row = (k0 + kk < p.S) ? src + (size_t)(k0 + kk) * BN : p.zero_row;
                                bulk_g2s(Bs + kk * BNP, row, BN * sizeof(double), full_bar + stage);
                            }
                            ++gp;
                        };
                        if (warp == 0) {
                            asm volatile("fence.proxy.async;\n" ::: "memory");   // slot rows were written through the generic proxy
                            for (int s = 0; s < n_stages - 1 && s < n_chunks; ++s) produce(s);
                        }
#pragma unroll
                        for (int i = 0; i < TMW; ++i)
#pragma unroll
                            for (int j = 0; j < TNW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
                        for (int chunk = 0; chunk < n_chunks; ++chunk) {
                            const unsigned stage = gc % (unsigned)n_stages;
                            mbar_wait(full_bar + stage, (gc / (unsigned)n_stages) & 1u);
                            const double* As = smem + (size_t)stage * Cfg::STAGE_DOUBLES;
                            const double* Bs = As + BK * BMP;
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
                            if (lane == 0) mbar_arrive(empty_bar + stage);       // this warp is done reading the stage
                            ++gc;
                            if (warp == 0 && chunk + n_stages - 1 < n_chunks) produce(chunk + n_stages - 1);
                        }
                        if (has_acc) {   // node_probs[i] *= result[i], children in descendant order (probability.cpp:215-217)
#pragma unroll
                            for (int i = 0; i < TMW; ++i)
#pragma unroll
                                for (int j = 0; j < TNW; ++j) {
                                    const double2 prev = *reinterpret_cast<const double2*>(out_slot + (size_t)(m0 + row_base + i * 8) * BN + col_base + j * 8);
                                    acc[i][j][0] = __dmul_rn(prev.x, acc[i][j][0]);
                                    acc[i][j][1] = __dmul_rn(prev.y, acc[i][j][1]);
                                }
                        }
                    }
                    has_acc = true;
                }
#pragma unroll
                for (int i = 0; i < TMW; ++i)
#pragma unroll
                    for (int j = 0; j < TNW; ++j)
                        *reinterpret_cast<double2*>(out_slot + (size_t)(m0 + row_base + i * 8) * BN + col_base + j * 8) =
                            make_double2(acc[i][j][0], acc[i][j][1]);
            }
            if (sp.is_root) {
                // ---- root epilogue: index j <-> root size j+1 (core.cpp:141), weighted by prior(j)
                __syncthreads();
                root_epilogue<BN, PRUNE_THREADS / BN, false>(p, out_slot, BN, red, tid, k, col0);
            }
            __syncthreads();   // the slot just written is read by a later step
        }
    }
}

}  // namespace cafe
