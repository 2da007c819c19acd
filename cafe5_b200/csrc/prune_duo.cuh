// prune_duo.cuh -- resident-vector pruning kernel with TWO column halves per CTA in strict alternation.
//
// Same algorithm as prune_resident.cuh (factors carried in DMMA accumulators, child vector resident in shared memory,
// matrix rows streamed by cp.async.bulk + mbarrier), re-organised around what the profile of that kernel showed
// (profiles/r01_prune_resident_ncu.txt): the tensor pipe was busy 69 % of the time because every non-contraction phase of a
// step -- leaf gathers from L2, Vres stores, factor parking, the root epilogue, ~30 k cycles of mostly latency per step
// whatever the tile width -- ran with the pipe idle, and two co-resident CTAs simply fell into lockstep.
//
// Here one 256-thread CTA owns TWO tiles of BNh = 16*TNW unique families ("halves", 4 warps each as 2 x 2, warp tile
// (8*TMW) x (8*TNW)).  The halves walk the same schedule but take the contraction phase in STRICT ALTERNATION
// (A0 B0 A1 B1 ...), handing over through a pair of mbarriers: while half A streams its matrix through the tensor
// pipe, half B gathers / multiplies / stores its next vector, and vice versa.  Because the order of contractions is
// therefore a fixed sequence, the matrix-chunk stream is ONE ring shared by both halves: whichever half is contracting
// refills it NS-1 chunks ahead, straight across the hand-over (the next half's matrix is known from the schedule), so
// the ring has the full leftover shared memory (3 stages of 8 rows at S = 171) instead of half of it per CTA.
#pragma once
#include "prune_dmma.cuh"

namespace cafe {

constexpr int DUO_MAX_STAGES = 8;

// HWN = column warps per half: 2 -> 256-thread CTA (1 warp per SM sub-partition while contracting, <= 255 registers),
//                              4 -> 512-thread CTA (2 warps per sub-partition while contracting, <= 128 registers).
template <int TMW, int TNW, int HWN, int BK>
struct DuoCfg {
    static constexpr int DUO_HALF = 64 * HWN;     // threads per half
    static constexpr int THREADS = 2 * DUO_HALF;
    static constexpr int BM = 16 * TMW;
    static constexpr int BNH = 8 * TNW * HWN;     // columns per half
    static constexpr int BMP = BM + 4;
    static constexpr int BNP = BNH + 4;
    static constexpr int STAGE_DOUBLES = BK * BMP;
    static constexpr int RED = 2 * 2 * BNH;       // per half: [2][2*BNH] epilogue reduction (two threads per column)
    static constexpr int TAIL_DOUBLES = 2 * RED + 2 * DUO_MAX_STAGES + 2;   // red of both halves, full/empty barriers, turn barriers
    static int vrows(int N) { return (N + 7) / 8 * 8; }
    static size_t fixed_bytes(int N) { return sizeof(double) * ((size_t)2 * vrows(N) * BNP + TAIL_DOUBLES); }
    static size_t smem_bytes(int N, int n_stages) { return fixed_bytes(N) + sizeof(double) * (size_t)n_stages * STAGE_DOUBLES; }
};

template <int NT>
__device__ __forceinline__ void half_sync(int half)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(half + 1), "n"(NT) : "memory");
}

template <int TMW, int TNW, int HWN, int BK>
__global__ void __launch_bounds__(128 * HWN, 1)
prune_duo_kernel(const PruneParams p, const int NS)
{
    using Cfg = DuoCfg<TMW, TNW, HWN, BK>;
    constexpr int BM = Cfg::BM, BN = Cfg::BNH, BMP = Cfg::BMP, BNP = Cfg::BNP, DUO_HALF = Cfg::DUO_HALF;
    extern __shared__ __align__(128) double smem_duo[];
    const int kpad = (p.S + BK - 1) / BK * BK;
    const int n_chunks = kpad / BK;
    const int vrows = (p.N + 7) / 8 * 8;

    const int tid = threadIdx.x;
    const int half = tid / DUO_HALF, htid = tid % DUO_HALF;
    const int lane = tid & 31, hw = htid >> 5;
    const int wm = hw / HWN, wn = hw % HWN;           // 2 x HWN warps per half
    const int g = lane >> 2, q = lane & 3;            // DMMA fragment coordinates
    const int row_base = wm * 8 * TMW + g;            // + i*8
    const int col_base = wn * 8 * TNW + 2 * q;        // + j*8 + e

    double* const Vres = smem_duo + (size_t)half * vrows * BNP;               // [vrows][BNP] of this half
    double* const stages = smem_duo + (size_t)2 * vrows * BNP;               // [NS][BK][BMP] shared ring
    double* const red = stages + (size_t)NS * Cfg::STAGE_DOUBLES + (size_t)half * Cfg::RED;   // [2][2*BN] of this half
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(stages + (size_t)NS * Cfg::STAGE_DOUBLES + 2 * Cfg::RED);
    uint64_t* const empty_bar = full_bar + DUO_MAX_STAGES;
    uint64_t* const turn_bar = empty_bar + DUO_MAX_STAGES;                    // [2]: turn_bar[h] completes when the OTHER half finished a contraction

    const int n_tiles = p.K * p.n_col_tiles;
    const int n_pairs = (n_tiles + 1) / 2;
    const int my_pairs = blockIdx.x < n_pairs ? (n_pairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    double* const my_slots = p.scratch + ((size_t)blockIdx.x * 2 + half) * p.n_fslots * p.slot_stride;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, DUO_HALF / 32); }
        mbar_init(turn_bar + 0, DUO_HALF / 32);
        mbar_init(turn_bar + 1, DUO_HALF / 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---- the chunk stream: slot s = (pair_i * n_gemm + gemm) * 2 + half, n_chunks chunks per slot.  Each half's warp 0 keeps a
    // producer cursor (no divisions on the critical path) and skips the n_chunks positions the other half produces in between.
    const int total_slots = my_pairs * p.n_gemm * 2;
    const int adv_stage = n_chunks % NS, adv_round = n_chunks / NS;   // effect of n_chunks positions on (stage, round parity)
    int p_slot = 0, p_h = 0, p_pair = 0, p_gi = 0, p_chunk = 0, p_stage = 0;
    unsigned p_phase = 1;                  // parity to wait for on the empty barrier (first round: free)
    const double* p_PT = nullptr;          // matrix of the cursor's slot (looked up once per slot, off the per-chunk path)
    auto p_lookup = [&]() {
        if (p_slot < total_slots) {
            int tile = 2 * (blockIdx.x + p_pair * (int)gridDim.x) + p_h;
            if (tile >= n_tiles) tile = n_tiles - 1;                        // odd tile count: the last B half replays the last tile
            const int kcat = tile / p.n_col_tiles;
            p_PT = p.arena + (size_t)p.mat_of[(size_t)kcat * p.n_nodes + p.gemm_nodes[p_gi]] * p.LD * p.LD;
        }
    };
    auto p_next_slot = [&]() {
        ++p_slot;
        if (p_h == 0) p_h = 1;
        else { p_h = 0; if (++p_gi == p.n_gemm) { p_gi = 0; ++p_pair; } }
        p_lookup();
    };
    auto produce_one = [&]() {
        if (p_slot < total_slots) {
            mbar_wait(empty_bar + p_stage, p_phase);
            if (p.LD == BMP) {                   // arena stride == smem stride: the stage is one contiguous copy
                if (lane == 0) {
                    mbar_expect_tx(full_bar + p_stage, (unsigned)(BK * BMP * sizeof(double)));
                    bulk_g2s(stages + (size_t)p_stage * Cfg::STAGE_DOUBLES, p_PT + (size_t)(p_chunk * BK) * p.LD,
                             BK * BMP * sizeof(double), full_bar + p_stage);
                }
            } else if (lane < BK) {
                if (lane == 0) mbar_expect_tx(full_bar + p_stage, (unsigned)(BK * BM * sizeof(double)));
                __syncwarp((1u << BK) - 1u);
                bulk_g2s(stages + (size_t)p_stage * Cfg::STAGE_DOUBLES + lane * BMP,
                         p_PT + (size_t)(p_chunk * BK + lane) * p.LD, BM * sizeof(double), full_bar + p_stage);
            }
            __syncwarp();
        }
        if (++p_stage == NS) { p_stage = 0; p_phase ^= 1u; }
        if (++p_chunk == n_chunks) { p_chunk = 0; p_next_slot(); }
    };
    auto p_skip_slot = [&]() {             // the other half produced one whole slot
        p_stage += adv_stage;
        unsigned r = (unsigned)adv_round;
        if (p_stage >= NS) { p_stage -= NS; ++r; }
        p_phase ^= (r & 1u);
        p_next_slot();
    };
    if (hw == 0) {
        p_lookup();
        if (half == 0) {
            for (int s = 0; s < NS - 1; ++s) produce_one();       // prefill; cursor now NS-1 ahead of A's first chunk
        } else {
            // B's cursor starts NS-1 ahead of ITS first chunk = position n_chunks + NS - 1
            for (int s = 0; s < NS - 1; ++s) {
                if (++p_stage == NS) { p_stage = 0; p_phase ^= 1u; }
                if (++p_chunk == n_chunks) { p_chunk = 0; p_next_slot(); }
            }
            p_skip_slot();
        }
    }
    // consumer cursor of this half: first chunk at position half * n_chunks
    int c_stage = 0;
    unsigned c_phase = 0;
    auto c_skip_slot = [&]() {
        c_stage += adv_stage;
        unsigned r = (unsigned)adv_round;
        if (c_stage >= NS) { c_stage -= NS; ++r; }
        c_phase ^= (r & 1u);
    };
    if (half == 1) c_skip_slot();
    unsigned m_idx = 0;                    // contractions this half has done

    for (int pair_i = 0; pair_i < my_pairs; ++pair_i) {
        int tile = 2 * (blockIdx.x + pair_i * (int)gridDim.x) + half;
        const bool live_tile = tile < n_tiles;
        if (!live_tile) tile = n_tiles - 1;
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
            half_sync<DUO_HALF>(half);          // every warp of this half finished reading Vres in the previous contraction / epilogue
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
            half_sync<DUO_HALF>(half);

            if (!sp.is_root) {
                // ---- 3. W_v = P_v . V_v on the FP64 tensor cores (matrix_cache.cpp:49-56), when it is this half's turn ----
#pragma unroll
                for (int i = 0; i < TMW; ++i)
#pragma unroll
                    for (int j = 0; j < TNW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
                if (half == 0) { if (m_idx > 0) mbar_wait(turn_bar + 0, (m_idx - 1u) & 1u); }
                else mbar_wait(turn_bar + 1, m_idx & 1u);
                for (int chunk = 0; chunk < n_chunks; ++chunk) {
                    if (hw == 0) produce_one();              // refill the stage the previous chunk released, NS-1 ahead of the consumer
                    mbar_wait(full_bar + c_stage, c_phase);
                    const double* As = stages + (size_t)c_stage * Cfg::STAGE_DOUBLES;
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
                    if (lane == 0) mbar_arrive(empty_bar + c_stage);
                    if (++c_stage == NS) { c_stage = 0; c_phase ^= 1u; }
                }
                if (lane == 0) mbar_arrive(turn_bar + (1 - half));   // hand the tensor pipe to the other half
                ++m_idx;
                c_skip_slot();                               // the other half's chunks
                if (hw == 0) p_skip_slot();
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
                constexpr int PARTS = 2;                     // two threads per column; the rest of the half only joins the barrier
                const int c = htid % BN, part = htid / BN;
                const bool active = part < PARTS;
                const int64_t u = col0 + c;
                const bool write = live_tile && u < p.U;
                const double* root = Vres + c;
                double best = 0.0;
                int any = 0;
                if (!active) {
                } else if (p.mode == MODE_BASE) {
                    best = -INFINITY;           // max_j log L_j + log prior_j   (base_model.cpp:82-91)
                    for (int j = part; j < p.R; j += PARTS) {
                        const double v = __dadd_rn(log(root[(size_t)(j + 1) * BNP]), p.logprior[j]);
                        if (v > best) best = v;
                    }
                } else if (p.mode == MODE_GAMMA) {
                    bool first = true;          // max_j L_j * prior_j ; failure iff sum_j L_j == 0   (gamma_core.cpp:151-160)
                    for (int j = part; j < p.R; j += PARTS) {
                        const double L = root[(size_t)(j + 1) * BNP];
                        any |= (L != 0.0);
                        const double v = __dmul_rn(L, p.prior_d[j]);
                        if (first || v > best) { best = v; first = false; }
                    }
                } else {
                    if (write && k == 0)
                        for (int j = part; j < p.R; j += PARTS) p.out_roots[(size_t)u * p.R + j] = root[(size_t)(j + 1) * BNP];
                }
                if (active) {
                    red[part * BN + c] = best;
                    red[PARTS * BN + part * BN + c] = (double)any;
                }
                half_sync<DUO_HALF>(half);
                if (part == 0 && write && p.mode != MODE_ROOTS) {
                    double bb = red[c];
                    const double b1 = red[BN + c];
                    if (b1 > bb) bb = b1;
                    const int aa = (red[PARTS * BN + c] != 0.0) | (red[PARTS * BN + BN + c] != 0.0);
                    p.out_best[(size_t)k * p.U_stride + u] = bb;
                    if (p.mode == MODE_GAMMA) p.out_ok[(size_t)k * p.U_stride + u] = (uint8_t)aa;
                }
            }
        }
        half_sync<DUO_HALF>(half);   // factor slots of this tile are dead; Vres / red are reused by the next tile
    }
}

}  // namespace cafe
