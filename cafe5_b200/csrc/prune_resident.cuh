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
//   3. W_v = P_v (BM x S) . Vres, FP64 tensor cores (mma.sync.m8n8k4.f64): B fragments come straight from Vres (XOR-swizzled,
//      no padding columns), the matrix streams through a cp.async.bulk + mbarrier ring that runs AHEAD ACROSS contractions
//      (the next node's matrix is known from the schedule), so the tensor pipe never waits for a pipeline fill.  The arena's
//      row stride equals the ring's padded row stride, so a stage is ONE contiguous 128-byte-aligned bulk copy, and the warps
//      take turns issuing it (a fixed producer warp was the pace-setter of the whole CTA).
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

// Geometry: 2 x WN warps (rows x columns); warp tile (8*TMW) x (8*TNW); CTA tile BM = 16*TMW rows x BN = 8*TNW*WN columns.
//   WN = 2 (default): 128-thread CTAs with BN = 64 and ~113 KB of shared memory, so TWO CTAs ARE RESIDENT PER SM and one CTA's leaf
//           gathers / epilogue overlap the other's DMMA stream (same registers per thread and accumulators per warp as WN = 4).
//   WN = 4: one 256-thread CTA per SM, BN = 128; used when two CTAs do not fit.
// The resident vector V_v is kept COLUMN-MAJOR, VT[column][state] with a column stride of BMP = BM + 4 doubles:
//   * BMP == 4 (mod 16) doubles, so the DMMA B fragments (8 columns x 4 consecutive states per warp) touch every bank pair exactly
//     twice: two wavefronts for 32 x 8 bytes, the minimum -- no swizzle, no padding columns;
//   * BMP is the row stride of the matrix arena and of the factor tables, so the factor of a gathered child (a leaf's matrix row
//     P(. -> count), or a table row) is ONE contiguous 16-byte-aligned cp.async.bulk per column straight into VT[column]: the gather
//     of the last child of a node no longer goes through registers (88 dependent 8-byte loads per thread in batches the register
//     file allows) but through the copy engine, in flight while the other children's factors are loaded; the running product is
//     then multiplied into VT in place, every thread touching only the elements it owns (no barrier between product and store).
template <int TMW, int TNW, int WN, int BK>
struct ResidentCfg {
    static constexpr int THREADS = 64 * WN;
    static constexpr int BM = 16 * TMW;
    static constexpr int BN = 8 * TNW * WN;
    static constexpr int BMP = BM + 4;
    static constexpr int VT_DOUBLES = BN * BMP;
    static constexpr int STAGE_DOUBLES = BK * BMP;
    static constexpr int MAX_STAGES = 8;
    static constexpr int PARTS = THREADS / BN < 2 ? THREADS / BN : 2;   // epilogue threads per column
    static constexpr int TAIL_DOUBLES = 2 * PARTS * BN + 2 * MAX_STAGES + 2;
    static size_t smem_bytes(int n_stages) { return sizeof(double) * ((size_t)VT_DOUBLES + (size_t)n_stages * STAGE_DOUBLES + TAIL_DOUBLES); }
};

// MINB = CTAs resident per SM (register cap 65536 / (64*WN*MINB)).  Co-resident CTAs drift out of phase on their own (starting
// them half a step apart on purpose changed nothing measurable), so one CTA's leaf gathers / stores overlap the other's DMMA stream.
// VARIANT 2 is a timing build: clock stamps around the phases of the chunk loop (tools/gpu_probe_chunks.py), not used by the product.
// JOBS = true is the TABLE build (subtree-pattern reuse): the launch computes the factor tables W_v = P_v . V_v of up to
// MAX_TABLE_JOBS nodes v, each over ITS OWN distinct patterns of leaf counts (columns = patterns, not families).  Job j is step j of
// the schedule: one step whose children are all gathers (leaves, or the tables of deeper nodes built by an earlier launch), one
// contraction, the result stored as table rows.  The main pass (JOBS = false) then gathers a table node's factor like a leaf column
// (StepChild::kind 3) instead of pruning the subtree again for every family: identical arithmetic per column, fewer columns.
template <int TMW, int TNW, int WN, int BK, int MINB, int VARIANT = 0, bool JOBS = false>
__global__ void __launch_bounds__(64 * WN, MINB)
prune_resident_kernel(const PruneParams p, const int NS, const __grid_constant__ InlineSchedule sched)
{
    constexpr bool PROBE = VARIANT == 2;
    constexpr bool MID = WN == 2;           // ring refill from the middle of the chunk's DMMA stream (see the chunk loop)
    int64_t* probe_out = nullptr;
    int probe_n = 0;
    if (PROBE && p.probe && (blockIdx.x % 148) == 0 && (threadIdx.x & 31) == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        probe_out = p.probe + ((size_t)(blockIdx.x / 148) * (blockDim.x / 32) + threadIdx.x / 32) * (PROBE_CHUNKS * 4 + 4);
        probe_out[0] = smid;
        probe_out += 4;
    }
    using Cfg = ResidentCfg<TMW, TNW, WN, BK>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BMP = Cfg::BMP, THREADS = Cfg::THREADS;
    // Where warp 0 refills the ring: before its own chunk (one more chunk of lead, but it first waits for the slowest warp) or after it.
    // Measured on B200: "after" wins with one warp per sub-partition and CTA (WN = 2), "before" with two.
    constexpr bool PRODUCE_FIRST = WN != 2;
    extern __shared__ __align__(128) double smem_rs[];
    const int kpad = (p.S + BK - 1) / BK * BK;
    const int n_chunks = kpad / BK;
    double* const VT = smem_rs;                                       // [BN][BMP]: V_v, column-major
    double* const stages = VT + Cfg::VT_DOUBLES;                      // [NS][BK][BMP]
    double* const red = stages + (size_t)NS * Cfg::STAGE_DOUBLES;     // [2][PARTS*BN]
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(red + 2 * Cfg::PARTS * BN);
    uint64_t* const empty_bar = full_bar + Cfg::MAX_STAGES;
    uint64_t* const v_full = empty_bar + Cfg::MAX_STAGES;             // the staged child's BN bulk copies have landed in VT
    uint64_t* const v_free = v_full + 1;                              // every warp is done reading VT (contraction / epilogue of the step before)

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;         // 2 x WN warps
    const int g = lane >> 2, q = lane & 3;            // DMMA fragment coordinates
    const int row_base = wm * 8 * TMW + g;            // + i*8
    const int col_base = wn * 8 * TNW + 2 * q;        // + j*8 + e
    const int n_tiles = JOBS ? sched.n_job_tiles : p.K * p.n_col_tiles;
    double* const my_slots = p.scratch + (size_t)blockIdx.x * p.n_fslots * p.slot_stride;
    auto job_word = [&](int j, int w) { return sched.w[sched.off_jobs + j * JOB_WORDS + w]; };
    const int n_gemm = JOBS ? 1 : p.n_gemm;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, THREADS / 32); }
        mbar_init(v_full, BN);
        mbar_init(v_free, THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---- producer: one continuous stream of matrix chunks over (tile, contraction, chunk).  EVERY warp tracks the cursor (it is a
    // deterministic function of the chunk count) and the warps take turns issuing the copy, so the ~150 cycles of empty-wait +
    // expect_tx + UBLKCP per chunk are spread over all sub-partitions instead of making warp 0 the pace-setter; the matrix
    // pointer is looked up once per contraction, off the per-chunk path. ----
    constexpr int NW = THREADS / 32;
    int p_stage = 0;
    unsigned p_phase = 1;                             // parity to wait for on the empty barrier (first pass: free)
    int p_tile = blockIdx.x, p_g = 0, p_chunk = 0, p_turn = 0, p_job = 0;
    const double* p_PT = nullptr;
    auto p_lookup = [&]() {
        if (n_gemm != 0 && p_tile < n_tiles) {
            if (JOBS)
                while (p_job + 1 < sched.n_jobs && p_tile >= job_word(p_job + 1, 0)) ++p_job;
            const int kcat = JOBS ? (p_tile - job_word(p_job, 0)) / job_word(p_job, 1) : p_tile / p.n_col_tiles;
            const int node = JOBS ? sched.w[p_job * 9] : sched.valid ? sched.w[sched.off_gemm + p_g] : p.gemm_nodes[p_g];
            const int mat = sched.valid ? sched.w[sched.off_mat_of + kcat * p.n_nodes + node] : p.mat_of[(size_t)kcat * p.n_nodes + node];
            p_PT = p.arena + (size_t)mat * p.LD * p.LD;
        }
    };
    auto produce_one = [&]() {
        if (n_gemm == 0 || p_tile >= n_tiles) return;
        if (warp == p_turn) {
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
        if (++p_turn == NW) p_turn = 0;
        if (++p_stage == NS) { p_stage = 0; p_phase ^= 1u; }
        if (++p_chunk == n_chunks) {
            p_chunk = 0;
            if (++p_g == n_gemm) { p_g = 0; p_tile += gridDim.x; }
            p_lookup();
        }
    };
    p_lookup();
    for (int s = 0; s < NS - 1; ++s) produce_one();
    int c_stage = 0;
    unsigned c_phase = 0;
    unsigned free_phase = 0, vfull_phase = 0;         // parities of v_free (one phase per step) and v_full (one per staged step)

    int c_job = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int k;
        int64_t col0, U, U_stride;
        const int32_t* __restrict__ ids;        // [rows][U_stride] leaf counts (and pattern ids of table children) per column
        if (JOBS) {
            while (c_job + 1 < sched.n_jobs && tile >= job_word(c_job + 1, 0)) ++c_job;
            const int local = tile - job_word(c_job, 0), ncol = job_word(c_job, 1);
            k = local / ncol;
            col0 = (int64_t)(local % ncol) * BN;
            U = job_word(c_job, 2);
            U_stride = job_word(c_job, 3);
            ids = p.counts_t + (((int64_t)job_word(c_job, 5) << 32) | (uint32_t)job_word(c_job, 4));
        } else {
            k = tile / p.n_col_tiles;
            col0 = (int64_t)(tile % p.n_col_tiles) * BN;
            U = p.U;
            U_stride = p.U_stride;
            ids = p.counts_t;
        }
        const int32_t* mat_of = p.mat_of + (size_t)k * p.n_nodes;
        double acc[TMW][TNW][2];
        int pf_id = -1;                         // table row this thread pulls into L2 for the next step's gathers (-1: none)

        for (int st = JOBS ? c_job : 0; st < (JOBS ? c_job + 1 : p.n_steps); ++st) {
            Step sp;
            if (sched.valid) {
                const int o = st * 9;
                sp.node = sched.w[o]; sp.is_root = sched.w[o + 1]; sp.out_slot = sched.w[o + 2]; sp.n_children = sched.w[o + 3];
                sp.child_begin = sched.w[o + 4]; sp.parent_step = sched.w[o + 5]; sp.carry_in = sched.w[o + 6];
                sp.dst_kind = sched.w[o + 7]; sp.f_slot = sched.w[o + 8];
            } else sp = p.steps[st];
            auto child_at = [&](int ci) {
                StepChild ch;
                if (sched.valid) {
                    const int o = sched.off_children + (sp.child_begin + ci) * 5;
                    ch.node = sched.w[o]; ch.leaf_row = sched.w[o + 1]; ch.slot = sched.w[o + 2]; ch.kind = sched.w[o + 3]; ch.f_slot = sched.w[o + 4];
                } else ch = p.children[sp.child_begin + ci];
                return ch;
            };
            auto gather_base = [&](const StepChild& ch) -> const double* {     // row 0 of the child's gather source (matrix or table)
                const int mat = sched.valid ? sched.w[sched.off_mat_of + k * p.n_nodes + ch.node] : mat_of[ch.node];
                return ch.kind == 3 ? p.tables + ((size_t)ch.slot * p.K + (size_t)k * ch.f_slot) * p.LD : p.arena + (size_t)mat * p.LD * p.LD;
            };
            // The LAST child of the product goes through the copy engine when it is a plain column gather (a leaf without an error model,
            // or a table node).  Products are formed in the schedule's child order either way (probability.cpp:215-217, 229-231).
            bool staged = false;
            StepChild sch{};
            if (sp.n_children > 0) {
                sch = child_at(sp.n_children - 1);
                staged = sch.kind == 3 || (sch.kind == 0 && p.em == nullptr);
            }
            int staged_id = 0;
            if (staged && tid < BN) {
                int64_t u = col0 + tid;
                if (u >= U) u = U - 1;                            // padding columns replay the last family; never written out
                staged_id = ids[(size_t)sch.leaf_row * U_stride + u];
            }
            // this warp no longer reads VT (previous contraction / epilogue): generic-proxy accesses before, async-proxy writes after
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(v_free);

            bool has_acc = sp.carry_in != 0;
            // ---- 1. the factors of the other children, through registers ----
            for (int ci = 0; ci < sp.n_children - (staged ? 1 : 0); ++ci) {
                const StepChild ch = child_at(ci);
                if (ch.kind == 1) continue;                       // carried: already in acc
                if (ch.kind == 0 || ch.kind == 3) {
                    // leaf: factor[s] = sum_d em[obs][d] * P(s -> obs-1+d)   (probability.cpp:187-202)
                    // subtree-pattern table (kind 3): the child's factor of pattern `id`, gathered like a leaf column of its table (no
                    // error model: it is already inside the table)
                    leaf_factor_into<TMW, TNW>(acc, has_acc, p, gather_base(ch), ids + (size_t)ch.leaf_row * U_stride, U,
                                               ch.kind == 3 ? nullptr : p.em, row_base, col0, col_base);
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

            // ---- 2. V_v -> VT.  Everyone waits until every warp has left the previous contraction; the staged child's columns are
            // then copied in by the copy engine (one 16-byte-aligned bulk copy of BM doubles per column, issued by BN threads) ----
            mbar_wait(v_free, free_phase);
            free_phase ^= 1u;
            if (staged) {
                if (tid < BN) {
                    mbar_expect_tx(v_full, (unsigned)(BM * sizeof(double)));
                    bulk_g2s(VT + (size_t)tid * BMP, gather_base(sch) + (size_t)staged_id * p.LD, BM * sizeof(double), v_full);
                }
                mbar_wait(v_full, vfull_phase);
                vfull_phase ^= 1u;
            }
            // the product, in place: a thread reads and writes only its own elements.  States >= S do not exist: zero rows (the matrix
            // rows there may be real).
#pragma unroll
            for (int j = 0; j < TNW; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double* const vt = VT + (size_t)(col_base + j * 8 + e) * BMP + row_base;
#pragma unroll
                    for (int i = 0; i < TMW; ++i) {
                        const bool live = sp.is_root || row_base + i * 8 < p.S;
                        double v = has_acc ? acc[i][j][e] : 1.0;
                        if (staged) {
                            const double f = vt[i * 8];
                            v = has_acc ? __dmul_rn(v, f) : f;
                        }
                        vt[i * 8] = live ? v : 0.0;
                    }
                }
            __syncthreads();

            // The next step's leaf counts come from DRAM (the count table is streamed once per category): pull their lines into L2 now, a
            // whole contraction ahead of the gather that needs them.
            if (!JOBS && sched.valid && tid < 2) {
                int nst = st + 1, ntile = tile;
                if (nst == p.n_steps) { nst = 0; ntile += gridDim.x; }
                if (ntile < n_tiles) {
                    const int64_t ncol = (int64_t)(ntile % p.n_col_tiles) * BN + tid * 32;
                    const int o = nst * 9, nch = sched.w[o + 3], cb = sched.w[o + 4];
                    for (int ci = 0; ci < nch; ++ci) {
                        const int lr = sched.w[sched.off_children + (cb + ci) * 5 + 1];
                        if (lr >= 0 && ncol < p.U_stride)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.counts_t + (size_t)lr * p.U_stride + ncol));
                    }
                }
            }
            if (!sp.is_root) {
                // ---- 3. W_v = P_v . V_v on the FP64 tensor cores (matrix_cache.cpp:49-56) ----
#pragma unroll
                for (int i = 0; i < TMW; ++i)
#pragma unroll
                    for (int j = 0; j < TNW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
                const double* const Bw = VT + (size_t)(wn * 8 * TNW + g) * BMP + q;     // + j*8*BMP + state
                for (int chunk = 0; chunk < n_chunks; ++chunk) {
                    if (PRODUCE_FIRST) produce_one();               // refill the stage the previous chunk released: NS-1 chunks of lead
                    long long t_a = 0, t_b = 0, t_c = 0;
                    if (PROBE) t_a = clock64();
                    mbar_wait(full_bar + c_stage, c_phase);
                    if (PROBE) t_b = clock64();
                    const double* As = stages + (size_t)c_stage * Cfg::STAGE_DOUBLES;
                    const double* Bs = Bw + chunk * BK;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 4; ++k4) {
                        double a[TMW], b[TNW];
                        const double* ap = As + (k4 * 4 + q) * BMP + row_base;
#pragma unroll
                        for (int i = 0; i < TMW; ++i) a[i] = ap[i * 8];
#pragma unroll
                        for (int j = 0; j < TNW; ++j) b[j] = Bs[(size_t)j * 8 * BMP + k4 * 4];
#pragma unroll
                        for (int i = 0; i < TMW; ++i) {
                            // A warp's DMMAs issue exactly 16 cycles apart (ptxas: stall 15 + NOP) and it cannot catch up on time spent elsewhere,
                            // so the ring refill sits INSIDE the DMMA stream rather than after it (59.6 -> 59.2 ms per launch; as a serial
                            // tail the refill + cursor bookkeeping was 15 % of a warp's time, profiles/r01_probe_chunks.txt).  Probing the
                            // next chunk's barrier from inside the stream as well lost time (60.1 ms), as did loading the fragments just in time.
                            if (MID && k4 == BK / 4 - 1 && i == TMW / 2) produce_one();
#pragma unroll
                            for (int j = 0; j < TNW; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                        }
                    }
                    if (PROBE) t_c = clock64();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar + c_stage);
                    if (++c_stage == NS) { c_stage = 0; c_phase ^= 1u; }
                    if (!PRODUCE_FIRST && !MID) produce_one();
                    // The table rows the NEXT step (or this CTA's next tile) gathers come from HBM: their ids are loaded early in this
                    // contraction and the rows pulled into L2 from its middle, so that the staged bulk copy and the register gather
                    // find them there.  Thread c of role r handles column c of the last (r = 0) / first (r = 1) child.
                    if (sched.valid && (chunk == 1 || chunk == (n_chunks >> 1)) && tid < 2 * BN) {
                        int nst = st + 1, ntile = tile, nk = k;
                        int64_t ncol0 = col0, nU = U, nUs = U_stride;
                        const int32_t* nids = ids;
                        bool have = true;
                        if (JOBS || nst == p.n_steps) {
                            ntile += gridDim.x;
                            have = ntile < n_tiles;
                            if (have) {
                                if (JOBS) {
                                    int nj = c_job;
                                    while (nj + 1 < sched.n_jobs && ntile >= job_word(nj + 1, 0)) ++nj;
                                    const int local = ntile - job_word(nj, 0), ncol = job_word(nj, 1);
                                    nst = nj;
                                    nk = local / ncol;
                                    ncol0 = (int64_t)(local % ncol) * BN;
                                    nU = job_word(nj, 2);
                                    nUs = job_word(nj, 3);
                                    nids = p.counts_t + (((int64_t)job_word(nj, 5) << 32) | (uint32_t)job_word(nj, 4));
                                } else {
                                    nst = 0;
                                    nk = ntile / p.n_col_tiles;
                                    ncol0 = (int64_t)(ntile % p.n_col_tiles) * BN;
                                }
                            }
                        }
                        if (have) {
                            const int nch = sched.w[nst * 9 + 3], cb = sched.w[nst * 9 + 4];
                            const int role = tid / BN;
                            const int ci = role == 0 ? nch - 1 : 0;
                            if (nch > role && ci >= 0) {
                                const int o = sched.off_children + (cb + ci) * 5;
                                if (sched.w[o + 3] == 3) {                      // a table node
                                    if (chunk == 1) {
                                        int64_t u = ncol0 + (tid % BN);
                                        if (u >= nU) u = nU - 1;
                                        pf_id = nids[(size_t)sched.w[o + 1] * nUs + u];
                                    } else if (pf_id >= 0) {
                                        const char* row = reinterpret_cast<const char*>(
                                            p.tables + ((size_t)sched.w[o + 2] * p.K + (size_t)nk * sched.w[o + 4] + pf_id) * p.LD);
#pragma unroll
                                        for (int b = 0; b < (BM * 8 + 127) / 128; ++b) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + b * 128));
                                        pf_id = -1;
                                    }
                                }
                            }
                        }
                    }
                    if (PROBE && probe_out && probe_n < PROBE_CHUNKS) {
                        probe_out[probe_n * 4 + 0] = t_a;
                        probe_out[probe_n * 4 + 1] = t_b;
                        probe_out[probe_n * 4 + 2] = t_c;
                        probe_out[probe_n * 4 + 3] = clock64();
                        ++probe_n;
                    }
                }
                // ---- 4. the factor stays in registers for the parent, is parked once in a global slot, or becomes table rows ----
                if (JOBS && sp.dst_kind == 3) {
                    // one table row per pattern: the transposed layout the gathers of the parent read
                    double* __restrict__ tb = p.tables + ((size_t)sp.f_slot * p.K + (size_t)k * U) * p.LD;
#pragma unroll
                    for (int j = 0; j < TNW; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int64_t col = col0 + col_base + j * 8 + e;
                            if (col < U) {
#pragma unroll
                                for (int i = 0; i < TMW; ++i) tb[(size_t)col * p.LD + row_base + i * 8] = acc[i][j][e];
                            }
                        }
                } else if (sp.dst_kind == 1) {
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
                root_epilogue<BN, Cfg::PARTS, true>(p, VT, BMP, red, tid, k, col0);
            }
        }
        __syncthreads();   // factor slots of this tile are dead; VT / red are reused by the next tile
    }
}

}  // namespace cafe
