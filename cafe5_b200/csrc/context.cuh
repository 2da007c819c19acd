// context.cuh -- the context behind the opaque cafe_b200_ctx handle and the host helpers the translation units of libcafe_b200.so
// share: error plumbing, device buffers, the per-evaluation matrix-key plan.  cafe_b200.cu owns the likelihood path (create, set_*,
// eval_*, reconstruct, test hooks); abi_analysis.cu builds the simulator, the p-values and the branch probabilities on top of it.
#pragma once
#include "../../include/cafe_b200.h"
#include "launchers.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace cafe {

constexpr int DM_BK_HOST = 16;   // DM_BK of prune_dmma.cuh
std::string& create_error();          // last failed cafe_b200_create on this thread

struct CudaError { std::string msg; };

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw cafe::CudaError{buf_};                                                                 \
        }                                                                                          \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n, bool zero = false)
    {
        if (n <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        CK(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
        if (zero) {
            // cudaMemset runs on the legacy default stream, which the contexts' non-blocking streams do NOT wait for: without the
            // synchronisation the zeros can land after a kernel on the context's stream has started writing the buffer (seen with
            // several contexts on one device: whole shards of factor tables wiped during their first evaluation)
            CK(cudaMemset(p, 0, n * sizeof(T)));
            CK(cudaStreamSynchronize(cudaStreamLegacy));
        }
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// A pruning schedule over the internal nodes of the tree (post-order, children before parents).
struct Schedule {
    std::vector<Step> steps;
    std::vector<StepChild> children;
    std::vector<int32_t> gemm_nodes;       // resident kernel: node of the g-th contraction of a tile
    int n_slots = 0, n_fslots = 1;
};

// Subtree-pattern reuse: an internal node whose subtree shows few DISTINCT patterns of leaf counts among the unique
// families gets its factor W_v = P_v . V_v computed once per pattern (a "table"), level by level from the cherries up; the main
// pass gathers it like a leaf column.  The reference prunes identical FAMILIES once (build_reference_list, base_model.cpp:27-51);
// this is the same idea below the root.
struct TableNode {
    int node = 0, level = 0;
    int64_t D = 0, D_stride = 0;           // distinct patterns; padded to a multiple of 64
    int64_t rows_before = 0;               // patterns of the table nodes before this one (table row offset per category)
    int64_t ids_off = 0;                   // this node's id table in d_ids: [n_children][D_stride] child ids of every pattern
    std::vector<int> kids;                 // children in the reference's product order (decreasing index)
};

struct KeyPlan {
    std::vector<MatParam> params;          // one per distinct matrix key
    std::vector<int32_t> mat_of;           // [K][n_nodes]
};

// One worker thread per device shard of a cafe_b200_create_multi context: a call on the group runs the single-device entry point
// of every shard at the same time (each worker owns its device's stream; CUDA's current device is per thread).
class ShardPool {
public:
    explicit ShardPool(int n) : rc_(n, 0)
    {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
    }
    ~ShardPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++generation_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // runs fn(i) on worker i for every shard and returns the first non-zero result (in shard order), 0 when all succeeded
    int run(const std::function<int(int)>& fn, int* failed_shard = nullptr)
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            pending_ = (int)workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
        for (size_t i = 0; i < rc_.size(); ++i)
            if (rc_[i] != 0) { if (failed_shard) *failed_shard = (int)i; return rc_[i]; }
        return 0;
    }
private:
    void loop(int i)
    {
        unsigned long seen = 0;
        for (;;) {
            const std::function<int(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                fn = fn_;
            }
            const int rc = (*fn)(i);
            {
                std::lock_guard<std::mutex> lk(m_);
                rc_[i] = rc;
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::vector<int> rc_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<int(int)>* fn_ = nullptr;
    unsigned long generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace cafe

struct cafe_b200_ctx {
    // cafe_b200_create_multi: a group context owns one single-device context per shard and nothing on any device itself
    std::vector<cafe_b200_ctx*> shards;
    std::vector<int64_t> shard_begin;      // [n_shards + 1] first family of every shard
    cafe::ShardPool* pool = nullptr;
    bool is_group() const { return !shards.empty(); }
    // cafe_b200_create_bucketed: the shards are BUCKETS of families by largest count (own, smaller state space each) on one device;
    // order[shard_begin[b] + j] is the caller's index of family j of bucket b (empty: shards are contiguous blocks in caller order)
    std::vector<int64_t> order;
    // page-locked scratch of a group whose shards are not contiguous in the caller's order: the shards write their outputs there
    // and scatter them into the caller's arrays (each shard worker its own rows)
    void* g_scratch = nullptr;
    size_t g_scratch_cap = 0, g_scratch_used = 0;
    void* take_scratch(size_t bytes)
    {
        bytes = (bytes + 255) / 256 * 256;
        if (g_scratch_used + bytes > g_scratch_cap) return nullptr;
        void* p = (char*)g_scratch + g_scratch_used;
        g_scratch_used += bytes;
        return p;
    }

    int device = 0;
    int n_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // tree (host copies)
    int n_nodes = 0;
    std::vector<int32_t> parent, leaf_col, lambda_class;
    std::vector<double> branch_length;
    int n_lambda_classes = 1;
    int S = 0, R = 0, N = 0, max_family_size = 0;

    // families
    int64_t F = 0, U = 0, U_stride = 0;
    int n_species = 0;
    std::vector<int32_t> max_count;        // [F] largest count of every family (find_best_pvalue, src/probability.cpp:513-526)
    std::vector<int64_t> f2u;

    // schedule
    std::vector<cafe::Step> steps;
    std::vector<cafe::StepChild> children;
    int n_slots = 0;
    std::vector<int32_t> leaf_row_of_node; // row in counts_t for leaf nodes (and, past the leaves, for the table nodes the main pass gathers)

    // subtree-pattern tables (resident wn2 kernel only; CAFE_B200_TABLES=0 disables, =force enables for any problem size)
    bool tables_on = false;
    std::vector<cafe::TableNode> tnodes;   // sorted by level
    std::vector<int> tnode_of;             // node -> index in tnodes, -1
    int n_table_levels = 0;
    int64_t table_rows = 0;                // sum of D over the table nodes (rows per category)
    cafe::Schedule tsched_main;            // the main pass with table nodes as pseudo-leaves
    cafe::InlineSchedule tsched{};         // its kernel-parameter form (per evaluation: carries the key index)
    std::vector<cafe::Schedule> tsched_jobs;   // one single-step-per-job schedule per table launch
    std::vector<std::vector<int>> tjobs;   // table nodes (indices into tnodes) of every table launch
    cafe::DevBuf<int32_t> d_ids;
    cafe::DevBuf<double> d_tables;
    // Pupko, second design (pupko2.cuh): built on first use
    std::vector<int32_t> tab_ids_host;     // [tnodes.size()][U_stride] pattern id of every unique family at every table node
    bool p2_ready = false;
    cafe::DevBuf<cafe::Step> d_p2_steps;             // main schedule (reduced when tables are on)
    cafe::DevBuf<cafe::StepChild> d_p2_children;
    std::vector<cafe::DevBuf<cafe::Step>> d_p2_job_steps;        // per table launch
    std::vector<cafe::DevBuf<cafe::StepChild>> d_p2_job_children;
    cafe::DevBuf<int32_t> d_p2_jobs, d_p2_tab_of, d_p2_tab_ids, d_p2_parent, d_p2_leaf_col, d_p2_root;
    cafe::DevBuf<int64_t> d_p2_coff;                 // [2][n_nodes]: offsets, columns
    cafe::DevBuf<uint16_t> d_p2_ctab;
    int pupko_version = 2;                           // CAFE_B200_PUPKO=1: the round-1 kernel (pupko.cuh)
    cafe::KeyPlan last_kp;                 // key plan of the evaluation being launched
    int last_table_launches = 0;
    int64_t last_columns = 0;              // contraction columns x categories executed by the last evaluation (tables + main pass)

    // tiling choice
    int TM = 11, TN = 4, n_mtiles = 1, LD = 176, n_col_tiles = 0, grid = 0;
    // pruning kernel: 2 = DMMA with the child vector resident in shared memory (default when it fits), 1 = DMMA streaming
    // both operands (large state spaces), 0 = DFMA register tiles.  CAFE_B200_PRUNE=resident|stream|dfma
    int prune_pref = 2, prune_kind = 2;
    bool use_dmma = true;
    std::vector<int32_t> gemm_nodes;
    cafe::InlineSchedule sched{};                // schedule + key index as a kernel parameter (resident kernel)
    int n_fslots = 1;
    int TNW = 4, WN = 2, resident_wn = 2, dmma_stages = 4;
    size_t smem_optin = 0, smem_per_sm = 0;

    // prior / error model
    bool have_prior = false;
    std::vector<float> prior;
    bool have_em = false;
    int em_rows = 0, em_maxcnt = 0;
    std::vector<double> em_host;           // the rows last given to set_error_model (the p-value path switches the model off and back on)

    // device buffers
    cafe::DevBuf<int32_t> d_counts_t, d_mat_of, d_gemm_nodes;
    cafe::DevBuf<int64_t> d_f2u;
    cafe::DevBuf<cafe::Step> d_steps;
    cafe::DevBuf<cafe::StepChild> d_children;
    cafe::DevBuf<double> d_zero, d_lg, d_arena, d_scratch, d_prior, d_logprior, d_em, d_best, d_cat_probs;
    cafe::DevBuf<uint8_t> d_ok;
    cafe::DevBuf<cafe::MatParam> d_params;
    cafe::DevBuf<double> d_powtab;               // [N][n_mats] pow(coeff, j) of every key (pow_table_kernel)
    bool resident_probe = false;                 // CAFE_B200_RESIDENT_PROBE=1: timing build with clock stamps (cafe_b200_debug_read_probe)
    cafe::DevBuf<int64_t> d_probe;
    int pupko_threads = 512;                     // CAFE_B200_PUPKO_THREADS=256: the two-warps-per-sub-partition comparison geometry
    bool matgen_entry = false;                   // CAFE_B200_MATGEN=entry: the one-thread-per-entry comparison kernel
    bool matgen_libexp = false;                  // CAFE_B200_MATGEN=rows: the default kernel with the library exp() per term
    cafe::DevBuf<double> d_family_lnl, d_cat_lk, d_family_lk, d_posterior, d_partial, d_partial_fail, d_result, d_roots;
    cafe::DevBuf<uint8_t> d_significant, d_failed;
    // pupko
    cafe::DevBuf<double> d_pupko_m;              // [grid][n_steps][kpad * BN]: M_v of every node of the tiles in flight
    cafe::DevBuf<int32_t> d_states, d_leaf_row, d_states_f, d_cat_states_f;
    cafe::DevBuf<double> d_avg_f;
    int lg_n = 0;

    // pinned staging
    void* h_stage = nullptr;
    size_t h_stage_cap = 0;
    double* h_result = nullptr;

    // stats
    int last_launches = 0, last_mats = 0;
    bool stats_valid = false;

    void* stage(size_t bytes)
    {
        if (bytes > h_stage_cap) {
            if (h_stage) cudaFreeHost(h_stage);
            h_stage = nullptr;
            h_stage_cap = 0;
            CK(cudaMallocHost(&h_stage, bytes));
            h_stage_cap = bytes;
        }
        return h_stage;
    }
};


namespace cafe {

// matrix_cache_key planning for one evaluation (src/matrix_cache.h:44-63; lambda::multiply, src/lambda.h:45-48,76-84)
KeyPlan plan_keys(const cafe_b200_ctx* c, const double* lambdas, const double* multipliers, int K);
// key parameters + key index -> device (and the kernel-parameter copy of the schedule); grows the matrix arena
void upload_plan(cafe_b200_ctx* c, const KeyPlan& kp);
void launch_matrices(cafe_b200_ctx* c, int n_mats);
// CudaError -> status code; stores the message in the context (or as the thread's create error)
int fail(cafe_b200_ctx* c, const CudaError& e);

template <typename T>
void d2h(cafe_b200_ctx* c, T* dst, const T* src, size_t n)
{
    if (dst && n) CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
}

}  // namespace cafe
