// cafe_b200.cu -- host side of libcafe_b200.so: context, per-evaluation key planning, launches, the likelihood-path C ABI.
// See include/cafe_b200.h for the boundary and the reference interfaces each entry point replaces.
#define CAFE_KERNELS_IMPL
#include "context.cuh"
#include "peak.cuh"

using namespace cafe;

namespace cafe {
std::string& create_error()
{
    thread_local std::string s;
    return s;
}
}  // namespace cafe

namespace {

// ---- schedule: post-order over internal nodes, heavier subtree first, so few vectors are live ----
// pseudo_leaf[v] != 0: v is a table node the pass gathers like a leaf (its subtree is not scheduled); empty = the whole tree.
void build_schedule(const cafe_b200_ctx* c, const std::vector<char>& pseudo_leaf, Schedule& out)
{
    const int n = c->n_nodes;
    auto leafish = [&](int v) { return c->leaf_col[v] >= 0 || (!pseudo_leaf.empty() && pseudo_leaf[v]); };
    std::vector<std::vector<int>> kids(n);
    for (int i = n - 1; i >= 0; --i)
        if (c->parent[i] >= 0) kids[c->parent[i]].push_back(i);   // decreasing index == reference descendant order
    std::vector<int> need(n, 0);
    for (int i = 0; i < n; ++i) {                                 // children precede parents
        if (leafish(i)) continue;
        std::vector<int> sub;
        for (int k : kids[i]) if (!leafish(k)) sub.push_back(need[k]);
        std::sort(sub.rbegin(), sub.rend());
        int best = 1;
        for (size_t j = 0; j < sub.size(); ++j) best = std::max(best, sub[j] + (int)j);
        need[i] = std::max(best, (int)sub.size() + 0) ;
        if (need[i] < 1) need[i] = 1;
    }
    std::vector<int> order;
    std::vector<int> stack{n - 1};
    // iterative post-order with children visited by decreasing need
    std::vector<int> state(n, 0);
    std::vector<std::vector<int>> visit(n);
    for (int i = 0; i < n; ++i) {
        if (leafish(i)) continue;
        for (int k : kids[i]) if (!leafish(k)) visit[i].push_back(k);
        std::stable_sort(visit[i].begin(), visit[i].end(), [&](int a, int b) { return need[a] > need[b]; });
    }
    while (!stack.empty()) {
        int v = stack.back();
        if (state[v] < (int)visit[v].size()) stack.push_back(visit[v][state[v]++]);
        else { order.push_back(v); stack.pop_back(); }
    }
    // slot assignment with reuse
    std::vector<int> slot_of(n, -1);
    std::vector<int> free_slots;
    int n_slots = 0;
    out.steps.clear();
    out.children.clear();
    for (int v : order) {
        Step st{};
        st.node = v;
        st.is_root = c->parent[v] < 0;
        st.child_begin = (int)out.children.size();
        st.n_children = (int)kids[v].size();
        int s;
        if (!free_slots.empty()) { s = free_slots.back(); free_slots.pop_back(); }
        else s = n_slots++;
        st.out_slot = s;
        slot_of[v] = s;
        std::vector<int> order_kids = kids[v];
        // A two-child product is commutative bit for bit, so the contraction (internal child) goes first and the
        // leaf factor is multiplied into the accumulators in registers: no parking of the running product.
        // Three or more children keep the reference's descendant order (the association matters there).
        if (order_kids.size() == 2 && leafish(order_kids[0]) && !leafish(order_kids[1]))
            std::swap(order_kids[0], order_kids[1]);
        for (int k : order_kids) {
            StepChild ch{};
            ch.node = k;
            ch.leaf_row = leafish(k) ? c->leaf_row_of_node[k] : -1;
            ch.slot = leafish(k) ? -1 : slot_of[k];
            out.children.push_back(ch);
        }
        out.steps.push_back(st);
        for (int k : kids[v]) if (!leafish(k)) free_slots.push_back(slot_of[k]);
    }
    std::vector<int> step_of(n, -1);
    for (size_t i = 0; i < out.steps.size(); ++i) step_of[out.steps[i].node] = (int)i;
    for (auto& st : out.steps) st.parent_step = c->parent[st.node] < 0 ? -1 : step_of[c->parent[st.node]];
    out.n_slots = n_slots;

    // ---- resident kernel: which factor travels how ----
    auto chain = [&](int v) {      // v's factor stays in registers: its parent is the very next step and has <= 2 children
        if (leafish(v) || c->parent[v] < 0) return false;
        int P = c->parent[v];
        return kids[P].size() <= 2 && step_of[P] == step_of[v] + 1;
    };
    std::vector<int> fslot_of(n, -1), free_f;
    int n_f = 0;
    out.gemm_nodes.clear();
    for (auto& st : out.steps) {
        const int v = st.node;
        st.carry_in = 0;
        for (int i = 0; i < st.n_children; ++i) {
            StepChild& ch = out.children[st.child_begin + i];
            if (c->leaf_col[ch.node] >= 0) { ch.kind = 0; ch.f_slot = -1; }
            else if (leafish(ch.node)) {                      // table node: gathered like a leaf column of its factor table
                const TableNode& t = c->tnodes[c->tnode_of[ch.node]];
                ch.kind = 3; ch.slot = (int32_t)t.rows_before; ch.f_slot = (int32_t)t.D;
            }
            else if (chain(ch.node)) { ch.kind = 1; ch.f_slot = -1; st.carry_in = 1; }
            else { ch.kind = 2; ch.f_slot = fslot_of[ch.node]; }
        }
        for (int i = 0; i < st.n_children; ++i) {      // slots of consumed factors are free again
            const StepChild& ch = out.children[st.child_begin + i];
            if (ch.kind == 2) free_f.push_back(ch.f_slot);
        }
        st.f_slot = -1;
        if (st.is_root) st.dst_kind = 2;
        else {
            out.gemm_nodes.push_back(v);
            if (chain(v)) st.dst_kind = 0;
            else {
                st.dst_kind = 1;
                if (!free_f.empty()) { st.f_slot = free_f.back(); free_f.pop_back(); }
                else st.f_slot = n_f++;
                fslot_of[v] = st.f_slot;
            }
        }
    }
    out.n_fslots = std::max(n_f, 1);
}

// ---- subtree-pattern tables: which nodes, their distinct patterns, the id tables (host, once per context) ----
// counts_t: [n_leaves][U_stride] unique count table.  Appends one id row per table node the main pass gathers (counts_t grows to
// [n_leaves + n_cut][U_stride]) and fills c->tnodes / d_ids / the schedules.  A node qualifies when every child is a leaf or a
// table node and its distinct patterns number at most frac * U.
void plan_tables(cafe_b200_ctx* c, std::vector<int32_t>& counts_t, int n_leaves)
{
    c->tables_on = false;
    c->tnodes.clear();
    c->tnode_of.assign(c->n_nodes, -1);
    const char* mode = std::getenv("CAFE_B200_TABLES");
    const bool force = mode && std::strcmp(mode, "force") == 0;
    if (mode && std::strcmp(mode, "0") == 0) return;
    double frac = 0.75;   // measured on the config-5 shard (B200): 0.25 -> 33.9 ms, 0.5 -> 31.2, 0.75 -> 30.1, 0.9 -> 30.7 (no tables: 61.3)
    if (const char* e = std::getenv("CAFE_B200_TABLE_FRAC")) frac = std::atof(e);
    if (force && !std::getenv("CAFE_B200_TABLE_FRAC")) frac = 1.0;
    const int64_t U = c->U, US = c->U_stride;
    if (!force && U < 8192) return;                  // small problems: the extra launches cost more than the columns they save
    if (c->n_mtiles != 1 || !c->use_dmma || c->prune_pref < 2 || c->resident_wn != 2) return;
    const int n = c->n_nodes;
    std::vector<std::vector<int>> kids(n);
    for (int i = n - 1; i >= 0; --i)
        if (c->parent[i] >= 0) kids[c->parent[i]].push_back(i);
    std::vector<std::vector<int32_t>> ids(n);        // per table node: pattern id of every unique family
    std::vector<int> level(n, -1);
    std::vector<std::vector<int32_t>> job_ids(n);    // per table node: [n_children][D_stride]
    auto col_of = [&](int v, int64_t u) -> int32_t {
        return c->leaf_col[v] >= 0 ? counts_t[(size_t)c->leaf_row_of_node[v] * US + u] : ids[v][u];
    };
    for (int v = 0; v < n - 1; ++v) {                // children precede parents; the root is never a table
        if (c->leaf_col[v] >= 0) continue;
        bool ok = true;
        int lv = 0;
        for (int k : kids[v]) {
            if (c->leaf_col[k] >= 0) continue;
            if (level[k] < 0) { ok = false; break; }
            lv = std::max(lv, level[k] + 1);
        }
        if (!ok) continue;
        const int m = (int)kids[v].size();
        const int64_t limit = (int64_t)std::floor(frac * (double)U);
        std::vector<int32_t> id(U);
        std::vector<int32_t> first;                  // child ids of every pattern, pattern-major
        int64_t D = 0;
        bool over = false;
        if (m == 2) {
            std::unordered_map<uint64_t, int32_t> seen;
            seen.reserve((size_t)std::min<int64_t>(U, 1 << 22));
            for (int64_t u = 0; u < U && !over; ++u) {
                const uint32_t a = (uint32_t)col_of(kids[v][0], u), b = (uint32_t)col_of(kids[v][1], u);
                auto it = seen.find(((uint64_t)a << 32) | b);
                if (it == seen.end()) {
                    if (D >= limit) { over = true; break; }
                    it = seen.emplace(((uint64_t)a << 32) | b, (int32_t)D).first;
                    first.push_back((int32_t)a); first.push_back((int32_t)b);
                    ++D;
                }
                id[u] = it->second;
            }
        } else {
            std::unordered_map<std::string, int32_t> seen;
            std::string key((size_t)m * sizeof(int32_t), '\0');
            for (int64_t u = 0; u < U && !over; ++u) {
                for (int j = 0; j < m; ++j) { const int32_t x = col_of(kids[v][j], u); memcpy(&key[(size_t)j * sizeof x], &x, sizeof x); }
                auto it = seen.find(key);
                if (it == seen.end()) {
                    if (D >= limit) { over = true; break; }
                    it = seen.emplace(key, (int32_t)D).first;
                    for (int j = 0; j < m; ++j) first.push_back(col_of(kids[v][j], u));
                    ++D;
                }
                id[u] = it->second;
            }
        }
        if (over || D == 0) continue;
        level[v] = lv;
        ids[v] = std::move(id);
        TableNode t;
        t.node = v; t.level = lv; t.D = D; t.D_stride = (D + 63) / 64 * 64; t.kids = kids[v];
        job_ids[v].assign((size_t)m * t.D_stride, 0);
        for (int64_t d = 0; d < D; ++d)
            for (int j = 0; j < m; ++j) job_ids[v][(size_t)j * t.D_stride + d] = first[(size_t)d * m + j];
        c->tnodes.push_back(std::move(t));
    }
    if (c->tnodes.empty()) return;
    std::stable_sort(c->tnodes.begin(), c->tnodes.end(), [](const TableNode& a, const TableNode& b) { return a.level < b.level; });
    c->table_rows = 0;
    c->n_table_levels = 0;
    std::vector<int32_t> all_ids;
    for (size_t i = 0; i < c->tnodes.size(); ++i) {
        TableNode& t = c->tnodes[i];
        c->tnode_of[t.node] = (int)i;
        t.rows_before = c->table_rows;
        c->table_rows += t.D;
        t.ids_off = (int64_t)all_ids.size();
        all_ids.insert(all_ids.end(), job_ids[t.node].begin(), job_ids[t.node].end());
        c->n_table_levels = std::max(c->n_table_levels, t.level + 1);
    }
    if (c->table_rows * 8 > (int64_t)INT32_MAX) { c->tnodes.clear(); c->tnode_of.assign(n, -1); return; }   // row offsets are int32 in the kernel (K <= 8)
    // the main pass gathers the table nodes whose parent is not a table: one more id row each
    std::vector<char> pseudo(n, 0);
    int extra = 0;
    for (const TableNode& t : c->tnodes)
        if (c->tnode_of[c->parent[t.node]] < 0) {
            pseudo[t.node] = 1;
            c->leaf_row_of_node[t.node] = n_leaves + extra++;
        }
    counts_t.resize((size_t)(n_leaves + extra) * US, 0);
    for (const TableNode& t : c->tnodes)
        if (pseudo[t.node]) std::copy(ids[t.node].begin(), ids[t.node].end(), counts_t.begin() + (size_t)c->leaf_row_of_node[t.node] * US);
    c->tab_ids_host.assign(c->tnodes.size() * (size_t)US, 0);        // the Pupko traceback needs every table node's pattern id per family
    for (size_t i = 0; i < c->tnodes.size(); ++i)
        std::copy(ids[c->tnodes[i].node].begin(), ids[c->tnodes[i].node].end(), c->tab_ids_host.begin() + i * (size_t)US);
    build_schedule(c, pseudo, c->tsched_main);
    // table launches: level by level, at most MAX_TABLE_JOBS nodes per launch; job j = step j (one step, all children gathers)
    c->tsched_jobs.clear();
    c->tjobs.clear();
    for (int lv = 0; lv < c->n_table_levels; ++lv) {
        std::vector<int> members;
        for (size_t i = 0; i < c->tnodes.size(); ++i) if (c->tnodes[i].level == lv) members.push_back((int)i);
        for (size_t b = 0; b < members.size(); b += MAX_TABLE_JOBS) {
            std::vector<int> group(members.begin() + b, members.begin() + std::min(members.size(), b + MAX_TABLE_JOBS));
            Schedule sc;
            for (int ti : group) {
                const TableNode& t = c->tnodes[ti];
                Step st{};
                st.node = t.node; st.is_root = 0; st.out_slot = 0; st.n_children = (int)t.kids.size(); st.child_begin = (int)sc.children.size();
                st.parent_step = -1; st.carry_in = 0; st.dst_kind = 3; st.f_slot = (int32_t)t.rows_before;
                for (size_t j = 0; j < t.kids.size(); ++j) {
                    const int k = t.kids[j];
                    StepChild ch{};
                    ch.node = k; ch.leaf_row = (int32_t)j;
                    if (c->leaf_col[k] >= 0) { ch.kind = 0; ch.slot = -1; ch.f_slot = -1; }
                    else { const TableNode& tk = c->tnodes[c->tnode_of[k]]; ch.kind = 3; ch.slot = (int32_t)tk.rows_before; ch.f_slot = (int32_t)tk.D; }
                    sc.children.push_back(ch);
                }
                sc.steps.push_back(st);
                sc.gemm_nodes.push_back(t.node);
            }
            c->tsched_jobs.push_back(std::move(sc));
            c->tjobs.push_back(std::move(group));
        }
    }
    c->d_ids.reserve(std::max<size_t>(all_ids.size(), 1));
    CK(cudaMemcpy(c->d_ids.p, all_ids.data(), all_ids.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    c->tables_on = true;
}

void choose_tiling(cafe_b200_ctx* c)
{
    // rows per pass BM = 16*TM, TM in 8..13; minimise padded rows, then passes
    int bestTM = 11, bestTiles = 1 << 30, bestPad = 1 << 30;
    for (int tm = 8; tm <= 13; ++tm) {
        int bm = 16 * tm;
        int tiles = (c->N + bm - 1) / bm;
        int pad = tiles * bm;
        if (pad < bestPad || (pad == bestPad && tiles < bestTiles)) { bestPad = pad; bestTiles = tiles; bestTM = tm; }
    }
    c->TM = bestTM;
    c->n_mtiles = bestTiles;
    // +4: the arena's row stride equals the kernels' padded shared-memory stride (== 4 mod 16 doubles, conflict-free DMMA fragment
    // loads), so a whole stage of matrix rows is ONE contiguous, 128-byte aligned bulk copy instead of one copy per row
    c->LD = bestPad + 4;
}

// shared-memory footprint of prune_resident_kernel<TM, TNW, WN, BK> (ResidentCfg::smem_bytes, prune_resident.cuh)
size_t resident_stage_bytes(int tm, int bk) { return sizeof(double) * (size_t)bk * (16 * tm + 4); }
size_t resident_fixed_bytes(int tm, int tnw, int wn)
{
    const int bn = 8 * tnw * wn;
    const int parts = std::min(64 * wn / bn, 2);
    return sizeof(double) * ((size_t)bn * (16 * tm + 4) + 2 * parts * bn + 2 * 8 + 2);
}

// DMMA kernels: pick the variant and its column tile.  Wide tiles amortise matrix traffic, narrow tiles fill the SMs.
//   resident, WN = 2: two 128-thread CTAs per SM (BN = 16*TNW, BK = 4)   -- default when it fits
//   resident, WN = 4: one 256-thread CTA per SM (BN = 32*TNW, BK = 8)
//   stream          : both operands streamed (any state-space size)
int choose_columns_dmma(cafe_b200_ctx* c, int K)
{
    c->prune_kind = 1;
    if (c->prune_pref >= 2 && c->n_mtiles == 1) {
        const int wn_first = c->resident_wn;           // 2 (default) or 4 (CAFE_B200_RESIDENT_WN)
        for (int pass = 0; pass < 2 && c->prune_kind == 1; ++pass) {
            const int mode = pass == 0 ? wn_first : (wn_first == 4 ? 2 : 4);   // 2: two 128-thread CTAs per SM, 4: one 256-thread CTA
            const int wn = mode;
            const int bk = mode == 4 ? 8 : 4;
            const int ctas = mode == 4 ? 1 : 2;        // resident CTAs per SM
            const size_t avail = std::min<size_t>(c->smem_optin, c->smem_per_sm / ctas - 1024);
            int tnw = 4;
            while (tnw > 1) {
                int64_t tiles = ((c->U + 8 * tnw * wn - 1) / (8 * tnw * wn)) * K;
                if (tiles >= 2 * (int64_t)c->n_sms * ctas) break;
                tnw >>= 1;
            }
            if (const char* e = std::getenv("CAFE_B200_TNW")) { const int v = std::atoi(e); if (v >= 1 && v <= 4) tnw = v; }   // experiment knob
            const size_t fixed = resident_fixed_bytes(c->TM, tnw, wn), stage = resident_stage_bytes(c->TM, bk);
            if (fixed + 2 * stage > avail) continue;
            int stages = (int)std::min<size_t>((avail - fixed) / stage, 8);
            if (const char* e = std::getenv("CAFE_B200_STAGES")) stages = std::max(2, std::min(stages, std::atoi(e)));
            stages = std::max(2, std::min(stages, (c->S + bk - 1) / bk + 1));   // the prefill never reaches past the first matrix
            c->prune_kind = 2;
            c->TNW = tnw;
            c->WN = mode;
            c->dmma_stages = stages;
            const int bn = 8 * tnw * wn;
            c->n_col_tiles = (int)((c->U + bn - 1) / bn);
            c->grid = (int)std::min<int64_t>((int64_t)c->n_col_tiles * K, (int64_t)c->n_sms * ctas);
            return bn;
        }
    }
    int tnw = 4;
    while (tnw > 1) {
        int64_t tiles = ((c->U + 32 * tnw - 1) / (32 * tnw)) * K;
        if (tiles >= 2 * (int64_t)c->n_sms) break;
        tnw >>= 1;
    }
    if (const char* e = std::getenv("CAFE_B200_TNW")) { const int v = std::atoi(e); if (v == 1 || v == 2 || v == 4) tnw = v; }
    c->TNW = tnw;
    c->WN = 4;
    const int bn = 32 * tnw, bm = 16 * c->TM;
    const size_t stage = sizeof(double) * (size_t)DM_BK_HOST * (bm + 4 + bn + 4);
    const size_t tail = sizeof(double) * (2 * PRUNE_THREADS + 8);
    int stages = (int)((c->smem_optin - tail) / stage);
    c->dmma_stages = stages > 4 ? 4 : stages;
    if (c->dmma_stages < 2) throw CudaError{"RANGE: state space too large for the pruning pipeline"};
    c->n_col_tiles = (int)((c->U + bn - 1) / bn);
    int64_t tiles = (int64_t)c->n_col_tiles * K;
    c->grid = (int)std::min<int64_t>(tiles, c->n_sms);
    return bn;
}

// Column-tile width by problem size: wide tiles amortise matrix traffic, narrow tiles fill the SMs
void choose_columns(cafe_b200_ctx* c, int K)
{
    int tn = 4;
    while (tn > 1) {
        int64_t tiles = ((c->U + 16 * tn - 1) / (16 * tn)) * K;
        if (tiles >= 2 * (int64_t)c->n_sms) break;
        tn >>= 1;
    }
    c->TN = tn;
    int bn = 16 * tn;
    c->n_col_tiles = (int)((c->U + bn - 1) / bn);
    int64_t tiles = (int64_t)c->n_col_tiles * K;
    c->grid = (int)std::min<int64_t>(tiles, c->n_sms);
}

}  // namespace

namespace cafe {
// ---- key planning (matrix_cache_key, src/matrix_cache.h:44-63; lambda::multiply, src/lambda.h:45-48,76-84) ----
KeyPlan plan_keys(const cafe_b200_ctx* c, const double* lambdas, const double* multipliers, int K)
{
    KeyPlan kp;
    kp.mat_of.assign((size_t)K * c->n_nodes, 0);
    std::map<std::pair<long, long>, int> index;
    for (int k = 0; k < K; ++k) {
        for (int i = 0; i < c->n_nodes; ++i) {
            if (c->parent[i] < 0) continue;
            double lam = lambdas[c->lambda_class[i]] * multipliers[k];
            long kl = long(lam * 1000000000);
            long kt = long(c->branch_length[i] * 1000);
            auto key = std::make_pair(kl, kt);
            auto it = index.find(key);
            int id;
            if (it == index.end()) {
                id = (int)kp.params.size();
                index.emplace(key, id);
                double ql = double(kl) / 1000000000.0;
                double qt = double(kt) / 1000.0;
                double alpha = ql * qt / (1 + ql * qt);
                double coeff = 1 - 2 * alpha;
                MatParam mp{};
                mp.coeff = coeff;
                mp.zero = !(coeff > 0 && coeff != 1);   // covers is_saturated (coeff < 0)
                mp.log_alpha = mp.zero ? 0.0 : std::log(alpha);
                kp.params.push_back(mp);
            } else id = it->second;
            kp.mat_of[(size_t)k * c->n_nodes + i] = id;
        }
    }
    return kp;
}
}  // namespace cafe

namespace {

// steps / children / key index / contraction nodes (and, for a table launch, its jobs) -> the kernel-parameter block
bool fill_inline(InlineSchedule& s, const std::vector<Step>& steps, const std::vector<StepChild>& children, const std::vector<int32_t>& gemm_nodes,
                 const KeyPlan& kp, int n_nodes, const std::vector<int32_t>* jobs = nullptr)
{
    const size_t words = steps.size() * 9 + children.size() * 5 + kp.mat_of.size() + gemm_nodes.size() + (jobs ? jobs->size() : 0);
    s.valid = 0;
    s.n_jobs = 0; s.off_jobs = 0; s.n_job_tiles = 0; s.pad = 0;
    if (words > (size_t)SCHED_WORDS || kp.mat_of.size() != (kp.mat_of.size() / n_nodes) * (size_t)n_nodes) return false;
    int o = 0;
    for (const Step& st : steps) {
        const int32_t f[9] = {st.node, st.is_root, st.out_slot, st.n_children, st.child_begin, st.parent_step, st.carry_in, st.dst_kind, st.f_slot};
        memcpy(s.w + o, f, sizeof f);
        o += 9;
    }
    s.off_children = o;
    for (const StepChild& ch : children) {
        const int32_t f[5] = {ch.node, ch.leaf_row, ch.slot, ch.kind, ch.f_slot};
        memcpy(s.w + o, f, sizeof f);
        o += 5;
    }
    s.off_mat_of = o;
    memcpy(s.w + o, kp.mat_of.data(), kp.mat_of.size() * sizeof(int32_t));
    o += (int)kp.mat_of.size();
    s.off_gemm = o;
    if (!gemm_nodes.empty()) memcpy(s.w + o, gemm_nodes.data(), gemm_nodes.size() * sizeof(int32_t));
    o += (int)gemm_nodes.size();
    if (jobs) {
        s.off_jobs = o;
        s.n_jobs = (int)(jobs->size() / JOB_WORDS);
        memcpy(s.w + o, jobs->data(), jobs->size() * sizeof(int32_t));
    }
    s.valid = 1;
    return true;
}

void fill_inline_schedule(cafe_b200_ctx* c, const KeyPlan& kp)
{
    fill_inline(c->sched, c->steps, c->children, c->gemm_nodes, kp, c->n_nodes);
}

}  // namespace

namespace cafe {
void upload_plan(cafe_b200_ctx* c, const KeyPlan& kp)
{
    fill_inline_schedule(c, kp);
    c->last_kp = kp;                        // the table launches build their own kernel-parameter blocks from it
    size_t n_mats = kp.params.size();
    size_t arena = n_mats * (size_t)c->LD * c->LD;
    if (arena > c->d_arena.cap) {
        c->d_arena.reserve(arena + arena / 4);
        CK(cudaMemsetAsync(c->d_arena.p, 0, c->d_arena.cap * sizeof(double), c->stream));   // zero padding once
    }
    c->d_params.reserve(n_mats);
    c->d_mat_of.reserve(kp.mat_of.size());
    size_t b1 = n_mats * sizeof(MatParam), b2 = kp.mat_of.size() * sizeof(int32_t);
    CK(cudaStreamSynchronize(c->stream));   // the staging buffer may still feed a previous async copy
    char* h = (char*)c->stage(b1 + b2);
    memcpy(h, kp.params.data(), b1);
    memcpy(h + b1, kp.mat_of.data(), b2);
    CK(cudaMemcpyAsync(c->d_params.p, h, b1, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_mat_of.p, h + b1, b2, cudaMemcpyHostToDevice, c->stream));
}

void launch_matrices(cafe_b200_ctx* c, int n_mats)
{
    const size_t rows_smem = matrix_gen_rows_smem(c->N);
    if (c->matgen_entry || rows_smem > 200 * 1024) {   // state spaces whose tables do not fit shared memory (N > 2300) take the per-entry kernel
        dim3 grid(c->N, n_mats);
        matrix_gen_kernel<<<grid, 192, c->lg_n * sizeof(double), c->stream>>>(c->d_params.p, c->d_lg.p, c->lg_n, c->N, c->LD, c->d_arena.p);
        CK(cudaGetLastError());
        return;
    }
    c->d_powtab.reserve((size_t)n_mats * c->N);
    pow_table_kernel<<<(n_mats + 127) / 128, 128, 0, c->stream>>>(c->d_params.p, n_mats, c->N, c->d_powtab.p);
    CK(cudaGetLastError());
    dim3 grid((c->N + MG_S - 1) / MG_S, n_mats);
    if (rows_smem > 48 * 1024) {
        CK(allow_max_smem(matrix_gen_rows_kernel<true>));
        CK(allow_max_smem(matrix_gen_rows_kernel<false>));
    }
    if (c->matgen_libexp)
        matrix_gen_rows_kernel<true><<<grid, 192, matrix_gen_rows_smem(c->N), c->stream>>>(c->d_params.p, c->d_powtab.p, n_mats, c->d_lg.p, c->N,
                                                                                           c->LD, c->d_arena.p);
    else
        matrix_gen_rows_kernel<false><<<grid, 192, matrix_gen_rows_smem(c->N), c->stream>>>(c->d_params.p, c->d_powtab.p, n_mats, c->d_lg.p, c->N,
                                                                                            c->LD, c->d_arena.p);
    CK(cudaGetLastError());
}
}  // namespace cafe

namespace {

bool tables_active(const cafe_b200_ctx* c)
{
    return c->tables_on && c->use_dmma && c->prune_kind == 2 && c->WN == 2 && !c->resident_probe;
}

// Subtree-pattern reuse: the factor tables level by level (JOBS build of the resident kernel, one launch per group of table nodes),
// then the main pass over the reduced schedule.  Returns false (nothing launched) when a schedule does not fit the kernel-parameter
// block; the caller then runs the plain pass.
bool launch_with_tables(cafe_b200_ctx* c, const PruneParams& p)
{
    const int K = p.K, bn = 8 * c->TNW * 2;
    const KeyPlan& kp = c->last_kp;
    std::vector<InlineSchedule> launches(c->tsched_jobs.size());
    for (size_t g = 0; g < c->tsched_jobs.size(); ++g) {
        std::vector<int32_t> jobs;
        int tile = 0;
        for (int ti : c->tjobs[g]) {
            const TableNode& t = c->tnodes[ti];
            const int ncol = (int)((t.D + bn - 1) / bn);
            const int32_t w[JOB_WORDS] = {tile, ncol, (int32_t)t.D, (int32_t)t.D_stride, (int32_t)(uint32_t)(t.ids_off & 0xffffffff),
                                          (int32_t)(t.ids_off >> 32), 0, 0};
            jobs.insert(jobs.end(), w, w + JOB_WORDS);
            tile += ncol * K;
        }
        const Schedule& sc = c->tsched_jobs[g];
        if (!fill_inline(launches[g], sc.steps, sc.children, sc.gemm_nodes, kp, c->n_nodes, &jobs)) return false;
        launches[g].n_job_tiles = tile;
    }
    if (!fill_inline(c->tsched, c->tsched_main.steps, c->tsched_main.children, c->tsched_main.gemm_nodes, kp, c->n_nodes)) return false;
    c->d_tables.reserve((size_t)c->table_rows * K * c->LD, true);
    c->last_table_launches = 0;
    for (size_t g = 0; g < launches.size(); ++g) {
        PruneParams pj = p;
        pj.counts_t = c->d_ids.p;
        pj.tables = c->d_tables.p;
        pj.n_steps = (int)c->tsched_jobs[g].steps.size();
        pj.n_gemm = 1;
        pj.n_fslots = 1;
        const int grid = std::min(launches[g].n_job_tiles, 2 * c->n_sms);
        CK(launch_prune_resident_wn2_tables(c->TM, c->TNW, grid, c->N, c->dmma_stages, c->stream, pj, launches[g]));
        ++c->last_table_launches;
    }
    PruneParams pm = p;
    pm.tables = c->d_tables.p;
    pm.n_steps = (int)c->tsched_main.steps.size();
    pm.n_gemm = (int)c->tsched_main.gemm_nodes.size();
    pm.n_fslots = c->tsched_main.n_fslots;
    CK(launch_prune_resident_wn2(c->TM, c->TNW, c->grid, c->N, c->dmma_stages, c->stream, pm, c->tsched));
    return true;
}

void launch_any_prune(cafe_b200_ctx* c, PruneParams& p)
{
    c->last_table_launches = 0;
    if (!c->use_dmma) CK(launch_prune_dfma(c->TM, c->TN, c->grid, c->S, c->stream, p));
    else if (c->prune_kind == 2) {
        if (tables_active(c) && launch_with_tables(c, p)) return;
        if (c->WN == 2 && c->resident_probe) {
            c->d_probe.reserve((size_t)2 * 4 * (PROBE_CHUNKS * 4 + 4));
            CK(cudaMemsetAsync(c->d_probe.p, 0, c->d_probe.cap * sizeof(int64_t), c->stream));
            p.probe = c->d_probe.p;
            CK(launch_prune_resident_wn2probe(c->TM, c->TNW, c->grid, c->N, c->dmma_stages, c->stream, p, c->sched));
        } else if (c->WN == 2) CK(launch_prune_resident_wn2(c->TM, c->TNW, c->grid, c->N, c->dmma_stages, c->stream, p, c->sched));
        else CK(launch_prune_resident_wn4(c->TM, c->TNW, c->grid, c->N, c->dmma_stages, c->stream, p, c->sched));
    } else CK(launch_prune_stream(c->TM, c->TNW, c->grid, c->dmma_stages, c->stream, p));
}

PruneParams base_params(cafe_b200_ctx* c, int K, int mode)
{
    int bn;
    if (c->use_dmma) bn = choose_columns_dmma(c, K);
    else { choose_columns(c, K); bn = 16 * c->TN; }
    const int bm = 16 * c->TM;
    PruneParams p{};
    p.steps = c->d_steps.p;
    p.children = c->d_children.p;
    p.mat_of = c->d_mat_of.p;
    p.arena = c->d_arena.p;
    p.counts_t = c->d_counts_t.p;
    p.em = c->have_em ? c->d_em.p : nullptr;
    p.prior_d = c->d_prior.p;
    p.logprior = c->d_logprior.p;
    p.slot_stride = (int64_t)c->n_mtiles * bm * bn;
    const bool resident = c->use_dmma && c->prune_kind == 2;
    c->d_scratch.reserve((size_t)c->grid * (resident ? std::max(c->n_fslots, c->tsched_main.n_fslots) : c->n_slots) * p.slot_stride);
    p.scratch = c->d_scratch.p;
    p.gemm_nodes = c->d_gemm_nodes.p;
    p.n_gemm = (int)c->gemm_nodes.size();
    p.n_fslots = c->n_fslots;
    c->d_best.reserve((size_t)K * c->U_stride);
    c->d_ok.reserve((size_t)K * c->U_stride);
    p.out_best = c->d_best.p;
    p.out_ok = c->d_ok.p;
    p.out_roots = nullptr;
    p.zero_row = c->d_zero.p;
    p.U = c->U;
    p.U_stride = c->U_stride;
    p.n_steps = (int)c->steps.size();
    p.n_nodes = c->n_nodes;
    p.n_slots = c->n_slots;
    p.LD = c->LD; p.S = c->S; p.R = c->R; p.N = c->N; p.K = K;
    p.n_col_tiles = c->n_col_tiles;
    p.n_mtiles = c->n_mtiles;
    p.em_rows = c->em_rows;
    p.mode = mode;
    return p;
}

// ---- Pupko, second design (pupko2.cuh): table launches level by level, the main pass, the traceback; matrices are already there ----
void reconstruct_v2(cafe_b200_ctx* c, int K)
{
    const int n = c->n_nodes;
    const bool tabs = c->tables_on;
    const Schedule* main_sched = tabs ? &c->tsched_main : nullptr;
    if (!c->p2_ready) {
        const std::vector<Step>& st = tabs ? main_sched->steps : c->steps;
        const std::vector<StepChild>& ch = tabs ? main_sched->children : c->children;
        c->d_p2_steps.reserve(st.size());
        CK(cudaMemcpy(c->d_p2_steps.p, st.data(), st.size() * sizeof(Step), cudaMemcpyHostToDevice));
        c->d_p2_children.reserve(std::max<size_t>(ch.size(), 1));
        CK(cudaMemcpy(c->d_p2_children.p, ch.data(), ch.size() * sizeof(StepChild), cudaMemcpyHostToDevice));
        c->d_p2_parent.reserve(n); c->d_p2_leaf_col.reserve(n); c->d_p2_tab_of.reserve(n);
        CK(cudaMemcpy(c->d_p2_parent.p, c->parent.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_p2_leaf_col.p, c->leaf_col.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
        std::vector<int32_t> tab_of(n, -1);
        if (tabs) for (int v = 0; v < n; ++v) tab_of[v] = c->tnode_of[v];
        CK(cudaMemcpy(c->d_p2_tab_of.p, tab_of.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
        c->d_p2_tab_ids.reserve(std::max<size_t>(c->tab_ids_host.size(), 1));
        if (tabs && !c->tab_ids_host.empty())
            CK(cudaMemcpy(c->d_p2_tab_ids.p, c->tab_ids_host.data(), c->tab_ids_host.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        if (tabs) {
            c->d_p2_job_steps.resize(c->tsched_jobs.size());
            c->d_p2_job_children.resize(c->tsched_jobs.size());
            for (size_t g = 0; g < c->tsched_jobs.size(); ++g) {
                const Schedule& sc = c->tsched_jobs[g];
                c->d_p2_job_steps[g].reserve(sc.steps.size());
                CK(cudaMemcpy(c->d_p2_job_steps[g].p, sc.steps.data(), sc.steps.size() * sizeof(Step), cudaMemcpyHostToDevice));
                c->d_p2_job_children[g].reserve(sc.children.size());
                CK(cudaMemcpy(c->d_p2_job_children[g].p, sc.children.data(), sc.children.size() * sizeof(StepChild), cudaMemcpyHostToDevice));
            }
        }
        c->p2_ready = true;
    }
    // geometry: 512 threads need at least 32 columns; the tile must fit shared memory
    while (c->TN > 1 && pupko2_smem_bytes(c->TM, c->TN) > c->smem_optin) c->TN >>= 1;
    if (pupko2_smem_bytes(c->TM, c->TN) > c->smem_optin) throw CudaError{"RANGE: state space too large for the shared memory of the Pupko kernel"};
    const int bn = 16 * c->TN, bm = 16 * c->TM;
    c->n_col_tiles = (int)((c->U + bn - 1) / bn);
    c->grid = (int)std::min<int64_t>((int64_t)c->n_col_tiles * K, c->n_sms);
    const int SP = (c->S + 7) / 8 * 8;
    // argmax tables: one per internal non-root node, over its own columns
    std::vector<int64_t> coff(2 * (size_t)n, 0);
    int64_t total = 0;
    for (int v = 0; v < n - 1; ++v) {
        if (c->leaf_col[v] >= 0) continue;
        const int64_t cols = tabs && c->tnode_of[v] >= 0 ? c->tnodes[c->tnode_of[v]].D : c->U;
        coff[v] = total;
        coff[n + v] = cols;
        total += cols * SP * K;
    }
    c->d_p2_coff.reserve(coff.size());
    CK(cudaMemcpyAsync(c->d_p2_coff.p, coff.data(), coff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));            // coff is a stack vector
    c->d_p2_ctab.reserve((size_t)std::max<int64_t>(total, 1));
    c->d_p2_root.reserve((size_t)K * c->U_stride);
    const int n_fslots = tabs ? main_sched->n_fslots : c->n_fslots;
    Pupko2Params p{};
    p.mat_of = c->d_mat_of.p;
    p.arena = c->d_arena.p;
    p.prior_d = c->d_prior.p;
    p.slot_stride = (int64_t)bm * bn;
    c->d_scratch.reserve((size_t)c->grid * n_fslots * p.slot_stride);
    p.scratch = c->d_scratch.p;
    if (tabs) c->d_tables.reserve((size_t)c->table_rows * K * c->LD, true);
    p.tables = c->d_tables.p;
    p.ctab = c->d_p2_ctab.p;
    p.c_off = c->d_p2_coff.p;
    p.c_cols = c->d_p2_coff.p + n;
    p.root_state = c->d_p2_root.p;
    p.U = c->U; p.U_stride = c->U_stride;
    p.n_nodes = n; p.n_fslots = n_fslots;
    p.LD = c->LD; p.S = c->S; p.SP = SP; p.R = c->R; p.N = c->N; p.K = K;
    p.root_len = std::min(c->max_family_size, c->R) + 1;
    p.n_col_tiles = c->n_col_tiles;
    if (tabs) {
        std::vector<int32_t> jobs;
        std::vector<size_t> job_off;
        std::vector<int> job_tiles;
        for (size_t g = 0; g < c->tjobs.size(); ++g) {
            job_off.push_back(jobs.size());
            int tile = 0;
            for (int ti : c->tjobs[g]) {
                const TableNode& t = c->tnodes[ti];
                const int ncol = (int)((t.D + bn - 1) / bn);
                const int32_t w[JOB_WORDS] = {tile, ncol, (int32_t)t.D, (int32_t)t.D_stride, (int32_t)(uint32_t)(t.ids_off & 0xffffffff),
                                              (int32_t)(t.ids_off >> 32), 0, 0};
                jobs.insert(jobs.end(), w, w + JOB_WORDS);
                tile += ncol * K;
            }
            job_tiles.push_back(tile);
        }
        c->d_p2_jobs.reserve(std::max<size_t>(jobs.size(), 1));
        CK(cudaMemcpyAsync(c->d_p2_jobs.p, jobs.data(), jobs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (size_t g = 0; g < c->tjobs.size(); ++g) {
            Pupko2Params pj = p;
            pj.steps = c->d_p2_job_steps[g].p;
            pj.children = c->d_p2_job_children[g].p;
            pj.jobs = c->d_p2_jobs.p + job_off[g];
            pj.n_jobs = (int)c->tjobs[g].size();
            pj.n_job_tiles = job_tiles[g];
            pj.ids = c->d_ids.p;
            pj.n_steps = pj.n_jobs;
            CK(launch_pupko2(c->TM, c->TN, std::min(job_tiles[g], c->n_sms), c->stream, pj, true));
        }
    }
    p.steps = c->d_p2_steps.p;
    p.children = c->d_p2_children.p;
    p.ids = c->d_counts_t.p;
    p.n_steps = (int)(tabs ? main_sched->steps.size() : c->steps.size());
    CK(launch_pupko2(c->TM, c->TN, c->grid, c->stream, p, false));
    c->d_states.reserve((size_t)K * c->U_stride * n);
    CK(launch_pupko_traceback(c->stream, c->d_p2_ctab.p, c->d_p2_coff.p, c->d_p2_coff.p + n, c->d_p2_parent.p, c->d_p2_leaf_col.p,
                              c->d_p2_tab_of.p, c->d_p2_tab_ids.p, c->d_p2_root.p, c->U, c->U_stride, n, K, SP, c->d_states.p));
}

bool lambdas_valid(const double* lambdas, int n)
{
    // single_lambda::is_valid: lambda > 0 (lambda.h:58-60); multiple_lambda::is_valid: none < 0 (lambda.cpp:59-62)
    if (n == 1) return lambdas[0] > 0;
    for (int i = 0; i < n; ++i) if (lambdas[i] < 0) return false;
    return true;
}

// Enqueue one evaluation (matrices + prune + finish) on the stream.  Returns false (and sets the
// host-side infinite result) when the parameters are rejected before any kernel runs.
bool enqueue_eval(cafe_b200_ctx* c, const double* lambdas, int n_lambda, double alpha, const double* multipliers,
                  const double* cat_probs, int n_cat)
{
    if (!c->have_prior) throw CudaError{"STATE: set_prior must be called before eval"};
    if (n_lambda < c->n_lambda_classes) throw CudaError{"ARG: fewer lambdas than lambda classes in the tree"};
    c->stats_valid = false;
    c->last_launches = 0;
    c->last_mats = 0;
    const bool gamma = n_cat > 0;
    const int K = gamma ? n_cat : 1;
    static const double one = 1.0;
    if (!gamma) multipliers = &one;
    bool ok = lambdas_valid(lambdas, n_lambda);
    if (ok && gamma) {
        // gamma_model::can_infer (gamma_core.cpp:123-141)
        if (alpha < 0) ok = false;
        double maxlam = *std::max_element(lambdas, lambdas + n_lambda);
        double maxbl = 0;
        for (int i = 0; i < c->n_nodes; ++i) if (c->branch_length[i] > 0) maxbl = std::max(maxbl, c->branch_length[i]);
        double a = maxlam * maxbl / (1 + maxlam * maxbl);
        if ((1 - 2 * a) < 0) ok = false;
    }
    if (!ok) return false;

    KeyPlan kp = plan_keys(c, lambdas, multipliers, K);
    upload_plan(c, kp);
    const int n_mats = (int)kp.params.size();
    c->last_mats = n_mats;
    CK(cudaEventRecord(c->ev[0], c->stream));
    launch_matrices(c, n_mats);
    CK(cudaEventRecord(c->ev[1], c->stream));
    PruneParams p = base_params(c, K, gamma ? MODE_GAMMA : MODE_BASE);
    CK(cudaEventRecord(c->ev[2], c->stream));
    launch_any_prune(c, p);
    CK(cudaEventRecord(c->ev[3], c->stream));
    const int nb = (int)((c->F + FIN_THREADS - 1) / FIN_THREADS);
    c->d_partial.reserve(nb);
    c->d_partial_fail.reserve(nb);
    c->d_result.reserve(2);
    c->d_family_lnl.reserve(c->F);
    if (!gamma) {
        finish_base_kernel<<<nb, FIN_THREADS, 0, c->stream>>>(c->d_best.p, c->d_f2u.p, c->F, c->d_family_lnl.p, c->d_partial.p);
        CK(cudaGetLastError());
        if (c->F <= SEQ_SUM_LIMIT) final_sum_sequential_kernel<<<1, FIN_THREADS, 0, c->stream>>>(c->d_family_lnl.p, c->F, nullptr, 0, c->d_result.p);
        else final_sum_kernel<<<1, FIN_THREADS, 0, c->stream>>>(c->d_partial.p, nullptr, nb, c->d_result.p);
    } else {
        c->d_cat_lk.reserve((size_t)c->F * K);
        c->d_posterior.reserve((size_t)c->F * K);
        c->d_significant.reserve((size_t)c->F * K);
        c->d_family_lk.reserve(c->F);
        c->d_failed.reserve(c->F);
        c->d_cat_probs.reserve(K);
        CK(cudaMemcpyAsync(c->d_cat_probs.p, cat_probs, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        finish_gamma_kernel<<<nb, FIN_THREADS, 0, c->stream>>>(c->d_best.p, c->d_ok.p, c->U_stride, c->d_f2u.p, c->F, K,
                                                               c->d_cat_probs.p, c->d_cat_lk.p, c->d_family_lk.p, c->d_posterior.p,
                                                               c->d_significant.p, c->d_failed.p, c->d_family_lnl.p, c->d_partial.p, c->d_partial_fail.p);
        CK(cudaGetLastError());
        if (c->F <= SEQ_SUM_LIMIT)
            final_sum_sequential_kernel<<<1, FIN_THREADS, 0, c->stream>>>(c->d_family_lnl.p, c->F, c->d_partial_fail.p, nb, c->d_result.p);
        else final_sum_kernel<<<1, FIN_THREADS, 0, c->stream>>>(c->d_partial.p, c->d_partial_fail.p, nb, c->d_result.p);
    }
    CK(cudaGetLastError());
    c->last_launches = (c->matgen_entry ? 4 : 5) + c->last_table_launches;   // [pow table +] matrices, [factor tables,] pruning, finish, final sum
    c->stats_valid = true;
    return true;
}


}  // namespace

namespace cafe {
int fail(cafe_b200_ctx* c, const CudaError& e)
{
    int code = CAFE_B200_ERR_CUDA;
    std::string m = e.msg;
    if (m.rfind("ARG: ", 0) == 0) { code = CAFE_B200_ERR_ARG; m = m.substr(5); }
    else if (m.rfind("RANGE: ", 0) == 0) { code = CAFE_B200_ERR_RANGE; m = m.substr(7); }
    else if (m.rfind("STATE: ", 0) == 0) { code = CAFE_B200_ERR_STATE; m = m.substr(7); }
    if (c) c->err = m; else create_error() = m;
    return code;
}
}  // namespace cafe

namespace {

}  // namespace

// ---- cafe_b200_create_multi: a group context forwards every call to its device shards (defined at the end of this file) ----
namespace group {
int set_prior(cafe_b200_ctx* g, const float* prior, int32_t n);
int set_error_model(cafe_b200_ctx* g, const double* probs, int32_t rows, int32_t max_cnt);
int eval_base(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double* neg_lnl, double* family_lnl);
int eval_gamma(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double alpha, const double* multipliers, const double* cat_probs,
               int32_t n_cat, double* neg_lnl, double* cat_lk, double* family_lk, double* posterior, uint8_t* significant, uint8_t* failed,
               int64_t* n_failed);
int enqueue_eval(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double alpha, const double* multipliers, const double* cat_probs,
                 int32_t n_cat);
int fetch_result(cafe_b200_ctx* g, double* neg_lnl, int64_t* n_failed);
int last_stats(cafe_b200_ctx* g, int32_t* n_launches, int32_t* n_matrices, float* ms_matrices, float* ms_prune);
int root_vectors(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double multiplier, double* out);
int reconstruct(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, const double* multipliers, const double* cat_probs, int32_t n_cat,
                int32_t* cat_states, int32_t* states, double* averaged);
}  // namespace group

// ================================================================================================
extern "C" {

int cafe_b200_create(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                     int32_t max_family_size, int32_t max_root_family_size, int32_t device, cafe_b200_ctx** out)
{
    if (out) *out = nullptr;
    cafe_b200_ctx* c = nullptr;
    try {
        if (!tree || !counts || !out || n_families <= 0 || n_species <= 0 || max_family_size < 1 || max_root_family_size < 1)
            throw CudaError{"ARG: null or non-positive argument"};
        const int n = tree->n_nodes;
        if (n < 3 || !tree->parent || !tree->branch_length || !tree->leaf_col || !tree->lambda_class)
            throw CudaError{"ARG: tree needs at least a root and two children"};
        if (tree->parent[n - 1] != -1) throw CudaError{"ARG: the root must be the last node (reverse level order)"};
        int n_leaves = 0, max_class = 0;
        for (int i = 0; i < n; ++i) {
            if (i < n - 1 && (tree->parent[i] <= i || tree->parent[i] >= n))
                throw CudaError{"ARG: nodes must be in reverse level order (children before parents)"};
            if (tree->leaf_col[i] >= n_species) throw CudaError{"ARG: leaf_col out of range"};
            if (tree->leaf_col[i] >= 0) ++n_leaves;
            if (tree->lambda_class[i] < 0) throw CudaError{"ARG: negative lambda class"};
            max_class = std::max(max_class, tree->lambda_class[i]);
        }
        std::vector<int> nkids(n, 0);
        for (int i = 0; i < n - 1; ++i) nkids[tree->parent[i]]++;
        for (int i = 0; i < n; ++i) {
            if ((tree->leaf_col[i] >= 0) != (nkids[i] == 0)) throw CudaError{"ARG: leaf_col must be >= 0 exactly for childless nodes"};
            if (tree->leaf_col[i] >= 0 && i == n - 1) throw CudaError{"ARG: root cannot be a leaf"};
        }
        int dev_count = 0;
        if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
            throw CudaError{"no CUDA device available: libcafe_b200 has no CPU fallback"};
        if (device < 0 || device >= dev_count) throw CudaError{"ARG: device ordinal out of range"};

        c = new cafe_b200_ctx();
        c->device = device;
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        c->n_sms = prop.multiProcessorCount;
        c->smem_optin = prop.sharedMemPerBlockOptin;
        c->smem_per_sm = prop.sharedMemPerMultiprocessor;
        if (const char* e = std::getenv("CAFE_B200_RESIDENT_WN")) { int v = std::atoi(e); c->resident_wn = v == 4 ? 4 : 2; }
        if (const char* e = std::getenv("CAFE_B200_RESIDENT_PROBE")) c->resident_probe = std::atoi(e) != 0;
        if (const char* e = std::getenv("CAFE_B200_PUPKO_THREADS")) c->pupko_threads = std::atoi(e) == 256 ? 256 : 512;
        if (const char* e = std::getenv("CAFE_B200_PUPKO")) c->pupko_version = std::atoi(e) == 1 ? 1 : 2;
        if (const char* e = std::getenv("CAFE_B200_MATGEN")) { c->matgen_entry = std::strcmp(e, "entry") == 0; c->matgen_libexp = std::strcmp(e, "rows") == 0; }
        if (const char* e = std::getenv("CAFE_B200_PRUNE")) {
            const std::string v(e);
            c->use_dmma = v != "dfma";
            c->prune_pref = v == "stream" ? 1 : 2;
        }
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto& e : c->ev) CK(cudaEventCreate(&e));
        CK(cudaMallocHost(&c->h_result, 2 * sizeof(double)));

        c->n_nodes = n;
        c->parent.assign(tree->parent, tree->parent + n);
        c->leaf_col.assign(tree->leaf_col, tree->leaf_col + n);
        c->lambda_class.assign(tree->lambda_class, tree->lambda_class + n);
        c->branch_length.assign(tree->branch_length, tree->branch_length + n);
        c->n_lambda_classes = max_class + 1;
        c->max_family_size = max_family_size;
        c->S = max_family_size + 1;
        c->R = max_root_family_size;
        c->N = std::max(max_root_family_size, max_family_size) + 1;   // base_model.cpp:65, gamma_core.cpp:185
        choose_tiling(c);
        // the narrowest Pupko geometry (16 columns) must fit: M_v tile + matrix stages + the traceback states of every internal node
        if (pupko_smem_bytes(c->TM, 1, c->S, (n - n_leaves)) > (size_t)prop.sharedMemPerBlockOptin)
            throw CudaError{"RANGE: max_family_size (with this many internal nodes) too large for the shared memory of the Pupko kernel"};

        // ---- families: validate, build the reference list (identical count vectors pruned once) ----
        c->F = n_families;
        c->n_species = n_species;
        c->max_count.assign((size_t)n_families, 0);
        std::vector<int> leaf_nodes;
        c->leaf_row_of_node.assign(n, -1);
        for (int i = 0; i < n; ++i)
            if (c->leaf_col[i] >= 0) { c->leaf_row_of_node[i] = (int)leaf_nodes.size(); leaf_nodes.push_back(i); }
        std::unordered_map<std::string, int64_t> seen;
        seen.reserve((size_t)n_families * 2);
        c->f2u.resize(n_families);
        std::vector<int64_t> uniq;
        std::string key((size_t)n_leaves * sizeof(int32_t), '\0');
        for (int64_t f = 0; f < n_families; ++f) {
            const int32_t* row = counts + (size_t)f * n_species;
            for (int j = 0; j < n_leaves; ++j) {
                int32_t v = row[c->leaf_col[leaf_nodes[j]]];
                if (v < 0 || v > max_family_size) throw CudaError{"RANGE: a count is negative or exceeds max_family_size"};
                c->max_count[f] = std::max(c->max_count[f], v);
                memcpy(&key[(size_t)j * sizeof(int32_t)], &v, sizeof v);
            }
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (int64_t)uniq.size()).first;
                uniq.push_back(f);
            }
            c->f2u[f] = it->second;
        }
        c->U = (int64_t)uniq.size();
        c->U_stride = (c->U + 63) / 64 * 64;
        // Column order of the distinct families: by total count (stable), not by first appearance.  A family's values do not depend on
        // its column, but a tile of 64 neighbouring columns then holds families of similar size: their leaf gathers hit the same few
        // matrix rows, the table plan numbers the patterns of a node in nearly sorted order (ids are given in column order), and
        // neighbouring columns gather neighbouring table rows.  CAFE_B200_SORT=0 keeps the order of first appearance.
        {
            const char* e = std::getenv("CAFE_B200_SORT");
            if (!(e && std::strcmp(e, "0") == 0) && c->U > 1) {
                std::vector<int64_t> total((size_t)c->U, 0), rank((size_t)c->U), by((size_t)c->U);
                for (int64_t u = 0; u < c->U; ++u) {
                    const int32_t* row = counts + (size_t)uniq[u] * n_species;
                    for (int j = 0; j < n_leaves; ++j) total[u] += row[c->leaf_col[leaf_nodes[j]]];
                    by[u] = u;
                }
                std::stable_sort(by.begin(), by.end(), [&](int64_t a, int64_t b) { return total[a] < total[b]; });
                std::vector<int64_t> sorted((size_t)c->U);
                for (int64_t pos = 0; pos < c->U; ++pos) { sorted[pos] = uniq[by[pos]]; rank[by[pos]] = pos; }
                uniq.swap(sorted);
                for (int64_t f = 0; f < n_families; ++f) c->f2u[f] = rank[c->f2u[f]];
            }
        }
        std::vector<int32_t> counts_t((size_t)n_leaves * c->U_stride, 0);
        for (int64_t u = 0; u < c->U; ++u) {
            const int32_t* row = counts + (size_t)uniq[u] * n_species;
            for (int j = 0; j < n_leaves; ++j) counts_t[(size_t)j * c->U_stride + u] = row[c->leaf_col[leaf_nodes[j]]];
        }
        plan_tables(c, counts_t, n_leaves);       // may append pattern-id rows for the table nodes the main pass gathers
        c->d_counts_t.reserve(counts_t.size());
        CK(cudaMemcpy(c->d_counts_t.p, counts_t.data(), counts_t.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        c->d_f2u.reserve(n_families);
        CK(cudaMemcpy(c->d_f2u.p, c->f2u.data(), (size_t)n_families * sizeof(int64_t), cudaMemcpyHostToDevice));

        c->d_zero.reserve(256, true);
        {
            Schedule full;
            build_schedule(c, std::vector<char>(), full);
            c->steps = std::move(full.steps);
            c->children = std::move(full.children);
            c->gemm_nodes = std::move(full.gemm_nodes);
            c->n_slots = full.n_slots;
            c->n_fslots = full.n_fslots;
        }
        c->d_steps.reserve(c->steps.size());
        CK(cudaMemcpy(c->d_steps.p, c->steps.data(), c->steps.size() * sizeof(Step), cudaMemcpyHostToDevice));
        c->d_children.reserve(c->children.size());
        CK(cudaMemcpy(c->d_children.p, c->children.data(), c->children.size() * sizeof(StepChild), cudaMemcpyHostToDevice));
        c->d_gemm_nodes.reserve(std::max<size_t>(c->gemm_nodes.size(), 1));
        if (!c->gemm_nodes.empty())
            CK(cudaMemcpy(c->d_gemm_nodes.p, c->gemm_nodes.data(), c->gemm_nodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));

        // lgamma table with the HOST libm, the same values the reference caches (probability.cpp:69-80)
        c->lg_n = 2 * c->N + 2;
        std::vector<double> lg(c->lg_n);
        for (int i = 0; i < c->lg_n; ++i) lg[i] = std::lgamma((double)i);
        c->d_lg.reserve(lg.size());
        CK(cudaMemcpy(c->d_lg.p, lg.data(), lg.size() * sizeof(double), cudaMemcpyHostToDevice));
        *out = c;
        return CAFE_B200_OK;
    } catch (const CudaError& e) {
        int code = fail(nullptr, e);
        if (c) cafe_b200_destroy(c);
        return code;
    } catch (const std::exception& e) {
        create_error() = e.what();
        if (c) cafe_b200_destroy(c);
        return CAFE_B200_ERR_ARG;
    }
}

int cafe_b200_destroy(cafe_b200_ctx* c)
{
    if (!c) return CAFE_B200_OK;
    if (c->is_group()) {
        delete c->pool;
        if (c->g_scratch) cudaFreeHost(c->g_scratch);
        for (cafe_b200_ctx* s : c->shards) if (s) cafe_b200_destroy(s);
        delete c;
        return CAFE_B200_OK;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->d_counts_t.release(); c->d_mat_of.release(); c->d_gemm_nodes.release(); c->d_f2u.release(); c->d_steps.release(); c->d_children.release();
    c->d_zero.release(); c->d_lg.release(); c->d_arena.release(); c->d_scratch.release(); c->d_prior.release(); c->d_logprior.release();
    c->d_em.release(); c->d_best.release(); c->d_cat_probs.release(); c->d_ok.release(); c->d_params.release(); c->d_powtab.release(); c->d_probe.release();
    c->d_family_lnl.release(); c->d_cat_lk.release(); c->d_family_lk.release(); c->d_posterior.release();
    c->d_partial.release(); c->d_partial_fail.release(); c->d_result.release(); c->d_roots.release();
    c->d_significant.release(); c->d_failed.release(); c->d_pupko_m.release(); c->d_states.release();
    c->d_leaf_row.release(); c->d_states_f.release(); c->d_cat_states_f.release(); c->d_avg_f.release();
    c->d_ids.release(); c->d_tables.release();
    c->d_p2_steps.release(); c->d_p2_children.release(); c->d_p2_jobs.release(); c->d_p2_tab_of.release(); c->d_p2_tab_ids.release();
    c->d_p2_parent.release(); c->d_p2_leaf_col.release(); c->d_p2_root.release(); c->d_p2_coff.release(); c->d_p2_ctab.release();
    for (auto& b : c->d_p2_job_steps) b.release();
    for (auto& b : c->d_p2_job_children) b.release();
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_result) cudaFreeHost(c->h_result);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CAFE_B200_OK;
}

const char* cafe_b200_last_error(const cafe_b200_ctx* c) { return c ? c->err.c_str() : create_error().c_str(); }

int cafe_b200_set_prior(cafe_b200_ctx* c, const float* prior, int32_t n)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::set_prior(c, prior, n);
    try {
        if (!prior || n < 0) throw CudaError{"ARG: null prior"};
        CK(cudaSetDevice(c->device));
        c->prior.assign(prior, prior + n);
        // Inference weights root index j by compute(j), j < R (base_model.cpp:84, gamma_core.cpp:156);
        // Pupko uses compute(j) for j < min(max_family_size, R) + 1 (gene_family_reconstructor.cpp:65,146).
        const int len = std::max(c->R, std::min(c->max_family_size, c->R) + 1);
        std::vector<double> pd(len), lp(len);
        for (int j = 0; j < len; ++j) {
            pd[j] = j < n ? (double)prior[j] : 0.0;           // float widened, as `double eq_freq = prior.compute(j)`
            lp[j] = std::log(pd[j]);                          // host libm, as std::log(eq_freq)
        }
        c->d_prior.reserve(len);
        c->d_logprior.reserve(len);
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemcpy(c->d_prior.p, pd.data(), len * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_logprior.p, lp.data(), len * sizeof(double), cudaMemcpyHostToDevice));
        c->have_prior = true;
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_set_error_model(cafe_b200_ctx* c, const double* probs, int32_t rows, int32_t max_cnt)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::set_error_model(c, probs, rows, max_cnt);
    try {
        CK(cudaSetDevice(c->device));
        if (!probs) { c->have_em = false; c->em_host.clear(); return CAFE_B200_OK; }
        if (rows < 1) throw CudaError{"ARG: error model needs at least one row"};
        c->em_host.assign(probs, probs + (size_t)rows * 3);
        c->d_em.reserve((size_t)rows * 3);
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemcpy(c->d_em.p, probs, (size_t)rows * 3 * sizeof(double), cudaMemcpyHostToDevice));
        c->em_rows = rows;
        c->em_maxcnt = max_cnt;
        c->have_em = true;
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_eval_base(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, double* neg_lnl, double* family_lnl)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::eval_base(c, lambdas, n_lambda, neg_lnl, family_lnl);
    try {
        if (!lambdas || n_lambda < 1 || !neg_lnl) throw CudaError{"ARG: null argument"};
        CK(cudaSetDevice(c->device));
        if (!enqueue_eval(c, lambdas, n_lambda, 0.0, nullptr, nullptr, 0)) {
            *neg_lnl = std::numeric_limits<double>::infinity();
            return CAFE_B200_OK;
        }
        d2h(c, c->h_result, c->d_result.p, 2);
        d2h(c, family_lnl, c->d_family_lnl.p, (size_t)c->F);
        CK(cudaStreamSynchronize(c->stream));
        *neg_lnl = c->h_result[0];
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_eval_gamma(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, double alpha,
                         const double* multipliers, const double* cat_probs, int32_t n_cat,
                         double* neg_lnl, double* cat_lk, double* family_lk, double* posterior,
                         uint8_t* significant, uint8_t* failed, int64_t* n_failed)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::eval_gamma(c, lambdas, n_lambda, alpha, multipliers, cat_probs, n_cat, neg_lnl, cat_lk, family_lk, posterior, significant, failed, n_failed);
    try {
        if (!lambdas || n_lambda < 1 || !neg_lnl || !multipliers || !cat_probs || n_cat < 1) throw CudaError{"ARG: null argument"};
        CK(cudaSetDevice(c->device));
        if (n_failed) *n_failed = 0;
        if (!enqueue_eval(c, lambdas, n_lambda, alpha, multipliers, cat_probs, n_cat)) {
            *neg_lnl = std::numeric_limits<double>::infinity();
            return CAFE_B200_OK;
        }
        const size_t FK = (size_t)c->F * n_cat;
        d2h(c, c->h_result, c->d_result.p, 2);
        d2h(c, cat_lk, c->d_cat_lk.p, FK);
        d2h(c, family_lk, c->d_family_lk.p, (size_t)c->F);
        d2h(c, posterior, c->d_posterior.p, FK);
        d2h(c, significant, c->d_significant.p, FK);
        d2h(c, failed, c->d_failed.p, (size_t)c->F);
        CK(cudaStreamSynchronize(c->stream));
        *neg_lnl = c->h_result[0];
        if (n_failed) *n_failed = (int64_t)c->h_result[1];
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_enqueue_eval(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, double alpha,
                           const double* multipliers, const double* cat_probs, int32_t n_cat)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::enqueue_eval(c, lambdas, n_lambda, alpha, multipliers, cat_probs, n_cat);
    try {
        CK(cudaSetDevice(c->device));
        if (!enqueue_eval(c, lambdas, n_lambda, alpha, multipliers, cat_probs, n_cat))
            throw CudaError{"ARG: parameters rejected before launch (invalid lambda / alpha / saturated)"};
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_fetch_result(cafe_b200_ctx* c, double* neg_lnl, int64_t* n_failed)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::fetch_result(c, neg_lnl, n_failed);
    try {
        CK(cudaSetDevice(c->device));
        d2h(c, c->h_result, c->d_result.p, 2);
        CK(cudaStreamSynchronize(c->stream));
        if (neg_lnl) *neg_lnl = c->h_result[0];
        if (n_failed) *n_failed = (int64_t)c->h_result[1];
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

void* cafe_b200_result_device(cafe_b200_ctx* c)
{
    if (!c) return nullptr;
    if (c->is_group()) return cafe_b200_result_device(c->shards[0]);
    try {
        CK(cudaSetDevice(c->device));
        c->d_result.reserve(2);
        return (void*)c->d_result.p;
    } catch (const CudaError& e) { fail(c, e); return nullptr; }
}

void* cafe_b200_stream(cafe_b200_ctx* c) { return !c ? nullptr : c->is_group() ? (void*)c->shards[0]->stream : (void*)c->stream; }

int cafe_b200_last_stats(cafe_b200_ctx* c, int32_t* n_launches, int32_t* n_matrices, float* ms_matrices, float* ms_prune)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::last_stats(c, n_launches, n_matrices, ms_matrices, ms_prune);
    try {
        if (!c->stats_valid) throw CudaError{"STATE: no evaluation has been launched"};
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
        CK(cudaEventElapsedTime(&b, c->ev[2], c->ev[3]));
        if (n_launches) *n_launches = c->last_launches;
        if (n_matrices) *n_matrices = c->last_mats;
        if (ms_matrices) *ms_matrices = a;
        if (ms_prune) *ms_prune = b;
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int64_t cafe_b200_unique_families(const cafe_b200_ctx* c)
{
    if (!c) return 0;
    if (!c->is_group()) return c->U;
    int64_t u = 0;
    for (const cafe_b200_ctx* s : c->shards) u += s->U;   // identical families in different shards are pruned once per shard
    return u;
}

int cafe_b200_node_columns(const cafe_b200_ctx* c, int64_t* columns)
{
    if (!c || !columns) return CAFE_B200_ERR_ARG;
    for (int v = 0; v < c->n_nodes; ++v) columns[v] = 0;
    const std::vector<cafe_b200_ctx*> one{const_cast<cafe_b200_ctx*>(c)};
    for (const cafe_b200_ctx* s : c->is_group() ? c->shards : one) {
        const bool tab = s->tables_on && s->use_dmma && s->prune_pref >= 2 && s->resident_wn == 2 && !s->resident_probe;
        for (int v = 0; v < s->n_nodes; ++v) {
            if (s->leaf_col[v] >= 0) continue;
            columns[v] += tab && s->tnode_of[v] >= 0 ? s->tnodes[s->tnode_of[v]].D : s->U;
        }
    }
    return CAFE_B200_OK;
}

int32_t cafe_b200_n_devices(const cafe_b200_ctx* c) { return !c ? 0 : c->is_group() ? (int32_t)c->shards.size() : 1; }

int cafe_b200_describe(const cafe_b200_ctx* c, int64_t* n_families, int32_t* n_nodes, int32_t* n_lambda_classes,
                       int32_t* max_family_size, int32_t* max_root_family_size, double* longest_branch)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (n_families) *n_families = c->F;
    if (n_nodes) *n_nodes = c->n_nodes;
    if (n_lambda_classes) *n_lambda_classes = c->n_lambda_classes;
    if (max_family_size) *max_family_size = c->max_family_size;
    if (max_root_family_size) *max_root_family_size = c->R;
    if (longest_branch) *longest_branch = *std::max_element(c->branch_length.begin(), c->branch_length.end());
    return CAFE_B200_OK;
}

int cafe_b200_debug_read_probe(cafe_b200_ctx* c, int64_t* out, int64_t n)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return cafe_b200_debug_read_probe(c->shards[0], out, n);
    try {
        if (!out || n < 0) throw CudaError{"ARG: bad argument"};
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        const size_t m = std::min<size_t>((size_t)n, c->d_probe.cap);
        if (m) CK(cudaMemcpy(out, c->d_probe.p, m * sizeof(int64_t), cudaMemcpyDeviceToHost));
        for (size_t i = m; i < (size_t)n; ++i) out[i] = 0;
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_host_alloc(size_t bytes, void** out)
{
    try {
        if (!out) throw CudaError{"ARG: null output"};
        *out = nullptr;
        CK(cudaMallocHost(out, bytes ? bytes : 1));
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(nullptr, e); }
}

int cafe_b200_host_free(void* p)
{
    try {
        if (p) CK(cudaFreeHost(p));
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(nullptr, e); }
}

int cafe_b200_measure_fp64_peak(int32_t device, int32_t use_dmma, double* tflops)
{
    try {
        if (!tflops) throw CudaError{"ARG: null output"};
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        double* d = nullptr;
        CK(cudaMalloc(&d, 64));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        // bits 8..15 of use_dmma: warps per SM for an issue-rate probe (one CTA per SM); 0 = saturating default
        const int probe_warps = (use_dmma >> 8) & 0xff;
        use_dmma &= 0xff;
        const int iters = 4096;
        const int blocks = probe_warps ? prop.multiProcessorCount : prop.multiProcessorCount * 8;
        const int threads = probe_warps ? probe_warps * 32 : 256;
        double best = 0.0;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaEventRecord(e0));
            if (use_dmma == 2) dmma_peak_kernel_k8<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            else if (use_dmma == 3) dmma_peak_kernel_k16<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            else if (use_dmma) dmma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            else dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            // DFMA: 16 fma per thread-iteration; DMMA: 8 tiles x (8*8*4) fma per warp-iteration
            // m16n8k8: 4 tiles x 1024 fma; m16n8k16: 4 tiles x 2048 fma
            double fma_count = use_dmma == 2 ? (double)blocks * (threads / 32) * iters * 4.0 * 1024.0
                             : use_dmma == 3 ? (double)blocks * (threads / 32) * iters * 4.0 * 2048.0
                             : use_dmma ? (double)blocks * (threads / 32) * iters * 8.0 * 256.0
                                        : (double)blocks * threads * iters * 16.0;
            double tf = 2.0 * fma_count / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best) best = tf;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(d);
        *tflops = best;
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(nullptr, e); }
}
int32_t cafe_b200_matrix_size(const cafe_b200_ctx* c) { return c ? c->N : 0; }

int cafe_b200_get_matrix(cafe_b200_ctx* c, double lambda, double branch_length, double* out)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) {
        for (cafe_b200_ctx* s : c->shards)          // a bucket may hold a smaller state space: take a shard with the full one
            if (s->N == c->N) return cafe_b200_get_matrix(s, lambda, branch_length, out);
        c->err = "no bucket holds the full state space";
        return CAFE_B200_ERR_STATE;
    }
    try {
        if (!out) throw CudaError{"ARG: null output"};
        CK(cudaSetDevice(c->device));
        // one-key plan through the same quantisation as an evaluation
        cafe_b200_ctx tmp_tree;   // only the fields plan_keys reads
        tmp_tree.n_nodes = 2;
        tmp_tree.parent = {1, -1};
        tmp_tree.lambda_class = {0, 0};
        tmp_tree.branch_length = {branch_length, 0.0};
        const double one = 1.0;
        KeyPlan kp = plan_keys(&tmp_tree, &lambda, &one, 1);
        // NB: matrix_cache::get_matrix does not refuse lambda <= 0; neither do we.
        kp.mat_of.assign((size_t)c->n_nodes, 0);
        upload_plan(c, kp);
        launch_matrices(c, 1);
        std::vector<double> pt((size_t)c->LD * c->LD);
        CK(cudaMemcpyAsync(pt.data(), c->d_arena.p, pt.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int s = 0; s < c->N; ++s)
            for (int ch = 0; ch < c->N; ++ch) out[(size_t)s * c->N + ch] = pt[(size_t)ch * c->LD + s];
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_root_vectors(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, double multiplier, double* out)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::root_vectors(c, lambdas, n_lambda, multiplier, out);
    try {
        if (!lambdas || !out || n_lambda < c->n_lambda_classes) throw CudaError{"ARG: bad argument"};
        if (!c->have_prior) throw CudaError{"STATE: set_prior must be called before eval"};
        CK(cudaSetDevice(c->device));
        KeyPlan kp = plan_keys(c, lambdas, &multiplier, 1);
        upload_plan(c, kp);
        launch_matrices(c, (int)kp.params.size());
        PruneParams p = base_params(c, 1, MODE_ROOTS);
        c->d_roots.reserve((size_t)c->U * c->R);
        p.out_roots = c->d_roots.p;
        launch_any_prune(c, p);
        std::vector<double> roots((size_t)c->U * c->R);
        CK(cudaMemcpyAsync(roots.data(), c->d_roots.p, roots.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int64_t f = 0; f < c->F; ++f)
            memcpy(out + (size_t)f * c->R, roots.data() + (size_t)c->f2u[f] * c->R, c->R * sizeof(double));
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_reconstruct(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda,
                          const double* multipliers, const double* cat_probs, int32_t n_cat,
                          int32_t* cat_states, int32_t* states, double* averaged)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) return group::reconstruct(c, lambdas, n_lambda, multipliers, cat_probs, n_cat, cat_states, states, averaged);
    try {
        if (!lambdas || n_lambda < c->n_lambda_classes || !states) throw CudaError{"ARG: bad argument"};
        if (n_cat > 0 && (!multipliers || !cat_probs)) throw CudaError{"ARG: gamma reconstruction needs multipliers and cat_probs"};
        if (!c->have_prior) throw CudaError{"STATE: set_prior must be called before reconstruct"};
        CK(cudaSetDevice(c->device));
        const int K = n_cat > 0 ? n_cat : 1;
        static const double one = 1.0;
        const double* mult = n_cat > 0 ? multipliers : &one;
        KeyPlan kp = plan_keys(c, lambdas, mult, K);
        upload_plan(c, kp);
        launch_matrices(c, (int)kp.params.size());

        choose_columns(c, K);
        if (c->pupko_version == 2 && c->n_mtiles == 1) reconstruct_v2(c, K);
        else {
        // the column tile must also fit shared memory (M_v tile + stages + traceback states of every step): narrow it until it does
        while (c->TN > 1 && pupko_smem_bytes(c->TM, c->TN, c->S, (int)c->steps.size()) > c->smem_optin) {
            c->TN >>= 1;
            c->n_col_tiles = (int)((c->U + 16 * c->TN - 1) / (16 * c->TN));
            c->grid = (int)std::min<int64_t>((int64_t)c->n_col_tiles * K, c->n_sms);
        }
        if (pupko_smem_bytes(c->TM, c->TN, c->S, (int)c->steps.size()) > c->smem_optin)
            throw CudaError{"RANGE: state space too large for the shared memory of the Pupko kernel"};
        PupkoParams p{};
        const int bn = 16 * c->TN, bm = 16 * c->TM;
        p.steps = c->d_steps.p;
        p.children = c->d_children.p;
        p.mat_of = c->d_mat_of.p;
        p.arena = c->d_arena.p;
        p.counts_t = c->d_counts_t.p;
        p.prior_d = c->d_prior.p;
        p.slot_stride = (int64_t)c->n_mtiles * bm * bn;
        c->d_scratch.reserve((size_t)c->grid * c->n_slots * p.slot_stride);
        p.scratch = c->d_scratch.p;
        p.n_steps = (int)c->steps.size();
        p.m_stride = (int64_t)((c->S + PRUNE_BK - 1) / PRUNE_BK * PRUNE_BK) * bn;
        c->d_pupko_m.reserve((size_t)c->grid * p.n_steps * p.m_stride);
        p.mstore = c->d_pupko_m.p;
        c->d_states.reserve((size_t)K * c->U_stride * c->n_nodes);
        p.states = c->d_states.p;
        p.U = c->U; p.U_stride = c->U_stride;
        p.n_nodes = c->n_nodes; p.n_slots = c->n_slots;
        p.LD = c->LD; p.S = c->S; p.R = c->R; p.N = c->N; p.K = K;
        p.root_len = std::min(c->max_family_size, c->R) + 1;
        p.n_col_tiles = c->n_col_tiles; p.n_mtiles = c->n_mtiles;
        CK(launch_pupko(c->TM, c->TN, c->grid, c->S, c->stream, p, c->pupko_threads));
        }
        const int n = c->n_nodes;
        const size_t Fn = (size_t)c->F * n;
        c->d_cat_probs.reserve(K);
        static const double one_prob = 1.0;
        CK(cudaMemcpyAsync(c->d_cat_probs.p, n_cat > 0 ? cat_probs : &one_prob, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c->d_leaf_row.reserve(n);
        // rows of the count table for the LEAVES only: leaf_row_of_node also names the pattern-id rows of the table nodes the main
        // pruning pass gathers (internal nodes), which must keep their reconstructed state
        std::vector<int32_t> true_leaf_row(n);
        for (int i = 0; i < n; ++i) true_leaf_row[i] = c->leaf_col[i] >= 0 ? c->leaf_row_of_node[i] : -1;
        CK(cudaMemcpyAsync(c->d_leaf_row.p, true_leaf_row.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));            // true_leaf_row is a stack vector
        c->d_states_f.reserve(Fn);
        if (cat_states) c->d_cat_states_f.reserve(Fn * K);
        if (averaged) c->d_avg_f.reserve(Fn);
        const int nb = (int)((Fn + 255) / 256);
        expand_states_kernel<<<nb, 256, 0, c->stream>>>(c->d_states.p, c->d_f2u.p, c->d_counts_t.p, c->d_leaf_row.p, c->d_cat_probs.p,
                                                        c->F, c->U_stride, n, K, n_cat > 0 ? 1 : 0,
                                                        cat_states ? c->d_cat_states_f.p : nullptr, c->d_states_f.p,
                                                        averaged ? c->d_avg_f.p : nullptr);
        CK(cudaGetLastError());
        d2h(c, states, c->d_states_f.p, Fn);
        d2h(c, cat_states, c->d_cat_states_f.p, Fn * K);
        d2h(c, averaged, c->d_avg_f.p, Fn);
        CK(cudaStreamSynchronize(c->stream));
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

}  // extern "C"

// ================================================================================================
// Clustered, work-balanced shards (strong scaling).  Every shard plans its own subtree-pattern tables, and a random
// block of the job shares fewer patterns than the whole job does (tools/probes/shard_clustering_study.py: 45 % of the columns left at 8
// random shards against 35 % for the whole job).  Families with similar counts share the most patterns, so the job is ordered by
// total count and cut into contiguous blocks of that order; blocks of large families keep more distinct patterns per family, so the
// cut points are moved until every block costs the same number of contraction columns under the table plan.
namespace {

// contraction columns per category the table plan leaves for the families idx[0 .. n): sum over the non-root internal nodes of
// (patterns of a table node | distinct families otherwise) -- the rule of plan_tables, evaluated on the host without a context
int64_t plan_cost(const cafe_b200_tree* tree, const int32_t* counts, int n_species, const int64_t* idx, int64_t n)
{
    const int nn = tree->n_nodes;
    const char* mode = std::getenv("CAFE_B200_TABLES");
    const bool off = mode && std::strcmp(mode, "0") == 0, force = mode && std::strcmp(mode, "force") == 0;
    double frac = force ? 1.0 : 0.75;
    if (const char* e = std::getenv("CAFE_B200_TABLE_FRAC")) frac = std::atof(e);
    std::vector<int> leaves;
    int n_internal = 0;
    for (int v = 0; v < nn; ++v) { if (tree->leaf_col[v] >= 0) leaves.push_back(v); else if (v != nn - 1) ++n_internal; }
    // distinct families
    std::unordered_map<std::string, int32_t> seen;
    seen.reserve((size_t)n * 2);
    std::vector<int64_t> uniq;
    std::string key(leaves.size() * sizeof(int32_t), '\0');
    for (int64_t i = 0; i < n; ++i) {
        const int32_t* row = counts + (size_t)idx[i] * n_species;
        for (size_t j = 0; j < leaves.size(); ++j) memcpy(&key[j * sizeof(int32_t)], &row[tree->leaf_col[leaves[j]]], sizeof(int32_t));
        if (seen.emplace(key, (int32_t)uniq.size()).second) uniq.push_back(idx[i]);
    }
    const int64_t U = (int64_t)uniq.size();
    if (off || (!force && U < 8192)) return U * n_internal;
    std::vector<std::vector<int>> kids(nn);
    for (int i = nn - 1; i >= 0; --i) if (tree->parent[i] >= 0) kids[tree->parent[i]].push_back(i);
    std::vector<std::vector<int32_t>> ids(nn);
    std::vector<char> is_table(nn, 0);
    auto id_of = [&](int v, int64_t u) -> int32_t {
        return tree->leaf_col[v] >= 0 ? counts[(size_t)uniq[u] * n_species + tree->leaf_col[v]] : ids[v][u];
    };
    int64_t cost = 0;
    const int64_t limit = (int64_t)std::floor(frac * (double)U);
    for (int v = 0; v < nn - 1; ++v) {
        if (tree->leaf_col[v] >= 0) continue;
        bool ok = true;
        for (int k : kids[v]) if (tree->leaf_col[k] < 0 && !is_table[k]) { ok = false; break; }
        int64_t D = 0;
        bool over = !ok;
        if (ok) {
            std::vector<int32_t> id(U);
            const int m = (int)kids[v].size();
            if (m == 2) {
                std::unordered_map<uint64_t, int32_t> pat;
                for (int64_t u = 0; u < U; ++u) {
                    const uint64_t kk = ((uint64_t)(uint32_t)id_of(kids[v][0], u) << 32) | (uint32_t)id_of(kids[v][1], u);
                    auto it = pat.find(kk);
                    if (it == pat.end()) {
                        if (D >= limit) { over = true; break; }
                        it = pat.emplace(kk, (int32_t)D++).first;
                    }
                    id[u] = it->second;
                }
            } else {
                std::unordered_map<std::string, int32_t> pat;
                std::string pk((size_t)m * sizeof(int32_t), '\0');
                for (int64_t u = 0; u < U; ++u) {
                    for (int j = 0; j < m; ++j) { const int32_t x = id_of(kids[v][j], u); memcpy(&pk[(size_t)j * sizeof x], &x, sizeof x); }
                    auto it = pat.find(pk);
                    if (it == pat.end()) {
                        if (D >= limit) { over = true; break; }
                        it = pat.emplace(pk, (int32_t)D++).first;
                    }
                    id[u] = it->second;
                }
            }
            if (!over && D > 0) { is_table[v] = 1; ids[v] = std::move(id); }
        }
        cost += is_table[v] ? D : U;
        for (int k : kids[v]) if (tree->leaf_col[k] < 0) std::vector<int32_t>().swap(ids[k]);     // children's ids are no longer needed
    }
    return cost;
}

}  // namespace

extern "C" int cafe_b200_plan_shards(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                                     int32_t n_shards, int64_t* order, int64_t* bounds)
{
    if (!tree || !counts || n_families <= 0 || n_species <= 0 || n_shards < 1 || !order || !bounds) return CAFE_B200_ERR_ARG;
    for (int i = 0; i <= n_shards; ++i) bounds[i] = n_families;        // more shards than families: the extra shards are empty
    n_shards = (int32_t)std::min<int64_t>(n_shards, n_families);
    std::vector<int64_t> total((size_t)n_families, 0);
    for (int64_t f = 0; f < n_families; ++f)
        for (int j = 0; j < n_species; ++j) total[f] += counts[(size_t)f * n_species + j];
    for (int64_t f = 0; f < n_families; ++f) order[f] = f;
    std::stable_sort(order, order + n_families, [&](int64_t a, int64_t b) { return total[a] < total[b]; });
    std::vector<double> size((size_t)n_shards, (double)n_families / n_shards);
    auto set_bounds = [&]() {
        double acc = 0;
        bounds[0] = 0;
        for (int i = 0; i < n_shards; ++i) {
            acc += size[i];
            bounds[i + 1] = i + 1 == n_shards ? n_families : std::max<int64_t>(bounds[i] + 1, std::min<int64_t>((int64_t)std::llround(acc), n_families - (n_shards - 1 - i)));
        }
    };
    set_bounds();
    if (n_shards == 1) return CAFE_B200_OK;
    for (int iter = 0; iter < 3; ++iter) {
        std::vector<int64_t> cost((size_t)n_shards, 0);
        std::vector<std::thread> workers;
        for (int i = 0; i < n_shards; ++i)
            workers.emplace_back([&, i] { cost[i] = plan_cost(tree, counts, n_species, order + bounds[i], bounds[i + 1] - bounds[i]); });
        for (auto& w : workers) w.join();
        // a block's cost per family is taken as constant while its size changes a little: sizes proportional to 1 / (cost per family)
        double norm = 0;
        std::vector<double> inv((size_t)n_shards);
        for (int i = 0; i < n_shards; ++i) { inv[i] = (double)(bounds[i + 1] - bounds[i]) / std::max<double>((double)cost[i], 1.0); norm += inv[i]; }
        const int64_t cmax = *std::max_element(cost.begin(), cost.end()), cmin = *std::min_element(cost.begin(), cost.end());
        if ((double)(cmax - cmin) <= 0.01 * (double)cmax) break;          // within 1 %: balanced
        for (int i = 0; i < n_shards; ++i) size[i] = 0.5 * size[i] + 0.5 * (double)n_families * inv[i] / norm;   // damped
        set_bounds();
    }
    return CAFE_B200_OK;
}

// ================================================================================================
// cafe_b200_create_multi: families sharded contiguously over the devices of one node.  Every group call runs the
// single-device entry point of all shards at once on the shard workers; per-family outputs land in the caller's buffers at the
// shard's offset; the scalar partials are added on the host in shard order.
extern "C" int cafe_b200_create_multi(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                                      int32_t max_family_size, int32_t max_root_family_size, const int32_t* devices, int32_t n_devices,
                                      cafe_b200_ctx** out)
{
    if (out) *out = nullptr;
    if (!devices || n_devices < 1 || !out) { create_error() = "null device list"; return CAFE_B200_ERR_ARG; }
    if (n_devices == 1) return cafe_b200_create(tree, counts, n_families, n_species, max_family_size, max_root_family_size, devices[0], out);
    if (!tree || !counts || n_families <= 0 || n_species <= 0) { create_error() = "null or non-positive argument"; return CAFE_B200_ERR_ARG; }
    const int n_shards = (int)std::min<int64_t>(n_devices, n_families);
    cafe_b200_ctx* g = new cafe_b200_ctx();
    g->shards.assign(n_shards, nullptr);
    g->shard_begin.resize(n_shards + 1);
    for (int i = 0; i <= n_shards; ++i) g->shard_begin[i] = n_families * i / n_shards;
    // Large jobs: clustered, work-balanced shards (cafe_b200_plan_shards) instead of blocks in the caller's order, so that the
    // per-shard subtree-pattern tables keep most of the whole job's reuse.  CAFE_B200_CLUSTER=0 keeps the caller's order.
    std::vector<int32_t> packed;
    const char* cl = std::getenv("CAFE_B200_CLUSTER");
    if (n_families >= (int64_t)8192 * n_shards && !(cl && std::strcmp(cl, "0") == 0)) {
        g->order.resize((size_t)n_families);
        if (cafe_b200_plan_shards(tree, counts, n_families, n_species, n_shards, g->order.data(), g->shard_begin.data()) != CAFE_B200_OK) {
            delete g;
            create_error() = "shard planning failed";
            return CAFE_B200_ERR_ARG;
        }
        // inside a shard the families keep the caller's relative order: the scatter of the per-family outputs then walks the caller's
        // arrays front to back instead of hopping at random
        for (int i = 0; i < n_shards; ++i) std::sort(g->order.begin() + g->shard_begin[i], g->order.begin() + g->shard_begin[i + 1]);
        packed.resize((size_t)n_families * n_species);
        for (int64_t p = 0; p < n_families; ++p)
            memcpy(&packed[(size_t)p * n_species], counts + (size_t)g->order[p] * n_species, n_species * sizeof(int32_t));
    }
    const int32_t* src = packed.empty() ? counts : packed.data();
    g->pool = new ShardPool(n_shards);
    std::vector<std::string> errs(n_shards);
    int bad = -1;
    const int rc = g->pool->run([&](int i) {
        const int64_t b = g->shard_begin[i], e = g->shard_begin[i + 1];
        const int r = cafe_b200_create(tree, src + (size_t)b * n_species, e - b, n_species, max_family_size, max_root_family_size,
                                       devices[i], &g->shards[i]);
        if (r != CAFE_B200_OK) errs[i] = cafe_b200_last_error(nullptr);   // create's error text is per thread
        return r;
    }, &bad);
    if (rc != CAFE_B200_OK) {
        create_error() = "device " + std::to_string(devices[bad]) + ": " + errs[bad];
        cafe_b200_destroy(g);
        return rc;
    }
    const cafe_b200_ctx* s0 = g->shards[0];
    g->device = s0->device;
    g->F = n_families;
    g->n_species = n_species;
    g->n_nodes = s0->n_nodes;
    g->n_lambda_classes = s0->n_lambda_classes;
    g->max_family_size = s0->max_family_size;
    g->S = s0->S; g->R = s0->R; g->N = s0->N;
    g->branch_length = s0->branch_length;
    g->parent = s0->parent; g->leaf_col = s0->leaf_col; g->lambda_class = s0->lambda_class;
    g->max_count.resize((size_t)n_families);
    for (int i = 0; i < n_shards; ++i)
        for (int64_t j = 0; j < g->shards[i]->F; ++j) {
            const int64_t pos = g->shard_begin[i] + j;
            g->max_count[(size_t)(g->order.empty() ? pos : g->order[pos])] = g->shards[i]->max_count[j];
        }
    *out = g;
    return CAFE_B200_OK;
}

// Bucketed mode (north_star: pruning "batched over families bucketed by max family size"; SURVEY.md 7 "No truncation in inference"):
// opt-in, labelled, never the default.  A family whose largest count is x is pruned over the states 0 .. m(x),
// m(x) = min(max_family_size, x + max(50, x / 5)) -- the truncation the reference itself applies to the simulated families of its p-value
// path (compute_family_probabilities, src/probability.cpp:394,416) -- rounded up to the next ceiling of `state_ceilings`; every bucket is
// a context of its own with max_family_size = its ceiling and max_root_family_size = min(R, ceiling), all on one device, run
// concurrently on their own streams.  Results differ from the exact path by the probability mass beyond the ceiling.
extern "C" int cafe_b200_create_bucketed(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                                         int32_t max_family_size, int32_t max_root_family_size, const int32_t* state_ceilings,
                                         int32_t n_ceilings, int32_t device, cafe_b200_ctx** out)
{
    if (out) *out = nullptr;
    if (!tree || !counts || !out || n_families <= 0 || n_species <= 0 || !state_ceilings || n_ceilings < 1) {
        create_error() = "null or non-positive argument";
        return CAFE_B200_ERR_ARG;
    }
    std::vector<int32_t> ceil(state_ceilings, state_ceilings + n_ceilings);
    std::sort(ceil.begin(), ceil.end());
    while (!ceil.empty() && ceil.back() >= max_family_size) ceil.pop_back();
    ceil.push_back(max_family_size);                               // the last bucket is the exact state space
    if (ceil.front() < 1) { create_error() = "state ceilings must be positive"; return CAFE_B200_ERR_ARG; }
    std::vector<std::vector<int64_t>> members(ceil.size());
    for (int64_t f = 0; f < n_families; ++f) {
        int32_t x = 0;
        for (int j = 0; j < n_species; ++j) {
            const int32_t v = counts[(size_t)f * n_species + j];
            if (v < 0 || v > max_family_size) { create_error() = "a count is negative or exceeds max_family_size"; return CAFE_B200_ERR_RANGE; }
            x = std::max(x, v);
        }
        const int32_t m = std::min(max_family_size, x + std::max(50, x / 5));
        members[std::lower_bound(ceil.begin(), ceil.end(), m) - ceil.begin()].push_back(f);
    }
    cafe_b200_ctx* g = new cafe_b200_ctx();
    std::vector<int32_t> used;
    std::vector<int32_t> packed;
    g->shard_begin.push_back(0);
    for (size_t b = 0; b < ceil.size(); ++b) {
        if (members[b].empty()) continue;
        used.push_back(ceil[b]);
        for (int64_t f : members[b]) {
            g->order.push_back(f);
            packed.insert(packed.end(), counts + (size_t)f * n_species, counts + (size_t)(f + 1) * n_species);
        }
        g->shard_begin.push_back((int64_t)g->order.size());
    }
    const int n_shards = (int)used.size();
    g->shards.assign(n_shards, nullptr);
    g->pool = new ShardPool(n_shards);
    std::vector<std::string> errs(n_shards);
    int bad = -1;
    const int rc = g->pool->run([&](int i) {
        const int64_t b = g->shard_begin[i], e = g->shard_begin[i + 1];
        const int r = cafe_b200_create(tree, packed.data() + (size_t)b * n_species, e - b, n_species, used[i],
                                       std::min(max_root_family_size, used[i]), device, &g->shards[i]);
        if (r != CAFE_B200_OK) errs[i] = cafe_b200_last_error(nullptr);
        return r;
    }, &bad);
    if (rc != CAFE_B200_OK) {
        create_error() = "bucket with ceiling " + std::to_string(used[bad]) + ": " + errs[bad];
        cafe_b200_destroy(g);
        return rc;
    }
    const cafe_b200_ctx* s0 = g->shards[0];
    g->device = device;
    g->F = n_families;
    g->n_species = n_species;
    g->n_nodes = s0->n_nodes;
    g->n_lambda_classes = s0->n_lambda_classes;
    g->max_family_size = max_family_size;
    g->S = max_family_size + 1; g->R = max_root_family_size; g->N = std::max(max_root_family_size, max_family_size) + 1;
    g->branch_length = s0->branch_length;
    g->parent = s0->parent; g->leaf_col = s0->leaf_col; g->lambda_class = s0->lambda_class;
    g->max_count.resize((size_t)n_families);
    for (int i = 0; i < n_shards; ++i)
        for (int64_t j = 0; j < g->shards[i]->F; ++j) g->max_count[(size_t)g->order[g->shard_begin[i] + j]] = g->shards[i]->max_count[j];
    *out = g;
    return CAFE_B200_OK;
}

namespace group {

// run fn on every shard; on failure the group's error text is the failing shard's
int each(cafe_b200_ctx* g, const std::function<int(int, cafe_b200_ctx*)>& fn)
{
    int bad = -1;
    const int rc = g->pool->run([&](int i) { return fn(i, g->shards[i]); }, &bad);
    if (rc != CAFE_B200_OK) g->err = "device " + std::to_string(g->shards[bad]->device) + ": " + g->shards[bad]->err;
    return rc;
}

// -sum over families, shards added in shard order (base_model.cpp:95, gamma_core.cpp:233); +inf when any shard rejected
double combine(const std::vector<double>& neg)
{
    double total = 0.0;
    for (double v : neg) total += v;
    return total;
}

// A per-family output of the caller (width values per family).  Contiguous shards write straight into it at their offset; shards that
// are not contiguous in the caller's order (buckets, clustered shards: g->order non-empty) write shard-major into page-locked scratch
// of the group and every shard worker scatters its own rows back to the caller's family order.
template <typename T>
struct Out {
    cafe_b200_ctx* g;
    T* user;
    size_t width;
    T* tmp = nullptr;
    Out(cafe_b200_ctx* g_, T* user_, size_t width_) : g(g_), user(user_), width(width_)
    {
        if (user && !g->order.empty()) {
            tmp = (T*)g->take_scratch((size_t)g->F * width * sizeof(T));
            if (!tmp) throw CudaError{"group scratch exhausted"};
        }
    }
    T* at(int shard) { return !user ? nullptr : (tmp ? tmp : user) + (size_t)g->shard_begin[shard] * width; }
    // rows [b, e) of the scratch -> the caller's order.  Called by the shard's worker; large shards split the rows over a few helper
    // threads (the copy is latency-bound: one thread moves ~5 GB/s of 1 ... 32-byte rows)
    void scatter(int shard)
    {
        if (!tmp) return;
        const int64_t b = g->shard_begin[shard], e = g->shard_begin[shard + 1];
        const int64_t rows = e - b;
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int helpers = (int)std::min<int64_t>(std::max<int64_t>(1, (int64_t)hw / (int64_t)g->shards.size()), std::min<int64_t>(8, rows / 32768));
        if (helpers <= 1) { scatter_rows(b, e); return; }
        std::vector<std::thread> th;
        for (int t = 1; t < helpers; ++t) th.emplace_back([=] { scatter_rows(b + rows * t / helpers, b + rows * (t + 1) / helpers); });
        scatter_rows(b, b + rows / helpers);
        for (auto& x : th) x.join();
    }
    void scatter_rows(int64_t b, int64_t e)
    {
        const int64_t* __restrict__ ord = g->order.data();
        const size_t bytes = width * sizeof(T);
        if (bytes == 1) {
            const uint8_t* s8 = (const uint8_t*)tmp; uint8_t* d8 = (uint8_t*)user;
            for (int64_t p = b; p < e; ++p) d8[ord[p]] = s8[p];
        } else if (bytes % 8 == 0 && bytes <= 64) {      // a family's row of doubles: word copies instead of a memcpy call per row
            const size_t w = bytes / 8;
            const uint64_t* s64 = (const uint64_t*)tmp; uint64_t* d64 = (uint64_t*)user;
            for (int64_t p = b; p < e; ++p) {
                const uint64_t* sp = s64 + (size_t)p * w; uint64_t* dp = d64 + (size_t)ord[p] * w;
                for (size_t i = 0; i < w; ++i) dp[i] = sp[i];
            }
        } else if (bytes <= 8) {
            const uint8_t* s8 = (const uint8_t*)tmp; uint8_t* d8 = (uint8_t*)user;
            for (int64_t p = b; p < e; ++p)
                for (size_t i = 0; i < bytes; ++i) d8[(size_t)ord[p] * bytes + i] = s8[(size_t)p * bytes + i];
        } else {
            for (int64_t p = b; p < e; ++p) memcpy(user + (size_t)ord[p] * width, tmp + (size_t)p * width, bytes);
        }
    }
};

// makes sure the group's scratch holds `bytes` (page-locked; contents live for one call) and resets it
void reset_scratch(cafe_b200_ctx* g, size_t bytes)
{
    g->g_scratch_used = 0;
    if (g->order.empty() || bytes <= g->g_scratch_cap) return;
    if (g->g_scratch) cudaFreeHost(g->g_scratch);
    g->g_scratch = nullptr;
    g->g_scratch_cap = 0;
    CK(cudaMallocHost(&g->g_scratch, bytes + 4096));
    g->g_scratch_cap = bytes + 4096;
}

int set_prior(cafe_b200_ctx* g, const float* prior, int32_t n)
{
    const int rc = each(g, [&](int, cafe_b200_ctx* s) { return cafe_b200_set_prior(s, prior, n); });
    if (rc == CAFE_B200_OK) { g->prior.assign(prior, prior + n); g->have_prior = true; }
    return rc;
}

int set_error_model(cafe_b200_ctx* g, const double* probs, int32_t rows, int32_t max_cnt)
{
    const int rc = each(g, [&](int, cafe_b200_ctx* s) { return cafe_b200_set_error_model(s, probs, rows, max_cnt); });
    if (rc == CAFE_B200_OK) {
        g->have_em = probs != nullptr;
        g->em_rows = rows; g->em_maxcnt = max_cnt;
        if (probs) g->em_host.assign(probs, probs + (size_t)rows * 3); else g->em_host.clear();
    }
    return rc;
}

int eval_base(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double* neg_lnl, double* family_lnl)
{
    if (!neg_lnl) { g->err = "null argument"; return CAFE_B200_ERR_ARG; }
    std::vector<double> neg(g->shards.size(), 0.0);
    try {
        reset_scratch(g, family_lnl ? (size_t)g->F * sizeof(double) + 256 : 0);
        Out<double> fam(g, family_lnl, 1);
        const int rc = each(g, [&](int i, cafe_b200_ctx* s) {
            const int r = cafe_b200_eval_base(s, lambdas, n_lambda, &neg[i], fam.at(i));
            if (r == CAFE_B200_OK && !std::isinf(neg[i])) fam.scatter(i);
            return r;
        });
        if (rc == CAFE_B200_OK) *neg_lnl = combine(neg);
        return rc;
    } catch (const CudaError& e) { return fail(g, e); }
}

int eval_gamma(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double alpha, const double* multipliers, const double* cat_probs,
               int32_t n_cat, double* neg_lnl, double* cat_lk, double* family_lk, double* posterior, uint8_t* significant, uint8_t* failed,
               int64_t* n_failed)
{
    if (!neg_lnl || n_cat < 1) { g->err = "null argument"; return CAFE_B200_ERR_ARG; }
    const size_t n = g->shards.size();
    std::vector<double> neg(n, 0.0);
    std::vector<int64_t> nf(n, 0);
    try {
        reset_scratch(g, (size_t)g->F * ((cat_lk ? n_cat * 8 : 0) + (family_lk ? 8 : 0) + (posterior ? n_cat * 8 : 0) + (significant ? n_cat : 0) +
                                         (failed ? 1 : 0)) + 5 * 256);
        Out<double> o_cat(g, cat_lk, n_cat), o_fam(g, family_lk, 1), o_post(g, posterior, n_cat);
        Out<uint8_t> o_sig(g, significant, n_cat), o_fail(g, failed, 1);
        const int rc = each(g, [&](int i, cafe_b200_ctx* s) {
            const int r = cafe_b200_eval_gamma(s, lambdas, n_lambda, alpha, multipliers, cat_probs, n_cat, &neg[i], o_cat.at(i), o_fam.at(i),
                                               o_post.at(i), o_sig.at(i), o_fail.at(i), &nf[i]);
            if (r == CAFE_B200_OK && !(std::isinf(neg[i]) && nf[i] == 0)) {     // (rejected before launch: nothing was written)
                o_cat.scatter(i); o_fam.scatter(i); o_post.scatter(i); o_sig.scatter(i); o_fail.scatter(i);
            }
            return r;
        });
        if (rc != CAFE_B200_OK) return rc;
        *neg_lnl = combine(neg);      // a shard with a failed family reports +inf, and so does the sum (gamma_core.cpp:216-225)
        if (n_failed) { *n_failed = 0; for (int64_t v : nf) *n_failed += v; }
        return rc;
    } catch (const CudaError& e) { return fail(g, e); }
}

int enqueue_eval(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double alpha, const double* multipliers, const double* cat_probs,
                 int32_t n_cat)
{
    return each(g, [&](int, cafe_b200_ctx* s) { return cafe_b200_enqueue_eval(s, lambdas, n_lambda, alpha, multipliers, cat_probs, n_cat); });
}

int fetch_result(cafe_b200_ctx* g, double* neg_lnl, int64_t* n_failed)
{
    const size_t n = g->shards.size();
    std::vector<double> neg(n, 0.0);
    std::vector<int64_t> nf(n, 0);
    const int rc = each(g, [&](int i, cafe_b200_ctx* s) { return cafe_b200_fetch_result(s, &neg[i], &nf[i]); });
    if (rc != CAFE_B200_OK) return rc;
    if (neg_lnl) *neg_lnl = combine(neg);
    if (n_failed) { *n_failed = 0; for (int64_t v : nf) *n_failed += v; }
    return rc;
}

int last_stats(cafe_b200_ctx* g, int32_t* n_launches, int32_t* n_matrices, float* ms_matrices, float* ms_prune)
{
    const size_t n = g->shards.size();
    std::vector<int32_t> nl(n, 0), nm(n, 0);
    std::vector<float> a(n, 0.f), b(n, 0.f);
    const int rc = each(g, [&](int i, cafe_b200_ctx* s) { return cafe_b200_last_stats(s, &nl[i], &nm[i], &a[i], &b[i]); });
    if (rc != CAFE_B200_OK) return rc;
    if (n_launches) { *n_launches = 0; for (int32_t v : nl) *n_launches += v; }       // kernels launched on all devices
    if (n_matrices) *n_matrices = nm[0];                                               // per device (every device regenerates them)
    if (ms_matrices) *ms_matrices = *std::max_element(a.begin(), a.end());             // slowest device
    if (ms_prune) *ms_prune = *std::max_element(b.begin(), b.end());
    return rc;
}

// out[F x R]; a bucket computes root sizes 1 .. R_b only (R_b <= R): the remaining entries are 0
int root_vectors(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, double multiplier, double* out)
{
    if (!out) { g->err = "bad argument"; return CAFE_B200_ERR_ARG; }
    const size_t R = (size_t)g->R;
    bool same = g->order.empty();
    for (const cafe_b200_ctx* s : g->shards) same = same && s->R == g->R;
    if (same) return each(g, [&](int i, cafe_b200_ctx* s) { return cafe_b200_root_vectors(s, lambdas, n_lambda, multiplier, out + (size_t)g->shard_begin[i] * R); });
    std::vector<std::vector<double>> part(g->shards.size());
    const int rc = each(g, [&](int i, cafe_b200_ctx* s) {
        part[i].resize((size_t)s->F * s->R);
        return cafe_b200_root_vectors(s, lambdas, n_lambda, multiplier, part[i].data());
    });
    if (rc != CAFE_B200_OK) return rc;
    for (size_t i = 0; i < g->shards.size(); ++i) {
        const size_t Rb = (size_t)g->shards[i]->R;
        for (int64_t j = 0; j < g->shards[i]->F; ++j) {
            const int64_t pos = g->shard_begin[i] + j;
            double* dst = out + (size_t)(g->order.empty() ? pos : g->order[pos]) * R;
            memcpy(dst, part[i].data() + (size_t)j * Rb, Rb * sizeof(double));
            std::fill(dst + Rb, dst + R, 0.0);
        }
    }
    return rc;
}

int reconstruct(cafe_b200_ctx* g, const double* lambdas, int32_t n_lambda, const double* multipliers, const double* cat_probs, int32_t n_cat,
                int32_t* cat_states, int32_t* states, double* averaged)
{
    if (!states) { g->err = "bad argument"; return CAFE_B200_ERR_ARG; }
    const size_t K = n_cat > 0 ? n_cat : 1, nn = (size_t)g->n_nodes;
    try {
        reset_scratch(g, (size_t)g->F * nn * ((cat_states ? K * 4 : 0) + 4 + (averaged ? 8 : 0)) + 3 * 256);
        Out<int32_t> o_cat(g, cat_states, K * nn), o_st(g, states, nn);
        Out<double> o_avg(g, averaged, nn);
        return each(g, [&](int i, cafe_b200_ctx* s) {
            const int r = cafe_b200_reconstruct(s, lambdas, n_lambda, multipliers, cat_probs, n_cat, o_cat.at(i), o_st.at(i), o_avg.at(i));
            if (r == CAFE_B200_OK) { o_cat.scatter(i); o_st.scatter(i); o_avg.scatter(i); }
            return r;
        });
    } catch (const CudaError& e) { return fail(g, e); }
}

}  // namespace group
