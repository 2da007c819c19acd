// abi_analysis.cu -- the analysis-side entry points of libcafe_b200.so built on top of the likelihood path: the simulator
// (SURVEY.md 8f row f4), the per-branch change probabilities of the report and the family-level Monte-Carlo p-values (row f2).
#include "context.cuh"
#include "simulate.cuh"
#include "viterbi.cuh"

using namespace cafe;

extern "C" {

int cafe_b200_simulate(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, const double* multipliers, const double* cat_probs,
                       int32_t n_cat, int32_t max_sim, int32_t max_redraws, const int32_t* root_sizes, int64_t n_families, uint64_t seed,
                       int32_t* counts, int32_t* node_sizes, int32_t* categories, int64_t* n_not_at_root)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) {   // the simulator needs the tree and the matrices only: first device
        const int rc = cafe_b200_simulate(c->shards[0], lambdas, n_lambda, multipliers, cat_probs, n_cat, max_sim, max_redraws, root_sizes,
                                          n_families, seed, counts, node_sizes, categories, n_not_at_root);
        if (rc != CAFE_B200_OK) c->err = c->shards[0]->err;
        return rc;
    }
    try {
        if (!lambdas || n_lambda < c->n_lambda_classes || !root_sizes || n_families <= 0 || !counts) throw CudaError{"ARG: bad argument"};
        if (n_cat > 0 && (!multipliers || !cat_probs)) throw CudaError{"ARG: gamma simulation needs multipliers and cat_probs"};
        if (max_sim < 2 || max_sim > c->N) throw CudaError{"RANGE: max_sim must be in [2, matrix size]"};
        CK(cudaSetDevice(c->device));
        const int K = n_cat > 0 ? n_cat : 1;
        static const double one = 1.0;
        const double* mult = n_cat > 0 ? multipliers : &one;
        const double* probs = n_cat > 0 ? cat_probs : &one;
        KeyPlan kp = plan_keys(c, lambdas, mult, K);
        upload_plan(c, kp);
        const int n_mats = (int)kp.params.size();
        launch_matrices(c, n_mats);
        const int n = c->n_nodes;
        const size_t F = (size_t)n_families;
        DevBuf<double> d_cdf, d_probs;
        DevBuf<int32_t> d_parent, d_leaf_col, d_root, d_sizes, d_counts, d_cat;
        DevBuf<uint8_t> d_has;
        DevBuf<unsigned long long> d_exh;
        struct Release {   // scratch of one call: freed on every exit path
            DevBuf<double>&a, &b; DevBuf<int32_t>&c1, &c2, &c3, &c4, &c5, &c6; DevBuf<uint8_t>& d; DevBuf<unsigned long long>& e;
            ~Release() { a.release(); b.release(); c1.release(); c2.release(); c3.release(); c4.release(); c5.release(); c6.release(); d.release(); e.release(); }
        } release{d_cdf, d_probs, d_parent, d_leaf_col, d_root, d_sizes, d_counts, d_cat, d_has, d_exh};
        d_cdf.reserve((size_t)n_mats * max_sim * c->N);
        d_probs.reserve(K);
        d_parent.reserve(n); d_leaf_col.reserve(n);
        d_root.reserve(F); d_sizes.reserve(F * n); d_counts.reserve(F * c->n_species); d_cat.reserve(F);
        d_has.reserve(F * n);
        d_exh.reserve(1, true);
        CK(cudaMemcpyAsync(d_probs.p, probs, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_parent.p, c->parent.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_leaf_col.p, c->leaf_col.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_root.p, root_sizes, F * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(d_counts.p, 0, F * c->n_species * sizeof(int32_t), c->stream));
        sim_cdf_kernel<<<n_mats, 256, 0, c->stream>>>(c->d_arena.p, c->LD, c->N, max_sim, d_cdf.p);
        CK(cudaGetLastError());
        SimParams sp{};
        sp.parent = d_parent.p; sp.leaf_col = d_leaf_col.p; sp.mat_of = c->d_mat_of.p; sp.cdf = d_cdf.p; sp.cat_probs = d_probs.p;
        sp.root_sizes = d_root.p; sp.sizes = d_sizes.p; sp.has = d_has.p; sp.counts = d_counts.p; sp.categories = d_cat.p;
        sp.exhausted = d_exh.p; sp.F = n_families; sp.seed = seed;
        sp.em = c->have_em ? c->d_em.p : nullptr; sp.em_rows = c->em_rows;
        sp.n_nodes = n; sp.n_species = c->n_species; sp.K = K; sp.N = c->N; sp.max_sim = max_sim;
        sp.max_attempts = 1 + std::max(max_redraws, 0);
        simulate_kernel<<<(unsigned)((F + 255) / 256), 256, 0, c->stream>>>(sp);
        CK(cudaGetLastError());
        d2h(c, counts, d_counts.p, F * c->n_species);
        d2h(c, categories, d_cat.p, F);
        unsigned long long exhausted = 0;
        CK(cudaMemcpyAsync(&exhausted, d_exh.p, sizeof exhausted, cudaMemcpyDeviceToHost, c->stream));
        std::vector<int32_t> sizes_t;
        if (node_sizes) {
            sizes_t.resize(F * n);
            CK(cudaMemcpyAsync(sizes_t.data(), d_sizes.p, F * n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        if (node_sizes)
            for (size_t f = 0; f < F; ++f)
                for (int i = 0; i < n; ++i) node_sizes[f * n + i] = sizes_t[(size_t)i * F + f];
        if (n_not_at_root) *n_not_at_root = (int64_t)exhausted;
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_branch_probabilities(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, const int32_t* states,
                                   const uint8_t* selected, double* probs)
{
    if (!c) return CAFE_B200_ERR_ARG;
    if (c->is_group()) {
        if (!states || !probs) { c->err = "bad argument"; return CAFE_B200_ERR_ARG; }
        const size_t nn = (size_t)c->n_nodes;
        std::vector<int32_t> st;
        std::vector<uint8_t> sel;
        std::vector<double> pr;
        if (!c->order.empty()) {                 // buckets: gather the caller's rows bucket-major, scatter the results back
            st.resize((size_t)c->F * nn);
            pr.resize((size_t)c->F * nn);
            if (selected) sel.resize((size_t)c->F);
            for (int64_t p = 0; p < c->F; ++p) {
                memcpy(&st[(size_t)p * nn], states + (size_t)c->order[p] * nn, nn * sizeof(int32_t));
                if (selected) sel[p] = selected[c->order[p]];
            }
        }
        const int32_t* st_in = st.empty() ? states : st.data();
        const uint8_t* sel_in = !selected ? nullptr : sel.empty() ? selected : sel.data();
        double* pr_out = pr.empty() ? probs : pr.data();
        int bad = -1;
        const int rc = c->pool->run([&](int i) {
            const size_t b = (size_t)c->shard_begin[i];
            return cafe_b200_branch_probabilities(c->shards[i], lambdas, n_lambda, st_in + b * nn, sel_in ? sel_in + b : nullptr, pr_out + b * nn);
        }, &bad);
        if (rc != CAFE_B200_OK) { c->err = c->shards[bad]->err; return rc; }
        if (!pr.empty())
            for (int64_t p = 0; p < c->F; ++p) memcpy(probs + (size_t)c->order[p] * nn, &pr[(size_t)p * nn], nn * sizeof(double));
        return rc;
    }
    try {
        if (!lambdas || n_lambda < c->n_lambda_classes || !states || !probs) throw CudaError{"ARG: bad argument"};
        CK(cudaSetDevice(c->device));
        const int n = c->n_nodes;
        const size_t Fn = (size_t)c->F * n;
        for (size_t i = 0; i < Fn; ++i)
            if (states[i] < 0 || states[i] > c->max_family_size) throw CudaError{"RANGE: a reconstructed state is outside the matrix"};
        static const double one = 1.0;
        KeyPlan kp = plan_keys(c, lambdas, &one, 1);          // the model's own lambda, no gamma multiplier (src/execute.cpp:160-168)
        upload_plan(c, kp);
        launch_matrices(c, (int)kp.params.size());
        DevBuf<int32_t> d_parent, d_st;
        DevBuf<uint8_t> d_sel;
        DevBuf<double> d_out;
        struct Release { DevBuf<int32_t>&a, &b; DevBuf<uint8_t>& s; DevBuf<double>& o; ~Release() { a.release(); b.release(); s.release(); o.release(); } }
            release{d_parent, d_st, d_sel, d_out};
        d_parent.reserve(n); d_st.reserve(Fn); d_out.reserve(Fn);
        CK(cudaMemcpyAsync(d_parent.p, c->parent.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_st.p, states, Fn * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        if (selected) {
            d_sel.reserve((size_t)c->F);
            CK(cudaMemcpyAsync(d_sel.p, selected, (size_t)c->F, cudaMemcpyHostToDevice, c->stream));
        }
        viterbi_sum_kernel<<<(unsigned)((Fn + 255) / 256), 256, 0, c->stream>>>(c->d_arena.p, c->d_mat_of.p, d_parent.p, d_st.p,
                                                                                 selected ? d_sel.p : nullptr, c->F, n, c->LD,
                                                                                 c->max_family_size, d_out.p);
        CK(cudaGetLastError());
        d2h(c, probs, d_out.p, Fn);
        CK(cudaStreamSynchronize(c->stream));
        return CAFE_B200_OK;
    } catch (const CudaError& e) { return fail(c, e); }
}

int cafe_b200_pvalues(cafe_b200_ctx* c, const double* lambdas, int32_t n_lambda, int32_t n_sims, uint64_t seed, double* pvalues)
{
    if (!c) return CAFE_B200_ERR_ARG;
    cafe_b200_ctx* helper = nullptr;
    cafe_b200_ctx* sim = nullptr;
    try {
        if (!lambdas || n_lambda < c->n_lambda_classes || n_sims < 1 || !pvalues) throw CudaError{"ARG: bad argument"};
        if (!c->have_prior) throw CudaError{"STATE: set_prior must be called first"};
        const int R = c->R, n = c->n_nodes, mfs = c->max_family_size;
        cafe_b200_tree t{n, c->parent.data(), c->branch_length.data(), c->leaf_col.data(), c->lambda_class.data()};
        // 1. conditional distributions: n_sims families per root size 1..R, no redraws, no error model (get_random_probabilities,
        //    create_family: src/probability.cpp:355-375,434-447), child sizes below max_family_size.  The simulator needs the tree and
        //    full-size matrices only: a one-family helper context (this context may be sharded, or bucketed with smaller state spaces).
        const size_t Fs = (size_t)R * n_sims;
        std::vector<int32_t> roots(Fs), sim_counts(Fs * c->n_species);
        for (int r = 0; r < R; ++r) std::fill(roots.begin() + (size_t)r * n_sims, roots.begin() + (size_t)(r + 1) * n_sims, r + 1);
        std::vector<int32_t> dummy((size_t)c->n_species, 1);
        int rc = cafe_b200_create(&t, dummy.data(), 1, c->n_species, mfs, R, c->device, &helper);
        if (rc != CAFE_B200_OK) throw CudaError{std::string("simulator context: ") + create_error()};
        rc = cafe_b200_simulate(helper, lambdas, n_lambda, nullptr, nullptr, 0, mfs, 0, roots.data(), (int64_t)Fs, seed, sim_counts.data(),
                                nullptr, nullptr, nullptr);
        if (rc != CAFE_B200_OK) throw CudaError{std::string("simulator: ") + helper->err};
        cafe_b200_destroy(helper);
        helper = nullptr;
        // 2. their likelihood at the root size they were generated from (compute_family_probabilities, :377-432).  Like the reference,
        //    every simulated family is pruned over a state space truncated at m = min(max_family_size, largest size + max(50, largest / 5))
        //    (:394,416) -- here rounded up to the next multiple of 16 states, the row tile of the pruning kernel (bucketed contexts).
        //    The root size r of a family is one of its sizes, so r <= m and its root row exists in its bucket.
        std::vector<int32_t> ceilings;
        for (int v = 63; v < mfs; v += 16) ceilings.push_back(v);
        ceilings.push_back(mfs);
        rc = cafe_b200_create_bucketed(&t, sim_counts.data(), (int64_t)Fs, c->n_species, mfs, R, ceilings.data(), (int32_t)ceilings.size(),
                                       c->device, &sim);
        if (rc != CAFE_B200_OK) throw CudaError{std::string("simulated context: ") + create_error()};
        rc = cafe_b200_set_prior(sim, c->prior.data(), (int32_t)c->prior.size());
        std::vector<double> vec(Fs * R);
        if (rc == CAFE_B200_OK) rc = cafe_b200_root_vectors(sim, lambdas, n_lambda, 1.0, vec.data());
        if (rc != CAFE_B200_OK) throw CudaError{std::string("simulated families: ") + sim->err};
        cafe_b200_destroy(sim);
        sim = nullptr;
        std::vector<std::vector<double>> cond(R, std::vector<double>(n_sims));
        for (int r = 0; r < R; ++r) {
            for (int i = 0; i < n_sims; ++i) cond[r][i] = vec[((size_t)r * n_sims + i) * R + r];   // index r <-> root size r+1
            std::sort(cond[r].begin(), cond[r].end());
        }
        // 3. observed families: root vectors without the error model (compute_pvalues passes NULL, :549) over the full state space,
        //    then find_best_pvalue (:513-526) over root sizes below rint(1.25 * largest count)
        const std::vector<double> em = c->em_host;
        const int em_rows = c->em_rows, em_maxcnt = c->em_maxcnt;
        if (!em.empty() && (rc = cafe_b200_set_error_model(c, nullptr, 0, 0)) != CAFE_B200_OK) return rc;
        vec.assign((size_t)c->F * R, 0.0);
        rc = cafe_b200_root_vectors(c, lambdas, n_lambda, 1.0, vec.data());
        if (!em.empty()) {
            const int rc2 = cafe_b200_set_error_model(c, em.data(), em_rows, em_maxcnt);
            if (rc == CAFE_B200_OK) rc = rc2;
        }
        if (rc != CAFE_B200_OK) return rc;
        for (int64_t f = 0; f < c->F; ++f) {
            const int limit = std::min((int)std::rint(c->max_count[(size_t)f] * 1.25), R);
            double best = 0.0;
            for (int j = 0; j < limit; ++j) {
                const std::vector<double>& d = cond[j];
                const double v = vec[(size_t)f * R + j];
                size_t idx = d.size() - 1;
                auto bound = std::upper_bound(d.begin(), d.end(), v);
                if (bound != d.end()) idx = (size_t)(bound - d.begin());
                best = std::max(best, (double)idx / (double)d.size());
            }
            pvalues[f] = best;
        }
        return CAFE_B200_OK;
    } catch (const CudaError& e) {
        if (helper) cafe_b200_destroy(helper);
        if (sim) cafe_b200_destroy(sim);
        return fail(c, e);
    }
}

}  // extern "C"
