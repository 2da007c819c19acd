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
        int bad = -1;
        const int rc = c->pool->run([&](int i) {
            const size_t b = (size_t)c->shard_begin[i], nn = (size_t)c->n_nodes;
            return cafe_b200_branch_probabilities(c->shards[i], lambdas, n_lambda, states + b * nn, selected ? selected + b : nullptr, probs + b * nn);
        }, &bad);
        if (rc != CAFE_B200_OK) c->err = c->shards[bad]->err;
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
    if (c->is_group()) {   // every device simulates the same conditional distributions (same seed) and scores its own shard of families
        if (!pvalues) { c->err = "bad argument"; return CAFE_B200_ERR_ARG; }
        int bad = -1;
        const int rc = c->pool->run([&](int i) { return cafe_b200_pvalues(c->shards[i], lambdas, n_lambda, n_sims, seed, pvalues + c->shard_begin[i]); }, &bad);
        if (rc != CAFE_B200_OK) c->err = c->shards[bad]->err;
        return rc;
    }
    cafe_b200_ctx* sim = nullptr;
    try {
        if (!lambdas || n_lambda < c->n_lambda_classes || n_sims < 1 || !pvalues) throw CudaError{"ARG: bad argument"};
        if (!c->have_prior) throw CudaError{"STATE: set_prior must be called first"};
        const int R = c->R, n = c->n_nodes;
        // 1. conditional distributions: n_sims families per root size 1..R, no redraws, no error model (get_random_probabilities,
        //    create_family: src/probability.cpp:355-375,434-447), child sizes below max_family_size
        const size_t Fs = (size_t)R * n_sims;
        std::vector<int32_t> roots(Fs), sim_counts(Fs * c->n_species);
        for (int r = 0; r < R; ++r) std::fill(roots.begin() + (size_t)r * n_sims, roots.begin() + (size_t)(r + 1) * n_sims, r + 1);
        int rc = cafe_b200_simulate(c, lambdas, n_lambda, nullptr, nullptr, 0, c->max_family_size, 0, roots.data(), (int64_t)Fs, seed,
                                    sim_counts.data(), nullptr, nullptr, nullptr);
        if (rc != CAFE_B200_OK) return rc;
        // 2. their likelihood at the root size they were generated from (compute_family_probabilities, :377-432; the reference
        //    truncates each family's state space at its largest size + max(50, size/5), we keep the full space)
        cafe_b200_tree t{n, c->parent.data(), c->branch_length.data(), c->leaf_col.data(), c->lambda_class.data()};
        rc = cafe_b200_create(&t, sim_counts.data(), (int64_t)Fs, c->n_species, c->max_family_size, R, c->device, &sim);
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
        // 3. observed families: root vectors without the error model (compute_pvalues passes NULL, :549), then
        //    find_best_pvalue (:513-526) over root sizes below rint(1.25 * largest count)
        const bool had_em = c->have_em;
        c->have_em = false;
        vec.assign((size_t)c->F * R, 0.0);
        rc = cafe_b200_root_vectors(c, lambdas, n_lambda, 1.0, vec.data());
        c->have_em = had_em;
        if (rc != CAFE_B200_OK) return rc;
        for (int64_t f = 0; f < c->F; ++f) {
            int mx = 0;
            for (int j = 0; j < c->n_species; ++j) mx = std::max(mx, c->counts[(size_t)f * c->n_species + j]);
            const int limit = std::min((int)std::rint(mx * 1.25), R);
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
        if (sim) cafe_b200_destroy(sim);
        return fail(c, e);
    }
}

}  // extern "C"
