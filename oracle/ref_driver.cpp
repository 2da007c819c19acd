// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" driver that is compiled TOGETHER WITH the unmodified reference
// sources where they lie (/root/reference/src/*.cpp) into oracle/_ref/libcafe_ref.so by
// oracle/build_ref.sh.  It builds the reference's own objects (clade, gene_family,
// base_model, gamma_model, error_model, root_equilibrium_distribution) and calls the
// reference's own functions for the hot path, so tests can pin the C restatement
// (oracle/cafe_oracle.c) and the CUDA path against the real thing:
//   matrix_cache::precalculate_matrices / get_matrix      src/matrix_cache.cpp:88-163
//   birthdeath_rate_with_log_alpha                         src/probability.cpp:104-148
//   inference_prune                                        src/core.cpp:134-145
//   base_model::infer_family_likelihoods                   src/base_model.cpp:53-100
//   gamma_model::infer_family_likelihoods / prune          src/gamma_core.cpp:143-237
//   base_model/gamma_model::reconstruct_ancestral_states   src/base_model.cpp:133-170, src/gamma_core.cpp:290-339
//   get_gamma                                              src/gamma.cpp:225-241
// Built with -fno-access-control so per-family side-channel outputs (model::results,
// gamma_model::_category_likelihoods) can be read without modifying the reference.
// No reference source is copied into this repository.
#include <cstring>
#include <fstream>
#include <cmath>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <set>
#include <omp.h>

#include "easylogging++.h"
#include "clade.h"
#include "gene_family.h"
#include "lambda.h"
#include "matrix_cache.h"
#include "probability.h"
#include "core.h"
#include "base_model.h"
#include "gamma_core.h"
#include "gamma.h"
#include "error_model.h"
#include "root_equilibrium_distribution.h"
#include "user_data.h"
#include "io.h"
#include "gene_family_reconstructor.h"
#include "report.h"
#include "simulator.h"
#include "optimizer.h"
#include "newick_ape_loader.h"
#include "optimizer_scorer.h"
#include "poisson.h"

#include "ref_optimize.hpp"

INITIALIZE_EASYLOGGINGPP

std::mt19937 randomizer_engine(10);  // the reference's main.cpp seeds from random_device; we seed explicitly

void init_lgamma_cache();

namespace {

bool g_inited = false;
void ensure_init()
{
    if (g_inited) return;
    init_lgamma_cache();
    el::Configurations conf;
    conf.setToDefault();
    conf.set(el::Level::Global, el::ConfigurationType::Enabled, "false");
    conf.set(el::Level::Global, el::ConfigurationType::ToFile, "false");
    conf.set(el::Level::Global, el::ConfigurationType::ToStandardOutput, "false");
    el::Loggers::reconfigureLogger("default", conf);
    g_inited = true;
}

std::vector<std::string> split(const std::string& s, char d)
{
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, d)) out.push_back(item);
    return out;
}


lambda* make_lambda(ref_ctx* c, const double* lambdas, int n)
{
    if (c->lambda_tree) {
        auto m = c->lambda_tree->get_lambda_index_map();
        return new multiple_lambda(m, std::vector<double>(lambdas, lambdas + n));
    }
    return new single_lambda(lambdas[0]);
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }

double ref_birthdeath_rate_with_log_alpha(int s, int c, double log_alpha, double coeff)
{
    ensure_init();
    return birthdeath_rate_with_log_alpha(s, c, log_alpha, coeff);
}

double ref_transition_probability(double lambda, double t, int s, int c)
{
    ensure_init();
    return the_probability_of_going_from_parent_fam_size_to_c(lambda, t, s, c);
}

double ref_chooseln(double n, double r) { ensure_init(); return chooseln(n, r); }

// Full N x N matrix via the reference's cache (so key quantisation applies).  out is row-major [s][c].
int ref_matrix(int N, double lambda, double t, double* out)
{
    ensure_init();
    try {
        matrix_cache cache(N);
        cache.precalculate_matrices(std::vector<double>{lambda}, std::set<double>{t});
        const matrix* m = cache.get_matrix(t, lambda);
        for (int s = 0; s < N; ++s)
            for (int c = 0; c < N; ++c) out[s * N + c] = m->get(s, c);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

int ref_is_saturated(double t, double lambda) { return matrix_cache::is_saturated(t, lambda) ? 1 : 0; }

int ref_get_gamma(int K, double alpha, double* probs, double* multipliers)
{
    std::vector<double> f(K), r(K);
    get_gamma(f, r, alpha);
    for (int i = 0; i < K; ++i) { probs[i] = f[i]; multipliers[i] = r[i]; }
    return 0;
}

// ---- tree flattening: nodes in the reference's reverse level order (src/clade.cpp:69-100) ----
int ref_tree_node_count(const char* newick)
{
    try {
        std::unique_ptr<clade> t(parse_newick(newick));
        return int(t->reverse_level_end() - t->reverse_level_begin());
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// names_out: '\t'-joined node names in reverse level order (caller buffer).  lambda_class filled
// from lambda_newick (0-based) when given, else zeros.
int ref_tree_flatten(const char* newick, const char* lambda_newick, int* parent, double* branch_length,
                     int* is_leaf, int* lambda_class, char* names_out, int names_cap)
{
    try {
        std::unique_ptr<clade> t(parse_newick(newick));
        std::vector<const clade*> order(t->reverse_level_begin(), t->reverse_level_end());
        std::map<const clade*, int> idx;
        for (size_t i = 0; i < order.size(); ++i) idx[order[i]] = int(i);
        std::map<std::string, int> lmap;
        if (lambda_newick && *lambda_newick) {
            std::unique_ptr<clade> lt(parse_newick(lambda_newick, true));
            t->validate_lambda_tree(lt.get());
            lmap = lt->get_lambda_index_map();
        }
        std::string names;
        for (size_t i = 0; i < order.size(); ++i) {
            const clade* c = order[i];
            parent[i] = c->is_root() ? -1 : idx.at(c->get_parent());
            branch_length[i] = c->get_branch_length();
            is_leaf[i] = c->is_leaf() ? 1 : 0;
            lambda_class[i] = lmap.empty() ? 0 : lmap.at(c->get_taxon_name());
            if (i) names += '\t';
            names += c->get_taxon_name();
        }
        if (int(names.size()) + 1 > names_cap) { g_err = "names buffer too small"; return 2; }
        std::strcpy(names_out, names.c_str());
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// The prior tables the reference builds (src/root_equilibrium_distribution.cpp:13-68, selected by user_data::create_prior,
// src/user_data.cpp:176-206): kind 0 = uniform over num_values sizes, 1 = user root distribution {sizes[i]: counts[i]} (`-f`),
// 2 = Poisson(poisson_lambda) over num_values simulated roots (`-p<lambda>`).  out[j] = compute(j) for j < cap (a float, .h:40);
// table_len = number of sizes the table covers (compute returns 0 beyond it, .cpp:81-87).
int ref_prior_table(int kind, double poisson_lambda, const int* sizes, const int* counts, int n, int num_values, float* out, int cap,
                    int* table_len)
{
    ensure_init();
    try {
        std::unique_ptr<root_equilibrium_distribution> p;
        if (kind == 0) p.reset(new root_equilibrium_distribution((size_t)num_values));
        else if (kind == 1) {
            std::map<int, int> m;
            for (int i = 0; i < n; ++i) m[sizes[i]] = counts[i];
            p.reset(new root_equilibrium_distribution(m));
        } else p.reset(new root_equilibrium_distribution(poisson_lambda, (size_t)num_values));
        for (int j = 0; j < cap; ++j) out[j] = p->compute(j);
        if (table_len) *table_len = (int)p->_frequency_percentage.size();
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// `-p` without a value: the reference estimates the Poisson mean of the root prior from the gene families
// (root_equilibrium_distribution(gene_families, num_values), src/root_equilibrium_distribution.cpp:42-54).  That constructor keeps
// neither the fitted mean nor the score, so its three statements are run here on the reference's own poisson_scorer and optimizer
// (seeded randomizer_engine), and the constructor itself builds the table that is returned.
int ref_fit_poisson_prior(const char* species, const int* counts, long F, unsigned seed, int num_values, double* poisson_lambda,
                          double* score, int* iterations, float* table, int cap, int* table_len)
{
    ensure_init();
    try {
        auto sp = split(species, '\t');
        std::vector<gene_family> fams(F);
        for (long f = 0; f < F; ++f) {
            fams[f].set_id(std::to_string(f));
            for (size_t j = 0; j < sp.size(); ++j) fams[f].set_species_size(sp[j], counts[f * (long)sp.size() + (long)j]);
        }
        randomizer_engine.seed(seed);
        poisson_scorer scorer(fams);
        optimizer opt(&scorer);
        opt.quiet = true;
        optimizer_parameters params;
        auto result = opt.optimize(params);
        *poisson_lambda = result.values[0];
        *score = result.score;
        *iterations = result.num_iterations;
        if (table) {
            randomizer_engine.seed(seed);
            root_equilibrium_distribution prior(fams, (size_t)num_values);
            for (int j = 0; j < cap; ++j) table[j] = prior.compute(j);
            if (table_len) *table_len = (int)prior._frequency_percentage.size();
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// ---- context ----
// species: '\t'-joined leaf names giving the column order of counts (F x n_species, row-major).
// prior: n_prior doubles (float-representable); installed verbatim as the prior's percentage table.
// error model: em_rows x 3 doubles or NULL.
void* ref_ctx_create(const char* newick, const char* lambda_newick, const char* species,
                     const int* counts, long F, int max_family_size, int max_root_family_size,
                     const double* prior, int n_prior,
                     const double* em_probs, int em_rows, int em_maxcnt)
{
    ensure_init();
    try {
        auto c = new ref_ctx();
        c->tree.reset(parse_newick(newick));
        if (lambda_newick && *lambda_newick) {
            c->lambda_tree.reset(parse_newick(lambda_newick, true));
            c->tree->validate_lambda_tree(c->lambda_tree.get());
        }
        c->order.assign(c->tree->reverse_level_begin(), c->tree->reverse_level_end());
        auto sp = split(species, '\t');
        c->ud.gene_families.resize(F);
        for (long f = 0; f < F; ++f) {
            c->ud.gene_families[f].set_id(std::to_string(f));
            for (size_t j = 0; j < sp.size(); ++j)
                c->ud.gene_families[f].set_species_size(sp[j], counts[f * sp.size() + j]);
        }
        c->max_family_size = max_family_size;
        c->max_root_family_size = max_root_family_size;
        root_equilibrium_distribution p((size_t)std::max(n_prior, 1));
        p._frequency_percentage.assign(prior, prior + n_prior);
        c->ud.prior = std::move(p);
        if (em_probs) {
            c->em.reset(new error_model());
            c->em->set_max_family_size(em_maxcnt);
            for (int i = 0; i < em_rows; ++i)
                c->em->set_probabilities(i, {em_probs[3 * i], em_probs[3 * i + 1], em_probs[3 * i + 2]});
        }
        c->ud.p_tree = c->tree.get();
        c->ud.max_family_size = max_family_size;
        c->ud.max_root_family_size = max_root_family_size;
        return c;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void ref_ctx_destroy(void* h) { delete (ref_ctx*)h; }

// Replace the error-model rows (the optimiser mutates epsilon every step).
int ref_ctx_set_error_model(void* h, const double* em_probs, int em_rows, int em_maxcnt)
{
    auto c = (ref_ctx*)h;
    try {
        c->em.reset(new error_model());
        c->em->set_max_family_size(em_maxcnt);
        for (int i = 0; i < em_rows; ++i)
            c->em->set_probabilities(i, {em_probs[3 * i], em_probs[3 * i + 1], em_probs[3 * i + 2]});
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Root vector (R doubles; index i <-> root size i+1) for one family: inference_prune.
int ref_prune(void* h, long family, const double* lambdas, int n_lambda, double multiplier, double* out)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        std::unique_ptr<lambda> mult(lam->multiply(multiplier));
        matrix_cache calc(std::max(c->max_root_family_size, c->max_family_size) + 1);
        calc.precalculate_matrices(get_lambda_values(mult.get()), c->tree->get_branch_lengths());
        auto v = inference_prune(c->ud.gene_families[family], calc, lam.get(), c->em.get(), c->tree.get(), multiplier,
                                 c->max_root_family_size, c->max_family_size);
        std::copy(v.begin(), v.end(), out);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// base_model::infer_family_likelihoods; family_lnl (F) may be NULL.
int ref_eval_base(void* h, const double* lambdas, int n_lambda, double* neg_lnl, double* family_lnl)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        base_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, c->em.get());
        *neg_lnl = m.infer_family_likelihoods(c->ud.prior, lam.get());
        if (family_lnl && !std::isinf(*neg_lnl))
            for (size_t i = 0; i < m.results.size(); ++i) family_lnl[i] = m.results[i].posterior_probability;
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Same, but skipping the O(F*U) dedup of the model constructor for timing large samples: not needed,
// the constructor is outside the timed call below.
void* ref_base_model_create(void* h, const double* lambdas, int n_lambda)
{
    auto c = (ref_ctx*)h;
    try {
        lambda* lam = make_lambda(c, lambdas, n_lambda);
        return new base_model(lam, c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, c->em.get());
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

// gamma_model::infer_family_likelihoods with explicit multipliers / category probabilities.
// cat_lk: F x K (may be NULL); failed: F bytes (may be NULL), from gamma_model::prune's return value.
int ref_eval_gamma(void* h, const double* lambdas, int n_lambda, const double* multipliers, const double* cat_probs,
                   int K, double* neg_lnl, double* cat_lk, unsigned char* failed)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        gamma_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size,
                      std::vector<double>(cat_probs, cat_probs + K), std::vector<double>(multipliers, multipliers + K), c->em.get());
        m._alpha = 1.0;  // explicit multipliers were supplied; can_infer only needs alpha >= 0
        *neg_lnl = m.infer_family_likelihoods(c->ud.prior, lam.get());
        size_t F = c->ud.gene_families.size();
        for (size_t i = 0; i < F; ++i) {
            const auto& v = m._category_likelihoods[i];
            bool ok = v.size() == size_t(K);
            if (failed) failed[i] = ok ? 0 : 1;
            if (cat_lk) for (int k = 0; k < K; ++k) cat_lk[i * K + k] = (k < int(v.size())) ? v[k] : 0.0;
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Family-level p-values with the reference's own compute_pvalues (src/probability.cpp:528-570), called as estimator::execute does
// (src/execute.cpp:164-171): matrices of size max(max_family_size, max_root) + 100 for the model's lambda.  randomizer_engine is
// seeded here (the reference never seeds it).
int ref_pvalues(void* h, const double* lambdas, int n_lambda, int n_sims, unsigned seed, double* out)
{
    auto c = (ref_ctx*)h;
    try {
        randomizer_engine.seed(seed);
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);
        cache.precalculate_matrices(get_lambda_values(lam.get()), c->tree->get_branch_lengths());
        pvalue_parameters p = {c->tree.get(), lam.get(), c->max_family_size, c->max_root_family_size, cache};
        auto pv = compute_pvalues(p, c->ud.gene_families, n_sims);
        for (size_t i = 0; i < pv.size(); ++i) out[i] = pv[i];
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// ---- the reference's own readers and writers (pins cafe5_b200/host/io.hpp) -----------------------------------------------------
static int put_text(const std::string& s, char* buf, long cap)
{
    if ((long)s.size() + 1 > cap) { g_err = "buffer too small"; return 1; }
    memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}

// read_gene_families (src/io.cpp:134-217) against the tree `newick`: counts F x n_leaves in the tree's reverse-level leaf order
int ref_read_families(const char* path, const char* newick, long* n_families, int* counts, long counts_cap, char* ids, long ids_cap)
{
    try {
        std::unique_ptr<clade> tree(parse_newick(newick));
        std::ifstream in(path);
        std::vector<gene_family> fams;
        read_gene_families(in, tree.get(), fams);
        std::vector<std::string> leaves;
        for (auto it = tree->reverse_level_begin(); it != tree->reverse_level_end(); ++it)
            if ((*it)->is_leaf()) leaves.push_back((*it)->get_taxon_name());
        *n_families = (long)fams.size();
        if ((long)(fams.size() * leaves.size()) > counts_cap) { g_err = "counts buffer too small"; return 1; }
        std::string all;
        for (size_t f = 0; f < fams.size(); ++f) {
            for (size_t j = 0; j < leaves.size(); ++j) counts[f * leaves.size() + j] = fams[f].get_species_size(leaves[j]);
            if (f) all += '\t';
            all += fams[f].id();
        }
        return put_text(all, ids, ids_cap);
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// read_error_model_file (src/io.cpp:228-274): every row the model holds, as get_probs returns them
int ref_read_error_model(const char* path, double* probs, int rows_cap, int* rows, int* max_count)
{
    try {
        std::ifstream in(path);
        error_model em;
        read_error_model_file(in, &em);
        const int n = (int)em.get_max_family_size();          // number of rows (error_model.h:55-57)
        *max_count = (int)em._max_family_size;                // the "maxcnt" line
        *rows = n;
        if (n > rows_cap) { g_err = "probs buffer too small"; return 1; }
        for (int i = 0; i < n; ++i) {
            auto r = em.get_probs(i);
            for (int d = 0; d < 3; ++d) probs[i * 3 + d] = r[d];
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// The text the reference writes for one evaluated model: *_family_likelihoods.txt (base_model.cpp:102-109 / gamma_core.cpp:52-58),
// *_results.txt (write_vital_statistics, core.cpp:97-112, gamma_core.cpp:46-50) and, for the gamma model,
// Gamma_category_likelihoods.txt (gamma_core.cpp:359-374).  K == 0: base model.
int ref_write_outputs(void* h, const double* lambdas, int n_lambda, const double* multipliers, const double* cat_probs, int K, double alpha,
                      char* fam, long fam_cap, char* res, long res_cap, char* cat, long cat_cap)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        std::ostringstream a, b, d;
        if (K == 0) {
            base_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, c->em.get());
            const double score = m.infer_family_likelihoods(c->ud.prior, lam.get());
            m.write_family_likelihoods(a);
            m.write_vital_statistics(b, score);
        } else {
            gamma_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size,
                          std::vector<double>(cat_probs, cat_probs + K), std::vector<double>(multipliers, multipliers + K), c->em.get());
            m._alpha = alpha;
            const double score = m.infer_family_likelihoods(c->ud.prior, lam.get());
            m.write_family_likelihoods(a);
            m.write_vital_statistics(b, score);
            matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);
            std::unique_ptr<reconstruction> rec(m.reconstruct_ancestral_states(c->ud, c->ui, &cache));
            cladevector order;
            dynamic_cast<gamma_model_reconstruction*>(rec.get())->print_category_likelihoods(d, order);
        }
        if (put_text(a.str(), fam, fam_cap) || put_text(b.str(), res, res_cap) || put_text(d.str(), cat, cat_cap)) return 1;
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// The reconstruction tables as the reference writes them (reconstruction::write_results, gene_family_reconstructor.cpp:352-379) for
// the base model (K == 0) or the gamma model at the given lambda: count.tab, change.tab, asr.tre, family_results, clade_results.
int ref_write_reconstruction(void* h, const double* lambdas, int n_lambda, const double* multipliers, const double* cat_probs, int K,
                             const double* pvalues, char* count, char* change, char* asr, char* famres, char* clade, long cap)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);
        std::unique_ptr<model> m;
        if (K == 0) m.reset(new base_model(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, nullptr));
        else {
            auto g = new gamma_model(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size,
                                     std::vector<double>(cat_probs, cat_probs + K), std::vector<double>(multipliers, multipliers + K), nullptr);
            g->_alpha = 1.0;
            m.reset(g);
            m->infer_family_likelihoods(c->ud.prior, lam.get());
        }
        std::unique_ptr<reconstruction> rec(m->reconstruct_ancestral_states(c->ud, c->ui, &cache));
        auto order = get_ape_order(c->tree.get());
        branch_probabilities none;
        std::vector<double> pv(pvalues, pvalues + c->ud.gene_families.size());
        std::ostringstream a, b, d, e, f;
        rec->print_node_counts(a, order);
        rec->print_node_change(b, order);
        rec->print_reconstructed_states(d, order, none);
        rec->print_increases_decreases_by_family(e, order, pv);
        rec->print_increases_decreases_by_clade(f, order);
        if (put_text(a.str(), count, cap) || put_text(b.str(), change, cap) || put_text(d.str(), asr, cap) || put_text(e.str(), famres, cap) ||
            put_text(f.str(), clade, cap)) return 1;
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Branch probabilities as estimator::execute computes them (src/execute.cpp:173-184) for the base model: compute_viterbi_sum for
// every node of every family whose p-value is below ui.pvalue.  probs: F x n_nodes (reverse level order), -1 = none.  Also returns
// the reference's _branch_probabilities.tab and _asr.tre (with significance stars) texts.
int ref_branch_probabilities(void* h, const double* lambdas, int n_lambda, const double* pvalues, double* probs, char* tab, char* asr, long cap)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);
        cache.precalculate_matrices(get_lambda_values(lam.get()), c->tree->get_branch_lengths());
        base_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, nullptr);
        std::unique_ptr<reconstruction> rec(m.reconstruct_ancestral_states(c->ud, c->ui, &cache));
        branch_probabilities bp;
        const size_t n = c->order.size();
        for (size_t i = 0; i < c->ud.gene_families.size(); ++i) {
            for (size_t k = 0; k < n; ++k) probs[i * n + k] = -1.0;
            if (pvalues[i] < c->ui.pvalue)
                for (size_t k = 0; k < n; ++k) {
                    auto p = compute_viterbi_sum(c->order[k], c->ud.gene_families[i], rec.get(), c->max_family_size, cache, lam.get());
                    bp.set(c->ud.gene_families[i], c->order[k], p);
                    if (p._is_valid) probs[i * n + k] = p._value;
                }
        }
        auto order = get_ape_order(c->tree.get());
        std::ostringstream a, b;
        print_branch_probabilities(a, order, c->ud.gene_families, bp);
        rec->print_reconstructed_states(b, order, bp);
        if (put_text(a.str(), tab, cap) || put_text(b.str(), asr, cap)) return 1;
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// <Model>_report.cafe as estimator::execute builds it for the base model (src/execute.cpp:167-197): reconstruction, branch
// probabilities for the families whose p-value is below ui.pvalue, Report::compute_expansion, one line item per family, operator<<.
int ref_write_report(void* h, const double* lambdas, int n_lambda, const double* pvalues, char* out, long cap)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);
        cache.precalculate_matrices(get_lambda_values(lam.get()), c->tree->get_branch_lengths());
        base_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, nullptr);
        std::unique_ptr<reconstruction> rec(m.reconstruct_ancestral_states(c->ud, c->ui, &cache));
        branch_probabilities bp;
        for (size_t i = 0; i < c->ud.gene_families.size(); ++i)
            if (pvalues[i] < c->ui.pvalue)
                for (auto node : c->order)
                    bp.set(c->ud.gene_families[i], node, compute_viterbi_sum(node, c->ud.gene_families[i], rec.get(), c->max_family_size, cache, lam.get()));
        Report r(c->tree.get(), c->lambda_tree.get(), lam.get());
        r.compute_expansion(c->ud.gene_families, *rec.get());
        for (size_t i = 0; i < c->ud.gene_families.size(); ++i) r.add_line_item(c->ud.gene_families[i], rec.get(), pvalues[i], bp);
        std::ostringstream ost;
        ost << r;
        return put_text(ost.str(), out, cap) ? 1 : 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// The error model file as the reference writes it (write_error_model_file, src/io.cpp:277-297) for a model built like ref_ctx_create
// builds it (set_max_family_size, set_deviations, set_probabilities).
int ref_write_error_model(const double* em_probs, int em_rows, int em_maxcnt, char* out, long cap)
{
    ensure_init();
    try {
        error_model em;
        em.set_max_family_size(em_maxcnt);
        em.set_deviations({"-1", "0", "1"});
        for (int i = 0; i < em_rows; ++i) em.set_probabilities(i, {em_probs[3 * i], em_probs[3 * i + 1], em_probs[3 * i + 2]});
        std::ostringstream ost;
        write_error_model_file(ost, em);
        return put_text(ost.str(), out, cap) ? 1 : 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// simulation.txt / simulation_truth.txt as simulator::print_simulations writes them (src/simulator.cpp:135-172) for n families whose
// node values (n x n_nodes, reverse level order) and lambdas are given.
int ref_print_simulations(void* h, long n, const int* node_sizes, const double* lambdas, int include_internal, char* out, long cap)
{
    auto c = (ref_ctx*)h;
    try {
        std::vector<simulated_family> results(n);
        const size_t nn = c->order.size();
        for (long f = 0; f < n; ++f) {
            results[f].lambda = lambdas[f];
            for (size_t k = 0; k < nn; ++k) results[f].values[c->order[k]] = node_sizes[f * nn + k];
        }
        simulator sim(c->ud, c->ui);
        std::ostringstream ost;
        sim.print_simulations(ost, include_internal != 0, results);
        return put_text(ost.str(), out, cap) ? 1 : 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Pupko reconstruction, base model.  states: F x n_nodes ints in reverse level order (leaves = observed).
int ref_reconstruct_base(void* h, const double* lambdas, int n_lambda, int* states)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        base_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size, c->em.get());
        matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);  // src/execute.cpp:167
        std::unique_ptr<reconstruction> rec(m.reconstruct_ancestral_states(c->ud, c->ui, &cache));
        size_t n = c->order.size();
        for (size_t f = 0; f < c->ud.gene_families.size(); ++f)
            for (size_t i = 0; i < n; ++i) states[f * n + i] = rec->get_node_count(c->ud.gene_families[f], c->order[i]);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Pupko reconstruction, gamma model: per-category states (F x K x n_nodes) and rounded weighted
// average (F x n_nodes, leaves = observed).  averaged_raw (F x n_nodes doubles) may be NULL.
int ref_reconstruct_gamma(void* h, const double* lambdas, int n_lambda, const double* multipliers,
                          const double* cat_probs, int K, int* cat_states, int* states, double* averaged_raw)
{
    auto c = (ref_ctx*)h;
    try {
        std::unique_ptr<lambda> lam(make_lambda(c, lambdas, n_lambda));
        gamma_model m(lam.get(), c->tree.get(), &c->ud.gene_families, c->max_family_size, c->max_root_family_size,
                      std::vector<double>(cat_probs, cat_probs + K), std::vector<double>(multipliers, multipliers + K), c->em.get());
        m._alpha = 1.0;
        matrix_cache cache(std::max(c->max_family_size, c->max_root_family_size) + 100);
        std::unique_ptr<reconstruction> rec(m.reconstruct_ancestral_states(c->ud, c->ui, &cache));
        auto g = dynamic_cast<gamma_model_reconstruction*>(rec.get());
        size_t n = c->order.size();
        for (size_t f = 0; f < c->ud.gene_families.size(); ++f) {
            const auto& r = g->_reconstructions.at(c->ud.gene_families[f].id());
            for (size_t i = 0; i < n; ++i) {
                const clade* node = c->order[i];
                states[f * n + i] = rec->get_node_count(c->ud.gene_families[f], node);
                for (int k = 0; k < K; ++k) {
                    int v = node->is_leaf() ? c->ud.gene_families[f].get_species_size(node->get_taxon_name())
                                            : r.category_reconstruction[k].at(node);
                    cat_states[(f * K + k) * n + i] = v;
                }
                if (averaged_raw)
                    averaged_raw[f * n + i] = node->is_leaf() ? double(states[f * n + i]) : r.reconstruction.at(node);
            }
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// ---- timing helpers for bench.py (cpu_baseline / --impl reference) ---------------------------------
// A full reference evaluation on the config-5 shard takes minutes (about 400 matrices at ~0.1 s each plus
// ~50 ms per family per thread), so the bench times BOUNDED SAMPLES of its two legs, each through the
// reference's own functions, and scales them linearly (documented in bench.py / DESIGN.md):
//   leg 1: matrix_cache::precalculate_matrices on every `stride`-th branch length  (src/matrix_cache.cpp:113-163)
//   leg 2: gamma_model::prune (+ the mixture of gamma_core.cpp:196-207) over the first n families under the
//          reference's `#pragma omp parallel for` shape (gamma_core.cpp:190), against a prebuilt full cache.
struct ref_session {
    std::unique_ptr<lambda> lam;
    std::unique_ptr<gamma_model> model;
    std::unique_ptr<matrix_cache> cache;
    int K = 0;
};

double ref_time_precalculate(void* h, const double* lambdas, int n_lambda, const double* multipliers, int K,
                             int stride, int* keys_done, int* keys_total)
{
    auto c = (ref_ctx*)h;
    try {
        std::vector<double> all;
        for (int k = 0; k < K; ++k) for (int i = 0; i < n_lambda; ++i) all.push_back(lambdas[i] * multipliers[k]);
        auto bls = c->tree->get_branch_lengths();
        std::set<double> sample;
        int idx = 0;
        for (double b : bls) if ((idx++ % stride) == 0) sample.insert(b);
        matrix_cache cache(std::max(c->max_root_family_size, c->max_family_size) + 1);
        double t0 = omp_get_wtime();
        cache.precalculate_matrices(all, sample);
        double t1 = omp_get_wtime();
        if (keys_done) *keys_done = cache.get_cache_size();
        if (keys_total) *keys_total = int(all.size() * bls.size());
        return t1 - t0;
    } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

void* ref_session_create(void* h, const double* lambdas, int n_lambda, const double* multipliers, const double* cat_probs, int K)
{
    auto c = (ref_ctx*)h;
    try {
        auto s = new ref_session();
        s->K = K;
        s->lam.reset(make_lambda(c, lambdas, n_lambda));
        std::vector<gene_family> none;
        s->model.reset(new gamma_model(s->lam.get(), c->tree.get(), nullptr, c->max_family_size, c->max_root_family_size,
                                       std::vector<double>(cat_probs, cat_probs + K), std::vector<double>(multipliers, multipliers + K), c->em.get()));
        s->model->_alpha = 1.0;
        s->cache.reset(new matrix_cache(std::max(c->max_root_family_size, c->max_family_size) + 1));
        s->model->prepare_matrices_for_simulation(*s->cache);   // the call infer_family_likelihoods makes (gamma_core.cpp:186)
        return s;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void ref_session_destroy(void* s) { delete (ref_session*)s; }

// Seconds to prune + mix the first n families; sum_lnl gets the summed log-likelihood (0 families failed) or NaN.
double ref_session_prune(void* h, void* sh, long n, double* sum_lnl)
{
    auto c = (ref_ctx*)h;
    auto s = (ref_session*)sh;
    try {
        std::vector<double> lnl(n, 0.0);
        std::vector<char> bad(n, 0);
        double t0 = omp_get_wtime();
#pragma omp parallel for
        for (long i = 0; i < n; ++i) {
            std::vector<double> cat;
            if (s->model->prune(c->ud.gene_families[i], c->ud.prior, *s->cache, s->lam.get(), cat)) {
                double fam = std::accumulate(cat.begin(), cat.end(), 0.0);
                auto post = s->model->get_posterior_probabilities(cat);
                (void)post;
                lnl[i] = std::log(fam);
            } else bad[i] = 1;
        }
        double t1 = omp_get_wtime();
        double tot = 0;
        bool any_bad = false;
        for (long i = 0; i < n; ++i) { tot += lnl[i]; any_bad |= bad[i] != 0; }
        if (sum_lnl) *sum_lnl = any_bad ? std::nan("") : tot;
        return t1 - t0;
    } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

// The reference's optimizer over the reference's own CPU models (the CUDA counterpart lives in libcafe_ref_shim.so, ref_gpu_model.cpp).
// trace_* (optional, trace_cap rows): every attempt the optimizer made, in order.
int ref_optimize(void* h, int n_cat, int optimize_epsilon, unsigned seed, double* values_out, int* n_values, double* score,
                 int* iterations, int* attempts, double* seconds, double* trace_values, double* trace_scores, int* trace_failed_family,
                 int* trace_n_failed, int trace_cap, int* trace_n)
{
    auto c = (ref_ctx*)h;
    ref_opt::trace_buffer tb;
    tb.values = trace_values; tb.scores = trace_scores; tb.failed_family = trace_failed_family; tb.n_failed = trace_n_failed;
    tb.cap = trace_values ? trace_cap : 0;
    auto make = [&](user_data& ud, int k, error_model* p_em) -> model* {
        if (k > 1) return new gamma_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size, ud.max_root_family_size, k, -1.0, p_em);
        return new base_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size, ud.max_root_family_size, p_em);
    };
    const int rc = ref_opt::run(c, make, n_cat, optimize_epsilon, seed, values_out, n_values, score, iterations, attempts, seconds, &tb);
    if (trace_n) *trace_n = tb.n;
    return rc;
}

const char* ref_ctx_error(void* h) { return ((ref_ctx*)h)->err.c_str(); }

// Viterbi branch p-value (src/gene_family_reconstructor.cpp:390-429) restated through the matrix only:
// kept out; "next" row f2.

}  // extern "C"

// ---- the reference's fminsearch over an arbitrary C callback: pins the product's simplex search (cafe5_b200/host/nelder_mead.hpp) ----
namespace {
class callback_scorer : public optimizer_scorer {
    double (*_cb)(const double*, void*);
    void* _user;
    std::vector<double> _x0;
public:
    callback_scorer(double (*cb)(const double*, void*), void* user, const double* x0, int n) : _cb(cb), _user(user), _x0(x0, x0 + n) {}
    std::vector<double> initial_guesses() override { return _x0; }
    double calculate_score(const double* values) override { return _cb(values, _user); }
};
}

extern "C" int ref_fminsearch(double (*cb)(const double*, void*), void* user, int n, const double* x0, int max_iterations,
                              double* x_out, double* f_out, int* iterations)
{
    try {
        callback_scorer scorer(cb, user, x0, n);
        FMinSearch* pfm = fminsearch_new_with_eq(&scorer, n);
        pfm->maxiters = max_iterations > 0 ? max_iterations : 300;   // optimizer_parameters::neldermead_iterations
        std::vector<double> start(x0, x0 + n);
        fminsearch_min(pfm, start.data());
        candidate* best = get_best_result(pfm);
        for (int i = 0; i < n; ++i) x_out[i] = best->values[i];
        *f_out = best->score;
        *iterations = pfm->iters;
        fminsearch_free(pfm);
        return 0;
    } catch (std::exception&) { return 1; }
}
