// io_abi.cpp -- C entry points over cafe5_b200/host/io.hpp, part of libcafe_b200.so.  Plain host code: usable
// (and tested) without a GPU.  Strings come back tab-separated in caller-owned buffers; every function returns 0 or an error code
// and leaves the message in cafe_b200_io_last_error().
#include "../../include/cafe_b200.h"
#include "io.hpp"

#include <cstring>
#include <fstream>

namespace {
thread_local std::string g_io_error;

int put(const std::string& s, char* buf, int64_t cap)
{
    if (!buf || (int64_t)s.size() + 1 > cap) { g_io_error = "output buffer too small"; return CAFE_B200_ERR_RANGE; }
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return CAFE_B200_OK;
}

std::string join(const std::vector<std::string>& v)
{
    std::string out;
    for (size_t i = 0; i < v.size(); ++i) { if (i) out += '\t'; out += v[i]; }
    return out;
}

std::vector<std::string> ids_from(const char* tabbed, int64_t n)
{
    std::vector<std::string> ids = cafe_b200_host::split_tabs(tabbed ? tabbed : "");
    ids.resize((size_t)n);
    return ids;
}
}  // namespace

extern "C" {

const char* cafe_b200_io_last_error(void) { return g_io_error.c_str(); }

int cafe_b200_io_parse_tree(const char* newick, const char* lambda_newick, int32_t capacity, int32_t* n_nodes, int32_t* parent,
                            double* branch_length, int32_t* is_leaf, int32_t* lambda_class, int32_t* n_lambda, char* names, int64_t names_cap)
{
    try {
        if (!newick || !n_nodes) throw std::runtime_error("null argument");
        const cafe_b200_host::Tree t = cafe_b200_host::parse_newick(newick, lambda_newick ? lambda_newick : "");
        *n_nodes = t.n_nodes();
        if (t.n_nodes() > capacity) throw std::runtime_error("tree has more nodes than the caller's arrays");
        for (int i = 0; i < t.n_nodes(); ++i) {
            if (parent) parent[i] = t.parent[i];
            if (branch_length) branch_length[i] = t.branch_length[i];
            if (is_leaf) is_leaf[i] = t.is_leaf[i];
            if (lambda_class) lambda_class[i] = t.lambda_class[i];
        }
        if (n_lambda) *n_lambda = t.n_lambda;
        return names ? put(join(t.name), names, names_cap) : CAFE_B200_OK;
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_read_families(const char* path, const char* newick, int64_t* n_families, int32_t* n_species, int32_t* counts,
                               int64_t counts_cap, char* species, int64_t species_cap, char* ids, int64_t ids_cap)
{
    try {
        if (!path) throw std::runtime_error("null path");
        std::ifstream in(path);
        if (!in) throw std::runtime_error(std::string("cannot open ") + path);
        cafe_b200_host::Tree tree;
        const bool have_tree = newick && *newick;
        if (have_tree) tree = cafe_b200_host::parse_newick(newick, "");
        const cafe_b200_host::FamilyTable ft = cafe_b200_host::read_gene_families(in, have_tree ? &tree : nullptr);
        if (n_families) *n_families = (int64_t)ft.n_families();
        if (n_species) *n_species = (int32_t)ft.species.size();
        if (counts) {
            if ((int64_t)ft.counts.size() > counts_cap) throw std::runtime_error("count table larger than the caller's array");
            std::memcpy(counts, ft.counts.data(), ft.counts.size() * sizeof(int32_t));
        }
        int rc = species ? put(join(ft.species), species, species_cap) : CAFE_B200_OK;
        if (rc == CAFE_B200_OK && ids) rc = put(join(ft.ids), ids, ids_cap);
        return rc;
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_read_error_model(const char* path, double* probs, int32_t rows_cap, int32_t* rows, int32_t* max_count)
{
    try {
        if (!path) throw std::runtime_error("null path");
        std::ifstream in(path);
        if (!in) throw std::runtime_error(std::string("cannot open ") + path);
        const cafe_b200_host::ErrorModelTable em = cafe_b200_host::read_error_model(in);
        if (rows) *rows = em.rows();
        if (max_count) *max_count = em.max_count;
        if (probs) {
            if (em.rows() > rows_cap) throw std::runtime_error("error model larger than the caller's array");
            std::memcpy(probs, em.probs.data(), em.probs.size() * sizeof(double));
        }
        return CAFE_B200_OK;
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_make_prior(int32_t kind, double poisson_lambda, const char* rootdist_path, int32_t num_values, float* prior, int32_t cap,
                            int32_t* n)
{
    try {
        if (!n) throw std::runtime_error("null argument");
        std::vector<float> t;
        if (kind == 0) {
            if (num_values < 1) throw std::runtime_error("uniform prior needs a positive number of root sizes");
            t = cafe_b200_host::uniform_prior(num_values);
        } else if (kind == 1) {
            if (!rootdist_path) throw std::runtime_error("null path");
            std::ifstream in(rootdist_path);
            if (!in) throw std::runtime_error(std::string("Failed to open ") + rootdist_path);
            t = cafe_b200_host::rootdist_prior(cafe_b200_host::read_rootdist(in));
        } else if (kind == 2) {
            if (!(poisson_lambda > 0) || num_values < 1) throw std::runtime_error("Poisson prior needs lambda > 0 and a positive number of values");
            t = cafe_b200_host::poisson_prior(poisson_lambda, (size_t)num_values);
        } else throw std::runtime_error("unknown prior kind");
        *n = (int32_t)t.size();
        if (prior) {
            if ((int32_t)t.size() > cap) { g_io_error = "output buffer too small"; return CAFE_B200_ERR_RANGE; }
            std::memcpy(prior, t.data(), t.size() * sizeof(float));
        }
        return CAFE_B200_OK;
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_derive_sizes(const int32_t* counts, int64_t n, int32_t* max_family_size, int32_t* max_root_family_size)
{
    if (!counts || !max_family_size || !max_root_family_size) return CAFE_B200_ERR_ARG;
    int a = 0, b = 0;
    cafe_b200_host::derive_sizes(std::vector<int32_t>(counts, counts + n), a, b);
    *max_family_size = a;
    *max_root_family_size = b;
    return CAFE_B200_OK;
}

int cafe_b200_io_format_results(const char* model_name, double neg_lnl, const double* lambdas, int32_t n_lambda, double epsilon,
                                double longest_branch, int32_t attempts, int32_t rejects, double alpha, char* out, int64_t out_cap)
{
    try {
        std::ostringstream ost;
        cafe_b200_host::MonitorCounts mon;
        mon.attempts = attempts;
        mon.rejects = rejects;
        cafe_b200_host::write_vital_statistics(ost, model_name ? model_name : "", neg_lnl, std::vector<double>(lambdas, lambdas + n_lambda), epsilon,
                                               longest_branch, mon, alpha);
        return put(ost.str(), out, out_cap);
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_format_family_likelihoods(const char* ids_tabbed, int64_t n_families, int32_t n_cat, const double* multipliers,
                                           const double* cat_lk, const double* family_values, const double* posterior,
                                           const uint8_t* significant, int32_t what, char* out, int64_t out_cap)
{
    try {
        std::ostringstream ost;
        const std::vector<std::string> ids = ids_from(ids_tabbed, n_families);
        if (what == 0) cafe_b200_host::write_base_family_likelihoods(ost, ids, family_values);
        else if (what == 1) cafe_b200_host::write_gamma_family_likelihoods(ost, ids, n_cat, multipliers, cat_lk, family_values, posterior, significant);
        else cafe_b200_host::write_category_likelihoods(ost, ids, n_cat, multipliers, cat_lk);
        return put(ost.str(), out, out_cap);
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_format_reconstruction(const char* newick, const char* ids_tabbed, int64_t n_families, const int32_t* states,
                                       const double* pvalues, double pvalue_threshold, const double* gamma_multipliers, int32_t n_cat,
                                       const double* branch_probs, int32_t what, char* out, int64_t out_cap)
{
    try {
        if (!newick || !states) throw std::runtime_error("null argument");
        const cafe_b200_host::Tree t = cafe_b200_host::parse_newick(newick);
        const std::vector<std::string> ids = ids_from(ids_tabbed, n_families);
        std::ostringstream ost;
        switch (what) {
        case 0: cafe_b200_host::write_node_table(ost, t, ids, states, false); break;
        case 1: cafe_b200_host::write_node_table(ost, t, ids, states, true); break;
        case 2: cafe_b200_host::write_asr_trees(ost, t, ids, states, std::vector<double>(gamma_multipliers, gamma_multipliers + (gamma_multipliers ? n_cat : 0)),
                                                branch_probs, pvalue_threshold); break;
        case 3: if (!pvalues) throw std::runtime_error("p-values required"); cafe_b200_host::write_family_results(ost, ids, pvalues, pvalue_threshold); break;
        case 5: if (!branch_probs) throw std::runtime_error("branch probabilities required"); cafe_b200_host::write_branch_probabilities(ost, t, ids, branch_probs); break;
        default: cafe_b200_host::write_clade_results(ost, t, (size_t)n_families, states); break;
        }
        return put(ost.str(), out, out_cap);
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_format_report(const char* newick, const char* lambda_newick, const double* lambdas, int32_t n_lambda,
                               const char* ids_tabbed, int64_t n_families, const int32_t* states, const double* pvalues,
                               const double* branch_probs, char* out, int64_t out_cap)
{
    try {
        if (!newick || !states || !pvalues) throw std::runtime_error("null argument");
        const bool lambda_tree = lambda_newick && *lambda_newick;
        const cafe_b200_host::Tree t = cafe_b200_host::parse_newick(newick, lambda_tree ? lambda_newick : "");
        const std::vector<std::string> ids = ids_from(ids_tabbed, n_families);
        std::ostringstream ost;
        cafe_b200_host::write_report(ost, t, std::vector<double>(lambdas, lambdas + (lambdas ? n_lambda : 0)), lambda_tree, ids, states,
                                     pvalues, branch_probs);
        return put(ost.str(), out, out_cap);
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_format_simulation(const char* newick, int64_t n_families, const int32_t* node_sizes, const double* family_lambda,
                                   int32_t include_internal, char* out, int64_t out_cap)
{
    try {
        if (!newick || !node_sizes || !family_lambda || n_families < 0) throw std::runtime_error("bad argument");
        const cafe_b200_host::Tree t = cafe_b200_host::parse_newick(newick);
        std::ostringstream ost;
        cafe_b200_host::write_simulations(ost, t, (size_t)n_families, node_sizes, family_lambda, include_internal != 0);
        return put(ost.str(), out, out_cap);
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

int cafe_b200_io_format_error_model(const double* probs, int32_t rows, char* out, int64_t out_cap)
{
    try {
        if (!probs || rows <= 0) throw std::runtime_error("bad argument");
        cafe_b200_host::ErrorModelTable em;
        em.probs.assign(probs, probs + 3 * (size_t)rows);
        std::ostringstream ost;
        cafe_b200_host::write_error_model(ost, em);
        return put(ost.str(), out, out_cap);
    } catch (const std::exception& e) { g_io_error = e.what(); return CAFE_B200_ERR_ARG; }
}

}  // extern "C"
