// fit.cpp -- host driver of a parameter search: scorers + Nelder-Mead over the C ABI.
//
// What the reference does above the hot-path boundary, restated from its behaviour (no reference code):
//   optimizer::optimize                 src/optimizer.cpp:540-569   initial guess, simplex search, result
//   optimizer::get_initial_guesses      src/optimizer.cpp:347-365   one retry when the first guess scores +inf, then give up
//   lambda_optimizer                    src/optimizer_scorer.cpp:38-66    values = one lambda per class
//   lambda_epsilon_optimizer            src/optimizer_scorer.cpp:68-107   values = lambdas..., epsilon (base model, `-e` without a file)
//   gamma_lambda_optimizer              src/optimizer_scorer.cpp:147-180  values = lambdas..., alpha
//   gamma_optimizer                     src/optimizer_scorer.cpp:109-145  values = alpha (lambda fixed)
//   inference_optimizer_scorer::calculate_score  :22-36   NaN -> +inf
// Every evaluation is one call of cafe_b200_eval_base / cafe_b200_eval_gamma; nothing else touches the GPU.
// Initial guesses draw from std::mt19937 exactly like the reference's scorers (normal(0.002 * longest, 0.2) / longest per lambda,
// redrawn while negative; gamma(4, 0.25) for alpha; the default error model's epsilon 0.05), so a run seeded like the reference's
// `randomizer_engine` starts from the same point.
#include "../../include/cafe_b200.h"
#include "discrete_gamma.hpp"
#include "nelder_mead.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <algorithm>
#include <cctype>
#include <random>
#include <string>
#include <vector>

namespace {

using namespace cafe_b200_host;

struct Problem {
    cafe_b200_ctx* ctx;
    cafe_b200_fit_options opt;
    int n_lambda = 1, max_family_size = 0;
    double longest_branch = 1.0;
    std::vector<double> fixed_lambdas;
    int evaluations = 0;
    int hard_error = 0;
    std::vector<double> cat_probs, multipliers;

    bool gamma() const { return opt.n_cat > 1; }
    bool fit_lambda() const { return opt.fixed_lambdas == nullptr; }
    bool fit_alpha() const { return gamma() && !(opt.fixed_alpha > 0); }
    bool fit_epsilon() const { return !gamma() && opt.optimize_epsilon != 0; }
    int n_values() const { return (fit_lambda() ? n_lambda : 0) + (fit_alpha() ? 1 : 0) + (fit_epsilon() ? 1 : 0); }

    // error_model::replace_epsilons (src/error_model.cpp:79-109) for the single-epsilon default model
    void set_epsilon(double eps)
    {
        const int rows = max_family_size + 1;
        std::vector<double> probs((size_t)rows * 3);
        probs[0] = 0.0; probs[1] = 1 - eps; probs[2] = eps;
        for (int i = 1; i < rows; ++i) { probs[i * 3] = eps; probs[i * 3 + 1] = 1 - (eps * 2); probs[i * 3 + 2] = eps; }
        if (cafe_b200_set_error_model(ctx, probs.data(), rows, max_family_size) != CAFE_B200_OK) hard_error = 1;
    }

    double score(const double* values)
    {
        ++evaluations;
        const double* lambdas = fit_lambda() ? values : fixed_lambdas.data();
        const double* rest = values + (fit_lambda() ? n_lambda : 0);
        double neg = std::numeric_limits<double>::infinity();
        int rc;
        if (gamma()) {
            const double alpha = fit_alpha() ? rest[0] : opt.fixed_alpha;
            discrete_gamma(opt.n_cat, alpha, cat_probs, multipliers);   // gamma_model::set_alpha (src/gamma_core.cpp:61-67)
            int64_t n_failed = 0;
            rc = cafe_b200_eval_gamma(ctx, lambdas, n_lambda, alpha, multipliers.data(), cat_probs.data(), opt.n_cat, &neg,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, &n_failed);
        } else {
            if (fit_epsilon()) set_epsilon(rest[0]);
            rc = cafe_b200_eval_base(ctx, lambdas, n_lambda, &neg, nullptr);
        }
        if (rc != CAFE_B200_OK) { hard_error = rc; return std::numeric_limits<double>::infinity(); }
        if (std::isnan(neg)) neg = std::numeric_limits<double>::infinity();
        return neg;
    }

    std::vector<double> initial_guess(std::mt19937& engine)
    {
        std::vector<double> v;
        if (fit_lambda()) {
            const double distmean = 0.002 / (1.0 / longest_branch);
            std::normal_distribution<double> distribution(distmean, 0.2);
            for (int i = 0; i < n_lambda; ++i) {
                double x = 1.0 / longest_branch * distribution(engine);
                while (x < 0) x = 1.0 / longest_branch * distribution(engine);
                v.push_back(x);
            }
        }
        if (fit_alpha()) {
            std::gamma_distribution<double> distribution(4.0, 0.25);
            v.push_back(distribution(engine));
        }
        if (fit_epsilon()) v.push_back(0.05);
        return v;
    }
};

}  // namespace

extern "C" int cafe_b200_discrete_gamma(int32_t n_cat, double alpha, double* cat_probs, double* multipliers)
{
    if (n_cat < 1 || !cat_probs || !multipliers) return CAFE_B200_ERR_ARG;
    std::vector<double> p, m;
    cafe_b200_host::discrete_gamma(n_cat, alpha, p, m);
    std::memcpy(cat_probs, p.data(), n_cat * sizeof(double));
    std::memcpy(multipliers, m.data(), n_cat * sizeof(double));
    return CAFE_B200_OK;
}

extern "C" int cafe_b200_minimize(double (*objective)(const double*, void*), void* user, int32_t n, const double* x0,
                                  int32_t max_iterations, double* x_out, double* f_out, int32_t* iterations)
{
    if (!objective || n < 1 || !x0 || !x_out) return CAFE_B200_ERR_ARG;
    cafe_b200_host::NelderMeadOptions o;
    if (max_iterations > 0) o.max_iters = max_iterations;
    cafe_b200_host::NelderMead nm([&](const double* x) { return objective(x, user); }, n, o);
    cafe_b200_host::NelderMeadResult r = nm.minimize(x0);
    std::memcpy(x_out, r.x.data(), n * sizeof(double));
    if (f_out) *f_out = r.f;
    if (iterations) *iterations = r.iterations;
    return CAFE_B200_OK;
}

// `-p` without a value: the Poisson mean of the root prior estimated from the gene families themselves
// (root_equilibrium_distribution(gene_families, num_values), src/root_equilibrium_distribution.cpp:42-54; poisson_scorer,
// src/poisson.cpp:21-78).  Every positive leaf count c contributes log pdf(c - 1; lambda), pdf(x; l) = exp(x log l - lgamma(x + 1) - l);
// terms whose pdf is 0, inf or NaN are skipped; lambda < 0 scores +inf.  The terms are added family by family, and inside a family in
// the order of the reference's species map (names compared case-insensitively) when the names are given, so that the sum -- and with
// it the simplex trajectory -- is the reference's to the last bit.  Start: one uniform(0, 1) draw of the seeded engine, one retry
// (optimizer::get_initial_guesses).  Host only.
extern "C" int cafe_b200_fit_poisson_prior(const int32_t* counts, int64_t n_families, int32_t n_species, const char* species, uint32_t seed,
                                           double* poisson_lambda, double* neg_lnl, int32_t* iterations)
{
    if (!counts || n_families < 1 || n_species < 1 || !poisson_lambda) return CAFE_B200_ERR_ARG;
    std::vector<int> order(n_species);
    for (int j = 0; j < n_species; ++j) order[j] = j;
    if (species && *species) {
        std::vector<std::string> names(1);
        for (const char* c = species; *c; ++c) {
            if (*c == '\t') names.emplace_back();
            else names.back().push_back((char)std::tolower((unsigned char)*c));
        }
        if ((int)names.size() != n_species) return CAFE_B200_ERR_ARG;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {   // ci_less, src/gene_family.h:10-25
            return std::lexicographical_compare(names[a].begin(), names[a].end(), names[b].begin(), names[b].end(),
                                                [](char x, char y) { return (unsigned char)x < (unsigned char)y; });
        });
    }
    std::vector<int> sizes;                    // poisson_scorer::leaf_family_sizes
    for (int64_t f = 0; f < n_families; ++f)
        for (int j = 0; j < n_species; ++j) {
            const int32_t c = counts[(size_t)f * n_species + order[j]];
            if (c > 0) sizes.push_back(c - 1);
        }
    auto score = [&](const double* v) {
        const double lambda = v[0];
        if (lambda < 0) return std::numeric_limits<double>::infinity();
        double sum = 0.0;
        for (int sz : sizes) {
            const double ll = std::exp(sz * std::log(lambda) - std::lgamma((double)(sz + 1)) - lambda);
            if (std::isnan(ll) || std::isinf(ll) || ll == 0) continue;
            sum += std::log(ll);
        }
        return -sum;
    };
    std::mt19937 engine(seed);
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    double start = distribution(engine);
    double first = score(&start);
    for (int attempt = 0; std::isinf(first) && attempt < 1; ++attempt) {
        start = distribution(engine);
        first = score(&start);
    }
    if (std::isinf(first)) {                   // OptimizerInitializationFailure
        *poisson_lambda = start;
        if (neg_lnl) *neg_lnl = first;
        if (iterations) *iterations = 0;
        return CAFE_B200_ERR_STATE;
    }
    NelderMead nm(score, 1, NelderMeadOptions());
    const NelderMeadResult r = nm.minimize(&start);
    *poisson_lambda = r.x[0];
    if (neg_lnl) *neg_lnl = r.f;
    if (iterations) *iterations = r.iterations;
    return CAFE_B200_OK;
}

extern "C" int cafe_b200_fit(cafe_b200_ctx* ctx, const cafe_b200_fit_options* options, cafe_b200_fit_result* result)
{
    if (!ctx || !options || !result) return CAFE_B200_ERR_ARG;
    std::memset(result, 0, sizeof *result);
    Problem pb;
    pb.ctx = ctx;
    pb.opt = *options;
    int64_t n_families = 0;
    int32_t n_nodes = 0, n_classes = 1, mfs = 0, mrs = 0;
    if (cafe_b200_describe(ctx, &n_families, &n_nodes, &n_classes, &mfs, &mrs, &pb.longest_branch) != CAFE_B200_OK) return CAFE_B200_ERR_ARG;
    pb.n_lambda = n_classes;
    pb.max_family_size = mfs;
    if (options->fixed_lambdas) pb.fixed_lambdas.assign(options->fixed_lambdas, options->fixed_lambdas + n_classes);
    const int n = pb.n_values();
    if (n < 1 || n > CAFE_B200_FIT_MAX_VALUES) return CAFE_B200_ERR_ARG;   // nothing to optimise (get_lambda_optimizer returns nullptr)

    const auto t0 = std::chrono::steady_clock::now();
    std::vector<double> start;
    if (options->start) start.assign(options->start, options->start + n);
    else {
        std::mt19937 engine(options->seed);
        start = pb.initial_guess(engine);
        double first = pb.score(start.data());
        for (int attempt = 0; std::isinf(first) && attempt < 1; ++attempt) {   // NUM_OPTIMIZER_INITIALIZATION_ATTEMPTS = 1
            start = pb.initial_guess(engine);
            first = pb.score(start.data());
        }
        if (pb.hard_error) return pb.hard_error;
        if (std::isinf(first)) {   // OptimizerInitializationFailure
            result->n_values = n;
            result->neg_lnl = first;
            result->evaluations = pb.evaluations;
            result->status = 1;
            return CAFE_B200_OK;
        }
    }
    NelderMeadOptions nm_opt;
    if (options->max_iterations > 0) nm_opt.max_iters = options->max_iterations;
    NelderMead nm([&](const double* x) { return pb.score(x); }, n, nm_opt);
    NelderMeadResult r = nm.minimize(start.data());
    if (pb.hard_error) return pb.hard_error;
    // scorer->finalize(): leave the context's error model at the fitted epsilon
    if (pb.fit_epsilon()) pb.set_epsilon(r.x[n - 1]);
    result->n_values = n;
    for (int i = 0; i < n; ++i) result->values[i] = r.x[i];
    result->neg_lnl = r.f;
    result->iterations = r.iterations;
    result->evaluations = pb.evaluations;
    result->status = r.hit_max ? 2 : 0;
    result->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return CAFE_B200_OK;
}
