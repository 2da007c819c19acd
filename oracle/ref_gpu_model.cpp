// oracle/ref_gpu_model.cpp -- TEST / MEASUREMENT INFRASTRUCTURE.
// Compiles the product's drop-in shim (cafe5_b200/host/gpu_model.hpp) against the UNMODIFIED reference
// headers and exposes `ref_optimize`: the reference's own optimizer + optimizer_scorer + Nelder-Mead
// (src/optimizer.cpp:540-569, src/optimizer_scorer.cpp:22-36) driving EITHER the reference's CPU models or
// the CUDA models through the same virtual call.  Both backends start from the same seeded
// randomizer_engine, so fitted lambda / alpha / epsilon and the trajectory length can be compared.
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "easylogging++.h"
#include "error_model.h"
#include "lambda.h"
#include "optimizer.h"
#include "optimizer_scorer.h"
#include "user_data.h"
#include "io.h"

#include "../cafe5_b200/host/gpu_model.hpp"

extern std::mt19937 randomizer_engine;

// mirrors the private struct of ref_driver.cpp (same layout, same translation-unit family)
struct ref_ctx_view {
    std::unique_ptr<clade> tree;
    std::unique_ptr<clade> lambda_tree;
    std::vector<const clade*> order;
    int max_family_size, max_root_family_size;
    std::unique_ptr<error_model> em;
    user_data ud;
    input_parameters ui;
    std::string err;
};

extern "C" {

// backend: 0 = reference CPU models, 1 = CUDA models (gpu_model.hpp).  n_cat: 0/1 = base model, > 1 = gamma model
// with alpha estimated.  optimize_epsilon: base model with the default error model and epsilon as a free
// parameter (`-e` without a file, src/core.cpp:39-45).  values_out: fitted parameters (lambdas..., alpha | epsilon).
int ref_optimize(void* h, int backend, int n_cat, int optimize_epsilon, unsigned seed, int device,
                 double* values_out, int* n_values, double* score, int* iterations, int* attempts, double* seconds)
{
    auto c = (ref_ctx_view*)h;
    try {
        randomizer_engine.seed(seed);
        user_data& ud = c->ud;
        ud.p_lambda = nullptr;
        ud.p_lambda_tree = c->lambda_tree.get();
        std::unique_ptr<error_model> em;
        error_model* p_em = c->em.get();
        if (optimize_epsilon) {
            em.reset(new error_model());
            em->set_probabilities(0, {0, .95, 0.05});
            em->set_probabilities(ud.max_family_size, {0.05, .9, 0.05});
            p_em = em.get();
            ud.p_error_model = nullptr;            // "no file given" -> epsilon is estimated (base_model.cpp:121-124)
        } else {
            ud.p_error_model = p_em;
        }
        std::unique_ptr<model> m;
        if (n_cat > 1) {
            if (backend) m.reset(new cafe_b200_shim::gpu_gamma_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size,
                                                                     ud.max_root_family_size, n_cat, -1.0, p_em, device));
            else m.reset(new gamma_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size, ud.max_root_family_size,
                                         n_cat, -1.0, p_em));
        } else {
            if (backend) m.reset(new cafe_b200_shim::gpu_base_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size,
                                                                    ud.max_root_family_size, p_em, device));
            else m.reset(new base_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size, ud.max_root_family_size, p_em));
        }
        std::unique_ptr<inference_optimizer_scorer> scorer(m->get_lambda_optimizer(ud));
        if (!scorer) { c->err = "nothing to optimise"; return 2; }
        scorer->quiet = true;
        optimizer opt(scorer.get());
        opt.quiet = true;
        optimizer_parameters params;
        auto t0 = std::chrono::steady_clock::now();
        auto result = opt.optimize(params);
        auto t1 = std::chrono::steady_clock::now();
        scorer->finalize(&result.values[0]);
        *n_values = int(result.values.size());
        for (size_t i = 0; i < result.values.size(); ++i) values_out[i] = result.values[i];
        *score = result.score;
        *iterations = result.num_iterations;
        *attempts = m->get_monitor().attempts;
        *seconds = std::chrono::duration<double>(t1 - t0).count();
        delete m->get_lambda();
        return 0;
    } catch (std::exception& e) { c->err = e.what(); return 1; }
}

const char* ref_ctx_error(void* h) { return ((ref_ctx_view*)h)->err.c_str(); }

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
