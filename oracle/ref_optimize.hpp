// oracle/ref_optimize.hpp -- TEST / MEASUREMENT INFRASTRUCTURE.
// Runs the reference's own optimizer + optimizer_scorer + Nelder-Mead (src/optimizer.cpp:540-569,
// src/optimizer_scorer.cpp:22-36) over a model built by the caller's factory: the reference's CPU models (ref_driver.cpp,
// libcafe_ref.so) or the CUDA drop-in models of cafe5_b200/host/gpu_model.hpp (ref_gpu_model.cpp, libcafe_ref_shim.so).
// Both start from the same seeded randomizer_engine, so fitted values and the whole evaluation trace can be compared.
#pragma once
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "easylogging++.h"
#include "core.h"
#include "optimizer.h"
#include "optimizer_scorer.h"

#include "ref_ctx.hpp"

extern std::mt19937 randomizer_engine;

namespace ref_opt {

// Every call the optimizer makes goes through here: (values, score, which family failed) of each attempt, in order.
struct trace_buffer {
    double* values = nullptr;      // [cap x n_values]
    double* scores = nullptr;      // [cap]
    int* failed_family = nullptr;  // [cap] index of a family whose likelihood was 0 in some category (-1: none)
    int* n_failed = nullptr;       // [cap] number of such families
    int cap = 0, n = 0;
};

class tracing_scorer : public optimizer_scorer {
    inference_optimizer_scorer* _inner;
    model* _model;
    trace_buffer* _trace;
    std::map<std::string, int> _seen;
    size_t _n_values = 0;
public:
    tracing_scorer(inference_optimizer_scorer* inner, model* m, trace_buffer* t) : _inner(inner), _model(m), _trace(t) {}
    std::vector<double> initial_guesses() override
    {
        auto g = _inner->initial_guesses();
        _n_values = g.size();
        return g;
    }
    double calculate_score(const double* values) override
    {
        const double s = _inner->calculate_score(values);
        if (_trace && _trace->n < _trace->cap) {
            const int i = _trace->n++;
            for (size_t j = 0; j < _n_values; ++j) _trace->values[i * _n_values + j] = values[j];
            _trace->scores[i] = s;
            int first = -1, cnt = 0;
            for (auto& kv : _model->get_monitor().failure_count) {    // -fno-access-control: event_monitor keeps this private
                const int before = _seen[kv.first];
                if (kv.second > before) {
                    ++cnt;
                    const int id = std::atoi(kv.first.c_str());     // ref_ctx_create names family f "f"
                    if (first < 0 || id < first) first = id;
                    _seen[kv.first] = kv.second;
                }
            }
            _trace->failed_family[i] = first;
            _trace->n_failed[i] = cnt;
        }
        return s;
    }
};

// n_cat: 0/1 = base model, > 1 = gamma model with alpha estimated.  optimize_epsilon: base model with the default error model and
// epsilon as a free parameter (`-e` without a file, src/core.cpp:39-45).  values_out: fitted parameters (lambdas..., alpha | epsilon).
template <typename MakeModel>
int run(ref_ctx* c, MakeModel make_model, int n_cat, int optimize_epsilon, unsigned seed, double* values_out, int* n_values,
        double* score, int* iterations, int* attempts, double* seconds, trace_buffer* trace)
{
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
        std::unique_ptr<model> m(make_model(ud, n_cat, p_em));
        std::unique_ptr<inference_optimizer_scorer> scorer(m->get_lambda_optimizer(ud));
        if (!scorer) { c->err = "nothing to optimise"; return 2; }
        scorer->quiet = true;
        tracing_scorer traced(scorer.get(), m.get(), trace);
        optimizer opt(&traced);
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

}  // namespace ref_opt
