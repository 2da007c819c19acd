// gpu_model.hpp -- drop-in models for the UNMODIFIED CAFE5 source tree (C++11, header-only).
//
// `gpu_base_model` / `gpu_gamma_model` subclass the reference's base_model / gamma_model
// (src/base_model.h, src/gamma_core.h) and override the two virtuals of the hot-path boundary,
//     double model::infer_family_likelihoods(const root_equilibrium_distribution&, const lambda*)   src/core.h:174
//     reconstruction* model::reconstruct_ancestral_states(const user_data&, const input_parameters&, matrix_cache*)  src/core.h:182
// forwarding to libcafe_b200.so through the C ABI of include/cafe_b200.h.  Everything above the boundary
// (optimizer, optimizer_scorer, estimator, reports) keeps running unchanged and reads the same side-channel
// members it reads today: model::results, gamma_model::_category_likelihoods, event_monitor counters.
//
// This file contains no reference code; it is compiled INSIDE a CAFE5 checkout (include path = CAFE5/src) and
// needs access to two private members of gamma_model (_gamma_cat_probs, _category_likelihoods): either add
// `friend class cafe_b200_shim::gpu_gamma_model;` to gamma_core.h or compile this translation unit with
// -fno-access-control (what oracle/build_ref.sh does).  See INTEGRATION.md.
#pragma once

#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "base_model.h"
#include "clade.h"
#include "error_model.h"
#include "gamma_core.h"   // NB: the reference's gamma_core.h has no include guard: include it only through this header
#include "gene_family.h"
#include "lambda.h"
#include "root_equilibrium_distribution.h"
#include "user_data.h"

#include "cafe_b200.h"

namespace cafe_b200_shim {

// Owns one cafe_b200_ctx built from the reference's own objects.
class device_context {
public:
    // devices: CUDA ordinals; more than one shards the families over them (cafe_b200_create_multi)
    device_context(const clade* p_tree, const lambda* p_lambda, const std::vector<gene_family>& families,
                   int max_family_size, int max_root_family_size, const std::vector<int>& devices)
        : _max_family_size(max_family_size), _max_root_family_size(max_root_family_size)
    {
        _order.assign(p_tree->reverse_level_begin(), p_tree->reverse_level_end());
        const int n = int(_order.size());
        std::map<const clade*, int> index;
        for (int i = 0; i < n; ++i) index[_order[i]] = i;
        std::vector<int32_t> parent(n), leaf_col(n), lambda_class(n, 0);
        std::vector<double> branch_length(n);
        std::vector<std::string> species;
        for (int i = 0; i < n; ++i) {
            const clade* c = _order[i];
            parent[i] = c->is_root() ? -1 : index.at(c->get_parent());
            branch_length[i] = c->get_branch_length();
            leaf_col[i] = -1;
            if (c->is_leaf()) { leaf_col[i] = int(species.size()); species.push_back(c->get_taxon_name()); }
        }
        // lambda class of every node: probe a clone of the caller's lambda with distinct values
        _n_lambda = p_lambda ? p_lambda->count() : 1;
        if (p_lambda && _n_lambda > 1) {
            std::unique_ptr<lambda> probe(p_lambda->clone());
            std::vector<double> marks(_n_lambda);
            for (int k = 0; k < _n_lambda; ++k) marks[k] = double(k + 1);
            probe->update(marks.data());
            for (int i = 0; i < n; ++i) lambda_class[i] = int(probe->get_value_for_clade(_order[i])) - 1;
        }
        _n_families = families.size();
        std::vector<int32_t> counts(_n_families * species.size());
        for (size_t f = 0; f < _n_families; ++f)
            for (size_t j = 0; j < species.size(); ++j) counts[f * species.size() + j] = families[f].get_species_size(species[j]);
        cafe_b200_tree t{n, parent.data(), branch_length.data(), leaf_col.data(), lambda_class.data()};
        std::vector<int32_t> devs(devices.begin(), devices.end());
        if (devs.empty()) devs.push_back(0);
        if (cafe_b200_create_multi(&t, counts.data(), int64_t(_n_families), int32_t(species.size()), max_family_size, max_root_family_size,
                                   devs.data(), int32_t(devs.size()), &_ctx) != CAFE_B200_OK)
            throw std::runtime_error(std::string("cafe_b200_create_multi: ") + cafe_b200_last_error(nullptr));
    }
    ~device_context() { cafe_b200_destroy(_ctx); }
    device_context(const device_context&) = delete;
    device_context& operator=(const device_context&) = delete;

    void check(int rc, const char* what) const
    {
        if (rc != CAFE_B200_OK) throw std::runtime_error(std::string(what) + ": " + cafe_b200_last_error(_ctx));
    }

    // Parameters arrive by mutation of shared objects (lambda::update, set_alpha, replace_epsilons): re-read them every call.
    void sync_inputs(const root_equilibrium_distribution& prior, const error_model* em)
    {
        std::vector<float> p(std::max(_max_root_family_size, _max_family_size) + 2);
        for (size_t j = 0; j < p.size(); ++j) p[j] = prior.compute(j);
        check(cafe_b200_set_prior(_ctx, p.data(), int32_t(p.size())), "cafe_b200_set_prior");
        if (em) {
            if (em->n_deviations() != 3) throw std::runtime_error("cafe_b200: only the -1 0 1 error classes are supported");
            const int rows = int(em->get_max_family_size());
            std::vector<double> probs(size_t(rows) * 3);
            for (int i = 0; i < rows; ++i) {
                auto r = em->get_probs(i);
                for (int d = 0; d < 3; ++d) probs[size_t(i) * 3 + d] = r[d];
            }
            check(cafe_b200_set_error_model(_ctx, probs.data(), rows, rows - 1), "cafe_b200_set_error_model");
        } else {
            check(cafe_b200_set_error_model(_ctx, nullptr, 0, 0), "cafe_b200_set_error_model");
        }
    }

    cafe_b200_ctx* get() const { return _ctx; }
    size_t n_families() const { return _n_families; }
    const std::vector<const clade*>& order() const { return _order; }

private:
    cafe_b200_ctx* _ctx = nullptr;
    std::vector<const clade*> _order;
    size_t _n_families = 0;
    int _n_lambda = 1;
    int _max_family_size, _max_root_family_size;
};

class gpu_base_model : public base_model {
    std::unique_ptr<device_context> _dev;
    std::vector<int> _devices;

    device_context& dev(const lambda* p_lambda)
    {
        if (!_dev) _dev.reset(new device_context(_p_tree, p_lambda, *_p_gene_families, _max_family_size, _max_root_family_size, _devices));
        return *_dev;
    }

public:
    gpu_base_model(lambda* p_lambda, const clade* p_tree, const std::vector<gene_family>* p_gene_families, int max_family_size,
                   int max_root_family_size, error_model* p_error_model, const std::vector<int>& devices = std::vector<int>(1, 0))
        : base_model(p_lambda, p_tree, p_gene_families, max_family_size, max_root_family_size, p_error_model), _devices(devices)
    {
    }

    double infer_family_likelihoods(const root_equilibrium_distribution& prior, const lambda* p_lambda) override
    {
        _monitor.Event_InferenceAttempt_Started();
        if (!_p_lambda->is_valid()) {
            _monitor.Event_InferenceAttempt_InvalidValues();
            return -log(0);
        }
        device_context& d = dev(p_lambda);
        d.sync_inputs(prior, _p_error_model);
        std::vector<double> lambdas = get_lambda_values(p_lambda);
        std::vector<double> family_lnl(d.n_families());
        double score = 0;
        d.check(cafe_b200_eval_base(d.get(), lambdas.data(), int32_t(lambdas.size()), &score, family_lnl.data()), "cafe_b200_eval_base");
        results.resize(d.n_families());
        if (!std::isinf(score))
            for (size_t i = 0; i < d.n_families(); ++i)
                results[i] = family_info_stash(_p_gene_families->at(i).id(), 0.0, 0.0, 0.0, family_lnl[i], false);
        return score;
    }

    reconstruction* reconstruct_ancestral_states(const user_data& ud, const input_parameters& ui, matrix_cache*) override
    {
        device_context d(_p_tree, _p_lambda, ud.gene_families, _max_family_size, _max_root_family_size, _devices);
        d.sync_inputs(ud.prior, nullptr);
        std::vector<double> lambdas = get_lambda_values(_p_lambda);
        const size_t n = d.order().size();
        std::vector<int32_t> states(d.n_families() * n);
        d.check(cafe_b200_reconstruct(d.get(), lambdas.data(), int32_t(lambdas.size()), nullptr, nullptr, 0, nullptr, states.data(), nullptr),
                "cafe_b200_reconstruct");
        auto result = new base_model_reconstruction(ud, ui);
        for (size_t f = 0; f < d.n_families(); ++f) {
            clademap<int>& rc = result->_reconstructions[ud.gene_families[f].id()];
            for (size_t i = 0; i < n; ++i) rc[d.order()[i]] = d.order()[i]->is_leaf() ? 0 : states[f * n + i];
        }
        return result;
    }
};

class gpu_gamma_model : public gamma_model {
    std::unique_ptr<device_context> _dev;
    std::vector<int> _devices;

    device_context& dev(const lambda* p_lambda)
    {
        if (!_dev) _dev.reset(new device_context(_p_tree, p_lambda, *_p_gene_families, _max_family_size, _max_root_family_size, _devices));
        return *_dev;
    }

public:
    gpu_gamma_model(lambda* p_lambda, clade* p_tree, std::vector<gene_family>* p_gene_families, int max_family_size,
                    int max_root_family_size, int n_gamma_cats, double fixed_alpha, error_model* p_error_model,
                    const std::vector<int>& devices = std::vector<int>(1, 0))
        : gamma_model(p_lambda, p_tree, p_gene_families, max_family_size, max_root_family_size, n_gamma_cats, fixed_alpha, p_error_model),
          _devices(devices)
    {
    }

    double infer_family_likelihoods(const root_equilibrium_distribution& prior, const lambda* p_lambda) override
    {
        _monitor.Event_InferenceAttempt_Started();
        results.clear();
        if (!can_infer()) {
            _monitor.Event_InferenceAttempt_InvalidValues();
            return -log(0);
        }
        device_context& d = dev(p_lambda);
        d.sync_inputs(prior, _p_error_model);
        std::vector<double> lambdas = get_lambda_values(p_lambda);
        const std::vector<double> multipliers = get_lambda_multipliers();
        const std::vector<double>& cat_probs = _gamma_cat_probs;
        const int K = int(multipliers.size());
        const size_t F = d.n_families();
        std::vector<double> cat_lk(F * K), family_lk(F), posterior(F * K);
        std::vector<uint8_t> significant(F * K), failed(F);
        int64_t n_failed = 0;
        double score = 0;
        d.check(cafe_b200_eval_gamma(d.get(), lambdas.data(), int32_t(lambdas.size()), get_alpha(), multipliers.data(), cat_probs.data(), K,
                                     &score, cat_lk.data(), family_lk.data(), posterior.data(), significant.data(), failed.data(), &n_failed),
                "cafe_b200_eval_gamma");
        for (size_t i = 0; i < F; ++i) _category_likelihoods[i].assign(cat_lk.begin() + i * K, cat_lk.begin() + (i + 1) * K);
        if (n_failed > 0) {
            for (size_t i = 0; i < F; ++i)
                if (failed[i]) _monitor.Event_InferenceAttempt_Saturation(_p_gene_families->at(i).id());
            return -log(0);
        }
        if (std::isinf(score)) return score;
        for (size_t i = 0; i < F; ++i)
            for (int k = 0; k < K; ++k)
                results.push_back(family_info_stash(_p_gene_families->at(i).id(), multipliers[k], cat_lk[i * K + k], family_lk[i],
                                                    posterior[i * K + k], significant[i * K + k] != 0));
        return score;
    }

    reconstruction* reconstruct_ancestral_states(const user_data& ud, const input_parameters& ui, matrix_cache*) override
    {
        device_context d(_p_tree, _p_lambda, ud.gene_families, _max_family_size, _max_root_family_size, _devices);
        d.sync_inputs(ud.prior, nullptr);
        std::vector<double> lambdas = get_lambda_values(_p_lambda);
        const std::vector<double> multipliers = get_lambda_multipliers();
        const int K = int(multipliers.size());
        const size_t n = d.order().size(), F = d.n_families();
        std::vector<int32_t> cat_states(F * K * n), states(F * n);
        std::vector<double> averaged(F * n);
        d.check(cafe_b200_reconstruct(d.get(), lambdas.data(), int32_t(lambdas.size()), multipliers.data(), _gamma_cat_probs.data(), K,
                                      cat_states.data(), states.data(), averaged.data()), "cafe_b200_reconstruct");
        auto result = new gamma_model_reconstruction(ud, ui, multipliers);
        for (size_t f = 0; f < F; ++f) {
            auto& rec = result->_reconstructions[ud.gene_families[f].id()];
            rec._category_likelihoods = _category_likelihoods[f];
            rec.category_reconstruction.resize(K);
            for (size_t i = 0; i < n; ++i) {
                const clade* c = d.order()[i];
                if (c->is_leaf()) continue;
                rec.reconstruction[c] = averaged[f * n + i];
                for (int k = 0; k < K; ++k) rec.category_reconstruction[k][c] = cat_states[(f * K + k) * n + i];
            }
        }
        return result;
    }
};

}  // namespace cafe_b200_shim
