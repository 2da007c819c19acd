// oracle/ref_gpu_model.cpp -- TEST / MEASUREMENT INFRASTRUCTURE, built into its OWN library oracle/_ref/libcafe_ref_shim.so so
// that libcafe_ref.so (the unmodified reference, the CPU baseline of bench.py) never maps the product library.
// Compiles the product's drop-in shim (cafe5_b200/host/gpu_model.hpp) against the UNMODIFIED reference headers and exposes
// `ref_optimize_gpu`: the reference's own optimizer + optimizer_scorer + Nelder-Mead (src/optimizer.cpp:540-569,
// src/optimizer_scorer.cpp:22-36) driving the CUDA models through the same virtual call the CPU models answer in
// ref_driver.cpp::ref_optimize.  Same seeded randomizer_engine, same trace format.
#include "ref_optimize.hpp"

#include "../cafe5_b200/host/gpu_model.hpp"

extern "C" {

// devices[n_devices]: CUDA ordinals the models shard their families over (one entry: a single GPU).
int ref_optimize_gpu(void* h, int n_cat, int optimize_epsilon, unsigned seed, const int* devices, int n_devices, double* values_out,
                     int* n_values, double* score, int* iterations, int* attempts, double* seconds, double* trace_values,
                     double* trace_scores, int* trace_failed_family, int* trace_n_failed, int trace_cap, int* trace_n)
{
    auto c = (ref_ctx*)h;
    ref_opt::trace_buffer tb;
    tb.values = trace_values; tb.scores = trace_scores; tb.failed_family = trace_failed_family; tb.n_failed = trace_n_failed;
    tb.cap = trace_values ? trace_cap : 0;
    const std::vector<int> devs(devices, devices + (n_devices > 0 ? n_devices : 0));
    auto make = [&](user_data& ud, int k, error_model* p_em) -> model* {
        if (k > 1) return new cafe_b200_shim::gpu_gamma_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size,
                                                              ud.max_root_family_size, k, -1.0, p_em, devs);
        return new cafe_b200_shim::gpu_base_model(nullptr, c->tree.get(), &ud.gene_families, ud.max_family_size, ud.max_root_family_size,
                                                  p_em, devs);
    };
    const int rc = ref_opt::run(c, make, n_cat, optimize_epsilon, seed, values_out, n_values, score, iterations, attempts, seconds, &tb);
    if (trace_n) *trace_n = tb.n;
    return rc;
}

}  // extern "C"
