/* cafe_b200.h -- C ABI of the B200-native CAFE5 likelihood hot path (libcafe_b200.so).
 *
 * Drop-in boundary (reference paths relative to /root/reference):
 *   model::infer_family_likelihoods(prior, lambda)        src/core.h:174
 *     base_model  implementation                          src/base_model.cpp:53-100
 *     gamma_model implementation                          src/gamma_core.cpp:168-237
 *   model::reconstruct_ancestral_states(...)              src/core.h:182
 *     base / gamma implementations                        src/base_model.cpp:133-170, src/gamma_core.cpp:290-339
 * Everything at or below those calls (matrix_cache, birthdeath_rate_with_log_alpha,
 * inference_prune / compute_node_probability, root-prior weighting, gamma mixture,
 * error-model leaf emission, Pupko reconstruction) runs on the GPU behind this header;
 * everything above (CLI, model classes, scorers, Nelder-Mead) stays on the host.
 * INTEGRATION.md shows the `class gpu_model : public model` shim that binds these entry points
 * into the unmodified reference.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns a status (0 = OK) and
 *     cafe_b200_last_error() describes the last failure of that context (or of create).
 *   - numerical failure is NOT an error: invalid lambda / saturated / underflowed families yield
 *     neg_lnl = +inf exactly as the reference returns -log(0) (base_model.cpp:56-60,
 *     gamma_core.cpp:174-178,216-225).
 *   - caller owns every output buffer; NULL means "not wanted".
 *   - one context per host thread; a context is bound to one CUDA device (cafe_b200_create) or shards its families over
 *     several devices of the node (cafe_b200_create_multi).
 *   - there is no CPU fallback: create fails with CAFE_B200_ERR_CUDA when no device is usable.
 *   - state-space limit: the default kernels hold the whole state space in one 208-row pass (max(max_family_size,
 *     max_root_family_size) <= 207: every bundled data set, up to counts of ~165); larger spaces take the streamed-operand kernels (tested
 *     up to max_family_size 720, where the reference switches to compute_node_probability_large_families at 1000,
 *     src/probability.cpp:317-330); create returns CAFE_B200_ERR_RANGE once the narrowest Pupko tile no longer fits shared memory
 *     (max_family_size beyond roughly 1,400).
 *
 * Tree layout: nodes in the reference's reverse level order (src/clade.cpp:69-100): children
 * precede parents, the root is the LAST node; the reference's descendant order of a node is
 * DEcreasing node index (products over children are formed in that order, probability.cpp:209-218).
 */
#ifndef CAFE_B200_H
#define CAFE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cafe_b200_ctx cafe_b200_ctx;

enum {
    CAFE_B200_OK = 0,
    CAFE_B200_ERR_ARG = 1,     /* malformed argument (tree order, sizes, NULL where required) */
    CAFE_B200_ERR_CUDA = 2,    /* CUDA runtime failure or no usable device */
    CAFE_B200_ERR_RANGE = 3,   /* a count exceeds max_family_size / state space too large for the kernels */
    CAFE_B200_ERR_STATE = 4    /* call sequence error (e.g. eval before set_prior) */
};

/* Flattened species tree; replaces `const clade*` (src/clade.h) at the boundary. */
typedef struct {
    int32_t n_nodes;
    const int32_t* parent;        /* [n_nodes] parent index, -1 for the root (must be node n_nodes-1) */
    const double* branch_length;  /* [n_nodes] clade::get_branch_length(); root's own length ignored */
    const int32_t* leaf_col;      /* [n_nodes] column of `counts` for a leaf, -1 for internal nodes */
    const int32_t* lambda_class;  /* [n_nodes] 0-based lambda class of the node (lambda tree, src/lambda.cpp:32-40); all 0 for a single lambda */
} cafe_b200_tree;

/* Replaces model::model (src/core.cpp:60-71): borrows nothing, copies the tree and the count table
 * (gene_family::get_species_size, src/gene_family.cpp:38-45) to the device, builds the identical-family
 * reference list (build_reference_list, src/base_model.cpp:27-51).
 * counts: [n_families x n_species] row-major.  max_family_size / max_root_family_size as derived in
 * src/user_data.cpp:40-48.  device: CUDA ordinal. */
int cafe_b200_create(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                     int32_t max_family_size, int32_t max_root_family_size, int32_t device, cafe_b200_ctx** out);

/* The same model spread over several GPUs of one node from ONE host thread (the reference's only parallel axis is the
 * family loop, `#pragma omp parallel for` at src/base_model.cpp:69 and src/gamma_core.cpp:190): families are split into n_devices
 * shards (contiguous blocks of the caller's order; for large jobs clustered, work-balanced blocks, see cafe_b200_plan_shards), shard i
 * lives on devices[i] with its own stream, every device regenerates the (small) matrices itself, and each
 * call below launches on all devices before it waits for any.  The only cross-device step is the final sum of base_model.cpp:95 /
 * gamma_core.cpp:233: the per-device partial sums {sum lnL, n_failed} (16 bytes each) are added on the host in device order, so a
 * score is bit-reproducible for a given device list.  Per-family outputs always arrive in the CALLER's family order: contiguous
 * shards write straight into the caller's buffers at their offset, clustered shards go through page-locked scratch of the group and are
 * scattered back by the shard's worker thread.  The returned handle is used with every other entry point of this header exactly like a single-device context
 * (simulate, get_matrix and the stream / stats hooks act on the first device).  n_devices == 1 is cafe_b200_create. */
int cafe_b200_create_multi(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                           int32_t max_family_size, int32_t max_root_family_size, const int32_t* devices, int32_t n_devices,
                           cafe_b200_ctx** out);

/* Bucketed mode -- OPT-IN, approximate, never the default (north_star: pruning "batched over families bucketed by max family size").
 * The reference's inference path prunes every family over the full state space 0 .. max_family_size; only its p-value path truncates:
 * a simulated family whose largest size is x is pruned over 0 .. m(x), m(x) = min(max_family_size, x + max(50, x / 5))
 * (compute_family_probabilities, src/probability.cpp:394,416).  Here the same rule assigns every family to the first of the
 * state_ceilings[n_ceilings] (values of max_family_size for a bucket; max_family_size itself is always the last) that is >= m(x); each
 * bucket is pruned with max_family_size = its ceiling and max_root_family_size = min(max_root_family_size, ceiling), all buckets
 * concurrently on `device`.  Per-family results differ from cafe_b200_create by the probability mass beyond the ceiling (measured
 * <= 1e-12 relative on the bundled data sets and the bench workload, tests/test_gpu_parity.py); root vectors are 0 beyond a bucket's
 * root sizes.  cafe_b200_pvalues uses it internally for the simulated families, as the reference does. */
int cafe_b200_create_bucketed(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species,
                              int32_t max_family_size, int32_t max_root_family_size, const int32_t* state_ceilings, int32_t n_ceilings,
                              int32_t device, cafe_b200_ctx** out);

/* Work accounting for the roofline: columns[n_nodes] = the number of count-vector columns node v's vector is computed for in one
 * category of one evaluation (0 for leaves).  The reference prunes every family (gamma model) or every distinct family (base model,
 * build_reference_list, src/base_model.cpp:27-51) through every node; here every distinct family, and a node whose subtree shows
 * few distinct patterns of leaf counts is computed once per PATTERN (its factor table) and gathered by the families that share it. */
int cafe_b200_node_columns(const cafe_b200_ctx* ctx, int64_t* columns);

/* How to cut a job into n_shards pieces for strong scaling (host-only; what cafe_b200_create_multi does for large jobs, exported for
 * hosts that run one process per GPU).  order[n_families] receives the family indices ordered by total count (families with similar
 * counts share the most subtree patterns), bounds[n_shards + 1] the cut points in that order, moved until every piece costs the same
 * number of contraction columns under the subtree-pattern table plan: shard i prunes the families order[bounds[i] .. bounds[i+1]). */
int cafe_b200_plan_shards(const cafe_b200_tree* tree, const int32_t* counts, int64_t n_families, int32_t n_species, int32_t n_shards,
                          int64_t* order, int64_t* bounds);

/* Number of device shards behind a context (1 for cafe_b200_create). */
int32_t cafe_b200_n_devices(const cafe_b200_ctx* ctx);

int cafe_b200_destroy(cafe_b200_ctx* ctx);

/* Last error text of ctx (ctx == NULL: of the last failed create on this thread). */
const char* cafe_b200_last_error(const cafe_b200_ctx* ctx);

/* Root prior table: prior[j] = root_equilibrium_distribution::compute(j) (returns float,
 * src/root_equilibrium_distribution.h:40, .cpp:81-87); 0 beyond n. */
int cafe_b200_set_prior(cafe_b200_ctx* ctx, const float* prior, int32_t n);

/* Error model (src/error_model.cpp:52-57 get_probs; consumed at leaves, src/probability.cpp:187-198):
 * probs[rows x 3] = P(deviation -1, 0, +1 | observed size); sizes >= rows use the last row.  Call
 * again whenever epsilon changes (error_model::replace_epsilons).  probs == NULL disables it. */
int cafe_b200_set_error_model(cafe_b200_ctx* ctx, const double* probs, int32_t rows, int32_t max_cnt);

/* base_model::infer_family_likelihoods (src/base_model.cpp:53-100).
 * lambdas[n_lambda]: value per lambda class.  neg_lnl: -sum_f max_j[log L_f(j) + log prior(j)].
 * family_lnl[n_families] (optional): model::results[f].posterior_probability. */
int cafe_b200_eval_base(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda,
                        double* neg_lnl, double* family_lnl);

/* gamma_model::infer_family_likelihoods (src/gamma_core.cpp:168-237) with the K multipliers and
 * category probabilities produced on the host by get_gamma (src/gamma.cpp:225-241); alpha is only
 * used for can_infer (gamma_core.cpp:123-141).
 * cat_lk[F x K]: gamma_model::_category_likelihoods ; family_lk[F]; posterior[F x K];
 * significant[F x K] (posterior > 0.95); failed[F] (a category's root vector summed to 0,
 * gamma_core.cpp:151); n_failed: count of failed families (neg_lnl = +inf when > 0). */
int cafe_b200_eval_gamma(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda, double alpha,
                         const double* multipliers, const double* cat_probs, int32_t n_cat,
                         double* neg_lnl, double* cat_lk, double* family_lk, double* posterior,
                         uint8_t* significant, uint8_t* failed, int64_t* n_failed);

/* Pupko joint reconstruction (src/gene_family_reconstructor.cpp:30-190) for the base model (n_cat == 0)
 * or per gamma category followed by the weighted average and rounding (src/gamma_core.cpp:271-288,351-357).
 * cat_states[F x max(n_cat,1) x n_nodes], states[F x n_nodes] (leaves = observed counts),
 * averaged[F x n_nodes] (optional, un-rounded). */
int cafe_b200_reconstruct(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda,
                          const double* multipliers, const double* cat_probs, int32_t n_cat,
                          int32_t* cat_states, int32_t* states, double* averaged);

/* Host driver -------------------------------------------------------- */

/* Shape of a context: what get_lambda_optimizer reads from the model (longest branch: src/base_model.cpp:117-118). */
int cafe_b200_describe(const cafe_b200_ctx* ctx, int64_t* n_families, int32_t* n_nodes, int32_t* n_lambda_classes,
                       int32_t* max_family_size, int32_t* max_root_family_size, double* longest_branch);

/* get_gamma (src/gamma.cpp:225-241): K equiprobable categories, multipliers = rescaled category medians. */
int cafe_b200_discrete_gamma(int32_t n_cat, double alpha, double* cat_probs, double* multipliers);

/* The reference's simplex search (fminsearch_min, src/optimizer.cpp:287-322) over an arbitrary objective. */
int cafe_b200_minimize(double (*objective)(const double* x, void* user), void* user, int32_t n, const double* x0,
                       int32_t max_iterations, double* x_out, double* f_out, int32_t* iterations);

#define CAFE_B200_FIT_MAX_VALUES 16
typedef struct {
    int32_t n_cat;                /* <= 1: base model; > 1: gamma model with n_cat categories */
    int32_t optimize_epsilon;     /* base model with the default error model and epsilon as a free parameter (`-e` without a file) */
    double fixed_alpha;           /* gamma model: > 0 keeps alpha fixed (`-k` with a given alpha) */
    const double* fixed_lambdas;  /* non-NULL: lambdas are given (one per class), only alpha is estimated */
    const double* start;          /* non-NULL: explicit start point instead of the seeded random guess */
    uint32_t seed;                /* std::mt19937 seed of the initial guess (the reference never seeds its engine) */
    int32_t max_iterations;       /* <= 0: the reference's 300 */
} cafe_b200_fit_options;
typedef struct {
    double values[CAFE_B200_FIT_MAX_VALUES];   /* lambdas..., then alpha or epsilon */
    int32_t n_values;
    double neg_lnl;
    int32_t iterations;           /* simplex iterations (optimizer::result::num_iterations) */
    int32_t evaluations;          /* calls of infer_family_likelihoods (event_monitor attempts) */
    int32_t status;               /* 0 converged, 1 no finite starting point (OptimizerInitializationFailure), 2 iteration cap */
    double seconds;
} cafe_b200_fit_result;

/* optimizer::optimize (src/optimizer.cpp:540-569) over the scorer get_lambda_optimizer would build
 * (src/base_model.cpp:111-131, src/gamma_core.cpp:239-269); set_prior (and set_error_model for a fixed error model) first. */
int cafe_b200_fit(cafe_b200_ctx* ctx, const cafe_b200_fit_options* options, cafe_b200_fit_result* result);

/* Simulator: simulator::create_trial (src/simulator.cpp:29-58) for n_families families on the context's
 * tree.  root_sizes[f] is the root size of family f (the reference reads its vectorised root distribution at index f); a family absent at the root is
 * redrawn up to max_redraws times (the reference: 50); child sizes
 * are drawn from matrix rows restricted to sizes < max_sim (select_random_y, src/matrix_cache.cpp:60-66; the reference passes its
 * max_family_size, 120 by default); n_cat > 0: each family first picks a rate category with cat_probs (gamma_core.cpp:91-95).
 * counts[F x n_species] (the context's species columns); node_sizes[F x n_nodes] and categories[F] optional; n_not_at_root:
 * families still absent at the root after 50 redraws (kept, as the reference keeps them with a warning).  Counter-based RNG
 * (Philox4x32-10 keyed by seed, one stream per family): reproducible per seed, distributional parity with the reference.
 * When the context holds an error model (cafe_b200_set_error_model) every simulated LEAF count is perturbed like
 * adjust_for_error_model (src/probability.cpp:478-499), as simulator::create_trial does with the user's error model; a perturbed count
 * may reach max_sim.  Every family draws its own rate category: the reference draws one per batch of LAMBDA_PERTURBATION_STEP_SIZE
 * families (src/simulator.cpp:74-90), and that constant is 1 in its build (CMakeLists.txt:15). */
int cafe_b200_simulate(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda, const double* multipliers, const double* cat_probs,
                       int32_t n_cat, int32_t max_sim, int32_t max_redraws, const int32_t* root_sizes, int64_t n_families, uint64_t seed,
                       int32_t* counts, int32_t* node_sizes, int32_t* categories, int64_t* n_not_at_root);

/* Family-level p-values: compute_pvalues (src/probability.cpp:528-570) with the model's own lambda.  For every
 * root size 1..R, n_sims families are simulated from that root size (no error model, no redraws) and their likelihood AT that root
 * size forms a sorted conditional distribution; a family's p-value is the largest, over root sizes below rint(1.25 * its largest
 * count), of the fraction of that distribution not above the family's own root-vector entry (pvalue / find_best_pvalue, :501-526).
 * The reference calls it with n_sims = 1000 (src/execute.cpp:171).  Monte-Carlo: agreement with the reference is statistical.  The
 * simulated families are pruned over truncated state spaces like the reference's (largest size + max(50, size/5), src/probability.cpp:
 * 394,416; here rounded up to a multiple of 16 states), the observed families over the full one. */
int cafe_b200_pvalues(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda, int32_t n_sims, uint64_t seed, double* pvalues);

/* Per-branch change probabilities of the report (compute_viterbi_sum, src/gene_family_reconstructor.cpp:388-429, as estimator::execute
 * calls it, src/execute.cpp:173-184): states[F x n_nodes] as cafe_b200_reconstruct returns them; selected[F] (NULL = every family) marks
 * the families whose p-value is below the threshold; probs[F x n_nodes], -1 for the root and for unselected families. */
int cafe_b200_branch_probabilities(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda, const int32_t* states,
                                   const uint8_t* selected, double* probs);

/* On-disk formats: cafe5_b200/host/io.hpp through a C interface.  Host-only (no GPU needed).  Lists of
 * strings come back tab-separated in caller-owned buffers. -------------------------------------------------------------------- */

const char* cafe_b200_io_last_error(void);

/* Newick species tree (+ optional lambda tree) -> the flattened arrays cafe_b200_create takes, nodes in the reference's reverse
 * level order (src/clade.cpp:69-100, 293-419; interior nodes named as src/clade.cpp:161-173; lambda classes src/clade.cpp:194-204). */
int cafe_b200_io_parse_tree(const char* newick, const char* lambda_newick, int32_t capacity, int32_t* n_nodes, int32_t* parent,
                            double* branch_length, int32_t* is_leaf, int32_t* lambda_class, int32_t* n_lambda, char* names, int64_t names_cap);

/* Gene-family table, CAFE or CAFExp header style (src/io.cpp:134-217): counts[n_families x n_species].  newick (NULL or "": none): the
 * species tree; in the CAFExp format the reference looks every "#name" header line up in the tree and keeps a column only for leaves
 * (interior-node columns are skipped, unknown names rejected, src/io.cpp:153-161). */
int cafe_b200_io_read_families(const char* path, const char* newick, int64_t* n_families, int32_t* n_species, int32_t* counts,
                               int64_t counts_cap, char* species, int64_t species_cap, char* ids, int64_t ids_cap);

/* Error-model file (src/io.cpp:228-274, src/error_model.cpp:31-50): probs[rows x 3]. */
int cafe_b200_io_read_error_model(const char* path, double* probs, int32_t rows_cap, int32_t* rows, int32_t* max_count);

/* The root prior table cafe_b200_set_prior takes, built as the reference builds it (user_data::create_prior, src/user_data.cpp:176-206;
 * root_equilibrium_distribution, src/root_equilibrium_distribution.cpp:13-87).  kind 0: uniform over num_values root sizes (the default,
 * num_values = max_root_family_size); 1: `-f` root distribution file ("size count" lines, src/user_data.cpp:105-117);
 * 2: `-p<lambda>` Poisson(poisson_lambda) with num_values = max_root_family_size.  prior[cap] receives the table (NULL: only its
 * length n); entries beyond n are 0 for the model. */
int cafe_b200_io_make_prior(int32_t kind, double poisson_lambda, const char* rootdist_path, int32_t num_values, float* prior, int32_t cap,
                            int32_t* n);

/* `-p` without a value (user_data::create_prior, src/user_data.cpp:193-197): the Poisson mean of the root prior estimated from the gene
 * families -- the reference's poisson_scorer (src/poisson.cpp:40-78: every positive leaf count c contributes log pdf(c - 1; lambda))
 * minimised by the same simplex search from a seeded uniform(0, 1) start (root_equilibrium_distribution(gene_families, num_values),
 * src/root_equilibrium_distribution.cpp:42-54).  counts[n_families x n_species]; species: the tab-joined column names as
 * cafe_b200_io_read_families returns them (the reference adds a family's terms in the order of its case-insensitive species map; NULL
 * or "": column order -- same minimum, the sum may differ in the last bits).  The table is then
 * cafe_b200_io_make_prior(2, *poisson_lambda, NULL, (int)(max_root_family_size * 0.8), ...).  Returns CAFE_B200_ERR_STATE when no
 * start point has a finite score (OptimizerInitializationFailure).  The reference draws the start from its one process-wide engine
 * before any model is fitted; a host reproducing a whole `-p` run seeds this call first.  Host only. */
int cafe_b200_fit_poisson_prior(const int32_t* counts, int64_t n_families, int32_t n_species, const char* species, uint32_t seed,
                                double* poisson_lambda, double* neg_lnl, int32_t* iterations);

/* max_family_size / max_root_family_size from the count table (src/user_data.cpp:40-48, floors src/user_data.h:26-27). */
int cafe_b200_io_derive_sizes(const int32_t* counts, int64_t n, int32_t* max_family_size, int32_t* max_root_family_size);

/* <Model>_results.txt (model::write_vital_statistics, src/core.cpp:97-112; gamma: src/gamma_core.cpp:46-50); epsilon / alpha = NaN
 * when the model has none. */
int cafe_b200_io_format_results(const char* model_name, double neg_lnl, const double* lambdas, int32_t n_lambda, double epsilon,
                                double longest_branch, int32_t attempts, int32_t rejects, double alpha, char* out, int64_t out_cap);

/* what = 0: Base_family_likelihoods.txt (src/base_model.cpp:102-109; family_values = family lnL); 1: Gamma_family_likelihoods.txt
 * (src/gamma_core.cpp:52-58, src/core.cpp:53-58; family_values = family likelihood); 2: Gamma_category_likelihoods.txt
 * (src/gamma_core.cpp:359-374). */
int cafe_b200_io_format_family_likelihoods(const char* ids_tabbed, int64_t n_families, int32_t n_cat, const double* multipliers,
                                           const double* cat_lk, const double* family_values, const double* posterior,
                                           const uint8_t* significant, int32_t what, char* out, int64_t out_cap);

/* Reconstruction tables (reconstruction::write_results, src/gene_family_reconstructor.cpp:352-379) from the states cafe_b200_reconstruct
 * returns (states[F x n_nodes], nodes in the order cafe_b200_io_parse_tree gives for `newick`), nodes labelled with the reference's
 * ape numbering.  what = 0: <Model>_count.tab; 1: <Model>_change.tab; 2: <Model>_asr.tre (gamma_multipliers: the gamma model's
 * LAMBDA_MULTIPLIERS block, or NULL; branch_probs, when given, star the branches below the threshold); 3: <Model>_family_results.txt
 * (needs pvalues); 4: <Model>_clade_results.txt (the reference orders its rows by pointer value; here they come in ape order);
 * 5: <Model>_branch_probabilities.tab (needs branch_probs[F x n_nodes] from cafe_b200_branch_probabilities). */
int cafe_b200_io_format_reconstruction(const char* newick, const char* ids_tabbed, int64_t n_families, const int32_t* states,
                                       const double* pvalues, double pvalue_threshold, const double* gamma_multipliers, int32_t n_cat,
                                       const double* branch_probs, int32_t what, char* out, int64_t out_cap);

/* <Model>_report.cafe (operator<<(ostream&, const Report&), src/report.cpp:45-165, as estimator::execute builds it,
 * src/execute.cpp:190-197): the tree, the fitted lambdas, the lambda tree (lambda_newick NULL / "": the reference's dummy tree of
 * 1s), the ape node IDs, per-branch average expansion and expansion / no-change / contraction counts over all families, and one line
 * (id, newick with reconstructed counts, p-value, newick with node IDs) per family that has branch probabilities
 * (branch_probs[F x n_nodes] from cafe_b200_branch_probabilities, -1 = none; NULL: no family lines). */
int cafe_b200_io_format_report(const char* newick, const char* lambda_newick, const double* lambdas, int32_t n_lambda,
                               const char* ids_tabbed, int64_t n_families, const int32_t* states, const double* pvalues,
                               const double* branch_probs, char* out, int64_t out_cap);

/* simulation.txt (include_internal = 0) / simulation_truth.txt (include_internal != 0) as simulator::print_simulations writes them
 * (src/simulator.cpp:135-172) from node_sizes[F x n_nodes] of cafe_b200_simulate (nodes in the order cafe_b200_io_parse_tree gives for
 * `newick`) and the lambda each family was simulated with (family_lambda[F]: first lambda x the family's multiplier). */
int cafe_b200_io_format_simulation(const char* newick, int64_t n_families, const int32_t* node_sizes, const double* family_lambda,
                                   int32_t include_internal, char* out, int64_t out_cap);

/* The error model file the reference writes once epsilon has been estimated (write_error_model_file, src/io.cpp:277-297): probs
 * [rows x 3] as cafe_b200_io_read_error_model returns them / as cafe_b200_set_error_model takes them ("maxcnt" is rows - 1). */
int cafe_b200_io_format_error_model(const double* probs, int32_t rows, char* out, int64_t out_cap);

/* Test hooks ------------------------------------------------------------------------------- */

/* matrix_cache::get_matrix (src/matrix_cache.cpp:88-105) for one (lambda, branch length) key after
 * the reference's quantisation (src/matrix_cache.h:44-63).  out: [N x N] row-major [parent][child],
 * N = max(max_root_family_size, max_family_size) + 1. */
int cafe_b200_get_matrix(cafe_b200_ctx* ctx, double lambda, double branch_length, double* out);
int32_t cafe_b200_matrix_size(const cafe_b200_ctx* ctx);

/* inference_prune (src/core.cpp:134-145) root vectors: out[F x R], index i <-> root size i+1,
 * for lambda * multiplier. */
int cafe_b200_root_vectors(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda, double multiplier,
                           double* out);

/* Measurement hooks ------------------------------------------------------------------------ */

/* Device-resident step for bench.py: runs exactly the kernels of eval_gamma / eval_base
 * (n_cat == 0) on the context's stream WITHOUT the final device->host copies or a host sync.
 * The scalar result stays on the device and is fetched with cafe_b200_fetch_result. */
int cafe_b200_enqueue_eval(cafe_b200_ctx* ctx, const double* lambdas, int32_t n_lambda, double alpha,
                           const double* multipliers, const double* cat_probs, int32_t n_cat);
int cafe_b200_fetch_result(cafe_b200_ctx* ctx, double* neg_lnl, int64_t* n_failed);

/* Device address of the two doubles {-lnL partial, failed families} cafe_b200_enqueue_eval leaves behind, valid in stream order
 * after it on cafe_b200_stream(): a multi-process host (one rank per GPU) exchanges them with a collective ENQUEUED ON THAT STREAM
 * (bench.py: NCCL all_gather) instead of a host round trip.  Fixed for the life of the context. */
void* cafe_b200_result_device(cafe_b200_ctx* ctx);

/* The CUDA stream (cudaStream_t) the context launches on, for event timing by the caller. */
void* cafe_b200_stream(cafe_b200_ctx* ctx);

/* Kernel accounting for the last eval: launches issued, and CUDA-event milliseconds of the
 * matrix-generation and pruning kernels (valid after a synchronising call). */
int cafe_b200_last_stats(cafe_b200_ctx* ctx, int32_t* n_launches, int32_t* n_matrices,
                         float* ms_matrices, float* ms_prune);

/* FP64 pipe microbenchmark on `device` (register-resident DFMA chains, or mma.sync m8n8k4 DMMA tiles when
 * use_dmma != 0): the roofline denominator of the pruning kernel, measured on the GPU the bench runs on. */
int cafe_b200_measure_fp64_peak(int32_t device, int32_t use_dmma, double* tflops);

/* Development hook: clock stamps written by the timing build of the pruning kernel (CAFE_B200_RESIDENT_PROBE=1; see
 * tools/gpu_probe_chunks.py).  Zeros when that build was not used. */
int cafe_b200_debug_read_probe(cafe_b200_ctx* ctx, int64_t* out, int64_t n);

/* Page-locked host memory for the caller-owned input / output buffers of the calls above.  Any host pointer works; buffers
 * obtained here are copied by the DMA engines directly (no staging copy, no page faults on a fresh allocation), which is what
 * a host that keeps its result vectors between evaluations - like the reference's model::results (src/core.h:139) and
 * gamma_model::_category_likelihoods (src/gamma_core.h:46) - should use.  Independent of any context. */
int cafe_b200_host_alloc(size_t bytes, void** out);
int cafe_b200_host_free(void* p);

/* Number of distinct count vectors actually pruned (the reference list's unique entries). */
int64_t cafe_b200_unique_families(const cafe_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CAFE_B200_H */
