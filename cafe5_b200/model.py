"""Host-side mirror of the reference's model interface for the hot path, on top of the C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to /root/reference):
  single_lambda / multiple_lambda   src/lambda.h:37-112
  error_model                       src/error_model.h, src/error_model.cpp
  base_model                        src/base_model.h, src/base_model.cpp:53-100, 133-170
  gamma_model                       src/gamma_core.h, src/gamma_core.cpp:61-67, 123-237, 290-339
Everything numerical happens in libcafe_b200.so (CUDA); this file only marshals arrays.  The product
C++ equivalent a CAFE5 maintainer would link is cafe5_b200/host/gpu_model.hpp.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from .gamma import get_gamma


class CafeError(RuntimeError):
    pass


class single_lambda:
    def __init__(self, lam):
        self._lambda = float(lam)

    def update(self, values):
        self._lambda = float(values[0])

    def count(self):
        return 1

    def values(self):
        return [self._lambda]

    def is_valid(self):
        return self._lambda > 0


class multiple_lambda:
    def __init__(self, lambdas):
        self._lambdas = [float(v) for v in lambdas]

    def update(self, values):
        self._lambdas = [float(v) for v in values[:len(self._lambdas)]]

    def count(self):
        return len(self._lambdas)

    def values(self):
        return list(self._lambdas)

    def is_valid(self):
        return not any(v < 0 for v in self._lambdas)


class error_model:
    """Deviation classes are fixed to (-1, 0, +1) as in every error model the reference writes."""

    def __init__(self, probs, max_family_size):
        self.probs = np.ascontiguousarray(probs, dtype=np.float64).reshape(-1, 3).copy()
        self.max_family_size = int(max_family_size)

    def get_probs(self, fam_size):
        return self.probs[min(fam_size, len(self.probs) - 1)]

    def get_epsilons(self):
        return sorted(set(float(r[2]) for r in self.probs))

    def replace_epsilons(self, new_epsilons):
        """src/error_model.cpp:79-109: rows whose last entry is within 1% of a key get the new epsilon."""
        def near(x, y):
            return abs(x - y) <= 0.01 * abs(x)
        for i, row in enumerate(self.probs):
            for old, new in new_epsilons.items():
                if near(old, row[2]):
                    if i == 0:
                        self.probs[0] = [row[0], 1 - new, new]
                    else:
                        self.probs[i] = [new, 1 - (new * 2), new]
                    row = self.probs[i]

    def update_single_epsilon(self, new_epsilon):
        eps = self.get_epsilons()
        assert len(eps) == 1
        self.replace_epsilons({eps[0]: new_epsilon})


class _PinnedBlock:
    """One cafe_b200_host_alloc block, released when the last array viewing it is garbage-collected."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.cafe_b200_host_free(self.ptr)
        except Exception:
            pass


class Context:
    """Owns one cafe_b200_ctx: one GPU (device=...), or the families sharded over several GPUs of the node from this one process
    (devices=[...], cafe_b200_create_multi)."""

    def __init__(self, tree, counts, max_family_size, max_root_family_size, device=0, devices=None, state_ceilings=None):
        self.lib = _lib.load()
        self._pinned_bufs = {}
        self.tree = tree
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        self.F, self.n_species = counts.shape
        self.n_nodes = tree.n_nodes
        self.max_family_size = int(max_family_size)
        self.R = int(max_root_family_size)
        self._keep = (np.ascontiguousarray(tree.parent, dtype=np.int32), np.ascontiguousarray(tree.branch_length, dtype=np.float64),
                      np.ascontiguousarray(tree.leaf_col, dtype=np.int32), np.ascontiguousarray(tree.lambda_class, dtype=np.int32))
        ct = _lib.CTree(tree.n_nodes, _lib.ip(self._keep[0]), _lib.dp(self._keep[1]), _lib.ip(self._keep[2]), _lib.ip(self._keep[3]))
        h = C.c_void_p()
        if state_ceilings is not None:      # opt-in bucketed mode (cafe_b200_create_bucketed): approximate, labelled
            ce = np.ascontiguousarray(state_ceilings, dtype=np.int32)
            rc = self.lib.cafe_b200_create_bucketed(C.byref(ct), _lib.ip(counts), self.F, self.n_species, self.max_family_size, self.R,
                                                    _lib.ip(ce), len(ce), int(device), C.byref(h))
        elif devices is None:
            rc = self.lib.cafe_b200_create(C.byref(ct), _lib.ip(counts), self.F, self.n_species, self.max_family_size, self.R,
                                           int(device), C.byref(h))
        else:
            devs = np.ascontiguousarray(devices, dtype=np.int32)
            rc = self.lib.cafe_b200_create_multi(C.byref(ct), _lib.ip(counts), self.F, self.n_species, self.max_family_size, self.R,
                                                 _lib.ip(devs), len(devs), C.byref(h))
        if rc:
            raise CafeError("cafe_b200_create: %s (status %d)" % (self.lib.cafe_b200_last_error(None).decode(), rc))
        self.h = h
        self.N = self.lib.cafe_b200_matrix_size(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cafe_b200_destroy(self.h)
            self.h = None
        # the page-locked blocks are freed when the LAST array viewing them dies (a result kept past close() stays valid)
        self._pinned_bufs = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise CafeError("%s: %s (status %d)" % (what, self.lib.cafe_b200_last_error(self.h).decode(), rc))

    def set_prior(self, prior):
        prior = np.ascontiguousarray(prior, dtype=np.float32)
        self._check(self.lib.cafe_b200_set_prior(self.h, _lib.fp(prior), len(prior)), "set_prior")

    def set_error_model(self, em):
        if em is None:
            self._check(self.lib.cafe_b200_set_error_model(self.h, None, 0, 0), "set_error_model")
        else:
            probs = np.ascontiguousarray(em.probs, dtype=np.float64)
            self._check(self.lib.cafe_b200_set_error_model(self.h, _lib.dp(probs), probs.shape[0], em.max_family_size), "set_error_model")

    def unique_families(self):
        return self.lib.cafe_b200_unique_families(self.h)

    def node_columns(self):
        """columns[n_nodes]: how many columns every internal node is computed for per category (subtree-pattern tables shrink it)."""
        out = np.zeros(self.n_nodes, dtype=np.int64)
        self._check(self.lib.cafe_b200_node_columns(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))), "node_columns")
        return out

    def n_devices(self):
        return self.lib.cafe_b200_n_devices(self.h)

    def eval_base(self, lambdas, want_family=True):
        lam = _lib.as_f64(lambdas)
        neg = C.c_double()
        fam = np.empty(self.F) if want_family else None
        self._check(self.lib.cafe_b200_eval_base(self.h, _lib.dp(lam), len(lam), C.byref(neg), _lib.dp(fam)), "eval_base")
        return neg.value, fam

    def _pinned(self, name, shape, dtype):
        """A page-locked output buffer owned by this context (cafe_b200_host_alloc), allocated once per (name, shape) and REUSED by
        later calls - like the reference's model::results / gamma_model::_category_likelihoods, which are members overwritten by
        every evaluation."""
        key = (name, tuple(shape), np.dtype(dtype).str)
        buf = self._pinned_bufs.get(key)
        if buf is None:
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            ptr = C.c_void_p()
            self._check(self.lib.cafe_b200_host_alloc(nbytes, C.byref(ptr)), "host_alloc")
            raw = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
            raw._owner = _PinnedBlock(self.lib, ptr)      # numpy keeps `raw` alive through the array's base; `raw` keeps the block
            buf = (np.frombuffer(raw, dtype=dtype, count=int(np.prod(shape))).reshape(shape), ptr)
            self._pinned_bufs[key] = buf
        return buf[0]

    def eval_gamma(self, lambdas, alpha, multipliers, cat_probs, want_family=True, pinned=False):
        """pinned=True: the per-family outputs land in page-locked buffers owned by the context and overwritten by the next
        pinned call (what a host that keeps its result vectors between evaluations does); default: fresh numpy arrays."""
        lam = _lib.as_f64(lambdas)
        mu = _lib.as_f64(multipliers)
        cp = _lib.as_f64(cat_probs)
        K = len(mu)
        neg = C.c_double()
        nf = C.c_int64()
        out = dict(cat_lk=None, family_lk=None, posterior=None, significant=None, failed=None)
        if want_family and pinned:
            out = dict(cat_lk=self._pinned("cat_lk", (self.F, K), np.float64), family_lk=self._pinned("family_lk", (self.F,), np.float64),
                       posterior=self._pinned("posterior", (self.F, K), np.float64),
                       significant=self._pinned("significant", (self.F, K), np.uint8), failed=self._pinned("failed", (self.F,), np.uint8))
        elif want_family:
            out = dict(cat_lk=np.zeros((self.F, K)), family_lk=np.zeros(self.F), posterior=np.zeros((self.F, K)),
                       significant=np.zeros((self.F, K), dtype=np.uint8), failed=np.zeros(self.F, dtype=np.uint8))
        self._check(self.lib.cafe_b200_eval_gamma(self.h, _lib.dp(lam), len(lam), float(alpha), _lib.dp(mu), _lib.dp(cp), K,
                                                  C.byref(neg), _lib.dp(out["cat_lk"]), _lib.dp(out["family_lk"]),
                                                  _lib.dp(out["posterior"]), _lib.up(out["significant"]), _lib.up(out["failed"]),
                                                  C.byref(nf)), "eval_gamma")
        out["neg_lnl"] = neg.value
        out["n_failed"] = nf.value
        if want_family and pinned and math.isinf(neg.value) and nf.value == 0:
            # rejected before any kernel ran (invalid lambda / alpha / saturated): the reused buffers still hold the previous call's
            # values; hand back zeros like the fresh-array path does
            for key in ("cat_lk", "family_lk", "posterior", "significant", "failed"):
                out[key][...] = 0
        return out

    def reconstruct(self, lambdas, multipliers=None, cat_probs=None, want_cat_states=True, want_averaged=True):
        lam = _lib.as_f64(lambdas)
        K = 0 if multipliers is None else len(multipliers)
        mu = None if K == 0 else _lib.as_f64(multipliers)
        cp = None if K == 0 else _lib.as_f64(cat_probs)
        cat_states = np.zeros((self.F, max(K, 1), self.n_nodes), dtype=np.int32) if want_cat_states else None
        states = np.zeros((self.F, self.n_nodes), dtype=np.int32)
        avg = np.zeros((self.F, self.n_nodes)) if want_averaged else None
        self._check(self.lib.cafe_b200_reconstruct(self.h, _lib.dp(lam), len(lam), _lib.dp(mu), _lib.dp(cp), K,
                                                   _lib.ip(cat_states), _lib.ip(states), _lib.dp(avg)), "reconstruct")
        return dict(cat_states=cat_states, states=states, averaged=avg)

    def get_matrix(self, lam, t):
        out = np.empty((self.N, self.N))
        self._check(self.lib.cafe_b200_get_matrix(self.h, float(lam), float(t), _lib.dp(out)), "get_matrix")
        return out

    def root_vectors(self, lambdas, multiplier=1.0):
        lam = _lib.as_f64(lambdas)
        out = np.empty((self.F, self.R))
        self._check(self.lib.cafe_b200_root_vectors(self.h, _lib.dp(lam), len(lam), float(multiplier), _lib.dp(out)), "root_vectors")
        return out

    # measurement hooks -----------------------------------------------------------------------
    def enqueue_eval(self, lambdas, alpha=0.0, multipliers=None, cat_probs=None):
        lam = _lib.as_f64(lambdas)
        K = 0 if multipliers is None else len(multipliers)
        mu = None if K == 0 else _lib.as_f64(multipliers)
        cp = None if K == 0 else _lib.as_f64(cat_probs)
        self._check(self.lib.cafe_b200_enqueue_eval(self.h, _lib.dp(lam), len(lam), float(alpha), _lib.dp(mu), _lib.dp(cp), K), "enqueue_eval")

    def fetch_result(self):
        neg = C.c_double()
        nf = C.c_int64()
        self._check(self.lib.cafe_b200_fetch_result(self.h, C.byref(neg), C.byref(nf)), "fetch_result")
        return neg.value, nf.value

    def stream(self):
        return self.lib.cafe_b200_stream(self.h)

    def result_device(self):
        """Device address of the {-lnL partial, failed families} pair enqueue_eval leaves behind (cafe_b200_result_device)."""
        ptr = self.lib.cafe_b200_result_device(self.h)
        if not ptr:
            raise CafeError("result_device: %s" % self.lib.cafe_b200_last_error(self.h).decode())
        return ptr

    def branch_probabilities(self, lambdas, states, selected=None):
        """Per-branch change probabilities (cafe_b200_branch_probabilities): [F, n_nodes], -1 for the root / unselected families."""
        lam = _lib.as_f64(lambdas)
        st = np.ascontiguousarray(states, dtype=np.int32)
        sel = None if selected is None else np.ascontiguousarray(selected, dtype=np.uint8)
        out = np.zeros((self.F, self.n_nodes))
        self._check(self.lib.cafe_b200_branch_probabilities(self.h, _lib.dp(lam), len(lam), _lib.ip(st), _lib.up(sel), _lib.dp(out)),
                    "branch_probabilities")
        return out

    def pvalues(self, lambdas, n_sims=1000, seed=1):
        """Family-level Monte-Carlo p-values (cafe_b200_pvalues; compute_pvalues, src/probability.cpp:528-570)."""
        lam = _lib.as_f64(lambdas)
        out = np.zeros(self.F)
        self._check(self.lib.cafe_b200_pvalues(self.h, _lib.dp(lam), len(lam), int(n_sims), int(seed), _lib.dp(out)), "pvalues")
        return out

    def simulate(self, lambdas, root_sizes, multipliers=None, cat_probs=None, max_sim=120, seed=1, want_nodes=False, max_redraws=50):
        """Simulate len(root_sizes) families on this context's tree (cafe_b200_simulate): dict(counts[F, n_species],
        categories[F], node_sizes[F, n_nodes] or None, n_not_at_root)."""
        lam = _lib.as_f64(lambdas)
        K = 0 if multipliers is None else len(multipliers)
        mu = None if K == 0 else _lib.as_f64(multipliers)
        cp = None if K == 0 else _lib.as_f64(cat_probs)
        roots = np.ascontiguousarray(root_sizes, dtype=np.int32)
        F = roots.shape[0]
        counts = np.zeros((F, self.n_species), dtype=np.int32)
        cats = np.zeros(F, dtype=np.int32)
        nodes = np.zeros((F, self.n_nodes), dtype=np.int32) if want_nodes else None
        bad = C.c_int64()
        self._check(self.lib.cafe_b200_simulate(self.h, _lib.dp(lam), len(lam), _lib.dp(mu), _lib.dp(cp), K, int(max_sim), int(max_redraws), _lib.ip(roots),
                                                F, int(seed), _lib.ip(counts), _lib.ip(nodes), _lib.ip(cats), C.byref(bad)), "simulate")
        return dict(counts=counts, categories=cats, node_sizes=nodes, n_not_at_root=bad.value)

    def describe(self):
        nf = C.c_int64()
        nn, nl, mfs, mrs = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        lb = C.c_double()
        self._check(self.lib.cafe_b200_describe(self.h, C.byref(nf), C.byref(nn), C.byref(nl), C.byref(mfs), C.byref(mrs), C.byref(lb)), "describe")
        return dict(n_families=nf.value, n_nodes=nn.value, n_lambda_classes=nl.value, max_family_size=mfs.value,
                    max_root_family_size=mrs.value, longest_branch=lb.value)

    def fit(self, n_cat=0, optimize_epsilon=False, fixed_alpha=0.0, fixed_lambdas=None, start=None, seed=10, max_iterations=0):
        """optimizer::optimize over the scorer the reference would build (lambda | lambda, epsilon | lambda, alpha | alpha):
        the library's C++ host driver (cafe5_b200/host/fit.cpp) running Nelder-Mead over eval_base / eval_gamma."""
        o = _lib.FitOptions()
        o.n_cat, o.optimize_epsilon, o.fixed_alpha = int(n_cat), 1 if optimize_epsilon else 0, float(fixed_alpha)
        fl = _lib.as_f64(fixed_lambdas) if fixed_lambdas is not None else None
        st = _lib.as_f64(start) if start is not None else None
        o.fixed_lambdas, o.start = _lib.dp(fl), _lib.dp(st)
        o.seed, o.max_iterations = int(seed), int(max_iterations)
        r = _lib.FitResult()
        self._check(self.lib.cafe_b200_fit(self.h, C.byref(o), C.byref(r)), "fit")
        return dict(values=np.array(r.values[:r.n_values]), neg_lnl=r.neg_lnl, iterations=r.iterations, evaluations=r.evaluations,
                    status=r.status, seconds=r.seconds)

    def last_stats(self):
        nl, nm = C.c_int32(), C.c_int32()
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.cafe_b200_last_stats(self.h, C.byref(nl), C.byref(nm), C.byref(a), C.byref(b)), "last_stats")
        return dict(launches=nl.value, matrices=nm.value, ms_matrices=a.value, ms_prune=b.value)


def plan_shards(tree, counts, n_shards):
    """(order[F], bounds[n_shards + 1]) of cafe_b200_plan_shards: families ordered by total count and cut into blocks of equal cost
    under the subtree-pattern table plan.  Host-only."""
    lib = _lib.load()
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    keep = (np.ascontiguousarray(tree.parent, dtype=np.int32), np.ascontiguousarray(tree.branch_length, dtype=np.float64),
            np.ascontiguousarray(tree.leaf_col, dtype=np.int32), np.ascontiguousarray(tree.lambda_class, dtype=np.int32))
    ct = _lib.CTree(tree.n_nodes, _lib.ip(keep[0]), _lib.dp(keep[1]), _lib.ip(keep[2]), _lib.ip(keep[3]))
    order = np.zeros(counts.shape[0], dtype=np.int64)
    bounds = np.zeros(int(n_shards) + 1, dtype=np.int64)
    rc = lib.cafe_b200_plan_shards(C.byref(ct), _lib.ip(counts), counts.shape[0], counts.shape[1], int(n_shards),
                                   order.ctypes.data_as(C.POINTER(C.c_int64)), bounds.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc:
        raise CafeError("plan_shards: bad argument")
    return order, bounds


def discrete_gamma(n_cat, alpha):
    """(cat_probs, multipliers) from the library's C++ host implementation of get_gamma (cafe5_b200/host/discrete_gamma.hpp)."""
    lib = _lib.load()
    p, m = np.zeros(n_cat), np.zeros(n_cat)
    rc = lib.cafe_b200_discrete_gamma(int(n_cat), float(alpha), _lib.dp(p), _lib.dp(m))
    if rc:
        raise CafeError("discrete_gamma: bad argument")
    return list(p), list(m)


def minimize(fn, x0, max_iterations=300):
    """The library's simplex search (cafe5_b200/host/nelder_mead.hpp) over a Python objective: (x, f, iterations)."""
    lib = _lib.load()
    n = len(x0)
    raised = []

    def guarded(x, _u):
        # ctypes prints and swallows an exception raised inside a callback and hands 0.0 to the caller -- the best score the search
        # has ever seen.  Keep the first exception, answer +inf (a rejected point) from then on, and re-raise after the search.
        if raised:
            return math.inf
        try:
            return float(fn([x[i] for i in range(n)]))
        except BaseException as e:      # noqa: BLE001 -- re-raised below
            raised.append(e)
            return math.inf

    cb = _lib.OBJECTIVE(guarded)
    x0 = _lib.as_f64(x0)
    out = np.zeros(n)
    f, it = C.c_double(), C.c_int32()
    rc = lib.cafe_b200_minimize(cb, None, n, _lib.dp(x0), int(max_iterations), _lib.dp(out), C.byref(f), C.byref(it))
    if raised:
        raise raised[0]
    if rc:
        raise CafeError("minimize: bad argument")
    return out, f.value, it.value


def measure_fp64_peak(device=0, use_dmma=False):
    """TFLOP/s of the FP64 pipe on `device` (DFMA chains or DMMA tiles), measured by a microbenchmark kernel."""
    lib = _lib.load()
    out = C.c_double()
    rc = lib.cafe_b200_measure_fp64_peak(int(device), int(use_dmma), C.byref(out))
    if rc:
        raise CafeError("measure_fp64_peak: %s" % lib.cafe_b200_last_error(None).decode())
    return out.value


class model:
    """Common part of base_model / gamma_model (reference src/core.h:125-189)."""

    def __init__(self, p_lambda, tree, counts, max_family_size, max_root_family_size, p_error_model=None, device=0):
        self._p_lambda = p_lambda
        self._p_error_model = p_error_model
        self._max_family_size = max_family_size
        self._max_root_family_size = max_root_family_size
        self.ctx = Context(tree, counts, max_family_size, max_root_family_size, device=device)
        self.results = None
        self.attempts = 0
        self.rejects = 0
        self.failure_count = {}
        self._prior_key = None

    def _sync_inputs(self, prior):
        key = np.asarray(prior, dtype=np.float32).tobytes()
        if key != self._prior_key:
            self.ctx.set_prior(prior)
            self._prior_key = key
        self.ctx.set_error_model(self._p_error_model)   # epsilon may have been mutated since the last call

    def get_lambda(self):
        return self._p_lambda

    def get_gene_family_count(self):
        return self.ctx.F


class base_model(model):
    def name(self):
        return "Base"

    def infer_family_likelihoods(self, prior, p_lambda):
        """-lnL (reference src/base_model.cpp:53-100); +inf for invalid lambda."""
        self.attempts += 1
        if not self._p_lambda.is_valid():   # the reference tests the model's own lambda (:56)
            self.rejects += 1
            return math.inf
        self._sync_inputs(prior)
        neg, fam = self.ctx.eval_base(p_lambda.values())
        self.results = fam
        return neg

    def reconstruct_ancestral_states(self, prior, p_lambda=None):
        self._sync_inputs(prior)
        lam = (p_lambda or self._p_lambda).values()
        return self.ctx.reconstruct(lam)


class gamma_model(model):
    def __init__(self, p_lambda, tree, counts, max_family_size, max_root_family_size, n_gamma_cats, fixed_alpha,
                 p_error_model=None, device=0):
        super().__init__(p_lambda, tree, counts, max_family_size, max_root_family_size, p_error_model, device)
        self._gamma_cat_probs = [0.0] * n_gamma_cats
        self._lambda_multipliers = [0.0] * n_gamma_cats
        self._category_likelihoods = None
        self._alpha = fixed_alpha
        if n_gamma_cats > 1 and fixed_alpha > 0:   # gamma_core.cpp:23-29 (constructor calls set_alpha)
            self.set_alpha(fixed_alpha)

    def name(self):
        return "Gamma"

    def set_alpha(self, alpha):
        self._alpha = alpha
        if len(self._gamma_cat_probs) > 1:
            self._gamma_cat_probs, self._lambda_multipliers = get_gamma(len(self._gamma_cat_probs), alpha)

    def get_alpha(self):
        return self._alpha

    def get_lambda_multipliers(self):
        return list(self._lambda_multipliers)

    def infer_family_likelihoods(self, prior, p_lambda):
        """-lnL (reference src/gamma_core.cpp:168-237); +inf when can_infer fails or any family fails."""
        self.attempts += 1
        self.results = None
        self._sync_inputs(prior)
        out = self.ctx.eval_gamma(p_lambda.values(), self._alpha, self._lambda_multipliers, self._gamma_cat_probs)
        self._category_likelihoods = out["cat_lk"]
        if math.isinf(out["neg_lnl"]):
            if out["n_failed"] == 0:
                self.rejects += 1
            for f in np.nonzero(out["failed"])[0]:
                self.failure_count[int(f)] = self.failure_count.get(int(f), 0) + 1
            return math.inf
        self.results = out
        return out["neg_lnl"]

    def reconstruct_ancestral_states(self, prior, p_lambda=None):
        self._sync_inputs(prior)
        lam = (p_lambda or self._p_lambda).values()
        return self.ctx.reconstruct(lam, self._lambda_multipliers, self._gamma_cat_probs)
