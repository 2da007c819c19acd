"""ctypes bindings for the two checkers -- TEST INFRASTRUCTURE ONLY.

* ``OracleLib``  -> oracle/liboracle.so   (plain-C restatement, oracle/cafe_oracle.c)
* ``RefLib``     -> oracle/_ref/libcafe_ref.so (the UNMODIFIED reference sources + oracle/ref_driver.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (cafe5_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcafe_ref.so")
REF_SHIM_SO = os.path.join(HERE, "_ref", "libcafe_ref_shim.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_fp = C.POINTER(C.c_float)
c_up = C.POINTER(C.c_ubyte)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_fp)


def _up(a):
    return None if a is None else a.ctypes.data_as(c_up)


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "cafe_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def build_ref():
    """Build oracle/_ref/libcafe_ref.so when the reference sources are present; otherwise keep the prebuilt one."""
    ref_dir = os.environ.get("CAFE_REF_DIR", "/root/reference")
    if os.path.isdir(os.path.join(ref_dir, "src")):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def have_ref():
    return os.path.exists(REF_SO)


class FlatTree:
    """Flattened tree in the reference's reverse level order (see include/cafe_b200.h)."""

    def __init__(self, parent, branch_length, leaf_col, lambda_class, names=None):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.branch_length = np.ascontiguousarray(branch_length, dtype=np.float64)
        self.leaf_col = np.ascontiguousarray(leaf_col, dtype=np.int32)
        self.lambda_class = np.ascontiguousarray(lambda_class, dtype=np.int32)
        self.names = names
        self.n_nodes = len(self.parent)
        self.n_leaves = int((self.leaf_col >= 0).sum())


class OracleLib:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.oracle_chooseln.restype = C.c_double
        L.oracle_chooseln.argtypes = [C.c_int, C.c_int]
        L.oracle_birthdeath_rate_with_log_alpha.restype = C.c_double
        L.oracle_birthdeath_rate_with_log_alpha.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.oracle_transition_probability.restype = C.c_double
        L.oracle_transition_probability.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.oracle_matrix.restype = None
        L.oracle_matrix.argtypes = [C.c_int, C.c_double, C.c_double, c_dp]
        L.oracle_is_saturated.restype = C.c_int
        L.oracle_is_saturated.argtypes = [C.c_double, C.c_double]
        L.oracle_key_lambda.restype = C.c_int64
        L.oracle_key_lambda.argtypes = [C.c_double]
        L.oracle_key_branch.restype = C.c_int64
        L.oracle_key_branch.argtypes = [C.c_double]
        L.oracle_get_gamma.restype = None
        L.oracle_get_gamma.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_max_threads.restype = C.c_int
        tree_args = [C.c_int, c_ip, c_dp, c_ip, c_ip]
        L.oracle_prune.restype = C.c_int
        L.oracle_prune.argtypes = tree_args + [c_ip, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, c_dp, C.c_double, c_dp]
        L.oracle_eval_base.restype = C.c_int
        L.oracle_eval_base.argtypes = tree_args + [c_ip, C.c_long, C.c_int, C.c_int, C.c_int, c_fp, C.c_int,
                                                   c_dp, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp]
        L.oracle_eval_gamma.restype = C.c_int
        L.oracle_eval_gamma.argtypes = tree_args + [c_ip, C.c_long, C.c_int, C.c_int, C.c_int, c_fp, C.c_int,
                                                    c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_double, c_dp, c_dp, C.c_int,
                                                    c_dp, c_dp, c_dp, c_dp, c_up, c_up, C.POINTER(C.c_long), c_dp]
        L.oracle_reconstruct.restype = C.c_int
        L.oracle_reconstruct.argtypes = tree_args + [c_ip, C.c_long, C.c_int, C.c_int, C.c_int, c_fp, C.c_int,
                                                     c_dp, C.c_int, c_dp, c_dp, C.c_int, c_ip, c_ip, c_dp]

    # scalar helpers
    def chooseln(self, n, r):
        return self.lib.oracle_chooseln(n, r)

    def birthdeath(self, s, c, log_alpha, coeff):
        return self.lib.oracle_birthdeath_rate_with_log_alpha(s, c, log_alpha, coeff)

    def transition(self, lam, t, s, c):
        return self.lib.oracle_transition_probability(lam, t, s, c)

    def matrix(self, N, lam, t):
        out = np.empty((N, N), dtype=np.float64)
        self.lib.oracle_matrix(N, lam, t, _dp(out))
        return out

    def get_gamma(self, K, alpha):
        p = np.empty(K)
        m = np.empty(K)
        self.lib.oracle_get_gamma(K, alpha, _dp(p), _dp(m))
        return p, m

    def set_threads(self, n):
        self.lib.oracle_set_threads(n)

    def max_threads(self):
        return self.lib.oracle_max_threads()

    @staticmethod
    def _tree(t):
        return [t.n_nodes, _ip(t.parent), _dp(t.branch_length), _ip(t.leaf_col), _ip(t.lambda_class)]

    @staticmethod
    def _em(em):
        if em is None:
            return None, 0, 0, None
        probs, maxcnt = em
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        return _dp(probs), probs.shape[0], int(maxcnt), probs

    def prune(self, tree, counts_row, max_family_size, max_root, lambdas, multiplier=1.0, em=None):
        counts_row = np.ascontiguousarray(counts_row, dtype=np.int32)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        out = np.empty(max_root)
        emp, emr, emm, _keep = self._em(em)
        rc = self.lib.oracle_prune(*self._tree(tree), _ip(counts_row), max_family_size, max_root, emp, emr, emm,
                                   _dp(lambdas), multiplier, _dp(out))
        if rc:
            raise RuntimeError("oracle_prune rc=%d" % rc)
        return out

    def eval_base(self, tree, counts, max_family_size, max_root, prior, lambdas, em=None, want_roots=False):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        F, nl = counts.shape
        prior = np.ascontiguousarray(prior, dtype=np.float32)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        neg = C.c_double()
        fam = np.empty(F)
        roots = np.empty((F, max_root)) if want_roots else None
        emp, emr, emm, _keep = self._em(em)
        rc = self.lib.oracle_eval_base(*self._tree(tree), _ip(counts), F, nl, max_family_size, max_root, _fp(prior), len(prior),
                                       emp, emr, emm, _dp(lambdas), len(lambdas), C.byref(neg), _dp(fam), _dp(roots))
        if rc:
            raise RuntimeError("oracle_eval_base rc=%d" % rc)
        return dict(neg_lnl=neg.value, family_lnl=fam, roots=roots)

    def eval_gamma(self, tree, counts, max_family_size, max_root, prior, lambdas, multipliers, cat_probs,
                   alpha=1.0, em=None, want_roots=False):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        F, nl = counts.shape
        K = len(multipliers)
        prior = np.ascontiguousarray(prior, dtype=np.float32)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        multipliers = np.ascontiguousarray(multipliers, dtype=np.float64)
        cat_probs = np.ascontiguousarray(cat_probs, dtype=np.float64)
        neg = C.c_double()
        nf = C.c_long()
        cat = np.zeros((F, K))
        fam = np.zeros(F)
        post = np.zeros((F, K))
        sig = np.zeros((F, K), dtype=np.uint8)
        failed = np.zeros(F, dtype=np.uint8)
        roots = np.empty((F, K, max_root)) if want_roots else None
        emp, emr, emm, _keep = self._em(em)
        rc = self.lib.oracle_eval_gamma(*self._tree(tree), _ip(counts), F, nl, max_family_size, max_root, _fp(prior), len(prior),
                                        emp, emr, emm, _dp(lambdas), len(lambdas), alpha, _dp(multipliers), _dp(cat_probs), K,
                                        C.byref(neg), _dp(cat), _dp(fam), _dp(post), _up(sig), _up(failed), C.byref(nf), _dp(roots))
        if rc:
            raise RuntimeError("oracle_eval_gamma rc=%d" % rc)
        return dict(neg_lnl=neg.value, cat_lk=cat, family_lk=fam, posterior=post, significant=sig, failed=failed,
                    n_failed=nf.value, roots=roots)

    def reconstruct(self, tree, counts, max_family_size, max_root, prior, lambdas, multipliers=None, cat_probs=None):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        F, nl = counts.shape
        K = 0 if multipliers is None else len(multipliers)
        KK = max(K, 1)
        prior = np.ascontiguousarray(prior, dtype=np.float32)
        lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
        mu = None if K == 0 else np.ascontiguousarray(multipliers, dtype=np.float64)
        cp = None if K == 0 else np.ascontiguousarray(cat_probs, dtype=np.float64)
        cat_states = np.zeros((F, KK, tree.n_nodes), dtype=np.int32)
        states = np.zeros((F, tree.n_nodes), dtype=np.int32)
        avg = np.zeros((F, tree.n_nodes))
        rc = self.lib.oracle_reconstruct(*self._tree(tree), _ip(counts), F, nl, max_family_size, max_root, _fp(prior), len(prior),
                                         _dp(lambdas), len(lambdas), _dp(mu), _dp(cp), K, _ip(cat_states), _ip(states), _dp(avg))
        if rc:
            raise RuntimeError("oracle_reconstruct rc=%d" % rc)
        return dict(cat_states=cat_states, states=states, averaged=avg)


class RefLib:
    """The unmodified reference, via oracle/ref_driver.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libcafe_ref.so missing (run oracle/build_ref.sh where /root/reference exists)")
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_birthdeath_rate_with_log_alpha.restype = C.c_double
        L.ref_birthdeath_rate_with_log_alpha.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_transition_probability.restype = C.c_double
        L.ref_transition_probability.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.ref_chooseln.restype = C.c_double
        L.ref_chooseln.argtypes = [C.c_double, C.c_double]
        L.ref_matrix.argtypes = [C.c_int, C.c_double, C.c_double, c_dp]
        L.ref_is_saturated.argtypes = [C.c_double, C.c_double]
        L.ref_get_gamma.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        L.ref_tree_node_count.argtypes = [C.c_char_p]
        L.ref_tree_flatten.argtypes = [C.c_char_p, C.c_char_p, c_ip, c_dp, c_ip, c_ip, C.c_char_p, C.c_int]
        L.ref_ctx_create.restype = C.c_void_p
        L.ref_ctx_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, c_ip, C.c_long, C.c_int, C.c_int,
                                     c_dp, C.c_int, c_dp, C.c_int, C.c_int]
        L.ref_ctx_destroy.argtypes = [C.c_void_p]
        L.ref_ctx_set_error_model.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int]
        L.ref_prune.argtypes = [C.c_void_p, C.c_long, c_dp, C.c_int, C.c_double, c_dp]
        L.ref_eval_base.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp]
        L.ref_eval_gamma.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_up]
        L.ref_reconstruct_base.argtypes = [C.c_void_p, c_dp, C.c_int, c_ip]
        L.ref_reconstruct_gamma.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_int, c_ip, c_ip, c_dp]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        L.ref_optimize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, c_dp, c_ip, c_dp, c_ip, c_ip, c_dp,
                                   c_dp, c_dp, c_ip, c_ip, C.c_int, c_ip]
        self._shim = None
        L.ref_ctx_error.restype = C.c_char_p
        L.ref_ctx_error.argtypes = [C.c_void_p]
        L.ref_time_precalculate.restype = C.c_double
        L.ref_time_precalculate.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.c_int, C.c_int, c_ip, c_ip]
        L.ref_session_create.restype = C.c_void_p
        L.ref_session_create.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_int]
        L.ref_session_destroy.argtypes = [C.c_void_p]
        L.ref_session_prune.restype = C.c_double
        L.ref_session_prune.argtypes = [C.c_void_p, C.c_void_p, C.c_long, c_dp]

    def _check(self, rc):
        if rc:
            raise RuntimeError("reference: " + self.lib.ref_last_error().decode())

    def shim(self):
        """oracle/_ref/libcafe_ref_shim.so: the product's drop-in models (cafe5_b200/host/gpu_model.hpp) compiled against the
        unmodified reference.  A separate library, loaded only when a test asks for the 'gpu' backend, so that libcafe_ref.so
        (the CPU baseline) never maps libcafe_b200.so."""
        if self._shim is None:
            if not os.path.exists(REF_SHIM_SO):
                raise RuntimeError("oracle/_ref/libcafe_ref_shim.so missing (run oracle/build_ref.sh after building libcafe_b200.so)")
            C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
            self._shim = C.CDLL(REF_SHIM_SO)
            self._shim.ref_optimize_gpu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, c_ip, C.c_int, c_dp, c_ip, c_dp, c_ip, c_ip,
                                                    c_dp, c_dp, c_dp, c_ip, c_ip, C.c_int, c_ip]
        return self._shim

    def birthdeath(self, s, c, log_alpha, coeff):
        return self.lib.ref_birthdeath_rate_with_log_alpha(s, c, log_alpha, coeff)

    def transition(self, lam, t, s, c):
        return self.lib.ref_transition_probability(lam, t, s, c)

    def chooseln(self, n, r):
        return self.lib.ref_chooseln(float(n), float(r))

    def matrix(self, N, lam, t):
        out = np.empty((N, N))
        self._check(self.lib.ref_matrix(N, lam, t, _dp(out)))
        return out

    def get_gamma(self, K, alpha):
        p = np.empty(K)
        m = np.empty(K)
        self.lib.ref_get_gamma(K, alpha, _dp(p), _dp(m))
        return p, m

    def prior_table(self, kind, num_values=0, poisson_lambda=0.0, rootdist=None, cap=512):
        """compute(j), j < cap, of the prior the reference builds: kind 'uniform' (num_values sizes), 'rootdist' ({size: count}),
        'poisson' (poisson_lambda, num_values).  Returns (float32 table trimmed to the sizes it covers, full cap-long array)."""
        out = np.zeros(cap, dtype=np.float32)
        tl = C.c_int()
        sizes = np.ascontiguousarray(sorted(rootdist) if rootdist else [0], dtype=np.int32)
        cnts = np.ascontiguousarray([rootdist[int(k)] for k in sizes] if rootdist else [0], dtype=np.int32)
        self.lib.ref_prior_table.argtypes = [C.c_int, C.c_double, c_ip, c_ip, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, c_ip]
        self._check(self.lib.ref_prior_table({"uniform": 0, "rootdist": 1, "poisson": 2}[kind], float(poisson_lambda), _ip(sizes), _ip(cnts),
                                             len(sizes) if rootdist else 0, int(num_values), _fp(out), cap, C.byref(tl)))
        return out[:tl.value].copy(), out

    def fit_poisson_prior(self, species, counts, seed=10, num_values=100, cap=512):
        """The reference's `-p` without a value: (poisson_lambda, score, iterations, float32 prior table) from its poisson_scorer,
        optimizer and root_equilibrium_distribution(gene_families, num_values)."""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        lam, score, it, tl = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        out = np.zeros(cap, dtype=np.float32)
        self.lib.ref_fit_poisson_prior.argtypes = [C.c_char_p, c_ip, C.c_long, C.c_uint, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                   c_ip, C.POINTER(C.c_float), C.c_int, c_ip]
        self._check(self.lib.ref_fit_poisson_prior("\t".join(species).encode(), _ip(counts), counts.shape[0], int(seed), int(num_values),
                                                   C.byref(lam), C.byref(score), C.byref(it), _fp(out), cap, C.byref(tl)))
        return lam.value, score.value, it.value, out[:tl.value].copy()

    def set_threads(self, n):
        self.lib.ref_set_threads(n)

    def max_threads(self):
        return self.lib.ref_max_threads()

    def read_families(self, path, newick, cap=4000000):
        """The reference's read_gene_families against `newick`: (ids, counts[F, n_leaves]) with leaves in reverse level order."""
        n = C.c_long()
        counts = np.zeros(cap, dtype=np.int32)
        ids = C.create_string_buffer(1 << 22)
        self.lib.ref_read_families.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_long), c_ip, C.c_long, C.c_char_p, C.c_long]
        self._check(self.lib.ref_read_families(str(path).encode(), newick.encode(), C.byref(n), _ip(counts), cap, ids, len(ids)))
        F = n.value
        ids_list = ids.value.decode().split("\t")
        # the number of leaves is whatever makes the flat buffer F rows long: count the tree's leaves through flatten()
        leaves = int(np.sum(self.flatten(newick)[2]))
        return ids_list, counts[:F * leaves].reshape(F, leaves).copy()

    def read_error_model(self, path, rows_cap=100000):
        probs = np.zeros((rows_cap, 3))
        rows, mx = C.c_int(), C.c_int()
        self.lib.ref_read_error_model.argtypes = [C.c_char_p, c_dp, C.c_int, c_ip, c_ip]
        self._check(self.lib.ref_read_error_model(str(path).encode(), _dp(probs), rows_cap, C.byref(rows), C.byref(mx)))
        return probs[:rows.value].copy(), mx.value

    def fminsearch(self, fn, x0, max_iterations=300):
        """The reference's fminsearch_min (src/optimizer.cpp:287-322) over a Python objective."""
        CB = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)
        n = len(x0)
        cb = CB(lambda x, _u: float(fn([x[i] for i in range(n)])))
        self.lib.ref_fminsearch.argtypes = [CB, C.c_void_p, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_ip]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        out = np.zeros(n)
        f, it = C.c_double(), C.c_int()
        self._check(self.lib.ref_fminsearch(cb, None, n, _dp(x0), int(max_iterations), _dp(out), C.byref(f), C.byref(it)))
        return out, f.value, it.value

    def flatten(self, newick, lambda_newick=None):
        n = self.lib.ref_tree_node_count(newick.encode())
        if n < 0:
            self._check(1)
        parent = np.empty(n, dtype=np.int32)
        bl = np.empty(n)
        is_leaf = np.empty(n, dtype=np.int32)
        lc = np.empty(n, dtype=np.int32)
        buf = C.create_string_buffer(1 << 20)
        self._check(self.lib.ref_tree_flatten(newick.encode(), (lambda_newick or "").encode(), _ip(parent), _dp(bl),
                                              _ip(is_leaf), _ip(lc), buf, len(buf)))
        names = buf.value.decode().split("\t")
        return parent, bl, is_leaf, lc, names

    class Ctx:
        def __init__(self, ref, newick, species, counts, max_family_size, max_root, prior, lambda_newick=None, em=None):
            self.ref = ref
            counts = np.ascontiguousarray(counts, dtype=np.int32)
            self.F, self.nl = counts.shape
            prior = np.ascontiguousarray(prior, dtype=np.float64)
            self.R = max_root
            self.n_nodes = ref.lib.ref_tree_node_count(newick.encode())
            emp, emr, emm = None, 0, 0
            if em is not None:
                probs = np.ascontiguousarray(em[0], dtype=np.float64)
                emp, emr, emm = _dp(probs), probs.shape[0], int(em[1])
            self.h = ref.lib.ref_ctx_create(newick.encode(), (lambda_newick or "").encode(), "\t".join(species).encode(),
                                            _ip(counts), self.F, max_family_size, max_root, _dp(prior), len(prior),
                                            emp, emr, emm)
            if not self.h:
                ref._check(1)

        def close(self):
            if self.h:
                self.ref.lib.ref_ctx_destroy(self.h)
                self.h = None

        def __del__(self):
            try:
                self.close()
            except Exception:
                pass

        def set_error_model(self, probs, maxcnt):
            probs = np.ascontiguousarray(probs, dtype=np.float64)
            self.ref._check(self.ref.lib.ref_ctx_set_error_model(self.h, _dp(probs), probs.shape[0], int(maxcnt)))

        def prune(self, family, lambdas, multiplier=1.0):
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            out = np.empty(self.R)
            self.ref._check(self.ref.lib.ref_prune(self.h, family, _dp(lambdas), len(lambdas), multiplier, _dp(out)))
            return out

        def eval_base(self, lambdas):
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            neg = C.c_double()
            fam = np.full(self.F, np.nan)
            self.ref._check(self.ref.lib.ref_eval_base(self.h, _dp(lambdas), len(lambdas), C.byref(neg), _dp(fam)))
            return dict(neg_lnl=neg.value, family_lnl=fam)

        def eval_gamma(self, lambdas, multipliers, cat_probs):
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            mu = np.ascontiguousarray(multipliers, dtype=np.float64)
            cp = np.ascontiguousarray(cat_probs, dtype=np.float64)
            K = len(mu)
            neg = C.c_double()
            cat = np.zeros((self.F, K))
            failed = np.zeros(self.F, dtype=np.uint8)
            self.ref._check(self.ref.lib.ref_eval_gamma(self.h, _dp(lambdas), len(lambdas), _dp(mu), _dp(cp), K,
                                                        C.byref(neg), _dp(cat), _up(failed)))
            return dict(neg_lnl=neg.value, cat_lk=cat, failed=failed)

        def optimize(self, backend, n_cat=0, optimize_epsilon=False, seed=10, device=0, devices=None, trace=False, trace_cap=4096):
            """Run the reference's optimizer (Nelder-Mead) with backend 'cpu' (the reference's models) or 'gpu' (the CUDA drop-in
            models on `devices`, default [device]).  trace=True also returns every attempt: values[n, n_values], scores[n],
            failed_family[n] (lowest index of a family whose likelihood was 0; -1 none), n_failed[n]."""
            vals = np.zeros(16)
            nv, iters, attempts, tn = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            score, secs = C.c_double(), C.c_double()
            cap = int(trace_cap) if trace else 0
            tv = np.zeros((max(cap, 1), 16))
            ts = np.zeros(max(cap, 1))
            tf = np.zeros(max(cap, 1), dtype=np.int32)
            tc = np.zeros(max(cap, 1), dtype=np.int32)
            targs = (_dp(tv), _dp(ts), _ip(tf), _ip(tc), cap, C.byref(tn)) if trace else (None, None, None, None, 0, C.byref(tn))
            if backend == "gpu":
                devs = np.ascontiguousarray(devices if devices is not None else [device], dtype=np.int32)
                rc = self.ref.shim().ref_optimize_gpu(self.h, int(n_cat), 1 if optimize_epsilon else 0, int(seed), _ip(devs), len(devs),
                                                      _dp(vals), C.byref(nv), C.byref(score), C.byref(iters), C.byref(attempts),
                                                      C.byref(secs), *targs)
            else:
                rc = self.ref.lib.ref_optimize(self.h, int(n_cat), 1 if optimize_epsilon else 0, int(seed), _dp(vals), C.byref(nv),
                                               C.byref(score), C.byref(iters), C.byref(attempts), C.byref(secs), *targs)
            if rc:
                raise RuntimeError("ref_optimize: " + self.ref.lib.ref_ctx_error(self.h).decode())
            out = dict(values=vals[:nv.value].copy(), score=score.value, iterations=iters.value, attempts=attempts.value, seconds=secs.value)
            if trace:
                n = tn.value
                # the tracing scorer stores rows of n_values doubles back to back
                flat = tv.reshape(-1)[:n * nv.value].reshape(n, nv.value).copy()
                out["trace"] = dict(values=flat, scores=ts[:n].copy(), failed_family=tf[:n].copy(), n_failed=tc[:n].copy())
            return out

        def write_outputs(self, lambdas, multipliers=None, cat_probs=None, alpha=0.0):
            """(family_likelihoods text, results text, category_likelihoods text) exactly as the reference writes them."""
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            K = 0 if multipliers is None else len(multipliers)
            mu = None if K == 0 else np.ascontiguousarray(multipliers, dtype=np.float64)
            cp = None if K == 0 else np.ascontiguousarray(cat_probs, dtype=np.float64)
            bufs = [C.create_string_buffer(1 << 24), C.create_string_buffer(1 << 16), C.create_string_buffer(1 << 24)]
            self.ref.lib.ref_write_outputs.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_int, C.c_double,
                                                       C.c_char_p, C.c_long, C.c_char_p, C.c_long, C.c_char_p, C.c_long]
            self.ref._check(self.ref.lib.ref_write_outputs(self.h, _dp(lambdas), len(lambdas), _dp(mu), _dp(cp), K, float(alpha),
                                                           bufs[0], len(bufs[0]), bufs[1], len(bufs[1]), bufs[2], len(bufs[2])))
            return tuple(b.value.decode() for b in bufs)

        def write_reconstruction(self, lambdas, pvalues, multipliers=None, cat_probs=None):
            """(count.tab, change.tab, asr.tre, family_results, clade_results) texts written by the reference's reconstruction."""
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            K = 0 if multipliers is None else len(multipliers)
            mu = None if K == 0 else np.ascontiguousarray(multipliers, dtype=np.float64)
            cp = None if K == 0 else np.ascontiguousarray(cat_probs, dtype=np.float64)
            pv = np.ascontiguousarray(pvalues, dtype=np.float64)
            cap = 1 << 24
            bufs = [C.create_string_buffer(cap) for _ in range(5)]
            self.ref.lib.ref_write_reconstruction.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp] + [C.c_char_p] * 5 + [C.c_long]
            self.ref._check(self.ref.lib.ref_write_reconstruction(self.h, _dp(lambdas), len(lambdas), _dp(mu), _dp(cp), K, _dp(pv), *bufs, cap))
            return tuple(b.value.decode() for b in bufs)

        def branch_probabilities(self, lambdas, pvalues):
            """(probs[F, n_nodes] with -1 = none, _branch_probabilities.tab text, _asr.tre text with stars) from the reference."""
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            pv = np.ascontiguousarray(pvalues, dtype=np.float64)
            out = np.zeros((self.F, self.n_nodes))
            cap = 1 << 24
            tab, asr = C.create_string_buffer(cap), C.create_string_buffer(cap)
            self.ref.lib.ref_branch_probabilities.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_char_p, C.c_char_p, C.c_long]
            self.ref._check(self.ref.lib.ref_branch_probabilities(self.h, _dp(lambdas), len(lambdas), _dp(pv), _dp(out), tab, asr, cap))
            return out, tab.value.decode(), asr.value.decode()

        def print_simulations(self, node_sizes, family_lambda, include_internal):
            """simulation.txt / simulation_truth.txt text of the reference's simulator::print_simulations for the given node values."""
            ns = np.ascontiguousarray(node_sizes, dtype=np.int32)
            fl = np.ascontiguousarray(family_lambda, dtype=np.float64)
            cap = 1 << 24
            out = C.create_string_buffer(cap)
            self.ref.lib.ref_print_simulations.argtypes = [C.c_void_p, C.c_long, c_ip, c_dp, C.c_int, C.c_char_p, C.c_long]
            self.ref._check(self.ref.lib.ref_print_simulations(self.h, ns.shape[0], _ip(ns), _dp(fl), int(bool(include_internal)), out, cap))
            return out.value.decode()

        def write_report(self, lambdas, pvalues):
            """<Model>_report.cafe text as the reference's estimator::execute builds it for the base model (src/execute.cpp:167-197)."""
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            pv = np.ascontiguousarray(pvalues, dtype=np.float64)
            cap = 1 << 24
            out = C.create_string_buffer(cap)
            self.ref.lib.ref_write_report.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.c_char_p, C.c_long]
            self.ref._check(self.ref.lib.ref_write_report(self.h, _dp(lambdas), len(lambdas), _dp(pv), out, cap))
            return out.value.decode()

        def pvalues(self, lambdas, n_sims=1000, seed=1):
            """compute_pvalues of the unmodified reference (src/probability.cpp:528-570), randomizer_engine seeded with `seed`."""
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            out = np.zeros(self.F)
            self.ref.lib.ref_pvalues.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int, C.c_uint, c_dp]
            self.ref._check(self.ref.lib.ref_pvalues(self.h, _dp(lambdas), len(lambdas), int(n_sims), int(seed), _dp(out)))
            return out

        def reconstruct_base(self, lambdas):
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            st = np.zeros((self.F, self.n_nodes), dtype=np.int32)
            self.ref._check(self.ref.lib.ref_reconstruct_base(self.h, _dp(lambdas), len(lambdas), _ip(st)))
            return st

        def reconstruct_gamma(self, lambdas, multipliers, cat_probs):
            lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
            mu = np.ascontiguousarray(multipliers, dtype=np.float64)
            cp = np.ascontiguousarray(cat_probs, dtype=np.float64)
            K = len(mu)
            cs = np.zeros((self.F, K, self.n_nodes), dtype=np.int32)
            st = np.zeros((self.F, self.n_nodes), dtype=np.int32)
            avg = np.zeros((self.F, self.n_nodes))
            self.ref._check(self.ref.lib.ref_reconstruct_gamma(self.h, _dp(lambdas), len(lambdas), _dp(mu), _dp(cp), K,
                                                               _ip(cs), _ip(st), _dp(avg)))
            return dict(cat_states=cs, states=st, averaged=avg)

    def ctx(self, *a, **k):
        return RefLib.Ctx(self, *a, **k)

    def write_error_model(self, probs, max_count):
        """Error model file text of the reference's write_error_model_file for probs[rows, 3] and max_count."""
        pr = np.ascontiguousarray(probs, dtype=np.float64).reshape(-1, 3)
        cap = 1 << 20
        out = C.create_string_buffer(cap)
        self.lib.ref_write_error_model.argtypes = [c_dp, C.c_int, C.c_int, C.c_char_p, C.c_long]
        self._check(self.lib.ref_write_error_model(_dp(pr), pr.shape[0], int(max_count), out, cap))
        return out.value.decode()

    # timing legs for bench.py (see oracle/ref_driver.cpp)
    def time_precalculate(self, ctx, lambdas, multipliers, stride):
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        mu = np.ascontiguousarray(multipliers, dtype=np.float64)
        done, total = C.c_int(), C.c_int()
        t = self.lib.ref_time_precalculate(ctx.h, _dp(lam), len(lam), _dp(mu), len(mu), int(stride), C.byref(done), C.byref(total))
        if t < 0:
            self._check(1)
        return t, done.value, total.value

    def session(self, ctx, lambdas, multipliers, cat_probs):
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        mu = np.ascontiguousarray(multipliers, dtype=np.float64)
        cp = np.ascontiguousarray(cat_probs, dtype=np.float64)
        h = self.lib.ref_session_create(ctx.h, _dp(lam), len(lam), _dp(mu), _dp(cp), len(mu))
        if not h:
            self._check(1)
        return h

    def session_prune(self, ctx, session, n):
        out = C.c_double()
        t = self.lib.ref_session_prune(ctx.h, session, int(n), C.byref(out))
        if t < 0:
            self._check(1)
        return t, out.value

    def session_destroy(self, session):
        self.lib.ref_session_destroy(session)
