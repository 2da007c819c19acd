"""ctypes view of the library's C++ readers / writers of the reference's on-disk formats (cafe5_b200/host/io.hpp).
Host-only: works without a GPU."""
import ctypes as C

import numpy as np

from . import _lib


class IoError(RuntimeError):
    pass


def _check(L, rc):
    if rc:
        raise IoError(L.cafe_b200_io_last_error().decode())


def parse_tree(newick, lambda_newick=None, capacity=8192):
    """dict(parent, branch_length, is_leaf, lambda_class, n_lambda, names): nodes in the reference's reverse level order."""
    L = _lib.load()
    n, nl = C.c_int32(), C.c_int32()
    parent = np.zeros(capacity, dtype=np.int32)
    bl = np.zeros(capacity)
    leaf = np.zeros(capacity, dtype=np.int32)
    cls = np.zeros(capacity, dtype=np.int32)
    buf = C.create_string_buffer(1 << 20)
    _check(L, L.cafe_b200_io_parse_tree(newick.encode(), (lambda_newick or "").encode(), capacity, C.byref(n), _lib.ip(parent), _lib.dp(bl),
                                       _lib.ip(leaf), _lib.ip(cls), C.byref(nl), buf, len(buf)))
    k = n.value
    return dict(parent=parent[:k].copy(), branch_length=bl[:k].copy(), is_leaf=leaf[:k].astype(bool), lambda_class=cls[:k].copy(),
                n_lambda=nl.value, names=buf.value.decode().split("\t"))


def read_gene_families(path, newick=None):
    """(species, ids, counts[F, n_species] int32).  newick: the species tree; with it the CAFExp header keeps only leaf columns."""
    L = _lib.load()
    nf, ns = C.c_int64(), C.c_int32()
    nw = None if not newick else newick.encode()
    _check(L, L.cafe_b200_io_read_families(str(path).encode(), nw, C.byref(nf), C.byref(ns), None, 0, None, 0, None, 0))
    counts = np.zeros((nf.value, ns.value), dtype=np.int32)
    sp = C.create_string_buffer(1 << 20)
    ids = C.create_string_buffer(max(1 << 20, 64 * nf.value))
    _check(L, L.cafe_b200_io_read_families(str(path).encode(), nw, C.byref(nf), C.byref(ns), _lib.ip(counts), counts.size, sp, len(sp), ids, len(ids)))
    return sp.value.decode().split("\t"), ids.value.decode().split("\t"), counts


def read_error_model(path, rows_cap=100000):
    L = _lib.load()
    probs = np.zeros((rows_cap, 3))
    rows, mx = C.c_int32(), C.c_int32()
    _check(L, L.cafe_b200_io_read_error_model(str(path).encode(), _lib.dp(probs), rows_cap, C.byref(rows), C.byref(mx)))
    return probs[:rows.value].copy(), mx.value


def fit_poisson_prior(counts, species=None, seed=10):
    """`-p` without a value: (poisson_lambda, -lnL, iterations) of the reference's poisson_scorer over the positive leaf counts."""
    L = _lib.load()
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    lam, neg, it = C.c_double(), C.c_double(), C.c_int32()
    L.cafe_b200_fit_poisson_prior.argtypes = [C.POINTER(C.c_int32), C.c_int64, C.c_int32, C.c_char_p, C.c_uint32, C.POINTER(C.c_double),
                                              C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    sp = None if species is None else "\t".join(species).encode()
    rc = L.cafe_b200_fit_poisson_prior(_lib.ip(counts), counts.shape[0], counts.shape[1], sp, int(seed), C.byref(lam), C.byref(neg), C.byref(it))
    if rc:
        raise RuntimeError("cafe_b200_fit_poisson_prior: status %d" % rc)
    return lam.value, neg.value, it.value


def make_prior(kind, num_values=0, poisson_lambda=0.0, rootdist_path=None):
    """float32 prior table built by the library as the reference builds it: kind 'uniform' | 'rootdist' | 'poisson'."""
    L = _lib.load()
    k = {"uniform": 0, "rootdist": 1, "poisson": 2}[kind]
    n = C.c_int32()
    path = None if rootdist_path is None else str(rootdist_path).encode()
    L.cafe_b200_io_make_prior.argtypes = [C.c_int32, C.c_double, C.c_char_p, C.c_int32, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]
    _check(L, L.cafe_b200_io_make_prior(k, float(poisson_lambda), path, int(num_values), None, 0, C.byref(n)))
    out = np.zeros(n.value, dtype=np.float32)
    _check(L, L.cafe_b200_io_make_prior(k, float(poisson_lambda), path, int(num_values), out.ctypes.data_as(C.POINTER(C.c_float)), n.value, C.byref(n)))
    return out


def derive_sizes(counts):
    L = _lib.load()
    c = np.ascontiguousarray(counts, dtype=np.int32)
    a, b = C.c_int32(), C.c_int32()
    _check(L, L.cafe_b200_io_derive_sizes(_lib.ip(c), c.size, C.byref(a), C.byref(b)))
    return a.value, b.value


def format_results(model_name, neg_lnl, lambdas, longest_branch, attempts, rejects, epsilon=float("nan"), alpha=float("nan")):
    L = _lib.load()
    lam = _lib.as_f64(lambdas)
    buf = C.create_string_buffer(1 << 16)
    _check(L, L.cafe_b200_io_format_results(model_name.encode(), float(neg_lnl), _lib.dp(lam), len(lam), float(epsilon), float(longest_branch),
                                           int(attempts), int(rejects), float(alpha), buf, len(buf)))
    return buf.value.decode()


def format_family_likelihoods(ids, what, family_values=None, multipliers=None, cat_lk=None, posterior=None, significant=None):
    """what: 'base' | 'gamma' | 'categories'."""
    L = _lib.load()
    K = 0 if multipliers is None else len(multipliers)
    arr = lambda a: None if a is None else _lib.as_f64(a)   # noqa: E731
    mu, cl, fv, po = arr(multipliers), arr(cat_lk), arr(family_values), arr(posterior)
    sg = None if significant is None else np.ascontiguousarray(significant, dtype=np.uint8)
    buf = C.create_string_buffer(256 * max(1, len(ids)) * max(1, K) + 4096)
    _check(L, L.cafe_b200_io_format_family_likelihoods("\t".join(ids).encode(), len(ids), K, _lib.dp(mu), _lib.dp(cl), _lib.dp(fv), _lib.dp(po),
                                                      _lib.up(sg), {"base": 0, "gamma": 1, "categories": 2}[what], buf, len(buf)))
    return buf.value.decode()


def format_reconstruction(newick, ids, states, what, pvalues=None, threshold=0.05, gamma_multipliers=None, branch_probs=None):
    """what: 'count' | 'change' | 'asr' | 'family_results' | 'clade_results' | 'branch_probabilities'; states[F, n_nodes] as
    cafe_b200_reconstruct returns them; branch_probs[F, n_nodes] from Context.branch_probabilities (-1 = none)."""
    L = _lib.load()
    st = np.ascontiguousarray(states, dtype=np.int32)
    pv = None if pvalues is None else _lib.as_f64(pvalues)
    mu = None if gamma_multipliers is None else _lib.as_f64(gamma_multipliers)
    bp = None if branch_probs is None else _lib.as_f64(branch_probs)
    buf = C.create_string_buffer(64 * st.size + (1 << 16))
    _check(L, L.cafe_b200_io_format_reconstruction(newick.encode(), "\t".join(ids).encode(), len(ids), _lib.ip(st), _lib.dp(pv), float(threshold),
                                                  _lib.dp(mu), 0 if mu is None else len(mu), _lib.dp(bp),
                                                  {"count": 0, "change": 1, "asr": 2, "family_results": 3, "clade_results": 4,
                                                   "branch_probabilities": 5}[what], buf, len(buf)))
    return buf.value.decode()


def format_report(newick, ids, states, pvalues, lambdas=(), lambda_newick=None, branch_probs=None):
    """<Model>_report.cafe text (src/report.cpp) from the reconstructed states[F, n_nodes], the family p-values and, for the per-family
    lines, branch_probs[F, n_nodes] from Context.branch_probabilities (-1 = none)."""
    L = _lib.load()
    st = np.ascontiguousarray(states, dtype=np.int32)
    pv = _lib.as_f64(pvalues)
    lam = _lib.as_f64(list(lambdas)) if len(lambdas) else None
    bp = None if branch_probs is None else _lib.as_f64(branch_probs)
    buf = C.create_string_buffer(64 * st.size + (1 << 16))
    _check(L, L.cafe_b200_io_format_report(newick.encode(), (lambda_newick or "").encode(), _lib.dp(lam), 0 if lam is None else len(lam),
                                          "\t".join(ids).encode(), len(ids), _lib.ip(st), _lib.dp(pv), _lib.dp(bp), buf, len(buf)))
    return buf.value.decode()


def format_simulation(newick, node_sizes, family_lambda, include_internal=False):
    """simulation.txt / simulation_truth.txt text (src/simulator.cpp:135-172) from node_sizes[F, n_nodes] of Context.simulate and the
    lambda each family was simulated with."""
    L = _lib.load()
    ns = np.ascontiguousarray(node_sizes, dtype=np.int32)
    fl = _lib.as_f64(family_lambda)
    buf = C.create_string_buffer(16 * ns.size + 64 * ns.shape[0] + (1 << 16))
    _check(L, L.cafe_b200_io_format_simulation(newick.encode(), ns.shape[0], _lib.ip(ns), _lib.dp(fl), int(bool(include_internal)), buf, len(buf)))
    return buf.value.decode()


def format_error_model(probs):
    """Error model file text (write_error_model_file, src/io.cpp:277-297) from probs[rows, 3]."""
    L = _lib.load()
    pr = np.ascontiguousarray(probs, dtype=np.float64).reshape(-1, 3)
    buf = C.create_string_buffer(96 * pr.shape[0] + (1 << 12))
    _check(L, L.cafe_b200_io_format_error_model(_lib.dp(pr), pr.shape[0], buf, len(buf)))
    return buf.value.decode()
