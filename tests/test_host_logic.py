"""Host-side logic that needs no GPU: tree flattening, family tables, derived sizes, gamma discretisation,
sharding, and that the C-ABI library loads and exports every symbol include/cafe_b200.h declares."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from cafe5_b200 import _lib, dist, families as fam
from cafe5_b200.gamma import get_gamma
from cafe5_b200.tree import FlatTree, parse_newick

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAMMALS = ("((((cat:68.710507,horse:68.710507):4.566782,cow:73.277289):20.722711,(((((chimp:4.444172,human:4.444172):6.682678,"
           "orang:11.126850):2.285855,gibbon:13.412706):7.211527,(macaque:4.567240,baboon:4.567240):16.056992):16.060702,"
           "marmoset:36.684935):57.315065):38.738021,(rat:36.302445,mouse:36.302445):96.435575);")


def test_reverse_level_order_and_names():
    t = FlatTree("((A:1,B:3):7,(C:11,D:17):23)")
    assert t.names == ["D", "C", "B", "A", "CD", "AB", "ABCD"]        # reference clade.cpp:69-100, :161-173
    assert list(t.parent) == [4, 4, 5, 5, 6, 6, -1]
    assert list(t.branch_length) == [17, 11, 3, 1, 23, 7, 0]
    assert t.children_of(6) == [5, 4] and t.children_of(5) == [3, 2]  # descendant order = decreasing index


def test_flatten_matches_reference(ref, golden):
    for nw, lnw in [(MAMMALS, None), (str(golden["mammals"]["newick"]), str(golden["mammals"]["lambda_newick"])),
                    (str(golden["hymenoptera"]["newick"]), None), ("((A:2.5,B:2.5,C:2.5):4,(D:3,(E:1,F:1):2):3.5)", None)]:
        t = FlatTree(nw, lnw)
        parent, bl, is_leaf, lc, names = ref.flatten(nw, lnw)
        assert t.names == names
        assert np.array_equal(t.parent, parent) and np.array_equal(t.branch_length, bl)
        assert np.array_equal(t.is_leaf, is_leaf.astype(bool)) and np.array_equal(t.lambda_class, lc)


def test_lambda_tree_classes(golden):
    g = golden["mammals"]
    t = FlatTree(str(g["newick"]), str(g["lambda_newick"]))
    cls = dict(zip(t.names, t.lambda_class))
    assert cls["chimp"] == cls["human"] == cls["chimphuman"] == 1 and cls["orang"] == 0 and t.n_lambda == 2
    with pytest.raises(ValueError):
        FlatTree("((A:1,B:1):1,C:2)", "((A:1,X:1):1,C:1)")
    with pytest.raises(ValueError):
        parse_newick("(A:1,B:0)")       # clade.cpp:411-414: non-root branch lengths must be > 0


def test_interior_labels_are_kept(golden):
    t = FlatTree(str(golden["hymenoptera"]["newick"]))
    assert t.n_leaves == 10 and t.n_nodes == 19


def test_family_table_formats(tmp_path):
    p = tmp_path / "cafe.txt"
    p.write_text("Desc\tFamily ID\tA\tB\n\t f1\t5\t10\n\n\t f2\t5\t7\n")
    species, ids, counts = fam.read_gene_families(str(p))
    assert species == ["A", "B"] and ids == [" f1", " f2"] and counts.tolist() == [[5, 10], [5, 7]]
    p2 = tmp_path / "cafexp.txt"
    p2.write_text("#A\n#B\n3\t4\tfamX\n1\t0\tfamY\n")
    species, ids, counts = fam.read_gene_families(str(p2))
    assert species == ["A", "B"] and ids == ["famX", "famY"] and counts.tolist() == [[3, 4], [1, 0]]
    p3 = tmp_path / "empty.txt"
    p3.write_text("Desc\tFamily ID\tA\tB\n")
    with pytest.raises(ValueError):
        fam.read_gene_families(str(p3))


def test_derived_sizes():
    # reference user_data.h:26-27 floors and user_data.cpp:40-48
    assert fam.derive_sizes(np.array([[3, 72]])) == (170, 150)
    assert fam.derive_sizes(np.array([[150, 2]])) == (200, 188)
    assert fam.derive_sizes(np.array([[200, 2]])) == (250, 250)


def test_root_filter(golden):
    t = FlatTree("((A:1,B:3):7,(C:11,D:17):23)")
    counts = np.array([[1, 0, 0, 0], [1, 0, 0, 2], [0, 0, 3, 3], [0, 1, 1, 0]])   # columns D, C, B, A
    assert fam.exists_at_root(t, counts).tolist() == [False, True, False, True]
    g = golden["mammals"]
    assert int(g["n_before_filter"]) == 12653 and g["counts"].shape[0] == 10956   # SURVEY 3.1 [measured]


def test_uniform_prior_is_float():
    p = fam.uniform_prior(150)
    assert p.dtype == np.float32 and float(np.float64(p[0])) == 0.0066666668280959129
    r = fam.rootdist_prior({1: 2, 2: 2, 3: 2, 4: 2, 5: 1})
    assert r[0] == 0 and r[5] == np.float32(1) / np.float32(9)


def test_prior_tables_match_the_reference_constructors(ref, golden, tmp_path):
    """Row a12: the three priors of user_data::create_prior (src/user_data.cpp:176-206) as float32 tables, from the Python
    restatement and from the library's C++ builders (cafe_b200_io_make_prior), against the tables the reference's own
    root_equilibrium_distribution constructors give (ref_prior_table) -- bit for bit, including the length of the table."""
    from cafe5_b200 import io_cpp
    for R in (8, 45, 150, 188):
        want = ref.prior_table("uniform", R)[0]
        assert np.array_equal(fam.uniform_prior(R), want) and np.array_equal(io_cpp.make_prior("uniform", R), want)
    for lam, n in ((0.8, 150), (5.3, 45), (6.5, 45), (12.0, 150), (40.0, 188), (0.05, 30)):
        want = ref.prior_table("poisson", n, lam)[0]
        assert np.array_equal(fam.poisson_prior(lam, n), want), (lam, n)
        assert np.array_equal(io_cpp.make_prior("poisson", n, lam), want), (lam, n)
    rng = np.random.default_rng(3)
    for case in range(6):
        rd = {int(s): int(rng.integers(1, 5000)) for s in rng.choice(np.arange(0, 120), size=int(rng.integers(1, 40)), replace=False)}
        want = ref.prior_table("rootdist", rootdist=rd)[0]
        assert np.array_equal(fam.rootdist_prior(rd), want)
        path = tmp_path / ("rootdist%d.txt" % case)
        path.write_text("".join("%d\t%d\n" % kv for kv in rd.items()))
        assert np.array_equal(io_cpp.make_prior("rootdist", rootdist_path=path), want)
    # the committed fixtures carry the reference's tables too (the GPU box has no reference sources)
    g = golden["priors"]
    assert np.array_equal(fam.poisson_prior(float(g["small_poisson_lambda"]), 45), g["small_poisson_prior"])
    assert np.array_equal(fam.poisson_prior(0.8, 150), g["mammals_poisson08_prior"])
    rd = dict(zip(g["mammals_rootdist_sizes"].tolist(), g["mammals_rootdist_counts"].tolist()))
    assert np.array_equal(fam.rootdist_prior(rd), g["mammals_rootdist_prior"])


def test_poisson_prior_estimated_from_the_families(golden):
    """`-p` without a value (src/user_data.cpp:193-197): cafe_b200_fit_poisson_prior restates the reference's poisson_scorer
    (src/poisson.cpp:40-78) under the host driver's simplex search.  With the reference present: the same fitted mean, score and
    iteration count to the last bit for several seeds (the terms are added in the order of the reference's case-insensitive species
    map), and the same prior table as root_equilibrium_distribution(gene_families, num_values).  Always: the committed answers of that
    comparison."""
    from cafe5_b200 import io_cpp
    from oracle import pyoracle
    known = {"mammals": (0.7823844312347201, 230742.16544235396, 38), "hymenoptera": (0.17864856014870018, 50673.78984391159, 25)}
    for name, (lam, score, iters) in known.items():
        g = golden[name]
        species = [str(x) for x in g["species"]]
        got = io_cpp.fit_poisson_prior(g["counts"], species, seed=10)
        assert abs(got[0] - lam) <= 1e-12 * lam and abs(got[1] - score) <= 1e-12 * score and got[2] == iters
        shuffled = io_cpp.fit_poisson_prior(g["counts"][:, ::-1], species[::-1], seed=10)      # column order does not matter with names
        assert shuffled == got
    if not pyoracle.have_ref():
        return
    ref = pyoracle.RefLib()
    for name in known:
        g = golden[name]
        species = [str(x) for x in g["species"]]
        num_values = int(int(g["max_root_family_size"]) * 0.8)
        for seed in (10, 3, 77):
            lam, score, iters, table = ref.fit_poisson_prior(species, g["counts"], seed=seed, num_values=num_values)
            assert io_cpp.fit_poisson_prior(g["counts"], species, seed=seed) == (lam, score, iters), (name, seed)
            assert np.array_equal(io_cpp.make_prior("poisson", num_values, lam), table)
    # a table whose only counts are 0 has no terms: every start scores 0, the search ends where it started -- like the reference
    species = ["a", "B", "c"]
    zeros = np.zeros((4, 3), dtype=np.int32)
    assert io_cpp.fit_poisson_prior(zeros, species, seed=5)[:2] == ref.fit_poisson_prior(species, zeros, seed=5, num_values=10)[:2]
    # mixed-case names: the reference's map orders them case-insensitively
    rng = np.random.default_rng(8)
    species = ["zeta", "Alpha", "beta", "GAMMA", "delta"]
    counts = rng.integers(0, 40, size=(300, 5)).astype(np.int32)
    for seed in (1, 2):
        lam, score, iters, _ = ref.fit_poisson_prior(species, counts, seed=seed, num_values=30)
        assert io_cpp.fit_poisson_prior(counts, species, seed=seed) == (lam, score, iters)


def test_error_model_file_and_epsilon_table(tmp_path):
    p = tmp_path / "em.txt"
    p.write_text("maxcnt: 20\ncntdiff -1 0 1\n0 0.0 0.8 0.2\n1 0.2 0.6 0.2\n20 0.2 0.6 0.2\n")
    probs, maxcnt = fam.read_error_model(str(p))
    assert maxcnt == 20 and probs.shape == (21, 3) and probs[7].tolist() == [0.2, 0.6, 0.2]
    rows, _ = fam.epsilon_error_model(0.05, 10)
    assert rows[0].tolist() == [0.0, 0.95, 0.05] and rows[3].tolist() == [0.05, 1 - 0.1, 0.05]
    from cafe5_b200.model import error_model
    em = error_model(rows, 10)
    em.update_single_epsilon(0.07)                       # error_model.cpp:70-109
    assert em.probs[0].tolist() == [0.0, 1 - 0.07, 0.07] and em.probs[5].tolist() == [0.07, 1 - 0.14, 0.07]


def test_get_gamma_matches_reference(ref):
    for K in (2, 3, 4, 8):
        for alpha in (0.05, 0.3, 0.6, 0.65, 0.7, 1.0, 1.7, 5.0, 40.0):
            p, m = get_gamma(K, alpha)
            pr, mr = ref.get_gamma(K, alpha)
            assert np.array_equal(m, mr) and np.array_equal(p, pr)


def test_get_gamma_known_answer():
    assert get_gamma(4, 0.65)[1] == [0.062015465425384449, 0.37328920830134099, 0.99805780528212318, 2.5666375209911516]


def test_shard_bounds_cover_everything():
    for F in (1, 7, 10956, 1000000):
        for W in (1, 2, 3, 8):
            spans = [dist.shard_bounds(F, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_combine_partials_rejects_failures():
    assert dist.combine_partials([(1.5, 0), (2.25, 0)]) == (3.75, 0)
    assert dist.combine_partials([(1.5, 0), (math.inf, 0)])[0] == math.inf
    assert dist.combine_partials([(1.5, 2), (2.0, 0)]) == (math.inf, 2)


def test_plan_shards_is_a_balanced_partition(monkeypatch):
    """cafe_b200_plan_shards (host-only): a permutation of the families ordered by total count, cut into n_shards non-empty
    contiguous blocks whose cost under the subtree-pattern table plan is balanced (tables forced so that the small test job plans any)."""
    from cafe5_b200.model import plan_shards
    monkeypatch.setenv("CAFE_B200_TABLES", "force")
    monkeypatch.setenv("CAFE_B200_TABLE_FRAC", "0.75")
    rng = np.random.default_rng(12)
    t = FlatTree("(((A:1,B:1):1,(C:1,D:1):1):1,((E:1,F:1):1,(G:1,H:1):1):1)")
    F = 6000
    size = rng.integers(1, 60, size=F)
    counts = np.clip(size[:, None] + rng.integers(-2, 3, size=(F, 8)) * (size[:, None] > 20), 0, 80).astype(np.int32)   # large families vary more
    for n_shards in (1, 3, 8):
        order, bounds = plan_shards(t, counts, n_shards)
        assert sorted(order.tolist()) == list(range(F))
        assert bounds[0] == 0 and bounds[-1] == F and (np.diff(bounds) > 0).all()
        tot = counts.sum(axis=1)[order]
        assert (np.diff(tot) >= 0).all()                                   # ordered by total count
    order, bounds = plan_shards(t, counts, 8)
    sizes = np.diff(bounds)
    assert sizes[0] > sizes[-2]                                            # blocks of small, repetitive families are longer
    order, bounds = plan_shards(t, counts[:5], 8)                          # more shards than families
    assert len(bounds) == 9 and bounds[5] == 5 and (bounds[5:] == 5).all()


def test_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cafe_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cafe_b200_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.SYMBOLS)
    assert os.path.exists(_lib.LIB_PATH), "libcafe_b200.so must be built in-tree (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_a_gpu():
    """There is no CPU fallback: without a CUDA device create must return an error, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cafe5_b200.model import CafeError, Context
    t = FlatTree("(A:1,B:1);")
    with pytest.raises(CafeError):
        Context(t, np.array([[1, 2]], dtype=np.int32), 56, 8)


def test_group_constructors_reject_bad_arguments_before_touching_a_device():
    """cafe_b200_create_multi / cafe_b200_create_bucketed: hard errors are status codes with text (include/cafe_b200.h), the handle
    stays NULL, and without a GPU a well-formed call fails with the CUDA status -- never a CPU computation."""
    import torch
    lib = _lib.load()
    t = FlatTree("((A:1,B:1):1,C:2);")
    keep = (np.ascontiguousarray(t.parent, dtype=np.int32), np.ascontiguousarray(t.branch_length, dtype=np.float64),
            np.ascontiguousarray(t.leaf_col, dtype=np.int32), np.ascontiguousarray(t.lambda_class, dtype=np.int32))
    ct = _lib.CTree(t.n_nodes, _lib.ip(keep[0]), _lib.dp(keep[1]), _lib.ip(keep[2]), _lib.ip(keep[3]))
    counts = np.array([[1, 2, 3], [4, 0, 2]], dtype=np.int32)
    devs = np.array([0, 0], dtype=np.int32)
    ceilings = np.array([16, 32], dtype=np.int32)

    def multi(counts_p, F, devices_p, n_dev):
        h = ctypes.c_void_p(1)
        rc = lib.cafe_b200_create_multi(ctypes.byref(ct), counts_p, F, 3, 40, 20, devices_p, n_dev, ctypes.byref(h))
        return rc, h.value, lib.cafe_b200_last_error(None).decode()

    def bucketed(counts_p, F, ceil_p, n_ceil, mfs=40):
        h = ctypes.c_void_p(1)
        rc = lib.cafe_b200_create_bucketed(ctypes.byref(ct), counts_p, F, 3, mfs, 20, ceil_p, n_ceil, 0, ctypes.byref(h))
        return rc, h.value, lib.cafe_b200_last_error(None).decode()

    ERR_ARG, ERR_RANGE = 1, 3
    for rc, h, msg in (multi(_lib.ip(counts), 2, None, 2), multi(_lib.ip(counts), 2, _lib.ip(devs), 0),
                       multi(None, 2, _lib.ip(devs), 2), multi(_lib.ip(counts), 0, _lib.ip(devs), 2),
                       bucketed(_lib.ip(counts), 2, None, 2), bucketed(_lib.ip(counts), 2, _lib.ip(ceilings), 0),
                       bucketed(None, 2, _lib.ip(ceilings), 2)):
        assert rc == ERR_ARG and h is None and msg
    rc, h, msg = bucketed(_lib.ip(counts), 2, _lib.ip(np.array([0, 16], dtype=np.int32)), 2)
    assert rc == ERR_ARG and h is None and "positive" in msg
    rc, h, msg = bucketed(_lib.ip(counts), 2, _lib.ip(ceilings), 2, mfs=3)           # count 4 > max_family_size 3
    assert rc == ERR_RANGE and h is None and "max_family_size" in msg
    if not torch.cuda.is_available():
        for rc, h, msg in (multi(_lib.ip(counts), 2, _lib.ip(devs), 2), bucketed(_lib.ip(counts), 2, _lib.ip(ceilings), 2)):
            assert rc != 0 and h is None and msg


# ---- C++ host driver pieces of libcafe_b200.so that need no GPU (cafe5_b200/host/) ------------------------------

def test_cpp_discrete_gamma_matches_python_and_reference(ref):
    from cafe5_b200.model import discrete_gamma
    for K in (1, 2, 3, 4, 8):
        for alpha in (0.05, 0.3, 0.62731793802343, 0.65, 1.0, 2.5, 17.0):
            p, m = discrete_gamma(K, alpha)
            pp, mp = get_gamma(K, alpha)
            pr, mr = ref.get_gamma(K, alpha)
            assert p == pp and m == mp, (K, alpha)
            assert np.array_equal(m, mr) and np.array_equal(p, pr), (K, alpha)


def _rosenbrock(x):
    return (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2


def _walled(x):
    """A surface with an infinite wall (like a rejected parameter vector) and a flat tie region."""
    if x[0] <= 0 or x[1] > 3:
        return math.inf
    return abs(x[0] - 0.7) + round((x[1] - 1.3) ** 2, 3)


@pytest.mark.parametrize("fn,x0", [(_rosenbrock, [-1.2, 1.0]), (_rosenbrock, [0.0, 0.0]), (_walled, [0.002, 1.0]),
                                   (lambda x: (x[0] - 0.0018) ** 2 * 1e6 + 3.0, [0.0021]),
                                   (lambda x: sum((v - i) ** 2 for i, v in enumerate(x)), [0.5, 0.0, -1.0])])
def test_cpp_nelder_mead_follows_the_reference_fminsearch(ref, fn, x0):
    """Same objective, same start: the product's simplex search must visit the same points as fminsearch_min
    (src/optimizer.cpp:287-322) -- identical evaluation trace, iteration count and result, bit for bit."""
    from cafe5_b200.model import minimize
    trace_ref, trace_ours = [], []
    xr, fr, itr = ref.fminsearch(lambda x: (trace_ref.append(tuple(x)), fn(x))[1], x0)
    xo, fo, ito = minimize(lambda x: (trace_ours.append(tuple(x)), fn(x))[1], x0)
    assert trace_ours == trace_ref
    assert ito == itr and fo == fr and np.array_equal(xo, xr)
