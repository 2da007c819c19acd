"""bench.py's accounting helpers (no GPU): the algorithmic-flop formula of SURVEY.md 8(d) with and without the subtree-pattern plan."""
import importlib.util
import os

import numpy as np

from cafe5_b200.tree import FlatTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_flops_formula():
    b = _bench()
    # mammals-like numbers from SURVEY 8(d): 8 internal-child branches and 14 leaf branches under S = 171, R = 150 give
    # 8 * 2 * 171^2 + ... ; here a 4-leaf tree by hand: ((A,B),(C,D)): two internal children under the root, four leaf branches
    t = FlatTree("((A:1,B:1):1,(C:1,D:1):1)")
    S, R = 171, 150
    by_hand = 2 * (2 * R * S) + 4 * (2 * 1 * S) + (2 * S * (2 - 1) + R * (2 - 1)) + 2 * R
    assert b.alg_flops_per_prune(t, S, R) == by_hand
    assert b.alg_flops(t, S, R, None, 1000) == 1000 * by_hand
    # with the table plan the two cherries are computed for their own patterns only; everything gathered by the root still costs U
    cols = np.zeros(t.n_nodes, dtype=np.int64)
    internal = [i for i in range(t.n_nodes) if t.leaf_col[i] < 0]
    root = t.n_nodes - 1
    for v in internal:
        cols[v] = 1000 if v == root else 37
    planned = 2 * (2 * R * S) * 37 + 4 * (2 * 1 * S) * 37 + 2 * S * 37 + R * 1000 + 2 * R * 1000
    assert b.alg_flops(t, S, R, cols, 1000) == planned
    assert b.alg_flops_per_prune(t, S, R, nnz=3) == by_hand + 4 * (2 * 2 * S)       # error model: three columns per leaf


def test_matrix_terms_and_json_cleaning():
    b = _bench()
    assert b.matrix_terms(171) == 1681215                                          # binomial-sum terms of one matrix
    assert b._finite({"a": float("inf"), "b": [1.0, float("nan")], "c": 2}) == {"a": None, "b": [1.0, None], "c": 2}
    lam, alpha, cp, mu = b.step_params(3, 4)
    assert abs(lam - 0.002) < 5e-5 and len(cp) == len(mu) == 4 and abs(sum(cp) - 1) < 1e-12
