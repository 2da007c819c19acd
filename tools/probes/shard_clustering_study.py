"""Design study (CPU, numpy): strong scaling shards the SAME job over N GPUs, and every shard plans its own subtree-pattern tables.
Random contiguous shards share fewer patterns than the whole job does.  How much of the reuse survives when the job is first sorted
(families with similar counts next to each other) and THEN cut into contiguous shards?  Prints, per sharding, the contraction columns
of the table plan (frac 0.75 rule of plan_tables) summed over shards and for the slowest shard, relative to pruning every family through
every node."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200.gamma import get_gamma
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import OracleLib
from tools.probes.subtree_pattern_study import distinct_patterns

FRAC = 0.75


def plan_columns(tree, counts):
    counts = np.unique(counts, axis=0)
    U = counts.shape[0]
    children, D = distinct_patterns(tree, counts)
    root = tree.n_nodes - 1
    table = {}
    for v in sorted(D):                       # children precede parents
        if v == root:
            continue
        ok = all(tree.leaf_col[c] >= 0 or table.get(c, False) for c in children[v])
        table[v] = ok and D[v] <= FRAC * U
    cols = sum(D[v] if table[v] else U for v in D if v != root)
    return cols, U * (len(D) - 1)


def main():
    o = OracleLib()
    tree = FlatTree(make_tree_newick(60, seed=20261017))
    cp, mu = get_gamma(4, 0.65)
    mats = {}
    def P(lam, t):
        key = (round(lam, 12), round(t, 6))
        if key not in mats:
            mats[key] = o.matrix(171, lam, t)
        return mats[key]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    counts = simulate_families(tree, n, 0.002, mu, P, seed=20261017)
    whole, full = plan_columns(tree, counts)
    print("whole job (%d families): %.1f%% of the columns" % (n, 100.0 * whole / full))
    orders = {"as generated": np.arange(n),
              "sorted by mean count": np.argsort(counts.mean(axis=1), kind="stable"),
              "sorted lexicographically (leaf order)": np.lexsort(counts.T[::-1])}
    for shards in (2, 4, 8):
        for name, order in orders.items():
            c = counts[order]
            per = [plan_columns(tree, c[i * n // shards:(i + 1) * n // shards]) for i in range(shards)]
            tot = sum(p[0] for p in per)
            worst = max(p[0] for p in per)
            print("  %d shards, %-38s: sum %.1f%% of the columns, slowest shard %.1f%% of an equal share" % (
                shards, name, 100.0 * tot / full, 100.0 * worst / (full / shards)))


if __name__ == "__main__":
    main()


def check_library_plan(n=160000, shards=8):
    """cafe_b200_plan_shards (the library's host-side planner) on the same data: columns per shard under the plan."""
    from cafe5_b200.model import plan_shards
    o = OracleLib()
    tree = FlatTree(make_tree_newick(60, seed=20261017))
    cp, mu = get_gamma(4, 0.65)
    mats = {}
    def P(lam, t):
        key = (round(lam, 12), round(t, 6))
        if key not in mats:
            mats[key] = o.matrix(171, lam, t)
        return mats[key]
    counts = simulate_families(tree, n, 0.002, mu, P, seed=20261017)
    import time
    t0 = time.time()
    order, bounds = plan_shards(tree, counts, shards)
    dt = time.time() - t0
    assert sorted(order.tolist()) == list(range(n)) and bounds[0] == 0 and bounds[-1] == n
    per = [plan_columns(tree, counts[order[bounds[i]:bounds[i + 1]]]) for i in range(shards)]
    rnd = [plan_columns(tree, counts[i * n // shards:(i + 1) * n // shards]) for i in range(shards)]
    full = n * 58
    print("cafe_b200_plan_shards: %d families, %d shards, %.1f s; sizes %s" % (n, shards, dt, np.diff(bounds).tolist()))
    print("  planned shards : columns %s, slowest %.1f%% of an equal share of the unplanned job" % ([p[0] for p in per], 100.0 * max(p[0] for p in per) / (full / shards)))
    print("  random blocks  : columns %s, slowest %.1f%%" % ([p[0] for p in rnd], 100.0 * max(p[0] for p in rnd) / (full / shards)))
