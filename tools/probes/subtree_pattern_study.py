"""Design study (CPU, numpy; uses the C oracle for the generator's matrices): how many of the contractions of one pruning pass
collapse when the factor W_v = P_v . V_v of a subtree is computed once per DISTINCT pattern of leaf counts below v instead of once
per family?  (base_model.cpp:27-51 already prunes identical families once; this extends the idea below the root.)
Prints, for every internal node, the number of distinct patterns D_v among the U unique families, and the contraction columns of
three strategies: per family (today), maximal compressible subtrees with D_v <= frac * U, full per-node compression."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200.gamma import get_gamma
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import OracleLib


def distinct_patterns(tree, counts):
    n = tree.n_nodes
    children = [[] for _ in range(n)]
    for v in range(n):
        if tree.parent[v] >= 0:
            children[tree.parent[v]].append(v)
    ids = {}
    D = {}
    for v in range(n):
        if tree.leaf_col[v] >= 0:
            ids[v] = counts[:, tree.leaf_col[v]].astype(np.int64)
            continue
        key = np.stack([ids[c] for c in children[v]], axis=1)
        _, inv = np.unique(key, axis=0, return_inverse=True)
        ids[v] = inv.reshape(-1).astype(np.int64)
        D[v] = int(ids[v].max()) + 1
    return children, D


def report(name, tree, counts):
    counts = np.unique(counts, axis=0)
    U = counts.shape[0]
    children, D = distinct_patterns(tree, counts)
    root = tree.n_nodes - 1
    internal = [v for v in D if v != root]
    n_int = {}
    for v in range(tree.n_nodes):
        if tree.leaf_col[v] >= 0:
            n_int[v] = 0
        else:
            n_int[v] = 1 + sum(n_int[c] for c in children[v])
    n_leaves = {}
    for v in range(tree.n_nodes):
        n_leaves[v] = 1 if tree.leaf_col[v] >= 0 else sum(n_leaves[c] for c in children[v])
    print("== %s: U = %d unique families, %d contractions per (family, category)" % (name, U, len(internal)))
    print("   distinct patterns by subtree size:", sorted((n_leaves[v], D[v]) for v in internal))
    per_family = U * len(internal)
    full = sum(D[v] for v in internal)
    print("   per family (today): %d columns; full per-node compression: %d (%.1f%%)" % (per_family, full, 100.0 * full / per_family))
    for frac in (0.05, 0.1, 0.25, 0.5):
        # maximal subtrees whose pattern count is <= frac * U; inside them every node runs over the subtree's D_v patterns... or its own
        cut = set()
        def walk(v):
            if tree.leaf_col[v] >= 0:
                return
            if v != root and D[v] <= frac * U:
                cut.add(v)
                return
            for c in children[v]:
                walk(c)
        walk(root)
        def inside(v):   # internal nodes of the subtree of v (v included)
            out = [v]
            for c in children[v]:
                if tree.leaf_col[c] < 0:
                    out += inside(c)
            return out
        covered = set()
        cost_one_level = 0       # pre-pass prunes each cut subtree over ITS D_v patterns (no recursion)
        cost_recursive = 0       # every node inside a cut subtree over its own patterns
        for v in cut:
            nodes = inside(v)
            covered.update(nodes)
            cost_one_level += D[v] * len(nodes)
            cost_recursive += sum(D[w] for w in nodes)
        rest = [v for v in internal if v not in covered]
        print("   frac %.2f: %d cut subtrees covering %d of %d contractions; columns one-level %d (%.1f%%), recursive %d (%.1f%%)"
              % (frac, len(cut), len(covered), len(internal), cost_one_level + U * len(rest), 100.0 * (cost_one_level + U * len(rest)) / per_family,
                 cost_recursive + U * len(rest), 100.0 * (cost_recursive + U * len(rest)) / per_family))


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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
    counts = simulate_families(tree, n, 0.002, mu, P, seed=20261017)
    report("config-5 shard (%d simulated families, 60 taxa)" % n, tree, counts)
    for nm in ("mammals", "hymenoptera"):
        g = np.load(os.path.join(ROOT, "tests", "golden", nm + ".npz"))
        t = FlatTree(str(g["newick"]), species=[str(s) for s in g["species"]])
        report(nm, t, g["counts"].astype(np.int32))


if __name__ == "__main__":
    main()
