"""Synthetic workload generator for BASELINE config 5 (families simulated on an ultrametric tree).

Semantics follow the reference simulator (src/simulator.cpp:29-58 create_trial,
src/probability.cpp:449-476 set_weighted_random_family_size, src/matrix_cache.cpp:60-66
select_random_y): the root size is drawn from a uniform root distribution, every child size is drawn
from row `parent size` of the branch's transition matrix restricted to sizes < max_sim (120, the
reference's default max_family_size, src/user_data.h:26), a lost family stays lost, and families
that do not exist at the root (src/gene_family.cpp:62-91) are rejected.  Each family picks one gamma
category with probability 1/K (gamma_model::get_simulation_lambda, src/gamma_core.cpp:91-95).

The transition matrices come from a caller-supplied provider `get_matrix(lambda, t) -> [N, N]`; bench.py
passes the CUDA library's cafe_b200_get_matrix, CPU tests pass the oracle's.  The reference binary
cannot generate this data set itself in useful time (0.3 s per family).
"""
import numpy as np

from .families import exists_at_root


def make_tree_newick(n_taxa, seed=20261017, root_height=100.0):
    """Random binary ultrametric tree (pure-birth / Yule shape): node heights on a 0.001 grid, every
    branch >= 0.001, root at `root_height`.  Returns a newick string with taxa t0..t{n-1}."""
    rng = np.random.default_rng(seed)
    n = int(n_taxa)
    # coalescent-style construction backwards in time with Yule waiting times
    waits = np.array([rng.exponential(1.0 / k) for k in range(n, 1, -1)])
    heights = np.cumsum(waits)
    heights = heights / heights[-1] * root_height
    grid = np.maximum(np.round(heights * 1000).astype(np.int64), 1)
    for i in range(1, len(grid)):          # strictly increasing on the grid -> positive branch lengths
        if grid[i] <= grid[i - 1]:
            grid[i] = grid[i - 1] + 1
    active = [("t%d" % i, 0) for i in range(n)]   # (newick text, height in grid units)
    for h in grid:
        i, j = sorted(rng.choice(len(active), size=2, replace=False))
        a, b = active[i], active[j]
        text = "(%s:%.3f,%s:%.3f)" % (a[0], (h - a[1]) / 1000.0, b[0], (h - b[1]) / 1000.0)
        active = [x for k, x in enumerate(active) if k not in (i, j)] + [(text, int(h))]
    return active[0][0] + ";"


def _sample_children(parent_sizes, matrix, max_sim, rng):
    """Draw child sizes given parent sizes from matrix rows restricted to [0, max_sim)."""
    out = np.zeros_like(parent_sizes)
    u = rng.random(parent_sizes.shape[0])
    for s in np.unique(parent_sizes):
        if s == 0:
            continue                        # probability.cpp:459: parent 0 -> child 0
        idx = np.nonzero(parent_sizes == s)[0]
        w = matrix[s, :max_sim]
        cdf = np.cumsum(w)
        out[idx] = np.minimum(np.searchsorted(cdf, u[idx] * cdf[-1], side="right"), max_sim - 1)
    return out


def simulate_families(tree, n_families, lam, multipliers, get_matrix, seed=20261017, max_sim=120, max_root=124):
    """Return counts[n_families, n_leaves] int32 (columns in tree.species order)."""
    rng = np.random.default_rng(seed)
    K = len(multipliers)
    mats = {}
    for k in range(K):
        for i in range(tree.n_nodes):
            if tree.parent[i] >= 0:
                key = (k, float(tree.branch_length[i]))
                if key not in mats:
                    mats[key] = np.asarray(get_matrix(lam * multipliers[k], float(tree.branch_length[i])))
    chunks = []
    have = 0
    while have < n_families:
        n = int((n_families - have) * 1.15) + 64
        cat = rng.integers(0, K, size=n)
        sizes = np.zeros((tree.n_nodes, n), dtype=np.int64)
        sizes[tree.n_nodes - 1] = rng.integers(1, max_root + 1, size=n)
        for i in range(tree.n_nodes - 2, -1, -1):          # parents have larger indices
            par = sizes[tree.parent[i]]
            for k in range(K):
                sel = np.nonzero(cat == k)[0]
                sizes[i, sel] = _sample_children(par[sel], mats[(k, float(tree.branch_length[i]))], max_sim, rng)
        counts = np.zeros((n, tree.n_leaves), dtype=np.int32)
        for i in range(tree.n_nodes):
            if tree.leaf_col[i] >= 0:
                counts[:, tree.leaf_col[i]] = sizes[i]
        keep = exists_at_root(tree, counts)
        chunks.append(counts[keep])
        have += int(keep.sum())
    return np.concatenate(chunks)[:n_families]
