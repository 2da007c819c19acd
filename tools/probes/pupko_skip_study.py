"""Design study for the Pupko kernel (CPU, numpy; test infrastructure - uses the C oracle for matrices): how much of the max-product
work could an exact bound test skip?  A (row block of 11 parent states) x (chunk of 8 child states) unit of one 64-family tile can be
skipped when  max_i,j P(i->j) * max_j,col M[j][col]  <=  min_i,col best[i][col]  over the block - then no product in the unit can
replace a running maximum (rounding is monotone, the comparison is strict).  Chunks are visited in ascending j like the kernel.
Prints the fraction of units that survive, for families in table order and sorted by their largest count."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200.gamma import get_gamma
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import OracleLib

N, S, BN, RB, CK = 171, 171, 64, 11, 8
o = OracleLib()
tree = FlatTree(make_tree_newick(60, seed=20261017))
cp, mu = get_gamma(4, 0.65)
mats = {}
def P(lam, t):
    key = (round(lam, 12), round(t, 6))
    if key not in mats:
        mats[key] = o.matrix(N, lam, t)
    return mats[key]
counts = simulate_families(tree, 4096, 0.002, mu, P, seed=20261017)
n = tree.n_nodes
children = [[] for _ in range(n)]
for v in range(n):
    if tree.parent[v] >= 0:
        children[tree.parent[v]].append(v)
for c in children:
    c.sort(reverse=True)
leaf_col = np.asarray(tree.leaf_col)

def study(tile_counts, lam):
    total = kept = 0
    L = {}
    for v in range(n):                       # reverse level order: children before parents
        if not children[v]:
            continue
        M = np.ones((S, tile_counts.shape[0]))
        for c in children[v]:
            if not children[c]:
                M *= P(lam, tree.branch_length[c])[:, tile_counts[:, leaf_col[c]]]      # L[j] = P(j -> obs)
            else:
                M *= L[c]
        if tree.parent[v] < 0:
            break
        Pm = P(lam, tree.branch_length[v])                                               # [i, j]
        best = np.full((S, tile_counts.shape[0]), -1.0)
        for j0 in range(0, S, CK):
            js = slice(j0, min(j0 + CK, S))
            bmax = M[js].max(axis=0)                                                     # per column
            for i0 in range(0, S, RB):
                isl = slice(i0, min(i0 + RB, S))
                total += 1
                bound = Pm[isl, js].max() * bmax.max()
                if bound <= best[isl].min():
                    continue
                kept += 1
                prod = Pm[isl, js][:, :, None] * M[js][None, :, :]                       # [i, j, col]
                best[isl] = np.maximum(best[isl], prod.max(axis=1))
        L[v] = best
    return kept, total

for name, order in (("table order", np.arange(counts.shape[0])), ("sorted by largest count", np.argsort(counts.max(axis=1), kind="stable"))):
    kept = total = 0
    for t in range(0, 512, BN):                                                          # eight tiles spread over the table
        idx = order[t * 8:t * 8 + BN]
        for k in range(4):
            a, b = study(counts[idx], 0.002 * mu[k])
            kept += a
            total += b
    print("%s: %d of %d (row block x chunk) units survive the bound test = %.1f %%" % (name, kept, total, 100.0 * kept / total), flush=True)
