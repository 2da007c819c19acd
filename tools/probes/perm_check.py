import sys, numpy as np
sys.path.insert(0, '/root/repo')
from cafe5_b200 import families as fam
from cafe5_b200.gamma import get_gamma
from cafe5_b200.model import Context
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree
tree = FlatTree(make_tree_newick(60, seed=20261017))
cp, mu = get_gamma(4, 0.65)
boot = Context(tree, np.ones((1, tree.n_leaves), dtype=np.int32), 170, 150)
counts = simulate_families(tree, 125000, 0.002, mu, boot.get_matrix, seed=20261017)
boot.close()
prior = fam.uniform_prior(150)
ctx = Context(tree, counts, 170, 150); ctx.set_prior(prior)
a = ctx.eval_gamma([0.002], 0.65, mu, cp)["cat_lk"]
for rep in range(3):
    b = ctx.eval_gamma([0.002], 0.65, mu, cp)["cat_lk"]
    d = a != b
    print("repeat", rep, "mismatching entries", int(d.sum()), "max rel", float(np.max(np.abs(a - b) / np.maximum(a, 1e-300))))
ctx.close()
rng = np.random.default_rng(11)
perm = rng.permutation(counts.shape[0])
ctx = Context(tree, counts[perm], 170, 150); ctx.set_prior(prior)
b = ctx.eval_gamma([0.002], 0.65, mu, cp)["cat_lk"]
d = b != a[perm]
print("permuted: mismatching entries", int(d.sum()), "families", int(d.any(axis=1).sum()), "max rel", float(np.max(np.abs(a[perm] - b) / np.maximum(b, 1e-300))))
idx = np.flatnonzero(d.any(axis=1))[:10]
print(idx, [(perm[i]) for i in idx])
print(np.abs(a[perm][idx] - b[idx]) / b[idx])
ctx.close()
