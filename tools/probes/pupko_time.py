import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.gamma import get_gamma
from cafe5_b200.model import Context
from cafe5_b200.synthetic import make_tree_newick
from cafe5_b200.tree import FlatTree
tree = FlatTree(make_tree_newick(60, seed=20261017))
cp, mu = get_gamma(4, 0.65)
boot = Context(tree, np.ones((1, tree.n_leaves), dtype=np.int32), 170, 150)
roots = np.random.default_rng(1).integers(1, 125, size=125000).astype(np.int32)
counts = boot.simulate([0.002], roots, mu, cp, seed=7)["counts"]
boot.close()
ctx = Context(tree, counts, 170, 150)
ctx.set_prior(fam.uniform_prior(150))
for _ in range(2):
    ctx.reconstruct([0.002], mu, cp, want_cat_states=False, want_averaged=False)
print("unique", ctx.unique_families())
ctx.close()
