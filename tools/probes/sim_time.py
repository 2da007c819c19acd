import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from cafe5_b200.gamma import get_gamma
from cafe5_b200.model import Context
from cafe5_b200.synthetic import make_tree_newick
from cafe5_b200.tree import FlatTree
tree = FlatTree(make_tree_newick(60, seed=20261017))
cp, mu = get_gamma(4, 0.65)
ctx = Context(tree, np.ones((1, tree.n_leaves), dtype=np.int32), 170, 150)
for F in (125000, 1000000):
    roots = np.random.default_rng(1).integers(1, 125, size=F).astype(np.int32)
    ctx.simulate([0.002], roots[:1000], mu, cp)
    t = time.time(); r = ctx.simulate([0.002], roots, mu, cp, seed=7); dt = time.time() - t
    print("simulate %d families x %d nodes, K=4: %.3f s (incl. 400 matrices, CDF tables, D2H of the count table), max count %d, not at root %d"
          % (F, tree.n_nodes, dt, r["counts"].max(), r["n_not_at_root"]))
ctx.close()
