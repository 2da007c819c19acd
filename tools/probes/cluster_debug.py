import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.gamma import get_gamma
from cafe5_b200.model import Context, plan_shards
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree
cp, mu = get_gamma(4, 0.65)
t60 = FlatTree(make_tree_newick(60, seed=20261017))
boot = Context(t60, np.ones((1, t60.n_leaves), dtype=np.int32), 170, 150)
c60 = simulate_families(t60, 125000, 0.002, mu, boot.get_matrix, seed=20261017)
boot.close()
one = Context(t60, c60, 170, 150)
one.set_prior(fam.uniform_prior(150))
a = one.eval_gamma([0.002], 0.65, mu, cp)
order, bounds = plan_shards(t60, c60, 3)
pos = np.empty(len(order), dtype=np.int64); pos[order] = np.arange(len(order))
print("bounds", bounds)
for trial in range(8):
    multi = Context(t60, c60, 170, 150, devices=[0, 0, 0])
    multi.set_prior(fam.uniform_prior(150))
    for rep in range(2):
        b = multi.eval_gamma([0.002], 0.65, mu, cp)
        bad = np.flatnonzero((a["cat_lk"] != b["cat_lk"]).any(axis=1))
        msg = "trial %d rep %d: neg %r nf %d mismatching rows %d" % (trial, rep, b["neg_lnl"], b["n_failed"], len(bad))
        if len(bad):
            pp = np.sort(pos[bad])
            shard = np.searchsorted(bounds, pp, side="right") - 1
            msg += " shards %s pos range %d..%d failed flags %d zero rows %d; pos mod 64 of first %s" % (
                np.unique(shard), pp[0], pp[-1], int(b["failed"][bad].sum()), int((b["cat_lk"][bad] == 0).all(axis=1).sum()), (pp[:8] - bounds[shard[:8]]) % 64)
        print(msg, flush=True)
    multi.close()
