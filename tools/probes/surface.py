import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from cafe5_b200 import families as fam
from cafe5_b200.model import Context, discrete_gamma
class A: taxa=60; families=125000; cats=4
for rank in (0, 1):
    tree, counts, mfs, mrs = bench.make_workload(A, rank, 0)
    ctx = Context(tree, counts, mfs, mrs); ctx.set_prior(fam.uniform_prior(mrs))
    for lam, al in ((0.002, 0.65), (0.00191, 0.649), (0.00244, 0.618), (0.0022, 0.63), (0.0018, 0.66)):
        cp, mu = discrete_gamma(4, al)
        o = ctx.eval_gamma([lam], al, mu, cp, want_family=False)
        print(rank, lam, al, o['neg_lnl'], o['n_failed'])
    r = ctx.fit(n_cat=4, start=[0.003, 1.0]); print(rank, 'fit', r)
    r = ctx.fit(n_cat=4, start=[0.002, 0.65]); print(rank, 'fit from truth', r)
    ctx.close()
