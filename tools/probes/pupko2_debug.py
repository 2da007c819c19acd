import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.gamma import get_gamma
from cafe5_b200.model import Context
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree
cp, mu = get_gamma(4, 0.65)
t60 = FlatTree(make_tree_newick(60, seed=20261017))
boot = Context(t60, np.ones((1, t60.n_leaves), dtype=np.int32), 170, 150)
c60 = simulate_families(t60, 125000, 0.002, mu, boot.get_matrix, seed=20261017)
boot.close()
def run(counts, env):
    for k, v in env.items(): os.environ[k] = v
    ctx = Context(t60, counts, 170, 150)
    for k in env: del os.environ[k]
    ctx.set_prior(fam.uniform_prior(150))
    out = ctx.reconstruct([0.002], mu, cp, want_averaged=False)
    ctx.close()
    return out["cat_states"]
for n in (1000, 5000, 20000, 60000, 125000):
    a = run(c60[:n], {"CAFE_B200_PUPKO": "1"})
    b = run(c60[:n], {"CAFE_B200_TABLES": "0"})
    bad = np.argwhere(a != b)
    print("n=%d mismatches %d" % (n, len(bad)), flush=True)
    if len(bad):
        fams = np.unique(bad[:, 0]); nodes = np.unique(bad[:, 2]); cats = np.unique(bad[:, 1])
        print("  families", fams[:10], "... count", len(fams), "min", fams.min(), "max", fams.max())
        print("  cats", cats, "nodes", nodes)
        f, k, v = bad[0]
        print("  first:", bad[0], "v1", a[f, k, v], "v2", b[f, k, v])
        break
