import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.model import Context
from cafe5_b200.tree import FlatTree
g = np.load(os.path.join(ROOT, "tests", "golden", "mammals.npz"))
species = [str(s) for s in g["species"]]
counts = g["counts"].astype(np.int32)
mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
tree = FlatTree(str(g["newick"]), species=species)
ctx = Context(tree, counts, mfs, mrs)
ctx.set_prior(fam.uniform_prior(mrs))
res = [ctx.pvalues([0.001817441759677361], 1000, seed=s) for s in range(1, 9)]
print("families below 0.05 by seed:", [int((r < 0.05).sum()) for r in res])
print("families below 0.01 by seed:", [int((r < 0.01).sum()) for r in res])
print("mean |diff| between seeds: %.4f" % np.mean([np.abs(res[i] - res[0]).mean() for i in range(1, 8)]))
ctx.close()
