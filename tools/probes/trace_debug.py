import os, sys, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200 import model as cafe
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import RefLib
name = sys.argv[1]
tr = np.load(os.path.join(ROOT, "tests", "golden", "trace_%s.npz" % name))
g = np.load(os.path.join(ROOT, "tests", "golden", str(tr["data"]) + ".npz"))
species = [str(s) for s in g["species"]]
lam_newick = str(g["lambda_newick"]) if int(tr["lambda_tree"]) else None
tree = FlatTree(str(g["newick"]), lam_newick, species=species)
counts = g["counts"].astype(np.int32)
mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
ctx = cafe.Context(tree, counts, mfs, mrs)
ctx.set_prior(fam.uniform_prior(mrs))
fit = ctx.fit(seed=int(tr["seed"]), n_cat=int(tr["n_cat"]), optimize_epsilon=bool(int(tr["optimize_epsilon"])))
print("trace : attempts", int(tr["attempts"]), "iterations", int(tr["iterations"]), "fitted", tr["fitted"], "score %.10f" % float(tr["score"]))
print("fit   : evaluations", fit["evaluations"], "iterations", fit["iterations"], "fitted", fit["values"], "score %.10f" % fit["neg_lnl"])
print("rel diff of fitted values", np.abs(fit["values"] - tr["fitted"]) / np.abs(tr["fitted"]))
ctx.close()
ref = RefLib()
rctx = ref.ctx(str(g["newick"]), species, counts, mfs, mrs, fam.uniform_prior(mrs), lambda_newick=lam_newick)
got = rctx.optimize("gpu", n_cat=int(tr["n_cat"]), optimize_epsilon=bool(int(tr["optimize_epsilon"])), seed=int(tr["seed"]), trace=True)
t = got["trace"]
print("shim  : attempts", got["attempts"], "iterations", got["iterations"], "fitted", got["values"], "score %.10f" % got["score"])
n = min(len(t["scores"]), len(tr["scores"]))
same = np.all(t["values"][:n] == tr["values"][:n], axis=1)
first = int(np.argmin(same)) if not same.all() else n
print("first differing point:", first, "of", len(tr["scores"]), "/", len(t["scores"]))
lo = max(0, first - 6)
for i in range(lo, min(n, first + 2)):
    print(i, tr["values"][i], "%.12f" % tr["scores"][i], "|", t["values"][i], "%.12f" % t["scores"][i], "diff %.3e" % (t["scores"][i] - tr["scores"][i]))
