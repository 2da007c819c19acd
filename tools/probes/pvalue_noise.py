"""Is cafe_b200_pvalues noisier or biased against the reference's compute_pvalues?  Test infrastructure
(uses oracle/_ref).  The 150 conditional distributions are SHARED by all 10,956 families, so the family-averaged |difference| of two
runs is a statistic with few degrees of freedom: one pair of reference seeds says little.  This probe runs the reference with
N_REF seeds and the CUDA path with N_GPU seeds on BASELINE config 1 (1000 simulations per root size) and prints
  * the family-averaged |p_a - p_b| for every pair within each implementation and across them (mean, min, max over pairs),
  * the signed difference of the seed-averaged p-values with its standard error over families,
  * the same per family for the families that decide significance (p < 0.1)."""
import itertools, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.model import Context
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import RefLib

N_REF = int(os.environ.get("N_REF", "5"))
N_GPU = int(os.environ.get("N_GPU", "8"))
LAM = 0.001817441759677361
g = np.load(os.path.join(ROOT, "tests", "golden", "mammals.npz"))
species = [str(s) for s in g["species"]]
counts = g["counts"].astype(np.int32)
mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
tree = FlatTree(str(g["newick"]), species=species)
ctx = Context(tree, counts, mfs, mrs)
ctx.set_prior(fam.uniform_prior(mrs))
t0 = time.time()
ours = np.stack([ctx.pvalues([LAM], 1000, seed=s) for s in range(1, N_GPU + 1)])
t_gpu = (time.time() - t0) / N_GPU
ctx.close()
ref = RefLib()
ref.set_threads(os.cpu_count() or 1)
rctx = ref.ctx(str(g["newick"]), species, counts, mfs, mrs, fam.uniform_prior(mrs))
t0 = time.time()
theirs = np.stack([rctx.pvalues([LAM], 1000, seed=s) for s in range(1, N_REF + 1)])
t_ref = (time.time() - t0) / N_REF
rctx.close()


def pair_stats(a, b=None):
    if b is None:
        d = [np.abs(a[i] - a[j]).mean() for i, j in itertools.combinations(range(len(a)), 2)]
    else:
        d = [np.abs(x - y).mean() for x in a for y in b]
    return "mean %.4f  min %.4f  max %.4f over %d pairs" % (np.mean(d), np.min(d), np.max(d), len(d))


print("cafe_b200_pvalues %.2f s per run, reference compute_pvalues %.1f s per run (%d threads)" % (t_gpu, t_ref, ref.max_threads()))
print("family-averaged |p_a - p_b|, 1000 simulations per root size:")
print("  reference seed vs reference seed :", pair_stats(theirs))
print("  CUDA seed      vs CUDA seed      :", pair_stats(ours))
print("  CUDA seed      vs reference seed :", pair_stats(ours, theirs))
mo, mt = ours.mean(axis=0), theirs.mean(axis=0)
d = mo - mt
# per-family noise of the two seed averages
se = np.sqrt(ours.var(axis=0, ddof=1) / N_GPU + theirs.var(axis=0, ddof=1) / N_REF)
print("signed difference of the seed averages (CUDA - reference): mean %.5f, |mean| over families %.5f; typical per-family s.e. %.5f"
      % (d.mean(), np.abs(d).mean(), np.median(se[se > 0])))
for lo, hi in ((0.0, 0.01), (0.01, 0.05), (0.05, 0.1), (0.1, 0.5), (0.5, 1.01)):
    m = (mt >= lo) & (mt < hi)
    if m.any():
        print("  reference p in [%.2f, %.2f): %5d families, signed diff %+.5f, reference seed sd %.5f, CUDA seed sd %.5f"
              % (lo, hi, int(m.sum()), d[m].mean(), np.sqrt(theirs[:, m].var(axis=0, ddof=1).mean()), np.sqrt(ours[:, m].var(axis=0, ddof=1).mean())))
print("families below 0.05 by seed: CUDA", [int((r < 0.05).sum()) for r in ours], "reference", [int((r < 0.05).sum()) for r in theirs])
# a z-test per family on the seed averages: fraction of families whose two averages differ by more than 3 s.e.
z = np.abs(d[se > 0]) / se[se > 0]
print("families whose seed averages differ by > 3 s.e.: %.2f%% (expected ~0.3%% for equal distributions, heavier tails with %d + %d seeds)"
      % (100.0 * (z > 3).mean(), N_GPU, N_REF))
