"""Wall time of the family-level p-values (BASELINE config 1 data: 10,956 mammals families, 150 root sizes x 1000 simulations):
the reference's compute_pvalues on the host cores vs cafe_b200_pvalues."""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.model import Context
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import RefLib
g = np.load(os.path.join(ROOT, "tests", "golden", "mammals.npz"))
species = [str(s) for s in g["species"]]
counts = g["counts"].astype(np.int32)
mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
lam = [0.001817441759677361]
tree = FlatTree(str(g["newick"]), species=species)
ctx = Context(tree, counts, mfs, mrs)
ctx.set_prior(fam.uniform_prior(mrs))
ctx.pvalues(lam, 10, seed=1)
t = time.time(); ours = ctx.pvalues(lam, 1000, seed=1); t_ours = time.time() - t
ctx.close()
print("cafe_b200_pvalues: %.3f s for %d families (150 x 1000 simulated families); %d families below 0.05, %d below 0.01"
      % (t_ours, counts.shape[0], int((ours < 0.05).sum()), int((ours < 0.01).sum())), flush=True)
if not os.environ.get("SKIP_CPU"):
    ref = RefLib()
    rctx = ref.ctx(str(g["newick"]), species, counts, mfs, mrs, fam.uniform_prior(mrs))
    t = time.time(); want = rctx.pvalues(lam, 1000, seed=1); t_ref = time.time() - t
    d = np.abs(ours - want)
    print("reference compute_pvalues (%d threads): %.1f s; %d below 0.05, %d below 0.01; |diff| mean %.4f, 95th pct %.4f, max %.3f, corr %.5f"
          % (ref.max_threads(), t_ref, int((want < 0.05).sum()), int((want < 0.01).sum()), d.mean(), np.quantile(d, 0.95), d.max(), np.corrcoef(ours, want)[0, 1]))
    want2 = rctx.pvalues(lam, 1000, seed=2)
    np.savez(os.path.join(ROOT, "gpurun_out", "pvalues.npz"), ours=ours, ref1=want, ref2=want2, maxcount=counts.max(axis=1))
    print("reference seed 1 vs seed 2: |diff| mean %.4f; signed mean ours - ref %.5f, ref2 - ref %.5f" % (np.abs(want - want2).mean(), (ours - want).mean(), (want2 - want).mean()))
