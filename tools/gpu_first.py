"""Development check run on the GPU box: prints parity statistics of the CUDA path against the golden
fixtures (reference outputs) and the C oracle, plus first timings.  Not part of the product."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cafe5_b200.tree import FlatTree
from cafe5_b200 import families as fam
from cafe5_b200.model import Context
from cafe5_b200.gamma import get_gamma

G = os.path.join(ROOT, "tests", "golden")

def ulps(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    ia = a.view(np.int64); ib = b.view(np.int64)
    return np.abs(ia - ib)

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    d = np.abs(a - b); m = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.where(b != 0, d / m, d)))

mats = np.load(os.path.join(G, "matrices.npz"))
mm = np.load(os.path.join(G, "mammals.npz"))
newick = str(mm["newick"]); species = [str(s) for s in mm["species"]]
counts = mm["counts"].astype(np.int32)
mfs, mrs = int(mm["max_family_size"]), int(mm["max_root_family_size"])
tree = FlatTree(newick, species=species)
t0 = time.time()
ctx = Context(tree, counts, mfs, mrs)
print("create %.3fs  F=%d U=%d N=%d" % (time.time() - t0, ctx.F, ctx.unique_families(), ctx.N))
ctx.set_prior(fam.uniform_prior(mrs))
for name in ("m171_a", "m171_b"):
    N, lam, t = mats[name + "_params"]
    got = ctx.get_matrix(lam, t)
    ref = mats["ref_" + name]
    u = ulps(got, ref)
    print(name, "max ulp", u.max(), "mean ulp %.3f" % u.mean(), "exact frac %.3f" % (u == 0).mean(), "rel", rel(got, ref))
rv = ctx.root_vectors([0.0018])
sub = mm["roots_sub"]
print("root vectors rel", rel(rv[sub], mm["ref_roots"]), "max ulp", ulps(rv[sub], mm["ref_roots"]).max())
for it in range(3):
    t0 = time.time(); neg, famlnl = ctx.eval_base([0.0018]); dt = time.time() - t0
    print("base: %.10f ref %.10f rel %.3e  fam rel %.3e  %.2f ms" % (neg, mm["ref_base_neg_lnl"], abs(neg - mm["ref_base_neg_lnl"]) / mm["ref_base_neg_lnl"], rel(famlnl, mm["ref_base_family_lnl"]), dt * 1e3), ctx.last_stats())
for it in range(3):
    t0 = time.time(); out = ctx.eval_gamma([0.0018], 0.7, mm["gamma_mult"], mm["gamma_probs"]); dt = time.time() - t0
    print("gamma: %.10f ref %.10f rel %.3e cat rel %.3e  %.2f ms" % (out["neg_lnl"], mm["ref_gamma_neg_lnl"], abs(out["neg_lnl"] - mm["ref_gamma_neg_lnl"]) / mm["ref_gamma_neg_lnl"], rel(out["cat_lk"][mm["gamma_sub"]], mm["ref_gamma_cat_lk"]), dt * 1e3), ctx.last_stats())
out = ctx.eval_gamma([0.0018], 0.6, mm["gamma_fail_mult"], mm["gamma_fail_probs"])
print("gamma fail:", out["neg_lnl"], out["n_failed"], np.nonzero(out["failed"])[0], "ref", mm["ref_gamma_fail_failed"])
# Pupko
rs = mm["rec_sub"]
ctx2 = Context(tree, counts[rs], mfs, mrs); ctx2.set_prior(fam.uniform_prior(mrs))
t0 = time.time(); rec = ctx2.reconstruct([0.0018]); dt = time.time() - t0
print("pupko base mismatches", int((rec["states"] != mm["ref_rec_base"]).sum()), "of", rec["states"].size, "%.1f ms" % (dt * 1e3))
rec = ctx2.reconstruct([0.0018], mm["gamma_mult"], mm["gamma_probs"])
print("pupko gamma mismatches", int((rec["states"] != mm["ref_rec_gamma_states"]).sum()), "cat", int((rec["cat_states"] != mm["ref_rec_gamma_cat_states"]).sum()))
# error model + lambda tree
tree3 = FlatTree(newick, str(mm["lambda_newick"]), species=species)
from cafe5_b200.model import error_model
ctx3 = Context(tree3, counts, mfs, mrs); ctx3.set_prior(fam.uniform_prior(mrs))
ctx3.set_error_model(error_model(mm["em_probs"], int(mm["em_maxcnt"])))
neg, famlnl = ctx3.eval_base(mm["em_lambdas"])
print("errormodel+lambdatree: %.10f ref %.10f rel %.3e fam rel %.3e" % (neg, mm["ref_em_neg_lnl"], abs(neg - mm["ref_em_neg_lnl"]) / mm["ref_em_neg_lnl"], rel(famlnl, mm["ref_em_family_lnl"])), ctx3.last_stats())
# synthetic big: 60 taxa
sys.path.insert(0, ROOT)
from cafe5_b200.synthetic import make_tree_newick, simulate_families
nw = make_tree_newick(60, seed=20261017)
tr = FlatTree(nw)
cp, mu = get_gamma(4, 0.65)
for F in (20000, 125000):
    t0 = time.time()
    boot = Context(tr, np.ones((1, tr.n_leaves), dtype=np.int32), 170, 150)
    cnt = simulate_families(tr, F, 0.002, mu, boot.get_matrix, seed=20261017)
    boot.close()
    print("simulate %d families %.2fs max count %d" % (F, time.time() - t0, cnt.max()))
    mfs5, mrs5 = fam.derive_sizes(cnt)
    c5 = Context(tr, cnt, mfs5, mrs5); c5.set_prior(fam.uniform_prior(mrs5))
    for it in range(3):
        t0 = time.time(); out = c5.eval_gamma([0.002], 0.65, mu, cp, want_family=False); dt = time.time() - t0
        st = c5.last_stats()
        flop = c5.unique_families() * 4 * 3377592.0
        print("F=%d U=%d gamma4: neg=%.6f nfail=%d  %.1f ms" % (F, c5.unique_families(), out["neg_lnl"], out["n_failed"], dt * 1e3), st, "prune TFLOP/s %.2f" % (flop / st["ms_prune"] / 1e9))
    c5.close()
