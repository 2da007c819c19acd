"""BASELINE config 1 as a whole analysis (what `cafe5 -i mammal_gene_families.txt -t mammals_tree.txt` does after reading its inputs):
lambda search -> family p-values (1000 simulations per root size) -> Pupko reconstruction -> per-branch probabilities -> output tables.
Stage wall times of the product (GPU + C++ host code) next to the unmodified reference's own functions on the host cores.
Development / measurement script (uses oracle/_ref, like the tests)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam, io_cpp
from cafe5_b200.model import Context
from cafe5_b200.tree import FlatTree
from oracle.pyoracle import RefLib

g = np.load(os.path.join(ROOT, "tests", "golden", "mammals.npz"))
species = [str(s) for s in g["species"]]
counts = g["counts"].astype(np.int32)
newick = str(g["newick"])
mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
prior = fam.uniform_prior(mrs)
ids = [str(i) for i in range(counts.shape[0])]
out = {}

tree = FlatTree(newick, species=species)
t0 = time.time()
ctx = Context(tree, counts, mfs, mrs)
ctx.set_prior(prior)
out["create_s"] = time.time() - t0
t = time.time(); fit = ctx.fit(seed=10); out["fit_s"] = time.time() - t
lam = list(fit["values"])
t = time.time(); neg, famlnl = ctx.eval_base(lam); out["final_eval_s"] = time.time() - t
t = time.time(); pv = ctx.pvalues(lam, 1000, seed=1); out["pvalues_s"] = time.time() - t
t = time.time(); states = ctx.reconstruct(lam)["states"]; out["reconstruct_s"] = time.time() - t
t = time.time(); bp = ctx.branch_probabilities(lam, states, selected=(pv < 0.05)); out["branch_probabilities_s"] = time.time() - t
t = time.time()
texts = [io_cpp.format_results("Base", neg, lam, ctx.describe()["longest_branch"], fit["evaluations"], 0),
         io_cpp.format_family_likelihoods(ids, "base", family_values=famlnl)]
texts += [io_cpp.format_reconstruction(newick, ids, states, w, pvalues=pv, branch_probs=bp)
          for w in ("count", "change", "asr", "family_results", "clade_results", "branch_probabilities")]
texts.append(io_cpp.format_report(newick, ids, states, pv, lambdas=lam, branch_probs=bp))      # Base_report.cafe
out["tables_s"] = time.time() - t
out["tables_bytes"] = sum(len(x) for x in texts)
ctx.close()
out["total_s"] = sum(v for k, v in out.items() if k.endswith("_s"))
out["lambda"] = lam[0]
out["significant_families"] = int((pv < 0.05).sum())
print("product:", json.dumps(out), flush=True)

if not os.environ.get("SKIP_CPU"):
    ref = RefLib()
    r = {"threads": ref.max_threads()}
    rctx = ref.ctx(newick, species, counts, mfs, mrs, prior)
    t = time.time(); o = rctx.optimize("cpu", seed=10, n_cat=0); r["fit_s"] = time.time() - t
    t = time.time(); rpv = rctx.pvalues(list(o["values"]), 1000, seed=1); r["pvalues_s"] = time.time() - t
    t = time.time(); rbp, tab, asr = rctx.branch_probabilities(list(o["values"]), rpv); r["reconstruct_and_branch_probabilities_s"] = time.time() - t
    rctx.close()
    r["total_s"] = r["fit_s"] + r["pvalues_s"] + r["reconstruct_and_branch_probabilities_s"]
    r["lambda"] = float(o["values"][0])
    r["significant_families"] = int((rpv < 0.05).sum())
    print("reference:", json.dumps(r), flush=True)
