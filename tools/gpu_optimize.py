"""Wall time of a whole parameter optimisation (BASELINE configs 1 and 2) with the REFERENCE's own optimizer
(src/optimizer.cpp, compiled unmodified into oracle/_ref) driving either the reference's CPU models or the CUDA
models through cafe5_b200/host/gpu_model.hpp.  Development / measurement script (uses oracle/, like the tests)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from oracle.pyoracle import RefLib

g = np.load(os.path.join(ROOT, "tests", "golden", "mammals.npz"))
species = [str(s) for s in g["species"]]
counts = g["counts"].astype(np.int32)
mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
ref = RefLib()
ctx = ref.ctx(str(g["newick"]), species, counts, mfs, mrs, fam.uniform_prior(mrs))
out = {"threads": ref.max_threads(), "families": int(counts.shape[0])}
for name, kw, backends in (("config1_base", dict(n_cat=0), ("gpu", "cpu")), ("config2_gamma_k4", dict(n_cat=4), ("gpu",))):
    for b in backends:
        if b == "cpu" and os.environ.get("SKIP_CPU"):
            continue
        t = time.time()
        r = ctx.optimize(b, seed=10, **kw)
        r["wall_s"] = time.time() - t
        r["values"] = [float(v) for v in r["values"]]
        out["%s_%s" % (name, b)] = r
        print(name, b, json.dumps(r), flush=True)
# the product's own host driver (cafe_b200_fit) on the same data, same seed
from cafe5_b200.model import Context
from cafe5_b200.tree import FlatTree
tree = FlatTree(str(g["newick"]), species=species)
pctx = Context(tree, counts, mfs, mrs)
pctx.set_prior(fam.uniform_prior(mrs))
for name, kw in (("config1_base", dict(n_cat=0)), ("config2_gamma_k4", dict(n_cat=4)), ("config1_base_epsilon", dict(n_cat=0, optimize_epsilon=True))):
    t = time.time()
    r = pctx.fit(seed=10, **kw)
    r["wall_s"] = time.time() - t
    r["values"] = [float(v) for v in r["values"]]
    out["%s_product" % name] = r
    print(name, "product driver", json.dumps(r), flush=True)
pctx.close()
# config 3: error model from file + two lambda classes (chimp/human separate); config 4: Hymenoptera, gamma K=8
from cafe5_b200.model import error_model
tree3 = FlatTree(str(g["newick"]), str(g["lambda_newick"]), species=species)
pctx = Context(tree3, counts, mfs, mrs)
pctx.set_prior(fam.uniform_prior(mrs))
pctx.set_error_model(error_model(g["em_probs"], int(g["em_maxcnt"])))
t = time.time(); r = pctx.fit(seed=10, n_cat=0); r["wall_s"] = time.time() - t; r["values"] = [float(v) for v in r["values"]]
out["config3_errormodel_two_lambdas_product"] = r
print("config3_errormodel_two_lambdas product driver", json.dumps(r), flush=True)
pctx.close()
if not os.environ.get("SKIP_CPU"):   # the reference's optimizer on its CPU models, same seed (~3 minutes on 16 threads)
    rctx3 = ref.ctx(str(g["newick"]), species, counts, mfs, mrs, fam.uniform_prior(mrs), lambda_newick=str(g["lambda_newick"]),
                    em=(g["em_probs"], int(g["em_maxcnt"])))
    t = time.time(); r = rctx3.optimize("cpu", seed=10, n_cat=0); r["wall_s"] = time.time() - t; r["values"] = [float(v) for v in r["values"]]
    out["config3_errormodel_two_lambdas_cpu"] = r
    print("config3_errormodel_two_lambdas reference cpu", json.dumps(r), flush=True)
    rctx3.close()
h = np.load(os.path.join(ROOT, "tests", "golden", "hymenoptera.npz"))
treeh = FlatTree(str(h["newick"]), species=[str(x) for x in h["species"]])
hm, hr = int(h["max_family_size"]), int(h["max_root_family_size"])
pctx = Context(treeh, h["counts"].astype(np.int32), hm, hr)
pctx.set_prior(fam.uniform_prior(hr))
t = time.time(); r = pctx.fit(seed=10, n_cat=8); r["wall_s"] = time.time() - t; r["values"] = [float(v) for v in r["values"]]
out["config4_hymenoptera_gamma_k8_product"] = r
print("config4_hymenoptera_gamma_k8 product driver", json.dumps(r), flush=True)
pctx.close()
# Pupko reconstruction (row a13) wall times through the host API (all states copied back)
from cafe5_b200.gamma import get_gamma
cp, mu = get_gamma(4, 0.65)
pctx = Context(tree, counts, mfs, mrs)
pctx.set_prior(fam.uniform_prior(mrs))
for label, args in (("base", ([0.0018], None, None)), ("gamma_k4", ([0.0018], mu, cp))):
    pctx.reconstruct(*args)
    t = time.time(); pctx.reconstruct(*args); dt = time.time() - t
    out["pupko_mammals_%s_s" % label] = dt
    print("pupko mammals %s: %.4f s for %d families" % (label, dt, counts.shape[0]), flush=True)
pctx.close()
from cafe5_b200.synthetic import make_tree_newick, simulate_families
t60 = FlatTree(make_tree_newick(60, seed=20261017))
boot = Context(t60, np.ones((1, t60.n_leaves), dtype=np.int32), 170, 150)
c60 = simulate_families(t60, 125000, 0.002, mu, boot.get_matrix, seed=20261017)
boot.close()
pctx = Context(t60, c60, 170, 150)
pctx.set_prior(fam.uniform_prior(150))
pctx.reconstruct([0.002], mu, cp)
t = time.time(); pctx.reconstruct([0.002], mu, cp); dt = time.time() - t
out["pupko_config5_shard_gamma_k4_s"] = dt
print("pupko config-5 shard (125000 families, 60 taxa, K=4): %.3f s" % dt, flush=True)
pctx.close()
print(json.dumps(out))
