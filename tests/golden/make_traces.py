"""Generates tests/golden/trace_*.npz: the COMPLETE evaluation trace of the reference's own optimizer (src/optimizer.cpp,
src/optimizer_scorer.cpp, Nelder-Mead) over the reference's own CPU models at FULL size for the BASELINE configs, from a seeded
randomizer_engine: every attempted parameter vector, its score (+inf for rejected points), and which family made a gamma
evaluation fail.  The GPU tests replay every point through the C ABI (identical accept / reject, scores within 1e-12) and run
cafe_b200_fit and the drop-in shim from the same seed to the same fitted values.

Run in the build container (needs /root/reference compiled into oracle/_ref by oracle/build_ref.sh); CPU only, no GPU:
    python tests/golden/make_traces.py config2      # mammals, gamma K=4, (lambda, alpha): ~1 h on 8 threads
    python tests/golden/make_traces.py config3e     # mammals, two lambda classes, `-e` without a file: (lambda1, lambda2, epsilon)
    python tests/golden/make_traces.py config1      # mammals, base model, lambda
    python tests/golden/make_traces.py config4      # Hymenoptera, gamma K=8, (lambda, alpha)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam   # noqa: E402
from oracle.pyoracle import RefLib       # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 10

CONFIGS = {
    "config1": dict(data="mammals", n_cat=0, eps=False, lambda_tree=False),
    "config2": dict(data="mammals", n_cat=4, eps=False, lambda_tree=False),
    "config3e": dict(data="mammals", n_cat=0, eps=True, lambda_tree=True),
    "config4": dict(data="hymenoptera", n_cat=8, eps=False, lambda_tree=False),
}


def main():
    name = sys.argv[1]
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cfg = CONFIGS[name]
    g = np.load(os.path.join(GOLD, cfg["data"] + ".npz"))
    species = [str(s) for s in g["species"]]
    counts = g["counts"].astype(np.int32)
    mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
    ref = RefLib()
    if threads:
        ref.set_threads(threads)
    ctx = ref.ctx(str(g["newick"]), species, counts, mfs, mrs, fam.uniform_prior(mrs),
                  lambda_newick=str(g["lambda_newick"]) if cfg["lambda_tree"] else None)
    t0 = time.time()
    r = ctx.optimize("cpu", n_cat=cfg["n_cat"], optimize_epsilon=cfg["eps"], seed=SEED, trace=True)
    wall = time.time() - t0
    tr = r["trace"]
    np.savez_compressed(os.path.join(GOLD, "trace_%s.npz" % name), values=tr["values"], scores=tr["scores"],
                        failed_family=tr["failed_family"], n_failed=tr["n_failed"], fitted=r["values"], score=r["score"],
                        iterations=r["iterations"], attempts=r["attempts"], seed=SEED, n_cat=cfg["n_cat"],
                        optimize_epsilon=int(cfg["eps"]), lambda_tree=int(cfg["lambda_tree"]), data=cfg["data"],
                        wall_s=wall, threads=ref.max_threads())
    print(name, "fitted", r["values"], "score", r["score"], "iterations", r["iterations"], "attempts", r["attempts"],
          "trace rows", len(tr["scores"]), "rejected", int(np.isinf(tr["scores"]).sum()), "wall %.1fs" % wall)


if __name__ == "__main__":
    main()
