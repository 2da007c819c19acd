"""Pupko reconstruction at the bench shard size (125,000 simulated families, 60 taxa, gamma K=4) through the host API.
Run it under `ncu --metrics gpu__time_duration.sum -k regex:pupko` for the kernel time; the wall time printed here includes the
matrices, the expansion kernel and the copy of all states to the host.  CAFE_B200_PUPKO_THREADS=256 selects the older geometry."""
import os, sys, time, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cafe5_b200 import families as fam
from cafe5_b200.gamma import get_gamma
from cafe5_b200.model import Context
from cafe5_b200.synthetic import make_tree_newick, simulate_families
from cafe5_b200.tree import FlatTree

F = int(os.environ.get("PUPKO_FAMILIES", "125000"))
cp, mu = get_gamma(4, 0.65)
t60 = FlatTree(make_tree_newick(60, seed=20261017))
boot = Context(t60, np.ones((1, t60.n_leaves), dtype=np.int32), 170, 150)
c60 = simulate_families(t60, F, 0.002, mu, boot.get_matrix, seed=20261017)
boot.close()
ctx = Context(t60, c60, 170, 150)
ctx.set_prior(fam.uniform_prior(150))
ctx.reconstruct([0.002], mu, cp)
t = time.time()
out = ctx.reconstruct([0.002], mu, cp)
dt = time.time() - t
print("pupko threads=%s families=%d wall %.3f s, crc32 states %08x cat_states %08x" % (
    os.environ.get("CAFE_B200_PUPKO_THREADS", "512"), F, dt, zlib.crc32(np.ascontiguousarray(out["states"]).tobytes()),
    zlib.crc32(np.ascontiguousarray(out["cat_states"]).tobytes())), flush=True)
ctx.close()
