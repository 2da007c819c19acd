"""Clock stamps of the pruning kernel's chunk loop (timing build, CAFE_B200_RESIDENT_PROBE=1) on the bench workload: for the two
CTAs that share SM 0 and each of their four warps, per chunk: before / after the full-barrier wait, after the last DMMA, after the
ring refill.  Writes gpurun_out/probe_chunks.npz; numbers from this build are diagnostics, not bench values."""
import ctypes as C, os, sys
from types import SimpleNamespace
import numpy as np
os.environ["CAFE_B200_RESIDENT_PROBE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cafe5_b200 import families as fam
from cafe5_b200.model import Context

args = SimpleNamespace(taxa=60, cats=4, families=int(os.environ.get("PROBE_FAMILIES", "125000")))
tree, counts, mfs, mrs = bench.make_workload(args, 0, 0)
ctx = Context(tree, counts, mfs, mrs, device=0)
ctx.set_prior(fam.uniform_prior(mrs))
for i in range(3):
    lam, alpha, cp, mu = bench.step_params(i, 4)
    ctx.enqueue_eval([lam], alpha, mu, cp)
    ctx.fetch_result()
    print("step", i, ctx.last_stats(), flush=True)
PC = 4096
n = 2 * 4 * (PC * 4 + 4)
buf = np.zeros(n, dtype=np.int64)
rc = ctx.lib.cafe_b200_debug_read_probe(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_int64)), n)
assert rc == 0
buf = buf.reshape(2, 4, PC * 4 + 4)
smid = buf[:, :, 0]
t = buf[:, :, 4:].reshape(2, 4, PC, 4)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "probe_chunks.npz"), smid=smid, t=t)
print("smid", smid.tolist())
for c in range(2):
    for w in range(4):
        x = t[c, w]
        ok = x[:, 0] > 0
        x = x[ok]
        wait, stream, tail = x[:, 1] - x[:, 0], x[:, 2] - x[:, 1], x[:, 3] - x[:, 2]
        gap = x[1:, 0] - x[:-1, 3]
        period = x[1:, 0] - x[:-1, 0]
        q = lambda v: "med %d p10 %d p90 %d mean %.0f" % (np.median(v), np.percentile(v, 10), np.percentile(v, 90), v.mean())
        print("cta %d warp %d: n=%d | wait %s | stream %s | tail %s | gap %s | period %s" % (c, w, len(x), q(wait), q(stream), q(tail), q(gap), q(period)))
ctx.close()
