"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: contiguous family shards, one 16-byte exchange per
step, partials combined in fixed rank order.  The per-shard evaluator here is the C oracle (tests may use it);
on GPUs the same helpers wrap the CUDA context (bench.py)."""
import math
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from cafe5_b200 import dist as cdist
from cafe5_b200 import families as fam
from cafe5_b200.tree import FlatTree

NEWICK = "((A:1,B:3):7,(C:11,D:17):23)"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _counts():
    rng = np.random.default_rng(123)
    return rng.integers(1, 25, size=(41, 4)).astype(np.int32)


def _worker(rank, world, port, fail, out_queue):
    import torch.distributed as dist
    from oracle.pyoracle import OracleLib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = OracleLib()
    o.set_threads(1)
    tree = FlatTree(NEWICK)
    counts = _counts()
    lo, hi = cdist.shard_bounds(counts.shape[0], world, rank)
    prior = fam.uniform_prior(45)
    if fail:   # a saturated category kills every family of every shard -> +inf everywhere
        r = o.eval_gamma(tree, counts[lo:hi], 60, 45, prior, [0.04], [0.1, 2.0], [0.5, 0.5])
    else:
        r = o.eval_gamma(tree, counts[lo:hi], 60, 45, prior, [0.01], [0.5, 1.5], [0.5, 0.5])
    total, nfail = cdist.allreduce_score(r["neg_lnl"], r["n_failed"])
    out_queue.put((rank, total, nfail, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail", [False, True])
def test_two_rank_sharded_score(fail):
    from oracle.pyoracle import OracleLib
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fail, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][3:] == (0, 21) and results[1][3:] == (21, 41)
    assert results[0][1:3] == results[1][1:3]            # every rank ends with the identical score
    o = OracleLib()
    tree = FlatTree(NEWICK)
    prior = fam.uniform_prior(45)
    if fail:
        assert results[0][1] == math.inf and results[0][2] > 0
    else:
        whole = o.eval_gamma(tree, _counts(), 60, 45, prior, [0.01], [0.5, 1.5], [0.5, 0.5])
        assert abs(results[0][1] - whole["neg_lnl"]) <= 1e-13 * whole["neg_lnl"] and results[0][2] == 0


def _fit_worker(rank, world, port, out_queue):
    import torch.distributed as dist
    from oracle.pyoracle import OracleLib
    from cafe5_b200.gamma import get_gamma
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = OracleLib()
    o.set_threads(1)
    tree = FlatTree(NEWICK)
    counts = _counts()
    lo, hi = cdist.shard_bounds(counts.shape[0], world, rank)
    prior = fam.uniform_prior(45)

    def local_score(v):   # the oracle stands in for the GPU context of this rank (no GPU in the CPU suite)
        if v[0] <= 0 or v[1] <= 0:
            return math.inf, 0
        cp, mu = get_gamma(2, v[1])
        r = o.eval_gamma(tree, counts[lo:hi], 60, 45, prior, [v[0]], mu, cp, alpha=v[1])
        return r["neg_lnl"], r["n_failed"]

    r = cdist.fit_sharded(local_score, [0.01, 1.0], max_iterations=25)
    out_queue.put((rank, tuple(r["values"]), r["neg_lnl"], r["iterations"], r["evaluations"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_fit_matches_single_rank():
    """Sharded optimisation: both ranks walk the same simplex; the result equals the unsharded search to rounding of the
    two-term partial sum."""
    from oracle.pyoracle import OracleLib
    from cafe5_b200.gamma import get_gamma
    from cafe5_b200.model import minimize
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1:] == results[1][1:]              # identical trajectory and result on every rank
    o = OracleLib()
    tree = FlatTree(NEWICK)
    prior = fam.uniform_prior(45)
    counts = _counts()

    def whole(v):
        if v[0] <= 0 or v[1] <= 0:
            return math.inf
        cp, mu = get_gamma(2, v[1])
        return o.eval_gamma(tree, counts, 60, 45, prior, [v[0]], mu, cp, alpha=v[1])["neg_lnl"]

    x, f, it = minimize(whole, [0.01, 1.0], 25)
    assert it == results[0][3]
    assert np.allclose(x, results[0][1], rtol=1e-9) and abs(f - results[0][2]) <= 1e-11 * f


def _planned_worker(rank, world, port, out_queue):
    """Shards from cafe_b200_plan_shards (families ordered by total count, unequal block sizes) instead of contiguous blocks."""
    import torch.distributed as dist
    from cafe5_b200.model import plan_shards
    from oracle.pyoracle import OracleLib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CAFE_B200_TABLES"] = "force"          # the planner balances table-plan cost; forced so that 41 families plan any
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = OracleLib()
    o.set_threads(1)
    tree = FlatTree(NEWICK)
    counts = _counts()
    order, bounds = plan_shards(tree, counts, world)
    members = order[bounds[rank]:bounds[rank + 1]]
    prior = fam.uniform_prior(45)
    r = o.eval_gamma(tree, counts[members], 60, 45, prior, [0.01], [0.5, 1.5], [0.5, 0.5])
    total, nfail = cdist.allreduce_score(r["neg_lnl"], r["n_failed"])
    out_queue.put((rank, total, nfail, order.tolist(), bounds.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_planned_shards_score():
    from oracle.pyoracle import OracleLib
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_planned_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][3:] == results[1][3:]                              # every rank computed the same plan
    order, bounds = results[0][3], results[0][4]
    assert sorted(order) == list(range(41)) and bounds[0] == 0 and bounds[-1] == 41 and 0 < bounds[1] < 41
    assert results[0][1:3] == results[1][1:3]                            # and ends with the identical score
    o = OracleLib()
    want = o.eval_gamma(FlatTree(NEWICK), _counts(), 60, 45, fam.uniform_prior(45), [0.01], [0.5, 1.5], [0.5, 0.5])
    assert results[0][2] == 0 and abs(results[0][1] - want["neg_lnl"]) <= 1e-12 * want["neg_lnl"]
