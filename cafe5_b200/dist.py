"""Multi-GPU plumbing: one process per GPU, gene families sharded contiguously across ranks.

Families are independent given the matrices (reference: `#pragma omp parallel for` over families,
src/base_model.cpp:69, src/gamma_core.cpp:190); the only cross-family operations are the final sum
(src/base_model.cpp:95, src/gamma_core.cpp:233) and the any-failure test (src/gamma_core.cpp:216).  Each
rank regenerates every matrix locally (<= 120 MB, cheaper than moving them), so a step exchanges exactly
16 bytes per rank: {partial sum of log-likelihoods, failed-family count}.  The partial sums are combined
in FIXED rank order on every rank (all_gather + sequential add) instead of an all_reduce tree, so the score
is bit-reproducible run to run for a given number of ranks.
"""
import math

import numpy as np


def shard_bounds(n_families, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n_families), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_partials(partials):
    """partials: iterable of (neg_lnl_partial, n_failed) in rank order -> (neg_lnl, n_failed).
    Any failed family or infinite partial rejects the whole evaluation (src/gamma_core.cpp:216-225)."""
    total, failed = 0.0, 0
    for neg, nf in partials:
        failed += int(nf)
        total += float(neg)
    if failed > 0 or math.isinf(total) or math.isnan(total):
        return math.inf, failed
    return total, failed


def allreduce_score(neg_lnl, n_failed, group=None):
    """Exchange this rank's partial with every rank (torch.distributed; NCCL on GPUs, gloo on CPU tests)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return combine_partials([(neg_lnl, n_failed)])
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    # +inf does not survive a float sum meaningfully; ship it as a failure flag
    bad = 1.0 if (math.isinf(neg_lnl) or math.isnan(neg_lnl)) else 0.0
    mine = torch.tensor([0.0 if bad else neg_lnl, float(n_failed), bad], dtype=torch.float64, device=dev)
    world = dist.get_world_size(group)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    rows = np.stack([t.cpu().numpy() for t in out])
    if rows[:, 2].any():
        return math.inf, int(rows[:, 1].sum())
    return combine_partials([(r[0], r[1]) for r in rows])


def fit_sharded(local_score, start, max_iterations=300, group=None):
    """Parameter search over a family table sharded across ranks: every rank runs the library's simplex search
    (cafe_b200_minimize, cafe5_b200/host/nelder_mead.hpp) over the SAME objective -- its own shard's partial
    (-lnL, n_failed) from `local_score(values)` combined with every other rank's by `allreduce_score`.  The combined score
    is bit-identical on every rank (fixed-order sum), so all ranks take the same simplex decisions and stay in step with
    no other communication: one 24-byte exchange per likelihood evaluation.

    local_score(values) -> (neg_lnl_partial, n_failed).  Returns dict(values, neg_lnl, iterations, evaluations, status)."""
    from .model import minimize
    n_eval = [0]

    failure = []

    def objective(x):
        n_eval[0] += 1
        try:
            neg, nf = local_score(list(x)) if not failure else (math.inf, 0)
        except Exception as e:      # keep answering the collective: a rank that stops calling all_gather hangs every other rank
            failure.append(e)
            neg, nf = math.inf, 0
        total, _ = allreduce_score(neg, nf, group)
        return total

    x, f, it = minimize(objective, list(start), max_iterations)
    if failure:
        raise failure[0]
    # status as cafe_b200_fit reports it: 0 converged (tolx and tolf 1e-6), 2 stopped at the iteration cap
    return dict(values=x, neg_lnl=f, iterations=it, evaluations=n_eval[0], status=2 if it >= max_iterations else 0)
