#!/usr/bin/env python
"""bench.py -- family-likelihood evaluations per second of the CAFE5 likelihood hot path on B200.

Workload (BASELINE.json configs[4], the configuration the metric and the north-star target are quoted on):
gamma model, K = 4 categories, families simulated with the reference simulator's semantics on a seeded
60-taxon ultrametric tree (S = 171 states, R = 150 root sizes, 400 distinct matrix keys per step), sharded
contiguously across GPUs: 125,000 families per GPU (1,000,000 at 8 GPUs, weak scaling).  One "step" is one
optimiser evaluation = one call of gamma_model::infer_family_likelihoods for a new (lambda, alpha): all
transition matrices are regenerated, every family is pruned under every category, the mixture and the score
are reduced.  Nothing is cached between steps (lambda and alpha change every step, as under Nelder-Mead).

  value  : whole-job evals/s with the count table resident in HBM, timed with CUDA events on the library's stream
  e2e    : the same through the public host API (cafe_b200_set_prior / set_error_model / eval_gamma) with host
           buffers: per-step parameter upload and ALL per-family outputs copied back to host memory
  --impl reference : the unmodified reference (oracle/_ref) on the host cores, bounded samples scaled linearly

Launch: `python bench.py --gpus N --steps K --warmup W` (N > 1: under torchrun, one rank per GPU).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
LAMBDA0, ALPHA0 = 0.002, 0.65


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--families", type=int, default=125000, help="families per GPU")
    ap.add_argument("--taxa", type=int, default=60)
    ap.add_argument("--cats", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true", help="skip the whole-optimisation wall-time leg")
    return ap.parse_args()


def step_params(i, n_cat):
    """(lambda, alpha, cat_probs, multipliers) of step i: a small deterministic walk, as an optimiser would make."""
    from cafe5_b200.gamma import get_gamma
    lam = LAMBDA0 * (1.0 + 0.002 * ((i * 7) % 11 - 5))
    alpha = ALPHA0 * (1.0 + 0.003 * ((i * 5) % 7 - 3))
    cp, mu = get_gamma(n_cat, alpha)
    return lam, alpha, cp, mu


def alg_flops_per_prune(tree, S, R, nnz=1):
    """ALGORITHMIC flops of one (family, category) prune: dense S-column contraction for internal
    children, nnz-column gather for leaf children, the child products, and the 2R root weighting.  Padding and
    the rows the kernel computes beyond S / R are NOT counted."""
    root = tree.n_nodes - 1
    n_children = np.zeros(tree.n_nodes, dtype=np.int64)
    total = 0
    for i in range(tree.n_nodes - 1):
        p = int(tree.parent[i])
        n_children[p] += 1
        rows = R if p == root else S
        total += 2 * nnz * rows if tree.leaf_col[i] >= 0 else 2 * rows * S
    for i in range(tree.n_nodes):
        if tree.leaf_col[i] < 0:
            total += (R if i == root else S) * (n_children[i] - 1)
    return int(total + 2 * R)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.lines = []
        self.proc = None
        self.device = device

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(args, rank, device):
    """Tree + this rank's shard of simulated families (matrices for the simulation come from the CUDA library)."""
    from cafe5_b200 import families as fam
    from cafe5_b200.gamma import get_gamma
    from cafe5_b200.model import Context
    from cafe5_b200.synthetic import make_tree_newick, simulate_families
    from cafe5_b200.tree import FlatTree

    tree = FlatTree(make_tree_newick(args.taxa, seed=SEED))
    cp, mu = get_gamma(args.cats, ALPHA0)
    boot = Context(tree, np.ones((1, tree.n_leaves), dtype=np.int32), 170, 150, device=device)
    counts = simulate_families(tree, args.families, LAMBDA0, mu, boot.get_matrix, seed=SEED + 1000 * rank)
    boot.close()
    mfs, mrs = fam.derive_sizes(counts)
    # The reference rejects a whole evaluation (+inf) when ONE family's root vector underflows to zero in some category
    # (gamma_core.cpp:151,216-225); among 125,000 simulated families on 118 branches a handful do, near the generating parameters.
    # A user has to remove such families before the reference returns a finite score; the generator does the same: families that
    # fail anywhere in the parameter box the bench steps walk through are replaced by copies of healthy ones (count unchanged).
    ctx = Context(tree, counts, mfs, mrs, device=device)
    ctx.set_prior(fam.uniform_prior(mrs))
    bad = np.zeros(counts.shape[0], dtype=bool)
    for fl, fa in ((1.0, 1.0), (0.988, 0.99), (0.988, 1.01), (1.012, 0.99), (1.012, 1.01), (1.5, 1.0 / ALPHA0)):
        cp_i, mu_i = get_gamma(args.cats, ALPHA0 * fa)
        bad |= ctx.eval_gamma([LAMBDA0 * fl], ALPHA0 * fa, mu_i, cp_i)["failed"].astype(bool)
    ctx.close()
    if bad.any():
        good = np.flatnonzero(~bad)
        counts[np.flatnonzero(bad)] = counts[good[:int(bad.sum())]]
    return tree, counts, mfs, mrs


def reference_rate(tree_newick, species, counts, mfs, mrs, prior, n_cat, shard_families, steps, warmup, log=None):
    """Evals/s of the UNMODIFIED reference on this host, all threads, from bounded samples scaled linearly:
    T_step(F) = T_matrices(sampled keys) * keys_total / keys_sampled + T_prune(sampled families) * F / n_sample."""
    from oracle.pyoracle import RefLib
    ref = RefLib()
    cores = ref.max_threads()
    n_sample = min(counts.shape[0], 24 * cores)
    ctx = ref.ctx(tree_newick, species, counts[:n_sample], mfs, mrs, prior)
    lam, alpha, cp, mu = step_params(0, n_cat)
    session = ref.session(ctx, [lam], mu, cp)      # full matrix cache for the prune leg (outside the timed legs)
    per_step = []
    for i in range(warmup + steps):
        lam_i, _, _, mu_i = step_params(i, n_cat)
        t_mat, keys_done, keys_total = ref.time_precalculate(ctx, [lam_i], mu_i, stride=8)
        t_pr, _ = ref.session_prune(ctx, session, n_sample)
        t_full = t_mat * keys_total / keys_done + t_pr * shard_families / n_sample
        if log:
            log("reference step %d: matrices %.2fs for %d/%d keys, prune %.2fs for %d families -> %.1fs per full step"
                % (i, t_mat, keys_done, keys_total, t_pr, n_sample, t_full))
        if i >= warmup:
            per_step.append(t_full)
    ref.session_destroy(session)
    ctx.close()
    t = float(np.mean(per_step))
    sample = ("per step: matrix_cache::precalculate_matrices on every 8th branch length (%d of %d keys) + gamma_model::prune "
              "over %d of %d families (omp parallel for, %d threads); times scaled linearly to the full step"
              % (keys_done, keys_total, n_sample, shard_families, cores))
    return shard_families / t, t, cores, sample


def port_rate(tree, counts, mfs, mrs, prior, n_cat, shard_families):
    """Fallback when oracle/_ref is absent: the C oracle (a port), same two-leg scaling."""
    from oracle.pyoracle import OracleLib
    o = OracleLib()
    cores = o.max_threads()
    lam, alpha, cp, mu = step_params(0, n_cat)
    n1, n2 = 2 * cores, 10 * cores
    t0 = time.time(); o.eval_gamma(tree, counts[:n1], mfs, mrs, prior, [lam], mu, cp, alpha=alpha); t1 = time.time() - t0
    t0 = time.time(); o.eval_gamma(tree, counts[:n2], mfs, mrs, prior, [lam], mu, cp, alpha=alpha); t2 = time.time() - t0
    per_family = max((t2 - t1) / (n2 - n1), 1e-9)
    fixed = max(t1 - n1 * per_family, 0.0)
    t = fixed + per_family * shard_families
    return shard_families / t, t, cores, "C oracle port: evals of %d and %d families; fixed + per-family cost scaled to the shard" % (n1, n2)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cafe5_b200 import families as fam
    from cafe5_b200.synthetic import make_tree_newick, simulate_families
    from cafe5_b200.tree import FlatTree
    from cafe5_b200.gamma import get_gamma
    from oracle import pyoracle

    tree = FlatTree(make_tree_newick(args.taxa, seed=SEED))
    cp, mu = get_gamma(args.cats, ALPHA0)
    have_ref = pyoracle.have_ref()
    # the sample's families: same generator; matrices for the simulation from the reference itself (or the oracle)
    if have_ref:
        ref = pyoracle.RefLib()
        provider = lambda l, t: ref.matrix(171, l, t)   # noqa: E731
    else:
        o = pyoracle.OracleLib()
        provider = lambda l, t: o.matrix(171, l, t)     # noqa: E731
    n_need = 24 * (os.cpu_count() or 8)
    counts = simulate_families(tree, n_need, LAMBDA0, mu, provider, seed=SEED)
    mfs, mrs = 170, 150
    prior = fam.uniform_prior(mrs)
    log = lambda m: print(m, file=sys.stderr, flush=True)   # noqa: E731
    if have_ref:
        value, t_step, cores, sample = reference_rate(tree_newick(tree), tree.species, counts, mfs, mrs, prior, args.cats,
                                                      args.families, args.steps, args.warmup, log)
        kind = "reference"
    else:
        value, t_step, cores, sample = port_rate(tree, counts, mfs, mrs, prior, args.cats, args.families)
        kind = "port"
    # one host runs the reference: the job's family count scales with --gpus (weak scaling), the host's rate does not
    line = {
        "impl": "reference", "metric": "family-likelihood evals/s", "value": value, "unit": "family-likelihood evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3 * args.gpus,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, tree),
        "cpu_baseline": {"value": value, "unit": "family-likelihood evals/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "family-likelihood evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def tree_newick(tree):
    return tree.newick


def workload_config(args, tree):
    return {"workload": "BASELINE configs[4] shard: gamma model K=%d, %d simulated families per GPU on a seeded %d-taxon "
                        "ultrametric tree (%d nodes), S=171 states, R=150 root sizes; one step = one (lambda, alpha) evaluation "
                        "(matrices + pruning + mixture + score)" % (args.cats, args.families, args.taxa, tree.n_nodes),
            "families_per_gpu": args.families, "categories": args.cats, "taxa": args.taxa,
            "l2": "flushed between timed steps (256 MiB write on the same stream; count table 30 MB < 126 MB L2)",
            "parallelism": "families sharded contiguously, one rank per GPU, one 24-byte all_gather per step"}


def run_ours(args):
    import torch
    from cafe5_b200 import dist as cdist, families as fam
    from cafe5_b200.model import Context, measure_fp64_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience relaunch: one rank per GPU under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    tree, counts, mfs, mrs = make_workload(args, rank, local)
    prior = fam.uniform_prior(mrs)
    ctx = Context(tree, counts, mfs, mrs, device=local)
    ctx.set_prior(prior)
    U = ctx.unique_families()
    S, R = mfs + 1, mrs
    flops_per_launch = alg_flops_per_prune(tree, S, R) * U * args.cats
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    peak_dfma = measure_fp64_peak(local, use_dmma=False)
    peak_dmma = measure_fp64_peak(local, use_dmma=True)

    # ---------------- device-resident throughput ("value") ----------------
    def one_step(i, timed):
        lam, alpha, cp, mu = step_params(i, args.cats)
        with torch.cuda.stream(stream):
            flush_buf.zero_()                       # evict L2 between steps (outside the event pair)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.enqueue_eval([lam], alpha, mu, cp)
        e1.record(stream)
        return e0, e1

    for i in range(args.warmup):
        one_step(i, False)
    ctx.fetch_result()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pairs, prune_ms, mat_ms = [], [], []
    t_wall0 = time.time()
    for i in range(args.steps):
        pairs.append(one_step(args.warmup + i, True))
        st = ctx.last_stats()                       # syncs the stream; per-kernel CUDA-event times of this step
        prune_ms.append(st["ms_prune"])
        mat_ms.append(st["ms_matrices"])
    neg, n_failed = ctx.fetch_result()
    total, failed_all = cdist.allreduce_score(neg, n_failed)
    barrier()
    t_wall = time.time() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in pairs]
    ms_per_step = float(np.mean(ms_steps))
    n_mats = ctx.last_stats()["matrices"]

    # ---------------- end to end through the host API ("e2e") ----------------
    K = args.cats
    h2d = len(prior) * 4 + n_mats * 24 + K * tree.n_nodes * 4 + K * 8
    d2h = counts.shape[0] * (K * 8 + 8 + K * 8 + K + 1) + 16
    for i in range(2):
        lam, alpha, cp, mu = step_params(100 + i, K)
        ctx.eval_gamma([lam], alpha, mu, cp, pinned=True)
    barrier()
    e2e_ms = []
    for i in range(args.steps):
        lam, alpha, cp, mu = step_params(200 + i, K)
        t0 = time.perf_counter()
        ctx.set_prior(prior)                        # host buffers in, every step
        ctx.set_error_model(None)
        out = ctx.eval_gamma([lam], alpha, mu, cp, pinned=True)  # all per-family outputs back to (page-locked) host memory
        tot_i, _ = cdist.allreduce_score(out["neg_lnl"], out["n_failed"])
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e_ms_per_step = float(np.mean(e2e_ms))

    # max over ranks
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_per_step, e2e_ms_per_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_ms_per_step = float(t[0]), float(t[1])
    families_total = args.families * world
    value = families_total / (ms_per_step * 1e-3)
    e2e_value = families_total / (e2e_ms_per_step * 1e-3)

    # ---------------- wall time of one whole (lambda, alpha) optimisation (the metric's second half) ----------------
    # the library's own host driver (cafe_b200_fit: seeded start, Nelder-Mead with the reference's constants) on this shard
    fit = None
    if not args.no_fit:
        # explicit start: with 125,000 families on 118 branches the reference's random start (normal(0.002 L, 0.2) / L) is usually
        # rejected by the all-or-nothing rule (one underflowed family => +inf), for the reference exactly as for us
        start = [1.5 * LAMBDA0, 1.0]
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            r = ctx.fit(n_cat=K, start=start)
            how = "cafe_b200_fit (C++ host driver)"
        else:
            from cafe5_b200.model import discrete_gamma

            def local_score(v):   # this rank's shard; the ranks exchange 24 bytes per evaluation inside fit_sharded
                if not (v[1] > 0):
                    return math.inf, 0
                cp_i, mu_i = discrete_gamma(K, v[1])
                o = ctx.eval_gamma([v[0]], v[1], mu_i, cp_i, want_family=False)
                return o["neg_lnl"], o["n_failed"]

            r = cdist.fit_sharded(local_score, start)
            how = "cafe5_b200.dist.fit_sharded (the same simplex search on every rank over the all-gathered score)"
        barrier()
        fit = {"wall_s": time.perf_counter() - t0, "iterations": r["iterations"], "evaluations": r["evaluations"], "status": r["status"],
               "lambda": float(r["values"][0]), "alpha": float(r["values"][1]), "neg_lnl": r["neg_lnl"],
               "what": "%s: gamma K=%d, (lambda, alpha) estimated by Nelder-Mead (tolx = tolf = 1e-6, reference constants) over %d "
                       "families from the start point (1.5 x true lambda, alpha = 1)" % (how, K, counts.shape[0] * world)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            if pyoracle.have_ref():
                v, t_step, cores, sample = reference_rate(tree_newick(tree), tree.species, counts, mfs, mrs, prior, K, args.families, 1, 0)
                cpu = {"value": v, "unit": "family-likelihood evals/s", "cores": cores, "kind": "reference", "sample": sample}
            else:
                v, t_step, cores, sample = port_rate(tree, counts, mfs, mrs, prior, K, args.families)
                cpu = {"value": v, "unit": "family-likelihood evals/s", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "family-likelihood evals/s", "cores": None, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        achieved = flops_per_launch / (float(np.mean(prune_ms)) * 1e-3) / 1e12
        peak = max(peak_dfma, peak_dmma)
        line = {
            "metric": "family-likelihood evals/s", "value": value, "unit": "family-likelihood evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, tree),
            "e2e": {"value": e2e_value, "unit": "family-likelihood evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_per_step},
            "gpu_launches": int(ctx.last_stats()["launches"]) * args.steps,
            "roofline": {"bound": "tensor", "kernel": prune_kernel_name(), "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                         "peak_source": "FP64 tensor pipe (DMMA; tcgen05 has no FP64 kind) measured live on this GPU by "
                                        "cafe_b200_measure_fp64_peak: DMMA m8n8k4 %.2f, DFMA %.2f TFLOP/s; MEASURED_PEAKS.json "
                                        "carries only HBM and bf16 figures" % (peak_dmma, peak_dfma),
                         "alg_flops_per_launch": flops_per_launch, "ms_per_launch": float(np.mean(prune_ms)),
                         "share_of_step": float(np.mean(prune_ms)) / ms_per_step,
                         "matrix_gen": {"ms_per_launch": float(np.mean(mat_ms)), "matrices": n_mats,
                                        "write_GBps": n_mats * 8.0 * (max(mfs, mrs) + 1) ** 2 / (float(np.mean(mat_ms)) * 1e-3) / 1e9,
                                        "terms_per_s": n_mats * matrix_terms(max(mfs, mrs) + 1) / (float(np.mean(mat_ms)) * 1e-3)}},
            "cpu_baseline": cpu,
            "optimisation": fit,
            "clocks": clocks,
            "result": {"neg_lnl": total, "n_failed": failed_all, "unique_families_rank0": int(U), "wall_s_timed_region": t_wall},
        }
        print(json.dumps(_finite(line)), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def prune_kernel_name():
    return {"dfma": "prune_kernel (DFMA register tiles)", "stream": "prune_dmma_kernel (DMMA, both operands streamed)"}.get(
        os.environ.get("CAFE_B200_PRUNE", ""), "prune_resident_kernel (FP64 tensor cores: mma.sync.m8n8k4.f64 / DMMA.8x8x4)")


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the pruning kernel in this exact configuration, from the
    committed `ncu --set full` capture (profiles/r01_prune_resident2_ncu.txt); None when another variant / size is benchmarked."""
    if os.environ.get("CAFE_B200_PRUNE") or os.environ.get("CAFE_B200_RESIDENT_WN") or sys.argv[1:] and any(
            a.startswith(("--families", "--taxa", "--cats")) for a in sys.argv[1:]):
        return None
    path = os.path.join(ROOT, "profiles", "r01_prune_resident2_ncu.txt")
    try:
        total = 0.0
        for line in open(path):
            f = [x.strip() for x in line.split("|")]
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
        return total or None
    except OSError:
        return None


def _finite(x):
    """JSON has no Infinity / NaN: a rejected evaluation's +inf score is reported as null."""
    if isinstance(x, dict):
        return {k: _finite(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_finite(v) for v in x]
    if isinstance(x, float) and not math.isfinite(x):
        return None
    return x


def matrix_terms(N):
    """Binomial-sum terms of one N x N matrix: sum_{s=1}^{N-1} sum_{c=0}^{N-1} (min(s,c)+1)."""
    s = np.arange(1, N)[:, None]
    c = np.arange(0, N)[None, :]
    return int((np.minimum(s, c) + 1).sum())


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
