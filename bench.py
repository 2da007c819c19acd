#!/usr/bin/env python
"""bench.py -- family-likelihood evaluations per second of the CAFE5 likelihood hot path on B200.

Workload (BASELINE.json configs[4], the configuration the metric and the north-star target are quoted on):
gamma model, K = 4 categories, 1,000,000 families simulated by the library's own simulator (cafe_b200_simulate,
the reference simulator's semantics) on a seeded 60-taxon ultrametric tree (S = 171 states, R = 150 root sizes,
400 distinct matrix keys per step).  `--gpus 1` holds the whole job on one GPU; N GPUs shard the SAME job
contiguously (strong scaling).  One "step" is one optimiser evaluation = one call of
gamma_model::infer_family_likelihoods for a new (lambda, alpha): all transition matrices are regenerated, every
distinct family is pruned under every category, the mixture and the score are reduced and (N > 1) the per-rank
partial scores are exchanged.  Nothing is cached between steps (lambda and alpha change every step, as under
Nelder-Mead).

  value  : whole-job evals/s with the count table resident in HBM, CUDA events on the library's stream around
           [matrices, pruning, mixture, score, and for N > 1 the NCCL all_gather of the partial scores on that stream]
  e2e    : the same through the reference-facing C ABI with HOST buffers (cafe_b200_set_prior / set_error_model /
           eval_gamma: per-step parameter upload, ALL per-family outputs copied back).  For N > 1 it goes through
           cafe_b200_create_multi: ONE process (rank 0) drives all N GPUs, as a CAFE5 process linked against the
           drop-in models would; `e2e_ranks` is the one-process-per-GPU variant with the host-side exchange.
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
WEAK_FAMILIES = 125000      # the per-GPU shard of the weak series (round 1's bench workload)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--families", type=int, default=1000000, help="families of the whole job (sharded over the GPUs)")
    ap.add_argument("--taxa", type=int, default=60)
    ap.add_argument("--cats", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true", help="skip the whole-optimisation wall-time legs")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling side measurement")
    return ap.parse_args()


def step_params(i, n_cat):
    """(lambda, alpha, cat_probs, multipliers) of step i: a small deterministic walk, as an optimiser would make."""
    from cafe5_b200.gamma import get_gamma
    lam = LAMBDA0 * (1.0 + 0.002 * ((i * 7) % 11 - 5))
    alpha = ALPHA0 * (1.0 + 0.003 * ((i * 5) % 7 - 3))
    cp, mu = get_gamma(n_cat, alpha)
    return lam, alpha, cp, mu


def alg_flops_per_prune(tree, S, R, nnz=1):
    """ALGORITHMIC flops of one (family, category) prune when every family goes through every node: dense
    S-column contraction for internal children, nnz-column gather for leaf children, the child products, the 2R root weighting.
    Padding and the rows the kernel computes beyond S / R are NOT counted."""
    return alg_flops(tree, S, R, None, 1, nnz)


def alg_flops(tree, S, R, columns, U, nnz=1):
    """ALGORITHMIC flops of one category of one evaluation, counting every node's work only for the columns it is actually
    computed for: columns[v] (cafe_b200_node_columns) is U for a node pruned per distinct family and the number of distinct
    leaf-count patterns below v for a node served by a factor table (subtree-pattern reuse) -- work that is skipped is not
    credited.  columns=None: U columns everywhere."""
    root = tree.n_nodes - 1
    col = (lambda v: U) if columns is None else (lambda v: int(columns[v]))
    n_children = np.zeros(tree.n_nodes, dtype=np.int64)
    total = 0
    for i in range(tree.n_nodes - 1):
        p = int(tree.parent[i])
        n_children[p] += 1
        rows = R if p == root else S
        if tree.leaf_col[i] >= 0:
            total += 2 * nnz * rows * col(p)          # gathered once per column of the parent
        else:
            total += 2 * rows * S * col(i)            # the factor P_i . V_i, once per column of i
    for i in range(tree.n_nodes):
        if tree.leaf_col[i] < 0:
            total += (R if i == root else S) * (n_children[i] - 1) * col(i)
    return int(total + 2 * R * U)


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


def bench_tree(args):
    from cafe5_b200.synthetic import make_tree_newick
    from cafe5_b200.tree import FlatTree
    return FlatTree(make_tree_newick(args.taxa, seed=SEED))


def make_workload(args, device):
    """The whole job's families, identical on every rank: simulated on the device by cafe_b200_simulate (row f4; a counter-based
    stream per family index, so the table does not depend on the number of ranks), root sizes uniform on 1..124, one gamma
    category per family.  Returns (tree, counts[F, n_species], max_family_size, max_root_family_size, info)."""
    from cafe5_b200 import families as fam
    from cafe5_b200.gamma import get_gamma
    from cafe5_b200.model import Context

    tree = bench_tree(args)
    cp, mu = get_gamma(args.cats, ALPHA0)
    boot = Context(tree, np.ones((1, tree.n_leaves), dtype=np.int32), 170, 150, device=device)
    rng = np.random.default_rng(SEED)
    n_draw = int(args.families * 1.02) + 64
    roots = rng.integers(1, 125, size=n_draw).astype(np.int32)
    t0 = time.perf_counter()
    sim = boot.simulate([LAMBDA0], roots, mu, cp, max_sim=120, seed=SEED, max_redraws=50)
    t_sim = time.perf_counter() - t0
    boot.close()
    counts = sim["counts"]
    keep = fam.exists_at_root(tree, counts)
    counts = counts[keep][:args.families]
    assert counts.shape[0] == args.families, "simulator returned too few families that exist at the root"
    mfs, mrs = fam.derive_sizes(counts)
    return tree, counts, mfs, mrs, {"simulate_s": t_sim, "simulated": int(n_draw), "not_at_root": int(sim["n_not_at_root"])}


def drop_failing_families(args, tree, counts, mfs, mrs, members, device):
    """The reference rejects a whole evaluation (+inf) when ONE family's root vector underflows to zero in some category
    (gamma_core.cpp:151,216-225); among a million simulated families on 118 branches a few dozen do, near the generating
    parameters.  A user has to remove such families before the reference returns a finite score; the bench does the same: a
    family of this rank's shard (`members`: its family indices) that fails anywhere in the parameter box the steps walk through is
    replaced by a copy of a healthy one (so the family count is unchanged; duplicates are not counted in the throughput numerator)."""
    from cafe5_b200 import families as fam
    from cafe5_b200.gamma import get_gamma
    from cafe5_b200.model import Context
    shard = counts[members].copy()
    ctx = Context(tree, shard, mfs, mrs, device=device)
    ctx.set_prior(fam.uniform_prior(mrs))
    bad = np.zeros(shard.shape[0], dtype=bool)
    for fl, fa in ((1.0, 1.0), (0.988, 0.99), (0.988, 1.01), (1.012, 0.99), (1.012, 1.01), (1.5, 1.0 / ALPHA0)):
        cp_i, mu_i = get_gamma(args.cats, ALPHA0 * fa)
        bad |= ctx.eval_gamma([LAMBDA0 * fl], ALPHA0 * fa, mu_i, cp_i)["failed"].astype(bool)
    ctx.close()
    if bad.any():
        good = np.flatnonzero(~bad)
        shard[np.flatnonzero(bad)] = shard[good[:int(bad.sum())]]
    return shard, int(bad.sum())


def reference_rate(tree_newick, species, counts, mfs, mrs, prior, n_cat, job_families, steps, warmup, log=None):
    """Evals/s of the UNMODIFIED reference on this host, all threads, from bounded samples scaled linearly:
    T_step(F) = T_matrices(sampled keys) * keys_total / keys_sampled + T_prune(sampled families) * F / n_sample."""
    from oracle.pyoracle import RefLib
    ref = RefLib()
    ref.set_threads(os.cpu_count() or 1)           # torchrun exports OMP_NUM_THREADS=1: the reference gets every host core anyway
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
        t_full = t_mat * keys_total / keys_done + t_pr * job_families / n_sample
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
              % (keys_done, keys_total, n_sample, job_families, cores))
    return job_families / t, t, cores, sample


def port_rate(tree, counts, mfs, mrs, prior, n_cat, job_families):
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
    t = fixed + per_family * job_families
    return job_families / t, t, cores, "C oracle port: evals of %d and %d families; fixed + per-family cost scaled to the job" % (n1, n2)


def load_real_config(name):
    """Count table, tree and sizes of BASELINE configs 1-4 from the committed fixtures (tests/golden/*.npz hold the reference's
    example data after its root filter; the GPU box has no /root/reference)."""
    from cafe5_b200.tree import FlatTree
    g = np.load(os.path.join(ROOT, "tests", "golden", ("hymenoptera" if name == "config4" else "mammals") + ".npz"))
    species = [str(s) for s in g["species"]]
    lam_newick = str(g["lambda_newick"]) if name == "config3" else None
    tree = FlatTree(str(g["newick"]), lam_newick, species=species)
    return g, tree, species, g["counts"].astype(np.int32), int(g["max_family_size"]), int(g["max_root_family_size"])


REAL_CONFIGS = (
    ("config1", "mammals, base model, lambda", dict(n_cat=0)),
    ("config2", "mammals, gamma K=4, (lambda, alpha)", dict(n_cat=4)),
    ("config3", "mammals, error model from file + two lambda classes, (lambda1, lambda2)", dict(n_cat=0)),
    ("config4", "Hymenoptera, gamma K=8, (lambda, alpha)", dict(n_cat=8)),
)


def real_config_fits(device):
    """Wall time per optimisation of BASELINE configs 1-4 through the library's own host driver (cafe_b200_fit: seeded start like
    the reference's scorers, Nelder-Mead with the reference's constants), context creation reported separately."""
    from cafe5_b200 import families as fam
    from cafe5_b200.model import Context, error_model
    out = {}
    for name, what, kw in REAL_CONFIGS:
        g, tree, species, counts, mfs, mrs = load_real_config(name)
        t0 = time.perf_counter()
        ctx = Context(tree, counts, mfs, mrs, device=device)
        ctx.set_prior(fam.uniform_prior(mrs))
        if name == "config3":
            ctx.set_error_model(error_model(g["em_probs"], int(g["em_maxcnt"])))
        t_create = time.perf_counter() - t0
        r = ctx.fit(seed=10, **kw)
        ctx.close()
        out[name] = {"what": what, "families": int(counts.shape[0]), "wall_s": r["seconds"], "create_s": t_create,
                     "evaluations": r["evaluations"], "iterations": r["iterations"], "status": r["status"],
                     "values": [float(v) for v in r["values"]], "neg_lnl": r["neg_lnl"], "seed": 10}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cafe5_b200 import families as fam
    from cafe5_b200.synthetic import simulate_families
    from cafe5_b200.gamma import get_gamma
    from oracle import pyoracle

    tree = bench_tree(args)
    cp, mu = get_gamma(args.cats, ALPHA0)
    have_ref = pyoracle.have_ref()
    # the sample's families: the numpy restatement of the reference simulator (no GPU in this arm); matrices for the simulation from
    # the reference itself (or the oracle)
    if have_ref:
        ref = pyoracle.RefLib()
        ref.set_threads(os.cpu_count() or 1)
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
        value, t_step, cores, sample = reference_rate(tree.newick, tree.species, counts, mfs, mrs, prior, args.cats,
                                                      args.families, args.steps, args.warmup, log)
        kind = "reference"
    else:
        value, t_step, cores, sample = port_rate(tree, counts, mfs, mrs, prior, args.cats, args.families)
        kind = "port"
    opt = None
    if have_ref and not args.no_fit:
        # the metric's second half beside ours: the reference's own optimizer over its own CPU models, BASELINE config 1
        g, rtree, species, rcounts, rmfs, rmrs = load_real_config("config1")
        rctx = ref.ctx(str(g["newick"]), species, rcounts, rmfs, rmrs, fam.uniform_prior(rmrs))
        t0 = time.perf_counter()
        r = rctx.optimize("cpu", n_cat=0, seed=10)
        opt = {"config1": {"what": "mammals, base model, lambda: the reference's optimizer over its own CPU models", "wall_s": time.perf_counter() - t0,
                           "evaluations": r["attempts"], "iterations": r["iterations"], "values": [float(v) for v in r["values"]],
                           "neg_lnl": r["score"], "seed": 10, "threads": cores}}
        rctx.close()
    # one host runs the reference whatever --gpus says: the job is the same 1,000,000 families
    line = {
        "impl": "reference", "metric": "family-likelihood evals/s", "value": value, "unit": "family-likelihood evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, tree),
        "cpu_baseline": {"value": value, "unit": "family-likelihood evals/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "family-likelihood evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "optimisations": opt,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, tree):
    return {"workload": "BASELINE configs[4]: gamma model K=%d, %d families simulated on a seeded %d-taxon ultrametric tree "
                        "(%d nodes), S=171 states, R=150 root sizes; one step = one (lambda, alpha) evaluation of the whole job "
                        "(matrices + pruning + mixture + score)" % (args.cats, args.families, args.taxa, tree.n_nodes),
            "families": args.families, "categories": args.cats, "taxa": args.taxa,
            "l2": "flushed between timed steps (256 MiB write on the same stream, outside the event pair)",
            "parallelism": "the job's families sharded over the GPUs (strong scaling: the job does not grow with --gpus) in clustered, "
                           "work-balanced shards (cafe_b200_plan_shards); one 16-byte-per-rank all_gather per step, inside the timed region"}


class DeviceResult:
    """The two doubles a context's evaluation leaves on the device, as an object torch can wrap without a copy."""

    def __init__(self, ptr):
        self.__cuda_array_interface__ = {"shape": (2,), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}


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
    dist = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a short collective timeout: a rank-dependent code path shows up as an abort after two minutes, not as a ten-minute hang
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
        # host-side waits that must not put a spinning kernel on a GPU (rank 0 works alone for up to a minute at a time)
        cpu_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=900))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup0 = time.perf_counter()
    tree, all_counts, mfs, mrs, gen_info = make_workload(args, local)
    # strong scaling: the job is cut into clustered, work-balanced shards (cafe_b200_plan_shards: families ordered by total count,
    # cut points moved until every shard costs the same under the subtree-pattern table plan); every rank computes the same plan
    t0 = time.perf_counter()
    if world > 1:
        from cafe5_b200.model import plan_shards
        order, bounds = plan_shards(tree, all_counts, world)
    else:
        order, bounds = np.arange(args.families), np.array([0, args.families])
    gen_info["plan_shards_s"] = time.perf_counter() - t0
    gen_info["shard_sizes"] = np.diff(bounds).tolist()
    members = order[bounds[rank]:bounds[rank + 1]]
    counts, n_replaced = drop_failing_families(args, tree, all_counts, mfs, mrs, members, local)
    prior = fam.uniform_prior(mrs)
    t0 = time.perf_counter()
    ctx = Context(tree, counts, mfs, mrs, device=local)
    t_create = time.perf_counter() - t0
    ctx.set_prior(prior)
    U = ctx.unique_families()
    S, R, K = mfs + 1, mrs, args.cats
    columns = ctx.node_columns()
    flops_per_step = alg_flops(tree, S, R, columns, U) * K              # this rank, executed-unique work only
    flops_per_step_plain = alg_flops(tree, S, R, None, U) * K           # every distinct family through every node
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    peak_dfma = measure_fp64_peak(local, use_dmma=False)
    peak_dmma = measure_fp64_peak(local, use_dmma=True)
    # device-side exchange of the partial scores: the library's own result words, gathered on the library's stream
    if world > 1:
        mine = torch.as_tensor(DeviceResult(ctx.result_device()), device=torch.device("cuda", local))
        gathered = torch.zeros(2 * world, dtype=torch.float64, device="cuda")

    # ---------------- device-resident throughput ("value") ----------------
    def one_step(i):
        lam, alpha, cp, mu = step_params(i, K)
        with torch.cuda.stream(stream):
            flush_buf.zero_()                       # evict L2 between steps (outside the event pair)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.enqueue_eval([lam], alpha, mu, cp)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, mine)      # the path's one exchange, in stream order after the score kernel
        e1.record(stream)
        return e0, e1

    def combined_score():
        if world == 1:
            return ctx.fetch_result()
        torch.cuda.synchronize()
        rows = gathered.cpu().numpy().reshape(world, 2)
        return cdist.combine_partials([(r[0], r[1]) for r in rows])     # fixed rank order

    for i in range(args.warmup):
        one_step(i)
    combined_score()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pairs, prune_ms, mat_ms = [], [], []
    t_wall0 = time.time()
    for i in range(args.steps):
        pairs.append(one_step(args.warmup + i))
        st = ctx.last_stats()                       # syncs the stream; per-kernel CUDA-event times of this step
        prune_ms.append(st["ms_prune"])
        mat_ms.append(st["ms_matrices"])
    total, failed_all = combined_score()
    barrier()
    t_wall = time.time() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in pairs]
    ms_per_step = float(np.mean(ms_steps))
    stats = ctx.last_stats()
    n_mats = stats["matrices"]

    # ---------------- end to end, one process per GPU ("e2e_ranks"; at N = 1 this is "e2e") ----------------
    h2d = len(prior) * 4 + n_mats * 24 + K * tree.n_nodes * 4 + K * 8
    d2h_per_family = K * 8 + 8 + K * 8 + K + 1

    def e2e_loop(c, n_steps, exchange):
        for i in range(2):
            lam, alpha, cp, mu = step_params(100 + i, K)
            c.eval_gamma([lam], alpha, mu, cp, pinned=True)
        ms = []
        for i in range(n_steps):
            lam, alpha, cp, mu = step_params(200 + i, K)
            t0 = time.perf_counter()
            c.set_prior(prior)                        # host buffers in, every step
            c.set_error_model(None)
            out = c.eval_gamma([lam], alpha, mu, cp, pinned=True)   # all per-family outputs back to (page-locked) host memory
            if exchange:
                cdist.allreduce_score(out["neg_lnl"], out["n_failed"])
            ms.append((time.perf_counter() - t0) * 1e3)
        return float(np.mean(ms))

    barrier()
    ranks_ms = e2e_loop(ctx, args.steps, world > 1)
    barrier()

    # max over ranks; whole-job numerators
    U_total, launches_total = U, stats["launches"]
    flops_total, flops_plain_total = flops_per_step, flops_per_step_plain
    if world > 1:
        t = torch.tensor([ms_per_step, ranks_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, ranks_ms = float(t[0]), float(t[1])
        t = torch.tensor([U, stats["launches"], flops_per_step, flops_per_step_plain], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        U_total, launches_total, flops_total, flops_plain_total = int(t[0]), int(t[1]), float(t[2]), float(t[3])
    # numerator: DISTINCT families of the job (copies that replaced failing families and chance duplicates are not counted)
    value = U_total / (ms_per_step * 1e-3)
    e2e_ranks = {"value": U_total / (ranks_ms * 1e-3), "unit": "family-likelihood evals/s", "ms_per_step": ranks_ms,
                 "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(args.families * d2h_per_family + 16 * world),
                 "what": "one process per GPU; per step every rank uploads its parameters, copies all per-family outputs of its shard back "
                         "and exchanges its partial score (all_gather) through the host"}

    # ---------------- end to end through cafe_b200_create_multi: ONE process drives all N GPUs ("e2e" for N > 1) ----------------
    e2e = dict(e2e_ranks)
    e2e["what"] = ("cafe_b200_set_prior / set_error_model / eval_gamma with host buffers: per step the parameters go up and all "
                   "per-family outputs come back into page-locked host memory")
    if world > 1:
        multi_ms = None
        if rank == 0:
            try:
                # every rank holds the same simulated table; the failing families were replaced shard by shard, so rebuild the job's
                # table from the shards' healthy view: rank 0 repeats the (deterministic) replacement for every shard
                parts = [drop_failing_families(args, tree, all_counts, mfs, mrs, order[bounds[r]:bounds[r + 1]], local)[0] for r in range(world)]
                job = np.concatenate(parts)
                mctx = Context(tree, job, mfs, mrs, devices=list(range(world)))
                mctx.set_prior(prior)
                multi_ms = e2e_loop(mctx, args.steps, False)
                multi_U = mctx.unique_families()
                mctx.close()
            except Exception as e:      # rank 0 must reach the barrier below whatever happens; e2e then stays the per-rank number
                print("one-process e2e failed: %r" % (e,), file=sys.stderr, flush=True)
                multi_ms = None
        dist.barrier(group=cpu_group)                 # the other ranks wait on the host: no kernel of theirs on the GPUs meanwhile
        if rank == 0 and multi_ms is not None:
            e2e = {"value": multi_U / (multi_ms * 1e-3), "unit": "family-likelihood evals/s", "ms_per_step": multi_ms,
                   "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(args.families * d2h_per_family + 16 * world),
                   "what": "cafe_b200_create_multi: one host process drives all %d GPUs (what a CAFE5 process linked against the drop-in "
                           "models does); the per-device partial scores are added on the host in device order" % world}

    # ---------------- weak-scaling side series: 125,000 families per GPU (round 1's workload) ----------------
    weak = None
    # (the decision must not depend on the rank: the shards have different sizes and the block below contains collectives)
    if not args.no_weak and int(np.diff(bounds).min()) > WEAK_FAMILIES:
        wctx = Context(tree, counts[:WEAK_FAMILIES], mfs, mrs, device=local)
        wctx.set_prior(prior)
        wstream = torch.cuda.ExternalStream(wctx.stream(), device=torch.device("cuda", local))
        wpairs = []
        for i in range(3 + min(args.steps, 10)):
            lam, alpha, cp, mu = step_params(i, K)
            with torch.cuda.stream(wstream):
                flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(wstream)
            wctx.enqueue_eval([lam], alpha, mu, cp)
            b.record(wstream)
            wpairs.append((a, b))
        wctx.fetch_result()
        wms = float(np.mean([a.elapsed_time(b) for a, b in wpairs[3:]]))
        wU = wctx.unique_families()
        wctx.close()
        if world > 1:
            t = torch.tensor([wms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wms = float(t[0])
            t = torch.tensor([wU], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wU = int(t[0])
        weak = {"families_per_gpu": WEAK_FAMILIES, "value": wU / (wms * 1e-3), "ms_per_step": wms, "unit": "family-likelihood evals/s",
                "what": "the same step on %d families per GPU (weak series; device-timed, no exchange)" % WEAK_FAMILIES}

    # ---------------- wall time per optimisation (the metric's second half) ----------------
    fit = None
    if not args.no_fit:
        fit = {}
        # explicit start: with this many families on 118 branches the reference's random start (normal(0.002 L, 0.2) / L) is usually
        # rejected by the all-or-nothing rule (one underflowed family => +inf), for the reference exactly as for us
        start = [1.5 * LAMBDA0, 1.0]
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            r = ctx.fit(n_cat=K, start=start)
            how = "cafe_b200_fit (C++ host driver)"
        else:
            from cafe5_b200.model import discrete_gamma

            def local_score(v):   # this rank's shard; the ranks exchange their partials inside fit_sharded
                if not (v[1] > 0):
                    return math.inf, 0
                cp_i, mu_i = discrete_gamma(K, v[1])
                o = ctx.eval_gamma([v[0]], v[1], mu_i, cp_i, want_family=False)
                return o["neg_lnl"], o["n_failed"]

            r = cdist.fit_sharded(local_score, start)
            how = "cafe5_b200.dist.fit_sharded (the same simplex search on every rank over the all-gathered score)"
        barrier()
        fit["config5"] = {"wall_s": time.perf_counter() - t0, "iterations": r["iterations"], "evaluations": r["evaluations"],
                          "status": r["status"], "values": [float(r["values"][0]), float(r["values"][1])], "neg_lnl": r["neg_lnl"],
                          "families": args.families,
                          "what": "%s: gamma K=%d, (lambda, alpha) estimated by Nelder-Mead (tolx = tolf = 1e-6, reference constants) over "
                                  "the whole job from the start point (1.5 x true lambda, alpha = 1)" % (how, K)}
        if rank == 0:
            try:
                fit.update(real_config_fits(local))
            except Exception as e:      # a reported side number, never a reason to lose the bench line
                fit["configs_1_4_error"] = repr(e)
        if world > 1:
            dist.barrier(group=cpu_group)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import pyoracle
            if pyoracle.have_ref():
                v, t_step, cores, sample = reference_rate(tree.newick, tree.species, counts, mfs, mrs, prior, K, args.families, 1, 0)
                cpu = {"value": v, "unit": "family-likelihood evals/s", "cores": cores, "kind": "reference", "sample": sample}
            else:
                v, t_step, cores, sample = port_rate(tree, counts, mfs, mrs, prior, K, args.families)
                cpu = {"value": v, "unit": "family-likelihood evals/s", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "family-likelihood evals/s", "cores": None, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        prune = float(np.mean(prune_ms))
        achieved = flops_per_step / (prune * 1e-3) / 1e12          # rank 0's launches against rank 0's work
        peak = max(peak_dfma, peak_dmma)
        line = {
            "metric": "family-likelihood evals/s", "value": value, "unit": "family-likelihood evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, tree),
            "e2e": e2e,
            "e2e_ranks": e2e_ranks,
            "gpu_launches": int(launches_total) * args.steps,
            "roofline": {"bound": "tensor", "kernel": prune_kernel_name(), "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(args, world),
                         "peak_source": "FP64 tensor pipe (DMMA; tcgen05 has no FP64 kind) measured live on this GPU by "
                                        "cafe_b200_measure_fp64_peak: DMMA m8n8k4 %.2f, DFMA %.2f TFLOP/s; MEASURED_PEAKS.json "
                                        "carries only HBM and bf16 figures" % (peak_dmma, peak_dfma),
                         "alg_flops_per_step": flops_per_step, "ms_per_step_kernel": prune,
                         "launches_per_step": int(stats["launches"]) - 4,
                         "what": "all launches of the pruning kernel in a step (factor tables of the subtree-pattern reuse + the main pass) "
                                 "against the algorithmic flops of the columns actually computed (rank 0)",
                         "share_of_step": prune / ms_per_step,
                         "pattern_reuse": {"alg_flops_without": flops_per_step_plain, "fraction_executed": flops_per_step / flops_per_step_plain,
                                           "equivalent_TFLOPs": flops_per_step_plain / (prune * 1e-3) / 1e12},
                         "matrix_gen": {"ms_per_launch": float(np.mean(mat_ms)), "matrices": n_mats,
                                        "write_GBps": n_mats * 8.0 * (max(mfs, mrs) + 1) ** 2 / (float(np.mean(mat_ms)) * 1e-3) / 1e9,
                                        "terms_per_s": n_mats * matrix_terms(max(mfs, mrs) + 1) / (float(np.mean(mat_ms)) * 1e-3)}},
            "cpu_baseline": cpu,
            "weak": weak,
            "optimisations": fit,
            "clocks": clocks,
            "result": {"neg_lnl": total, "n_failed": failed_all, "families": args.families, "distinct_families": int(U_total),
                       "failing_families_replaced_rank0": n_replaced, "wall_s_timed_region": t_wall, "create_s_rank0": t_create,
                       "setup_s_rank0": time.perf_counter() - t_setup0, "generator": gen_info},
        }
        print(json.dumps(_finite(line)), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def prune_kernel_name():
    return {"dfma": "prune_kernel (DFMA register tiles)", "stream": "prune_dmma_kernel (DMMA, both operands streamed)"}.get(
        os.environ.get("CAFE_B200_PRUNE", ""), "prune_resident_kernel (FP64 tensor cores: mma.sync.m8n8k4.f64 / DMMA.8x8x4)")


def ncu_traffic_bytes(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pruning kernel's launches of ONE step in this exact configuration, from
    the committed `ncu --set full` capture (profiles/r02_prune_ncu_traffic.json, written by tools/ncu_summary.py); None when another
    variant / size is benchmarked."""
    if any(os.environ.get(k) for k in ("CAFE_B200_PRUNE", "CAFE_B200_RESIDENT_WN", "CAFE_B200_TABLES", "CAFE_B200_TABLE_FRAC")):
        return None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_prune_ncu_traffic.json")))
        key = "families_per_gpu_%d" % (args.families // world)
        return d.get(key)
    except (OSError, ValueError):
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
