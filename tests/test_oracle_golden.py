"""The C oracle against the reference's OWN known-answer tests (test.cpp, cited per test) and against the
fixtures produced by the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

from cafe5_b200 import families as fam
from cafe5_b200.tree import FlatTree

from conftest import max_rel, random_prior


class approx:
    """doctest::Approx semantics: |a - b| < epsilon * (scale + max(|a|, |b|)), epsilon = 100 * FLT_EPSILON, scale = 1."""

    def __init__(self, want, eps=1.1920929e-5, scale=1.0):
        self.want, self.eps, self.scale = np.asarray(want, dtype=float), eps, scale

    def __eq__(self, got):
        got = np.asarray(got, dtype=float)
        return bool(np.all(np.abs(got - self.want) < self.eps * (self.scale + np.maximum(np.abs(got), np.abs(self.want)))))

    matches = __eq__


def counts_for(tree, **by_name):
    row = np.zeros((1, tree.n_leaves), dtype=np.int32)
    for name, v in by_name.items():
        row[0, tree.species.index(name)] = v
    return row


def test_probability_of_some_values(oracle):
    # test.cpp:549-562
    assert approx(0.0152237).matches(oracle.transition(0.05, 5, 5, 9))
    assert approx(0.17573).matches(oracle.transition(0.05, 5, 10, 9))
    assert approx(0.182728).matches(oracle.transition(0.05, 5, 10, 10))
    assert approx(0.465565).matches(oracle.transition(0.05, 1, 10, 10))
    # 17-digit value from the compiled reference
    assert oracle.transition(0.05, 5, 10, 9) == 0.17573649449589199


def test_fractional_branch_lengths_and_key_truncation(oracle):
    # test.cpp:564-577; SURVEY 8c: the cache computes at t=68.71 (key truncation), the free function at 68.7105
    m = oracle.matrix(141, 0.006335, 68.7105)
    assert approx(0.194661, eps=1e-4).matches(m[5, 5])
    assert m[5, 5] == 0.19466119864519926
    assert oracle.matrix(141, 0.006335, 68)[5, 5] == 0.19579137469510155
    assert m[100, 120] == 0.0045868502540086847
    assert m[140, 140] == 0.036152834093856973
    assert oracle.transition(0.006335, 68.7105, 5, 5) == 0.19466040947681443


def test_probability_of_matrix(oracle):
    # test.cpp:595-617
    expected = np.array([[1, 0, 0, 0, 0],
                         [0.2, 0.64, 0.128, 0.0256, 0.00512],
                         [0.04, 0.256, 0.4608, 0.17408, 0.0512],
                         [0.008, 0.0768, 0.26112, 0.36352, 0.187392],
                         [0.0016, 0.02048, 0.1024, 0.249856, 0.305562]])
    assert np.abs(oracle.matrix(5, 0.05, 5) - expected).max() < 1e-5


def test_birthdeath_rate_with_log_alpha(oracle):
    # test.cpp:1329-1343
    for s, c, la, co, want in [(46, 45, -3.672556, 0.949177, -1.55455), (44, 46, -2.617970, 0.854098, -2.20436),
                               (43, 43, -1.686354, 0.629613, -2.39974), (43, 44, -1.686354, 0.629613, -2.44301),
                               (13, 14, -2.617970, 0.854098, -1.58253)]:
        assert approx(want).matches(math.log(oracle.birthdeath(s, c, la, co)))
    assert approx(0.107933).matches(oracle.birthdeath(40, 42, -1.37, 0.5))
    assert approx(0.005714).matches(oracle.birthdeath(41, 34, -1.262, 0.4))
    assert approx(0.194661).matches(oracle.birthdeath(5, 5, -1.1931291703283662, 0.39345841643135504))
    assert oracle.birthdeath(46, 45, -3.672556, 0.949177) == 0.21128389159772853


def test_matrix_cache_key_quantisation(oracle):
    # test.cpp:1313-1327: 31 distinct lambda keys for t = 0.1 .. 3.1, and 3.0 is one of them
    keys, t = set(), 0.0
    for _ in range(31):
        t += 0.1
        keys.add(oracle.lib.oracle_key_lambda(t))
    assert len(keys) == 31
    assert oracle.lib.oracle_key_lambda(3.0) in keys
    assert oracle.lib.oracle_key_branch(68.710507) == 68710


def test_saturation(oracle):
    # test.cpp:1630-1644
    assert oracle.lib.oracle_is_saturated(25, 0.05)
    assert not oracle.lib.oracle_is_saturated(25, 0.01)
    m = oracle.matrix(10, 0.05, 25)
    assert m[0, 0] == 1.0 and m.sum() == 1.0      # matrix_cache.cpp:144-146: only (0,0) is set


def test_inference_prune(oracle):
    # test.cpp:1662-1683
    tree = FlatTree("(A:1,B:3):7")
    got = oracle.prune(tree, counts_for(tree, A=3, B=6)[0], 20, 20, [0.03], 1.5)
    want = [-17.2771, -10.0323, -5.0695, -4.91426, -5.86062, -7.75163, -10.7347, -14.2334, -18.0458, -22.073, -26.2579,
            -30.5639, -34.9663, -39.4472, -43.9935, -48.595, -53.2439, -57.9338, -62.6597, -67.4173]
    assert approx(want).matches(np.log(got))
    assert got[0] == 3.1380342367040701e-08 and got[19] == 5.2604531182852501e-30


def test_root_node_probabilities(oracle):
    # test.cpp:1729-1763
    tree = FlatTree("(A:1,B:3):7")
    got = oracle.prune(tree, counts_for(tree, A=3, B=6)[0], 20, 20, [0.03], 1.0)
    want = [-19.7743, -11.6688, -5.85672, -5.66748, -6.61256, -8.59725, -12.2301, -16.4424, -20.9882, -25.7574, -30.6888,
            -35.7439, -40.8971, -46.1299, -51.4289, -56.7837, -62.1863, -67.6304, -73.1106, -78.6228]
    assert approx(want).matches(np.log(got))


def test_error_model_leaf_emission(oracle, ref):
    # test.cpp:1765-1804: leaf vector {0,0,.2,.6,.2,0...}; through the root vector, oracle == reference bit for bit
    tree = FlatTree("(A:1,B:3):7")
    em = (np.array([[0.0, 0.8, 0.2]] + [[0.2, 0.6, 0.2]] * 20), 20)
    cnt = counts_for(tree, A=3, B=6)
    got = oracle.prune(tree, cnt[0], 20, 20, [0.03], 1.5, em=em)
    ctx = ref.ctx("(A:1,B:3):7", tree.species, cnt, 20, 20, fam.uniform_prior(20), em=em)
    assert np.array_equal(got, ctx.prune(0, [0.03], 1.5))
    plain = oracle.prune(tree, cnt[0], 20, 20, [0.03], 1.5)
    assert not np.array_equal(got, plain)


def test_infer_processes(oracle):
    # test.cpp:461-488: -lnL = 46.56632 (17 digits: SURVEY 8c)
    tree = FlatTree("(A:1,B:1);")
    rows = np.concatenate([counts_for(tree, A=a, B=b) for a, b in [(1, 2), (2, 1), (3, 6), (6, 3)]])
    out = oracle.eval_base(tree, rows, 56, 8, fam.uniform_prior(100), [0.01])
    assert approx(46.56632).matches(out["neg_lnl"])
    assert out["neg_lnl"] == 46.566319823357453


def test_gamma_model_prune(oracle):
    # test.cpp:1269-1294
    tree = FlatTree("(A:1,B:3):7")
    prior = fam.rootdist_prior({1: 2, 2: 2, 3: 2, 4: 2, 5: 1})
    out = oracle.eval_gamma(tree, counts_for(tree, A=3, B=6), 10, 8, prior, [0.005], [0.1, 0.5], [0.01, 0.05])
    assert approx([-23.04433, -16.68005]).matches(np.log(out["cat_lk"][0]))
    assert out["cat_lk"][0, 0] == 9.8169154483081593e-11 and out["cat_lk"][0, 1] == 5.7009161811906484e-08


def test_gamma_model_prune_fails_if_saturated(oracle):
    # test.cpp:1296-1311: lambda 0.9 x {0.1, 0.5} on branches 1,3,7 -> a dead category
    tree = FlatTree("(A:1,B:3):7")
    out = oracle.eval_gamma(tree, counts_for(tree, A=3, B=6), 10, 8, fam.uniform_prior(100), [0.9], [0.1, 0.5], [1.0, 1.0])
    assert math.isinf(out["neg_lnl"])


def test_invalid_parameters_give_inf(oracle):
    tree = FlatTree("(A:1,B:1);")
    rows = counts_for(tree, A=1, B=2)
    assert math.isinf(oracle.eval_base(tree, rows, 56, 8, fam.uniform_prior(100), [-0.01])["neg_lnl"])   # base_model.cpp:56-60
    assert math.isinf(oracle.eval_gamma(tree, rows, 56, 8, fam.uniform_prior(100), [0.01], [0.5, 1.5], [0.5, 0.5], alpha=-1)["neg_lnl"])


def test_reconstruct_gene_family(oracle):
    # test.cpp:1090-1118: AB = 4
    tree = FlatTree("(A:1,B:3):7")
    prior = fam.rootdist_prior({i: v for i, v in enumerate([1, 2, 3, 4, 5, 4, 3, 2, 1])})
    out = oracle.reconstruct(tree, counts_for(tree, A=3, B=6), 10, 8, prior, [0.005])
    assert out["states"][0, tree.names.index("AB")] == 4


def test_get_gamma_known_answer(oracle):
    p, m = oracle.get_gamma(4, 0.65)
    assert list(m) == [0.062015465425384449, 0.37328920830134099, 0.99805780528212318, 2.5666375209911516]
    assert list(p) == [0.25] * 4


# ---- fixtures produced by the unmodified reference -------------------------------------------------

def test_matrices_match_reference_fixture(oracle, golden):
    g = golden["matrices"]
    for name in ("m171_a", "m171_b", "m141", "m5", "msat", "m201"):
        N, lam, t = g[name + "_params"]
        assert np.array_equal(oracle.matrix(int(N), lam, t), g["ref_" + name]), name
    got = [oracle.birthdeath(int(r[0]), int(r[1]), r[2], r[3]) for r in g["bd_in"]]
    assert np.array_equal(got, g["ref_bd"])


def test_small_problems_match_reference_fixture(oracle, golden):
    g = golden["small"]
    mfs, mrs, lam = int(g["max_family_size"]), int(g["max_root_family_size"]), float(g["lambda"])
    p3, m3 = oracle.get_gamma(3, float(g["gamma_alpha"]))
    for ti in range(4):
        tree = FlatTree(str(g["t%d_newick" % ti]))
        counts = g["t%d_counts" % ti]
        prior = fam.uniform_prior(mrs)
        b = oracle.eval_base(tree, counts, mfs, mrs, prior, [lam], want_roots=True)
        assert np.array_equal(b["roots"], g["t%d_ref_roots" % ti])
        assert np.array_equal(b["family_lnl"], g["t%d_ref_family_lnl" % ti])
        assert b["neg_lnl"] == float(g["t%d_ref_base" % ti])
        gm = oracle.eval_gamma(tree, counts, mfs, mrs, prior, [lam], m3, p3)
        assert np.array_equal(gm["cat_lk"], g["t%d_ref_cat_lk" % ti])
        assert gm["neg_lnl"] == float(g["t%d_ref_gamma" % ti])
        rb = oracle.reconstruct(tree, counts, mfs, mrs, prior, [lam])
        assert np.array_equal(rb["states"], g["t%d_ref_rec" % ti])
        rg = oracle.reconstruct(tree, counts, mfs, mrs, prior, [lam], m3, p3)
        assert np.array_equal(rg["states"], g["t%d_ref_rec_gamma" % ti])
        assert np.array_equal(rg["cat_states"], g["t%d_ref_rec_gamma_cat" % ti])


def test_mammals_subset_matches_reference_fixture(oracle, golden):
    g = golden["mammals"]
    tree = FlatTree(str(g["newick"]), species=[str(s) for s in g["species"]])
    counts = g["counts"].astype(np.int32)
    mfs, mrs = int(g["max_family_size"]), int(g["max_root_family_size"])
    assert (mfs, mrs) == (170, 150) and counts.shape[0] == 10956      # size floors, root filter
    prior = fam.uniform_prior(mrs)
    sub = np.arange(0, counts.shape[0], 37)
    b = oracle.eval_base(tree, counts[sub], mfs, mrs, prior, [0.0018], want_roots=True)
    assert np.array_equal(b["family_lnl"], g["ref_base_family_lnl"][sub])
    assert np.array_equal(b["roots"][:64], g["ref_roots"])
    gm = oracle.eval_gamma(tree, counts[sub], mfs, mrs, prior, [0.0018], g["gamma_mult"], g["gamma_probs"])
    assert np.array_equal(gm["cat_lk"], g["ref_gamma_cat_lk"])
    # the reference's underflow cliff: family 73 dies at alpha = 0.6 (SURVEY 7 "hard parts")
    gf = oracle.eval_gamma(tree, counts[:128], mfs, mrs, prior, [0.0018], g["gamma_fail_mult"], g["gamma_fail_probs"])
    assert list(np.nonzero(gf["failed"])[0]) == list(g["ref_gamma_fail_failed"]) == [73]
    assert math.isinf(gf["neg_lnl"]) and math.isinf(float(g["ref_gamma_fail_neg_lnl"]))
    # error model + two lambda classes (config 3)
    from cafe5_b200.tree import FlatTree as FT
    tree3 = FT(str(g["newick"]), str(g["lambda_newick"]), species=[str(s) for s in g["species"]])
    b3 = oracle.eval_base(tree3, counts[sub], mfs, mrs, prior, g["em_lambdas"], em=(g["em_probs"], int(g["em_maxcnt"])))
    assert np.array_equal(b3["family_lnl"], g["ref_em_family_lnl"][sub])
    # Pupko
    rs = g["rec_sub"][:60]
    rb = oracle.reconstruct(tree, counts[rs], mfs, mrs, prior, [0.0018])
    assert np.array_equal(rb["states"], g["ref_rec_base"][:60])
    rg = oracle.reconstruct(tree, counts[rs], mfs, mrs, prior, [0.0018], g["gamma_mult"], g["gamma_probs"])
    assert np.array_equal(rg["states"], g["ref_rec_gamma_states"][:60])
    assert max_rel(rg["averaged"], g["ref_rec_gamma_avg"][:60]) == 0.0


def test_non_uniform_priors_match_reference_fixture(oracle, golden):
    """Row a12: the prior's index conventions are invisible under a uniform prior.  Fixtures made by the reference with a spiky user
    root distribution (`-f`; holes inside the table, table shorter than max_root_family_size so compute() returns 0 beyond it,
    src/root_equilibrium_distribution.cpp:81-87) and Poisson priors (`-p<lambda>`, :56-68): inference weights root size j+1 with
    compute(j) (base_model.cpp:84, gamma_core.cpp:156), Pupko's root picks argmax over size j of L[j] * compute(j)
    (gene_family_reconstructor.cpp:65)."""
    g, small, m = golden["priors"], golden["small"], golden["mammals"]
    mfs, mrs, lam = int(small["max_family_size"]), int(small["max_root_family_size"]), float(small["lambda"])
    p3, m3 = oracle.get_gamma(3, float(small["gamma_alpha"]))
    for pname in ("rootdist", "poisson"):
        prior = g["small_%s_prior" % pname]
        assert prior.dtype == np.float32 and len(prior) < mrs
        for ti in range(4):
            tree = FlatTree(str(small["t%d_newick" % ti]))
            counts = small["t%d_counts" % ti]
            k = "small_%s_t%d_" % (pname, ti)
            b = oracle.eval_base(tree, counts, mfs, mrs, prior, [lam])
            assert np.array_equal(b["family_lnl"], g[k + "ref_family_lnl"]) and b["neg_lnl"] == float(g[k + "ref_base"])
            gm = oracle.eval_gamma(tree, counts, mfs, mrs, prior, [lam], m3, p3)
            assert np.array_equal(gm["failed"].astype(bool), g[k + "ref_failed"].astype(bool))
            assert np.array_equal(gm["cat_lk"], g[k + "ref_cat_lk"])
            assert gm["neg_lnl"] == float(g[k + "ref_gamma"]) or (math.isinf(gm["neg_lnl"]) and math.isinf(float(g[k + "ref_gamma"])))
            assert np.array_equal(oracle.reconstruct(tree, counts, mfs, mrs, prior, [lam])["states"], g[k + "ref_rec"])
            rg = oracle.reconstruct(tree, counts, mfs, mrs, prior, [lam], m3, p3)
            assert np.array_equal(rg["states"], g[k + "ref_rec_gamma"]) and np.array_equal(rg["cat_states"], g[k + "ref_rec_gamma_cat"])
            # the fixtures DO depend on the convention: the uniform-prior reconstruction differs somewhere
            if pname == "rootdist":
                assert not np.array_equal(g[k + "ref_rec"], small["t%d_ref_rec" % ti])
    tree = FlatTree(str(m["newick"]), species=[str(s) for s in m["species"]])
    counts = m["counts"].astype(np.int32)[g["mammals_sub"]][:80]
    mfs, mrs = int(m["max_family_size"]), int(m["max_root_family_size"])
    for pname in ("rootdist", "poisson08", "poisson12"):
        prior = g["mammals_%s_prior" % pname]
        k = "mammals_%s_" % pname
        b = oracle.eval_base(tree, counts, mfs, mrs, prior, [0.0018])
        assert np.array_equal(b["family_lnl"], g[k + "ref_family_lnl"][:80])
        gm = oracle.eval_gamma(tree, counts, mfs, mrs, prior, [0.0018], m["gamma_mult"], m["gamma_probs"])
        assert np.array_equal(gm["cat_lk"], g[k + "ref_cat_lk"][:80])
        rg = oracle.reconstruct(tree, counts[:30], mfs, mrs, prior, [0.0018], m["gamma_mult"], m["gamma_probs"])
        assert np.array_equal(rg["states"], g[k + "ref_rec_gamma"][:30]) and np.array_equal(rg["cat_states"], g[k + "ref_rec_gamma_cat"][:30])
        assert np.array_equal(oracle.reconstruct(tree, counts[:30], mfs, mrs, prior, [0.0018])["states"], g[k + "ref_rec"][:30])


@pytest.mark.parametrize("seed", range(8))
def test_oracle_is_bit_identical_to_the_live_reference_on_random_problems(oracle, ref, seed):
    """Beyond the committed fixtures: random trees (every third with multifurcations), 1-3 lambda classes, an error model on odd
    seeds -- the C restatement against the unmodified reference compiled into oracle/_ref, bit for bit."""
    import re
    rng = np.random.default_rng(500 + seed)
    n_taxa = int(rng.integers(3, 11))
    nodes = ["t%d" % i for i in range(n_taxa)]
    lam_nodes = list(nodes)
    while len(nodes) > 1:
        k = 3 if (seed % 3 == 0 and len(nodes) >= 3 and rng.random() < 0.4) else 2
        idx = sorted(rng.choice(len(nodes), size=k, replace=False), reverse=True)
        parts, lparts = [], []
        for i in idx:
            parts.append("%s:%g" % (nodes.pop(i), round(float(rng.uniform(0.2, 8.0)), 3)))
            lparts.append("%s:%d" % (lam_nodes.pop(i), int(rng.integers(1, 4))))
        nodes.append("(" + ",".join(parts) + ")")
        lam_nodes.append("(" + ",".join(lparts) + ")")
    newick, lam_newick = nodes[0], lam_nodes[0]
    used = sorted({int(x) for x in re.findall(r":(\d+)", lam_newick)})
    remap = {c: i + 1 for i, c in enumerate(used)}
    lam_newick = re.sub(r":(\d+)", lambda m: ":%d" % remap[int(m.group(1))], lam_newick)
    n_classes = len(used)
    lam_newick = lam_newick if n_classes > 1 else None
    tree = FlatTree(newick, lam_newick)
    F = int(rng.integers(1, 12))
    base = rng.integers(0, 26, size=F)
    counts = np.clip(base[:, None] + rng.integers(-4, 5, size=(F, tree.n_leaves)), 0, 30).astype(np.int32)
    mfs, mrs = 50, 38
    prior = random_prior(np.random.default_rng(900 + seed), mrs)       # uniform / root distribution / Poisson (own stream: the problems stay as they were)
    lambdas = [float(rng.uniform(0.001, 0.03)) for _ in range(n_classes)]
    em = (fam.epsilon_error_model(float(rng.uniform(0.01, 0.2)), mfs)) if seed % 2 == 1 else None
    rctx = ref.ctx(newick, tree.species, counts, mfs, mrs, prior, lambda_newick=lam_newick, em=em)
    rb = rctx.eval_base(lambdas)
    ob = oracle.eval_base(tree, counts, mfs, mrs, prior, lambdas, em=em)
    assert np.array_equal(ob["family_lnl"], rb["family_lnl"]) and (ob["neg_lnl"] == rb["neg_lnl"] or (math.isinf(ob["neg_lnl"]) and math.isinf(rb["neg_lnl"])))
    for f in range(min(F, 3)):
        assert np.array_equal(oracle.prune(tree, counts[f], mfs, mrs, lambdas, em=em), rctx.prune(f, lambdas))
    cp, mu = oracle.get_gamma(int(rng.integers(2, 5)), float(rng.uniform(0.4, 2.5)))
    rg = rctx.eval_gamma(lambdas, mu, cp)
    og = oracle.eval_gamma(tree, counts, mfs, mrs, prior, lambdas, mu, cp, em=em)
    assert np.array_equal(og["failed"].astype(bool), rg["failed"].astype(bool))
    ok = ~rg["failed"].astype(bool)
    assert np.array_equal(og["cat_lk"][ok], rg["cat_lk"][ok])
    assert og["neg_lnl"] == rg["neg_lnl"] or (math.isinf(og["neg_lnl"]) and math.isinf(rg["neg_lnl"]))
    rctx.close()
    rctx = ref.ctx(newick, tree.species, counts, mfs, mrs, prior, lambda_newick=lam_newick)   # reconstruction ignores the error model
    assert np.array_equal(oracle.reconstruct(tree, counts, mfs, mrs, prior, lambdas)["states"], rctx.reconstruct_base(lambdas))
    rr = rctx.reconstruct_gamma(lambdas, mu, cp)
    orr = oracle.reconstruct(tree, counts, mfs, mrs, prior, lambdas, mu, cp)
    assert np.array_equal(orr["cat_states"], rr["cat_states"]) and np.array_equal(orr["states"], rr["states"])
    rctx.close()
