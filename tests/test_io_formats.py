"""the library's C++ readers and writers of the reference's on-disk formats (cafe5_b200/host/io.hpp, reached through
the C ABI) against the reference's own parsers and writers (oracle/_ref, compiled from the unmodified sources) and against the files
the reference ships.  Host-only: runs without a GPU."""
import math
import os

import numpy as np
import pytest

from cafe5_b200 import families as fam
from cafe5_b200 import io_cpp
from cafe5_b200.gamma import get_gamma
from cafe5_b200.tree import FlatTree

REF = os.environ.get("CAFE_REF_DIR", "/root/reference")
needs_files = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference example files absent")

TREES = [
    "((A:1,B:1):1,(C:1,D:1):1);",
    "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)",
    "((E:0.36,D:0.30)H:1.00,(C:0.85,(A:0.59,B:0.35)F:0.42)G:0.45)I;",           # interior labels are kept
    "((A:1,B:1,C:2.5):4,(D:3,(E:1,F:1):2):3.5)",                                   # multifurcation
    "A:1,B:3",                                                                    # no outer parentheses
]


@pytest.mark.parametrize("newick", TREES)
def test_cpp_tree_parser_matches_reference(ref, newick):
    parent, bl, is_leaf, _, names = ref.flatten(newick)
    t = io_cpp.parse_tree(newick)
    assert np.array_equal(t["parent"], parent) and np.array_equal(t["branch_length"], bl)
    assert np.array_equal(t["is_leaf"], is_leaf.astype(bool)) and t["names"] == names
    py = FlatTree(newick)
    assert np.array_equal(t["parent"], py.parent) and t["names"] == py.names


def test_cpp_lambda_tree_classes_match_reference(ref):
    newick = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"
    lam = "(((chimp:2,human:2):2,(mouse:1,rat:1):1):1,dog:1)"
    _, _, _, cls, _ = ref.flatten(newick, lam)
    t = io_cpp.parse_tree(newick, lam)
    assert np.array_equal(t["lambda_class"], cls) and t["n_lambda"] == 2
    with pytest.raises(io_cpp.IoError):
        io_cpp.parse_tree(newick, "((chimp:2,human:2):2,dog:1)")               # structure mismatch (clade.cpp:247-262)
    with pytest.raises(io_cpp.IoError):
        io_cpp.parse_tree("((A:1,B:0):1,C:2)")                                   # non-positive branch length (clade.cpp:411-414)


@needs_files
def test_cpp_family_reader_on_the_reference_examples(ref):
    newick = open(os.path.join(REF, "examples", "mammals_tree.txt")).readline().strip()
    path = os.path.join(REF, "examples", "mammal_gene_families.txt")
    species, ids, counts = io_cpp.read_gene_families(path)
    rids, rcounts = ref.read_families(path, newick)
    t = io_cpp.parse_tree(newick)
    leaves = [n for n, leaf in zip(t["names"], t["is_leaf"]) if leaf]
    col = {s.lower(): j for j, s in enumerate(species)}
    assert ids == rids
    assert np.array_equal(counts[:, [col[n.lower()] for n in leaves]], rcounts)
    assert io_cpp.derive_sizes(counts) == fam.derive_sizes(counts) == (170, 150)


def test_cpp_family_reader_both_header_styles(tmp_path, ref):
    newick = "((A:1,B:1):1,(C:1,D:1):1)"
    cafe = tmp_path / "cafe.txt"
    cafe.write_text("Desc\tFamily ID\tA\tB\tC\tD\n\n(null)\tfam1\t5\t10\t2\t6\n(null)\tfam2\t5\t10abc\t-2\t6\r\n")
    cafexp = tmp_path / "cafexp.txt"
    cafexp.write_text("#A\n#B\n#C\n#D\n5\t10\t2\t6\tfam1\n1\t0\t3\t7\tfam2\n")
    for path in (cafe, cafexp):
        species, ids, counts = io_cpp.read_gene_families(path)
        sp2, ids2, c2 = fam.read_gene_families(path)
        assert species == sp2 == ["A", "B", "C", "D"] and ids == ids2 and np.array_equal(counts, c2)
        rids, rcounts = ref.read_families(path, newick)
        # the reference's tree order of the leaves of this newick is D C B A (reverse level order)
        t = io_cpp.parse_tree(newick)
        leaves = [n for n, leaf in zip(t["names"], t["is_leaf"]) if leaf]
        assert ids == rids and np.array_equal(counts[:, [species.index(n) for n in leaves]], rcounts)
    with pytest.raises(io_cpp.IoError):
        empty = tmp_path / "empty.txt"
        empty.write_text("Desc\tFamily ID\tA\tB\n")
        io_cpp.read_gene_families(empty)
    # CAFExp header naming INTERIOR nodes too (src/io.cpp:153-161): every '#name' advances the column index, only leaves get a column;
    # a name that is not in the tree is rejected
    mixed = tmp_path / "cafexp_interior.txt"
    mixed.write_text("#A\n#AB\n#B\n#C\n#CD\n#D\n5\t99\t10\t2\t77\t6\tfam1\n1\t98\t0\t3\t76\t7\tfam2\n")
    species, ids, counts = io_cpp.read_gene_families(mixed, newick)
    assert species == ["A", "B", "C", "D"] and ids == ["fam1", "fam2"] and counts.tolist() == [[5, 10, 2, 6], [1, 0, 3, 7]]
    rids, rcounts = ref.read_families(mixed, newick)
    t = io_cpp.parse_tree(newick)
    leaves = [n for n, leaf in zip(t["names"], t["is_leaf"]) if leaf]
    assert ids == rids and np.array_equal(counts[:, [species.index(n) for n in leaves]], rcounts)
    assert io_cpp.derive_sizes(counts) == fam.derive_sizes(counts)          # 99 / 98 never reach the size derivation
    unknown = tmp_path / "cafexp_unknown.txt"
    unknown.write_text("#A\n#Z\n5\t1\tfam1\n")
    with pytest.raises(io_cpp.IoError, match="Z not located in tree"):
        io_cpp.read_gene_families(unknown, newick)
    with pytest.raises(io_cpp.IoError):
        io_cpp.read_error_model(tmp_path / "does_not_exist.txt")


@needs_files
def test_cpp_error_model_reader_matches_reference(ref, tmp_path):
    path = os.path.join(REF, "examples", "errormodel_0.1.txt")
    probs, mx = io_cpp.read_error_model(path)
    rprobs, rmx = ref.read_error_model(path)
    pprobs, pmx = fam.read_error_model(path)
    assert mx == rmx == pmx
    assert np.array_equal(probs, rprobs) and np.array_equal(probs, pprobs)
    sparse = tmp_path / "em.txt"
    sparse.write_text("maxcnt:10\ncntdiff -1 0 1\n0 0.0 0.8 0.2\n1 0.2 0.6 0.2\n5 0.3 0.5 0.2\n")
    probs, mx = io_cpp.read_error_model(sparse)
    rprobs, _ = ref.read_error_model(sparse)
    assert mx == 10 and np.array_equal(probs, rprobs) and np.array_equal(probs[2], probs[1])   # missing sizes inherit the previous row
    bad = tmp_path / "bad.txt"
    bad.write_text("maxcnt:10\ncntdiff -1 0 1\n0 0.2 0.6 0.2\n")
    with pytest.raises(io_cpp.IoError):
        io_cpp.read_error_model(bad)


def _small_problem():
    newick = "(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25)"
    rng = np.random.default_rng(21)
    base = rng.integers(1, 25, size=12)
    counts = np.clip(base[:, None] + rng.integers(-3, 4, size=(12, 7)), 0, 40).astype(np.int32)
    return newick, counts


def test_cpp_writers_reproduce_the_reference_text(ref, oracle):
    """Feed the writers the C oracle's numbers (equal to the reference's to ~1e-14) and compare with what the reference's own
    write_family_likelihoods / write_vital_statistics / print_category_likelihoods print for the same model."""
    newick, counts = _small_problem()
    tree = FlatTree(newick)
    mfs, mrs = 60, 45
    prior = fam.uniform_prior(mrs)
    ids = [str(i) for i in range(counts.shape[0])]
    rctx = ref.ctx(newick, tree.species, counts, mfs, mrs, prior)
    lam = [0.0123456789]
    # base model
    fam_txt, res_txt, _ = rctx.write_outputs(lam)
    want = oracle.eval_base(tree, counts, mfs, mrs, prior, lam)
    assert io_cpp.format_family_likelihoods(ids, "base", family_values=want["family_lnl"]) == fam_txt
    longest = float(tree.branch_length.max())
    assert io_cpp.format_results("Base", want["neg_lnl"], lam, longest, 1, 0) == res_txt
    # gamma model
    alpha = 0.7
    cp, mu = get_gamma(3, alpha)
    fam_txt, res_txt, cat_txt = rctx.write_outputs(lam, mu, cp, alpha=alpha)
    og = oracle.eval_gamma(tree, counts, mfs, mrs, prior, lam, mu, cp, alpha=alpha)
    assert io_cpp.format_family_likelihoods(ids, "gamma", family_values=og["family_lk"], multipliers=mu, cat_lk=og["cat_lk"],
                                            posterior=og["posterior"], significant=og["significant"]) == fam_txt
    assert io_cpp.format_results("Gamma", og["neg_lnl"], lam, longest, 1, 0, alpha=alpha) == res_txt
    assert io_cpp.format_family_likelihoods(ids, "categories", multipliers=mu, cat_lk=og["cat_lk"]) == cat_txt
    rctx.close()


def test_cpp_results_writer_with_two_lambdas_and_epsilon():
    txt = io_cpp.format_results("Base", 154503.34090635672, [0.0008193138839916892, 0.0096118070615455], 93.0, 141, 7, epsilon=0.05)
    assert txt.splitlines() == ["Model Base Final Likelihood (-lnL): 154503",
                                "Lambda: 0.00081931388399169, 0.0096118070615455",
                                "Epsilon: 0.05",
                                "Maximum possible lambda for this topology: 0.0107527",
                                "141 values were attempted (5% rejected)"]
    assert not math.isnan(float(txt.split()[5]))


@pytest.mark.parametrize("newick", [
    "(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25)",
    "((A:2.5,B:2.5,C:2.5):4,(D:3,(E:1,F:1):2):3.5)",                               # multifurcation
    "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)",
])
@pytest.mark.parametrize("gamma", [False, True])
def test_cpp_reconstruction_tables_reproduce_the_reference_text(ref, oracle, newick, gamma):
    """<Model>_count.tab, _change.tab, _asr.tre and _family_results.txt: the C++ writers fed with the C oracle's Pupko states (equal to
    the reference's, tests above) against reconstruction::write_results' own printers; node labels follow the reference's ape numbering.
    _clade_results.txt is compared as a set of lines (the reference orders its rows by pointer value)."""
    tree = FlatTree(newick)
    rng = np.random.default_rng(len(newick))
    F = 9
    base = rng.integers(1, 20, size=F)
    counts = np.clip(base[:, None] + rng.integers(-4, 5, size=(F, tree.n_leaves)), 0, 40).astype(np.int32)
    mfs, mrs = 60, 45
    prior = fam.uniform_prior(mrs)
    lam = [0.0123]
    ids = [str(i) for i in range(F)]
    pv = rng.random(F)
    cp, mu = get_gamma(3, 0.7) if gamma else (None, None)
    rctx = ref.ctx(newick, tree.species, counts, mfs, mrs, prior)
    want = rctx.write_reconstruction(lam, pv, mu, cp)
    rctx.close()
    rec = oracle.reconstruct(tree, counts, mfs, mrs, prior, lam, mu, cp)
    states = rec["states"]
    got = [io_cpp.format_reconstruction(newick, ids, states, w, pvalues=pv, threshold=0.05, gamma_multipliers=mu)
           for w in ("count", "change", "asr", "family_results", "clade_results")]
    assert got[0] == want[0]
    assert got[1] == want[1]
    assert got[2] == want[2]
    assert got[3] == want[3]
    assert sorted(got[4].splitlines()) == sorted(want[4].splitlines())


def test_cpp_branch_probability_tables_reproduce_the_reference_text(ref):
    """_branch_probabilities.tab and the starred _asr.tre: the C++ writers fed with the REFERENCE's own compute_viterbi_sum values and
    reconstruction (the values themselves are checked on the GPU, tests/test_gpu_parity.py) against the reference's printers."""
    newick = "(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25)"
    tree = FlatTree(newick)
    rng = np.random.default_rng(4)
    F = 10
    base = rng.integers(1, 20, size=F)
    counts = np.clip(base[:, None] + rng.integers(-6, 7, size=(F, tree.n_leaves)), 0, 40).astype(np.int32)
    mfs, mrs = 60, 45
    pv = np.where(np.arange(F) % 3 == 0, 0.2, 0.01)                          # two thirds of the families are "significant"
    rctx = ref.ctx(newick, tree.species, counts, mfs, mrs, fam.uniform_prior(mrs))
    probs, tab_txt, asr_txt = rctx.branch_probabilities([0.0123], pv)
    states = rctx.reconstruct_base([0.0123])
    rctx.close()
    ids = [str(i) for i in range(F)]
    assert (probs[pv >= 0.05] == -1).all() and (probs[pv < 0.05][:, :-1] >= 0).all() and (probs[:, -1] == -1).all()
    assert io_cpp.format_reconstruction(newick, ids, states, "branch_probabilities", branch_probs=probs) == tab_txt
    assert io_cpp.format_reconstruction(newick, ids, states, "asr", branch_probs=probs, threshold=0.05) == asr_txt
    assert "*" in asr_txt


def test_cpp_report_reproduces_the_reference_text(ref):
    """<Model>_report.cafe (src/report.cpp): the C++ writer fed with the reference's own reconstruction, p-values and branch
    probabilities against the text the reference's Report streams, with and without a lambda tree, including the reference's missing
    line break after the 'ID' / 'Newick' header."""
    newick = "(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25)"
    lambda_newick = "(((A:1,B:1):1,(C:1,D:1):1):1,(E:2,(F:2,G:2):2):2)"
    rng = np.random.default_rng(5)
    F = 12
    mfs, mrs = 60, 45
    ids = [str(i) for i in range(F)]
    for lam_newick, lambdas in ((None, [0.0123]), (lambda_newick, [0.01, 0.03])):
        tree = FlatTree(newick, lam_newick) if lam_newick else FlatTree(newick)
        base = rng.integers(1, 20, size=F)
        counts = np.clip(base[:, None] + rng.integers(-6, 7, size=(F, tree.n_leaves)), 0, 40).astype(np.int32)
        pv = np.where(np.arange(F) % 3 == 0, 0.2, 0.0125)
        rctx = ref.ctx(newick, tree.species, counts, mfs, mrs, fam.uniform_prior(mrs), lambda_newick=lam_newick)
        want = rctx.write_report(lambdas, pv)
        probs, _, _ = rctx.branch_probabilities(lambdas, pv)
        states = rctx.reconstruct_base(lambdas)
        rctx.close()
        got = io_cpp.format_report(newick, ids, states, pv, lambdas=lambdas, lambda_newick=lam_newick, branch_probs=probs)
        assert got == want
        assert got.startswith("Tree:(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25):0\nLambda:\t")
        assert "'ID'\t'Newick'1\t" in got                  # family 1 is the first with branch probabilities; no line break before it
        assert got.count("\n") == 9 + int((pv < 0.05).sum())
    # no branch probabilities at all: the header block only
    head = io_cpp.format_report(newick, ids, states, pv, lambdas=[0.01, 0.03], lambda_newick=lambda_newick)
    assert head.endswith("'ID'\t'Newick'") and head == want[:len(head)]


def test_cpp_simulation_tables_reproduce_the_reference_text(ref):
    """simulation.txt and simulation_truth.txt (simulator::print_simulations, src/simulator.cpp:135-172): the C++ writer against the
    reference's own printer on the same node values, leaves only and with the interior columns."""
    newick = "(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25)"
    tree = FlatTree(newick)
    rng = np.random.default_rng(6)
    F = 9
    sizes = rng.integers(0, 60, size=(F, tree.n_nodes)).astype(np.int32)
    lambdas = 0.0123 * rng.choice([0.062015465425384449, 0.37328920830134099, 0.99805780528212318, 2.5666375209911516], size=F)
    counts = np.ones((1, tree.n_leaves), dtype=np.int32)
    rctx = ref.ctx(newick, tree.species, counts, 60, 45, fam.uniform_prior(45))
    for internal in (False, True):
        want = rctx.print_simulations(sizes, lambdas, internal)
        got = io_cpp.format_simulation(newick, sizes, lambdas, include_internal=internal)
        assert got == want
        assert got.startswith("DESC\tFID\t") and got.count("\n") == F + 1
        assert ("\t%d\t" % (tree.n_nodes - 1) in got.splitlines()[0] + "\t") == internal       # the root's column exists only in the truth table
    rctx.close()


def test_cpp_error_model_writer_reproduces_the_reference_text(ref):
    """The error model file written after epsilon has been estimated (write_error_model_file, src/io.cpp:277-297): one line per size
    whose probabilities differ from the previous size's; "maxcnt" is the number of table rows - 1 whatever the model's size limit."""
    for eps, rows, max_count in ((0.1, 91, 91), (0.0417, 3, 60), (0.25, 1, 10)):
        probs = np.tile([eps, 1 - 2 * eps, eps], (rows, 1))
        probs[0] = [0.0, 1 - eps, eps]
        want = ref.write_error_model(probs, max_count)
        got = io_cpp.format_error_model(probs)
        assert got == want
        assert got.startswith("maxcnt: %d\ncntdiff: -1 0 1\n0 0 " % (rows - 1))
    # the reference's own example file round-trips through its reader and both writers
    path = os.path.join(REF, "examples", "errormodel_0.1.txt")
    if os.path.exists(path):                                  # the reference's files are not present on the GPU box
        em = io_cpp.read_error_model(path)
        assert io_cpp.format_error_model(em[0]) == ref.write_error_model(em[0], em[1])


def test_cpp_report_on_a_multifurcating_tree(ref):
    """The report lists the children of every interior node: a node with three children gets a three-entry tuple in every block."""
    newick = "((A:1,B:1,C:1):2,(D:1.5,E:1.5):1.5)"
    tree = FlatTree(newick)
    rng = np.random.default_rng(8)
    F = 7
    counts = rng.integers(0, 12, size=(F, tree.n_leaves)).astype(np.int32)
    counts[:, 0] = np.maximum(counts[:, 0], 1)
    counts[:, 3] = np.maximum(counts[:, 3], 1)
    pv = np.full(F, 0.01)
    ids = [str(i) for i in range(F)]
    rctx = ref.ctx(newick, tree.species, counts, 40, 30, fam.uniform_prior(30))
    want = rctx.write_report([0.02], pv)
    probs, _, _ = rctx.branch_probabilities([0.02], pv)
    states = rctx.reconstruct_base([0.02])
    rctx.close()
    got = io_cpp.format_report(newick, ids, states, pv, lambdas=[0.02], branch_probs=probs)
    assert got == want
    assert "(1,2,3) " in got.splitlines()[4]
