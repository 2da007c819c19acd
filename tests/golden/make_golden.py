"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcafe_ref.so built from
/root/reference by oracle/build_ref.sh).  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so the inputs (count tables read from the reference's example data,
trees, error model) and the reference's outputs travel as these fixtures.  Every value stored under a
`ref_` key was produced by reference code, not by this repository's oracle or kernels.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cafe5_b200 import families as fam  # noqa: E402
from cafe5_b200.tree import FlatTree  # noqa: E402
from oracle.pyoracle import RefLib, build_ref  # noqa: E402

REF = os.environ.get("CAFE_REF_DIR", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_dataset(fam_path, tree_path, lambda_tree_path=None):
    newick = open(tree_path).readline().strip()
    lnewick = open(lambda_tree_path).readline().strip() if lambda_tree_path else ""
    species, ids, counts = fam.read_gene_families(fam_path)
    tree = FlatTree(newick, lnewick or None, species=species)
    keep = fam.exists_at_root(tree, counts)
    counts = counts[keep]
    mfs, mrs = fam.derive_sizes(counts)
    return newick, lnewick, species, counts, mfs, mrs, int(keep.size)


def spiky_rootdist(max_size, seed):
    """A root distribution whose neighbouring sizes differ by orders of magnitude, with holes: an off-by-one in the prior index
    (inference weights root size j+1 with compute(j), Pupko's root weights size j with compute(j)) changes every result."""
    rng = np.random.default_rng(seed)
    rd = {}
    for s in range(1, max_size + 1):
        if rng.random() < 0.15:
            continue                                   # a hole: prior 0 inside the table
        rd[s] = int(rng.choice([1, 3, 40, 1000, 25000]))
    rd[max_size] = 7
    return rd


def make_priors():
    """tests/golden/priors.npz: the reference's likelihoods and Pupko reconstructions under NON-UNIFORM priors -- a user root
    distribution (`-f`) whose table is shorter than max_root_family_size (so compute() returns 0 beyond it) and Poisson priors
    (`-p<lambda>`), tables built by the reference's own constructors (ref_prior_table)."""
    build_ref()
    ref = RefLib()
    out = {}
    small = np.load(os.path.join(OUT, "small.npz"))
    mfs, mrs, lam = int(small["max_family_size"]), int(small["max_root_family_size"]), float(small["lambda"])
    p3, m3 = ref.get_gamma(3, float(small["gamma_alpha"]))
    rd_small = spiky_rootdist(30, 1)
    priors = {"rootdist": ref.prior_table("rootdist", rootdist=rd_small)[0],
              "poisson": ref.prior_table("poisson", mrs, 6.5)[0]}
    out["small_rootdist_sizes"] = np.array(sorted(rd_small))
    out["small_rootdist_counts"] = np.array([rd_small[k] for k in sorted(rd_small)])
    out["small_poisson_lambda"] = 6.5
    for pname, prior in priors.items():
        assert len(prior) < mrs                                  # "0 beyond the table" must fire
        out["small_%s_prior" % pname] = prior
        for ti in range(4):
            nw = str(small["t%d_newick" % ti])
            tree = FlatTree(nw)
            counts = small["t%d_counts" % ti]
            c = ref.ctx(nw, tree.species, counts, mfs, mrs, prior.astype(np.float64))
            b = c.eval_base([lam])
            g = c.eval_gamma([lam], m3, p3)
            rg = c.reconstruct_gamma([lam], m3, p3)
            k = "small_%s_t%d_" % (pname, ti)
            out[k + "ref_base"] = b["neg_lnl"]
            out[k + "ref_family_lnl"] = b["family_lnl"]
            out[k + "ref_gamma"] = g["neg_lnl"]
            out[k + "ref_cat_lk"] = g["cat_lk"]
            out[k + "ref_failed"] = g["failed"]
            out[k + "ref_rec"] = c.reconstruct_base([lam])
            out[k + "ref_rec_gamma"] = rg["states"]
            out[k + "ref_rec_gamma_cat"] = rg["cat_states"]
            c.close()
    m = np.load(os.path.join(OUT, "mammals.npz"))
    species = [str(x) for x in m["species"]]
    counts = m["counts"].astype(np.int32)
    sub = np.arange(0, counts.shape[0], 37)
    mfs, mrs = int(m["max_family_size"]), int(m["max_root_family_size"])
    rd = spiky_rootdist(100, 2)
    priors = {"rootdist": ref.prior_table("rootdist", rootdist=rd)[0],
              "poisson08": ref.prior_table("poisson", mrs, 0.8)[0],
              "poisson12": ref.prior_table("poisson", mrs, 12.0)[0]}
    out["mammals_sub"] = sub
    out["mammals_rootdist_sizes"] = np.array(sorted(rd))
    out["mammals_rootdist_counts"] = np.array([rd[k] for k in sorted(rd)])
    for pname, prior in priors.items():
        assert len(prior) < mrs
        out["mammals_%s_prior" % pname] = prior
        c = ref.ctx(str(m["newick"]), species, counts[sub], mfs, mrs, prior.astype(np.float64))
        b = c.eval_base([0.0018])
        g = c.eval_gamma([0.0018], m["gamma_mult"], m["gamma_probs"])
        rg = c.reconstruct_gamma([0.0018], m["gamma_mult"], m["gamma_probs"])
        k = "mammals_%s_" % pname
        out[k + "ref_base"] = b["neg_lnl"]
        out[k + "ref_family_lnl"] = b["family_lnl"]
        out[k + "ref_gamma"] = g["neg_lnl"]
        out[k + "ref_cat_lk"] = g["cat_lk"]
        out[k + "ref_failed"] = g["failed"]
        out[k + "ref_rec"] = c.reconstruct_base([0.0018])
        out[k + "ref_rec_gamma"] = rg["states"]
        out[k + "ref_rec_gamma_cat"] = rg["cat_states"]
        c.close()
        print(pname, "table", len(prior), "base", b["neg_lnl"], "gamma", g["neg_lnl"])
    np.savez_compressed(os.path.join(OUT, "priors.npz"), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "priors":
        make_priors()
        return
    build_ref()
    ref = RefLib()
    ex = os.path.join(REF, "examples")

    # ---- config 1 / 2 / 3: mammals ------------------------------------------------------------
    newick, _, species, counts, mfs, mrs, n_before = load_dataset(os.path.join(ex, "mammal_gene_families.txt"),
                                                                  os.path.join(ex, "mammals_tree.txt"))
    prior = fam.uniform_prior(mrs)
    ctx = ref.ctx(newick, species, counts, mfs, mrs, prior)
    base = ctx.eval_base([0.0018])
    p4, m4 = ref.get_gamma(4, 0.7)
    g_ok = ctx.eval_gamma([0.0018], m4, p4)
    p4b, m4b = ref.get_gamma(4, 0.6)
    g_fail = ctx.eval_gamma([0.0018], m4b, p4b)
    sub = np.arange(0, counts.shape[0], 37)
    roots = np.stack([ctx.prune(int(f), [0.0018], 1.0) for f in sub[:64]])
    rec_sub = np.arange(0, counts.shape[0], 23)[:400]
    ctx_rec = ref.ctx(newick, species, counts[rec_sub], mfs, mrs, prior)
    rec_base = ctx_rec.reconstruct_base([0.0018])
    rec_gamma = ctx_rec.reconstruct_gamma([0.0018], m4, p4)
    # error model + lambda tree (config 3)
    lnewick = open(os.path.join(ex, "chimphuman_separate_lambda.txt")).readline().strip()
    em_probs, em_max = fam.read_error_model(os.path.join(ex, "errormodel_0.1.txt"))
    ctx3 = ref.ctx(newick, species, counts, mfs, mrs, prior, lambda_newick=lnewick, em=(em_probs, em_max))
    base3 = ctx3.eval_base([0.0018, 0.0042])
    np.savez_compressed(
        os.path.join(OUT, "mammals.npz"),
        newick=newick, lambda_newick=lnewick, species=np.array(species), counts=counts.astype(np.uint8),
        max_family_size=mfs, max_root_family_size=mrs, n_before_filter=n_before,
        em_probs=em_probs, em_maxcnt=em_max,
        ref_base_neg_lnl=base["neg_lnl"], ref_base_family_lnl=base["family_lnl"],
        gamma_alpha=0.7, gamma_mult=m4, gamma_probs=p4,
        ref_gamma_neg_lnl=g_ok["neg_lnl"], ref_gamma_cat_lk=g_ok["cat_lk"][sub], gamma_sub=sub,
        gamma_fail_alpha=0.6, gamma_fail_mult=m4b, gamma_fail_probs=p4b,
        ref_gamma_fail_neg_lnl=g_fail["neg_lnl"], ref_gamma_fail_failed=np.nonzero(g_fail["failed"])[0],
        roots_sub=sub[:64], ref_roots=roots,
        rec_sub=rec_sub, ref_rec_base=rec_base, ref_rec_gamma_states=rec_gamma["states"],
        ref_rec_gamma_cat_states=rec_gamma["cat_states"], ref_rec_gamma_avg=rec_gamma["averaged"],
        ref_em_neg_lnl=base3["neg_lnl"], ref_em_family_lnl=base3["family_lnl"], em_lambdas=np.array([0.0018, 0.0042]),
    )
    print("mammals: F=%d (of %d) S=%d R=%d base=%.8f gamma=%.8f em=%.8f" % (
        counts.shape[0], n_before, mfs + 1, mrs, base["neg_lnl"], g_ok["neg_lnl"], base3["neg_lnl"]))

    # ---- config 4: Hymenoptera, K = 8 ---------------------------------------------------------
    hd = os.path.join(REF, "Test_data", "Hymenoptera_Data")
    newick, _, species, counts, mfs, mrs, n_before = load_dataset(os.path.join(hd, "10Hymenoptera_genefamilies.tab"),
                                                                  os.path.join(hd, "10Hymenoptera.tree"))
    prior = fam.uniform_prior(mrs)
    ctx = ref.ctx(newick, species, counts, mfs, mrs, prior)
    base = ctx.eval_base([0.0018])
    p8, m8 = ref.get_gamma(8, 0.6)
    g8 = ctx.eval_gamma([0.0018], m8, p8)
    sub = np.arange(0, counts.shape[0], 41)
    np.savez_compressed(
        os.path.join(OUT, "hymenoptera.npz"),
        newick=newick, species=np.array(species), counts=counts.astype(np.uint8),
        max_family_size=mfs, max_root_family_size=mrs, n_before_filter=n_before,
        ref_base_neg_lnl=base["neg_lnl"], ref_base_family_lnl=base["family_lnl"][sub], base_sub=sub,
        gamma_alpha=0.6, gamma_mult=m8, gamma_probs=p8, ref_gamma_neg_lnl=g8["neg_lnl"],
        ref_gamma_cat_lk=g8["cat_lk"][sub], gamma_sub=sub, ref_gamma_failed=np.nonzero(g8["failed"])[0],
    )
    print("hymenoptera: F=%d (of %d) S=%d R=%d base=%.8f gamma8=%.8f" % (counts.shape[0], n_before, mfs + 1, mrs,
                                                                        base["neg_lnl"], g8["neg_lnl"]))

    # ---- matrices and scalar known answers ----------------------------------------------------
    mats = {}
    for name, (N, lam, t) in dict(m171_a=(171, 0.0018, 68.710507), m171_b=(171, 0.0018 * 2.5666375209911516, 96.435575),
                                  m141=(141, 0.006335, 68.7105), m5=(5, 0.05, 5.0), msat=(30, 0.02, 60.0),
                                  m201=(201, 0.0031, 12.3456)).items():
        mats[name + "_params"] = np.array([N, lam, t])
        mats["ref_" + name] = ref.matrix(N, lam, t)
    bd_in = np.array([[46, 45, -3.672556, 0.949177], [41, 34, -1.0986122886681098, 0.33333333333333337],
                      [10, 9, -1.5040773967762742, 0.5555555555555556], [170, 170, -2.2, 0.7784],
                      [1, 0, -0.5, 0.1], [1, 170, -3.0, 0.9], [120, 3, -4.1, 0.96]])
    bd_out = np.array([ref.birthdeath(int(r[0]), int(r[1]), r[2], r[3]) for r in bd_in])
    np.savez_compressed(os.path.join(OUT, "matrices.npz"), bd_in=bd_in, ref_bd=bd_out, **mats)

    # ---- small random problems: root vectors, likelihoods and Pupko states --------------------
    rng = np.random.default_rng(20261017)
    trees = ["((A:1,B:3):7,(C:11,D:17):23)", "(A:1,B:3):7", "((A:2.5,B:2.5,C:2.5):4,(D:3,(E:1,F:1):2):3.5)",
             "(((A:1.25,B:1.25):2,(C:2,D:2):1.25):3,(E:5,(F:0.5,G:0.5):4.5):1.25)"]
    small = {}
    for ti, nw in enumerate(trees):
        tree = FlatTree(nw)
        F = 96
        rootsz = rng.integers(1, 25, size=F)
        counts = np.clip(rootsz[:, None] + rng.integers(-6, 7, size=(F, tree.n_leaves)), 0, 40).astype(np.int32)
        counts[0] = 0
        counts[0, 0] = 1          # a nearly-extinct family
        counts[1] = 40            # everything at the table maximum
        mfs, mrs = 60, 45
        prior = fam.uniform_prior(mrs)
        c = ref.ctx(nw, tree.species, counts, mfs, mrs, prior)
        lam = [0.0123]
        b = c.eval_base(lam)
        p3, m3 = ref.get_gamma(3, 0.9)
        g = c.eval_gamma(lam, m3, p3)
        small["t%d_newick" % ti] = nw
        small["t%d_counts" % ti] = counts
        small["t%d_ref_roots" % ti] = np.stack([c.prune(f, lam, 1.0) for f in range(F)])
        small["t%d_ref_base" % ti] = b["neg_lnl"]
        small["t%d_ref_family_lnl" % ti] = b["family_lnl"]
        small["t%d_ref_gamma" % ti] = g["neg_lnl"]
        small["t%d_ref_cat_lk" % ti] = g["cat_lk"]
        small["t%d_ref_rec" % ti] = c.reconstruct_base(lam)
        rg = c.reconstruct_gamma(lam, m3, p3)
        small["t%d_ref_rec_gamma" % ti] = rg["states"]
        small["t%d_ref_rec_gamma_cat" % ti] = rg["cat_states"]
    small["lambda"] = 0.0123
    small["max_family_size"] = 60
    small["max_root_family_size"] = 45
    small["gamma_alpha"] = 0.9
    np.savez_compressed(os.path.join(OUT, "small.npz"), **small)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
