"""Host-side input handling for the hot path: gene-family count tables, derived state-space sizes,
the root filter, the error-model table and root priors.  Behaviour follows the reference (cited per
function); the numbers produced here are what crosses the C ABI.
"""
import math

import numpy as np


def read_gene_families(path):
    """Parse a CAFE / CAFExp family table (reference src/io.cpp:134-217).
    Returns (species, ids, counts[F, n_species] int32)."""
    species, ids, rows = None, [], []
    leaf_cols = {}  # CAFExp format: '#taxon' header lines
    header = True
    index = 0
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if not line:
                continue
            tok = line.split("\t")
            if leaf_cols and line[0] != "#":
                header = False
            if header:
                if line[0] == "#":
                    leaf_cols[index] = line[1:].rstrip("\r")
                    index += 1
                else:
                    header = False
                    if not leaf_cols:
                        species = [t.rstrip("\r") for t in tok[2:]]
                continue
            if leaf_cols:
                if species is None:
                    species = [leaf_cols[i] for i in sorted(leaf_cols)]
                rows.append([_atoi(tok[i]) for i in sorted(leaf_cols)])
                ids.append(tok[-1].rstrip("\r"))
            else:
                ids.append(tok[1])
                rows.append([_atoi(t) for t in tok[2:2 + len(species)]])
    if not rows:
        raise ValueError("No families found")
    return species, ids, np.asarray(rows, dtype=np.int32)


def _atoi(s):
    s = s.strip()
    n, i = 0, 0
    neg = False
    if i < len(s) and s[i] in "+-":
        neg = s[i] == "-"
        i += 1
    while i < len(s) and s[i].isdigit():
        n = n * 10 + ord(s[i]) - 48
        i += 1
    return -n if neg else n


def derive_sizes(counts, floor_family=120):
    """(max_family_size, max_root_family_size) as the reference derives them
    (src/user_data.h:26-27 floors, src/user_data.cpp:40-48)."""
    m = max(int(floor_family), int(np.max(counts)) if counts.size else 0)
    max_root = max(30, int(np.rint(m * 1.25)))
    max_family = m + max(50, m // 5)
    return max_family, max_root


def exists_at_root(tree, counts):
    """Boolean mask of families kept by the reference's root filter: every child of the root must have a
    descendant leaf with a positive count (src/gene_family.cpp:62-91, src/core.cpp:190-204)."""
    counts = np.asarray(counts)
    F = counts.shape[0]
    present = np.zeros((tree.n_nodes, F), dtype=bool)
    for i in range(tree.n_nodes):  # children precede parents
        if tree.leaf_col[i] >= 0:
            present[i] = counts[:, tree.leaf_col[i]] > 0
        p = tree.parent[i]
        if p >= 0:
            present[p] |= present[i]
    root = tree.n_nodes - 1
    keep = np.ones(F, dtype=bool)
    for c in range(tree.n_nodes):
        if tree.parent[c] == root:
            keep &= present[c]
    return keep


def uniform_prior(max_root_family_size):
    """Uniform prior table: sizes 0..R-1 each float(1)/float(R) (src/root_equilibrium_distribution.cpp:30-35,
    71-79).  Returned as float32 because the reference's compute() returns float (.h:40)."""
    R = int(max_root_family_size)
    return np.full(R, np.float32(1.0) / np.float32(R), dtype=np.float32)


def rootdist_prior(rootdist):
    """Prior from a {size: count} map (src/root_equilibrium_distribution.cpp:13-28, 71-79)."""
    total = sum(rootdist.values())
    mx = max(rootdist) + 1
    out = np.zeros(mx, dtype=np.float32)
    for s, c in rootdist.items():
        out[s] = np.float32(c) / np.float32(total)
    return out


def poisson_prior(poisson_lambda, num_values):
    """Poisson prior table (`-p<lambda>`; src/root_equilibrium_distribution.cpp:56-68, poisspdf src/poisson.cpp:21-24): entry i is
    pdf(i) for every i the reference visits while it fills `num_values` simulated roots (each pdf(i) contributes
    ceil(pdf(i) * num_values) of them), plus five more.  NB the reference weights root size j+1 with entry j in inference
    (base_model.cpp:84, gamma_core.cpp:156) and root size j with entry j in Pupko (gene_family_reconstructor.cpp:65): the table is
    uploaded as it is, the conventions live in the kernels.  float32 because compute() returns float."""
    def pdf(x):
        return math.exp(x * math.log(poisson_lambda) - math.lgamma(x + 1) - poisson_lambda)
    table, filled, i = [], 0, 0
    while filled < num_values:
        pct = pdf(i)
        j = 0
        while j < pct * num_values:       # `for (size_t j = 0; j < pct * num_values; ++j)`
            j += 1
        filled += j
        table.append(pct)
        i += 1
    for _ in range(5):
        table.append(pdf(len(table)))
    return np.asarray(table, dtype=np.float64).astype(np.float32)


def read_error_model(path):
    """Parse an error-model file (src/io.cpp:228-274, src/error_model.cpp:31-50): returns
    (probs[rows, 3] float64, maxcnt).  Missing sizes inherit the previous row."""
    maxcnt = 0
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith("max"):
                maxcnt = int(line.split(":")[1].strip())
            elif line.startswith("cnt"):
                devs = [int(t) for t in line.split()[1:]]
                if devs != [-1, 0, 1]:
                    raise ValueError("only the -1 0 1 error classes are supported on the device path")
            else:
                tok = line.split()
                sz = int(tok[0])
                vals = [float(t) for t in tok[1:]]
                if (sz == 0 or not rows) and abs(vals[0]) > 0.0:
                    raise ValueError("Cannot have a non-zero probability for family size 0 for negative deviation")
                if abs(sum(vals) - 1.0) > 0.01 * abs(sum(vals)):
                    raise ValueError("Sum of probabilities must be equal to one")
                if not rows:
                    rows.append(vals)
                while len(rows) <= sz:
                    rows.append(list(rows[-1]))
                rows[sz] = vals
    return np.asarray(rows, dtype=np.float64), maxcnt


def epsilon_error_model(epsilon, max_family_size):
    """The table the reference builds when epsilon is a free parameter: row 0 = {0, 1-eps, eps}, other rows
    {eps, 1-2eps, eps} (src/core.cpp:39-45 default + src/error_model.cpp:79-109 replace_epsilons)."""
    rows = np.empty((max_family_size + 1, 3), dtype=np.float64)
    rows[:] = [epsilon, 1 - (epsilon * 2), epsilon]
    rows[0] = [0.0, 1 - epsilon, epsilon]
    return rows, max_family_size
