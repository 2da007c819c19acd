"""Species-tree handling for the host side of the hot path.

The C ABI (include/cafe_b200.h) takes a *flattened* tree: nodes listed in the reference's reverse
level order (reference src/clade.cpp:69-100: breadth-first order from the root, reversed), so that
children precede parents, the root is last, and the reference's descendant order of a node equals
decreasing node index.  This module parses newick (reference src/clade.cpp:293-419 behaviour: names,
``:length`` or ``:lambda-class`` annotations, interior labels kept, unlabelled interior nodes named by
the sorted concatenation of their leaf names, src/clade.cpp:161-173) and produces those arrays.
"""
import re

import numpy as np

_TOKEN = re.compile(r"\(|\)|[^\s\(\)\:\;\,]+|\:[+-]?[0-9]*\.?[0-9]+(?:[eE][+-]?[0-9]+)?|\,|\;")


class Node:
    __slots__ = ("parent", "children", "name", "value", "labelled")

    def __init__(self, parent=None):
        self.parent = parent
        self.children = []
        self.name = ""
        self.value = 0.0
        self.labelled = False

    def is_leaf(self):
        return not self.children

    def leaf_names(self):
        if not self.children:
            return [self.name]
        out = []
        for c in self.children:
            out.extend(c.leaf_names())
        return out


def parse_newick(text, lambda_tree=False):
    """Return the root Node.  ``value`` is the branch length, or the 1-based lambda class when
    ``lambda_tree`` (root defaults to class 1, reference src/clade.cpp:398-402)."""
    root = Node()
    cur = root
    for m in _TOKEN.finditer(text):
        tok = m.group(0)
        if tok == "(":
            child = Node(cur)
            cur.children.append(child)
            cur = child
        elif tok == ",":
            if cur is root:  # newick without the outer parentheses
                new_root = Node()
                root.parent = new_root
                new_root.children.append(root)
                root = new_root
            sib = Node(cur.parent)
            cur.parent.children.append(sib)
            cur = sib
        elif tok == ")":
            cur = cur.parent
        elif tok == ";":
            break
        elif tok[0] == ":":
            cur.value = int(tok[1:], 0) if lambda_tree else float(tok[1:])
        else:
            cur.name = tok
            cur.labelled = True
    if lambda_tree and root.value == 0:
        root.value = 1
    for n in bfs(root):
        if n.children and not n.labelled:
            n.name = "".join(sorted(n.leaf_names()))
        if lambda_tree:
            if n.value < 1:
                raise ValueError("Invalid lambda index set for " + n.name)
        elif n.parent is not None and n.value <= 0:
            raise ValueError("Invalid branch length set for " + n.name)
    return root


def bfs(root):
    order = [root]
    i = 0
    while i < len(order):
        order.extend(order[i].children)
        i += 1
    return order


class FlatTree:
    """Arrays handed to ``cafe_b200_create``; node i of every array is the i-th node of the reference's
    reverse level order."""

    def __init__(self, newick, lambda_newick=None, species=None):
        self.newick = newick
        self.lambda_newick = lambda_newick
        root = parse_newick(newick)
        order = list(reversed(bfs(root)))
        index = {id(n): i for i, n in enumerate(order)}
        self.names = [n.name for n in order]
        self.n_nodes = len(order)
        self.parent = np.array([-1 if n.parent is None else index[id(n.parent)] for n in order], dtype=np.int32)
        self.branch_length = np.array([n.value for n in order], dtype=np.float64)
        self.is_leaf = np.array([n.is_leaf() for n in order], dtype=bool)
        leaves = [n.name for n in order if n.is_leaf()]
        # column order of the count matrix: caller-supplied species list, else leaves in node order
        self.species = list(species) if species is not None else leaves
        lower = {s.lower(): j for j, s in enumerate(self.species)}  # reference lookup is case-insensitive (gene_family.h:10-33)
        self.leaf_col = np.full(self.n_nodes, -1, dtype=np.int32)
        for i, n in enumerate(order):
            if n.is_leaf():
                if n.name.lower() not in lower:
                    raise ValueError(n.name + " was not found in gene family")  # gene_family.cpp:40-42
                self.leaf_col[i] = lower[n.name.lower()]
        self.n_leaves = len(leaves)
        self.lambda_class = np.zeros(self.n_nodes, dtype=np.int32)
        self.n_lambda = 1
        if lambda_newick:
            lroot = parse_newick(lambda_newick, lambda_tree=True)
            lmap = {n.name: int(n.value) - 1 for n in bfs(lroot)}
            if set(lmap) != set(self.names):
                raise ValueError("The lambda tree structure does not match that of the tree")  # clade.cpp:247-262
            self.lambda_class = np.array([lmap[nm] for nm in self.names], dtype=np.int32)
            self.n_lambda = len(set(lmap.values()))
        self.root = root

    def branch_lengths(self):
        """Distinct positive branch lengths (reference clade::get_branch_lengths, src/clade.cpp:236-245)."""
        return sorted(set(float(b) for b in self.branch_length if b > 0))

    def children_of(self, i):
        """Children in the reference's descendant order (decreasing index)."""
        return [c for c in range(self.n_nodes - 1, -1, -1) if self.parent[c] == i]
