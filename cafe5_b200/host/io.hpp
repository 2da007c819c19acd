// io.hpp -- the reference's on-disk formats on the host side: species tree and lambda tree (newick),
// gene-family tables (both header styles), error-model files, root distributions, and the text tables that expose hot-path
// results (<Model>_results.txt, *_family_likelihoods.txt, Gamma_category_likelihoods.txt).
// Written from the behaviour of src/io.cpp:107-297, src/clade.cpp:69-100,161-173,293-419, src/error_model.cpp:31-50,
// src/user_data.cpp:40-48,105-117, src/gene_family.cpp:62-91, src/core.cpp:53-58,97-112,152-174, src/base_model.cpp:102-109,
// src/gamma_core.cpp:46-58,359-381 and src/lambda.cpp:25-58 -- no reference code.  Numbers are printed through std::ostream with
// the reference's manipulators, so the text is identical whenever the values are.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iomanip>
#include <istream>
#include <map>
#include <memory>
#include <ostream>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cafe_b200_host {

// ---- trees ------------------------------------------------------------------------------------------------------------------
struct Tree {
    // nodes in the reference's reverse level order (breadth-first from the root, reversed): children precede parents, the root is
    // last, a node's descendants in the reference's order are its children by DEcreasing index
    std::vector<int32_t> parent;
    std::vector<double> branch_length;
    std::vector<std::string> name;        // taxon name, interior label, or the sorted concatenation of the leaf names below
    std::vector<uint8_t> is_leaf;
    std::vector<int32_t> lambda_class;    // 0-based; all 0 without a lambda tree
    int n_lambda = 1;
    int n_nodes() const { return (int)parent.size(); }
    std::vector<std::string> leaf_names() const
    {
        std::vector<std::string> out;
        for (int i = 0; i < n_nodes(); ++i) if (is_leaf[i]) out.push_back(name[i]);
        return out;
    }
};

namespace detail {
struct Node {
    Node* parent = nullptr;
    std::vector<Node*> children;
    std::string name;
    double value = 0.0;
    bool labelled = false;
};

inline void leaf_names_below(const Node* n, std::vector<std::string>& out)
{
    if (n->children.empty()) { out.push_back(n->name); return; }
    for (const Node* c : n->children) leaf_names_below(c, out);
}

// tokens: ( ) , ; :number name   (the reference's regex, src/clade.cpp:312)
inline std::vector<std::string> tokenize(const std::string& s)
{
    std::vector<std::string> out;
    size_t i = 0;
    const auto stop = [](char ch) { return std::isspace((unsigned char)ch) || ch == '(' || ch == ')' || ch == ':' || ch == ';' || ch == ','; };
    while (i < s.size()) {
        const char ch = s[i];
        if (std::isspace((unsigned char)ch)) { ++i; continue; }
        if (ch == '(' || ch == ')' || ch == ',' || ch == ';') { out.emplace_back(1, ch); ++i; continue; }
        size_t j = i + 1;
        while (j < s.size() && !stop(s[j])) ++j;
        out.push_back(s.substr(i, j - i));   // ":12.5" or a name
        i = j;
    }
    return out;
}

struct Parsed {
    std::vector<std::unique_ptr<Node>> pool;
    Node* root = nullptr;
    std::vector<Node*> level_order;   // breadth first from the root
};

inline Parsed parse(const std::string& text, bool lambda_tree)
{
    Parsed p;
    auto make = [&](Node* parent) { p.pool.emplace_back(new Node()); p.pool.back()->parent = parent; return p.pool.back().get(); };
    Node* root = make(nullptr);
    Node* cur = root;
    for (const std::string& tok : tokenize(text)) {
        if (tok == "(") { Node* c = make(cur); cur->children.push_back(c); cur = c; }
        else if (tok == ",") {
            if (cur == root) {   // newick without the outer parentheses
                Node* nr = make(nullptr);
                root->parent = nr;
                nr->children.push_back(root);
                root = nr;
            }
            Node* sib = make(cur->parent);
            cur->parent->children.push_back(sib);
            cur = sib;
        }
        else if (tok == ")") cur = cur->parent;
        else if (tok == ";") break;
        else if (tok[0] == ':') cur->value = lambda_tree ? (double)std::strtol(tok.c_str() + 1, nullptr, 0) : std::strtod(tok.c_str() + 1, nullptr);
        else { cur->name = tok; cur->labelled = true; }
    }
    if (lambda_tree && root->value == 0) root->value = 1;   // src/clade.cpp:398-402
    p.root = root;
    p.level_order.push_back(root);
    for (size_t i = 0; i < p.level_order.size(); ++i)
        for (Node* c : p.level_order[i]->children) p.level_order.push_back(c);
    for (Node* n : p.level_order) {
        if (!n->children.empty() && !n->labelled) {   // src/clade.cpp:161-173
            std::vector<std::string> names;
            leaf_names_below(n, names);
            std::sort(names.begin(), names.end());
            n->name.clear();
            for (const auto& s : names) n->name += s;
        }
        if (lambda_tree) { if (n->value < 1) throw std::runtime_error("Invalid lambda index set for " + n->name); }
        else if (n->parent != nullptr && n->value <= 0) throw std::runtime_error("Invalid branch length set for " + n->name);   // :411-414
    }
    return p;
}
}  // namespace detail

inline Tree parse_newick(const std::string& newick, const std::string& lambda_newick = std::string())
{
    detail::Parsed p = detail::parse(newick, false);
    const int n = (int)p.level_order.size();
    std::map<const detail::Node*, int> index;
    for (int i = 0; i < n; ++i) index[p.level_order[n - 1 - i]] = i;
    Tree t;
    t.parent.resize(n); t.branch_length.resize(n); t.name.resize(n); t.is_leaf.resize(n); t.lambda_class.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        const detail::Node* nd = p.level_order[n - 1 - i];
        t.parent[i] = nd->parent ? index.at(nd->parent) : -1;
        t.branch_length[i] = nd->value;
        t.name[i] = nd->name;
        t.is_leaf[i] = nd->children.empty();
    }
    if (!lambda_newick.empty()) {
        detail::Parsed lp = detail::parse(lambda_newick, true);
        std::map<std::string, int> cls;
        for (const detail::Node* nd : lp.level_order) cls[nd->name] = (int)nd->value - 1;
        std::set<std::string> a(t.name.begin(), t.name.end()), b;
        for (const auto& kv : cls) b.insert(kv.first);
        if (a != b) throw std::runtime_error("The lambda tree structure does not match that of the tree");   // src/clade.cpp:247-262
        std::set<int> distinct;
        for (int i = 0; i < n; ++i) { t.lambda_class[i] = cls.at(t.name[i]); distinct.insert(t.lambda_class[i]); }
        t.n_lambda = (int)distinct.size();
    }
    return t;
}

// ---- gene families ------------------------------------------------------------------------------------------------------------
struct FamilyTable {
    std::vector<std::string> species, ids;
    std::vector<int32_t> counts;          // [ids.size() x species.size()] row-major
    size_t n_families() const { return ids.size(); }
};

inline std::vector<std::string> split_tabs(const std::string& line)
{
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t b = line.find('\t', a);
        out.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}

inline std::string strip_cr(std::string s) { while (!s.empty() && s.back() == '\r') s.pop_back(); return s; }

// CAFE format: "Desc<TAB>Family ID<TAB>sp1..." header, rows "desc<TAB>id<TAB>c1..."; CAFExp format: "#taxon" header lines, rows of
// counts with the id in the last column (src/io.cpp:134-217).  Counts go through atoi, blank lines are skipped.
// tree (optional): in the CAFExp format the reference looks every "#name" up in the tree, advances the column index for each of
// them but records a column only for LEAVES (src/io.cpp:153-161) -- interior-node columns are skipped -- and rejects a name that is
// not in the tree.  Without a tree every "#name" line is taken as a species column.
inline FamilyTable read_gene_families(std::istream& in, const Tree* tree = nullptr)
{
    FamilyTable ft;
    std::map<int, std::string> leaf_cols;
    bool header = true;
    int index = 0;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        const std::vector<std::string> tok = split_tabs(line);
        if (!leaf_cols.empty() && line[0] != '#') header = false;
        if (header) {
            if (line[0] == '#') {
                const std::string taxon = strip_cr(line.substr(1));
                bool keep = true;
                if (tree) {
                    int found = -1;
                    for (int i = 0; i < tree->n_nodes() && found < 0; ++i) if (tree->name[i] == taxon) found = i;
                    if (found < 0) throw std::runtime_error(taxon + " not located in tree");
                    keep = tree->is_leaf[found] != 0;
                }
                if (keep) leaf_cols[index] = taxon;
                index++;
            } else {
                header = false;
                if (leaf_cols.empty())
                    for (size_t i = 2; i < tok.size(); ++i) ft.species.push_back(strip_cr(tok[i]));
            }
            continue;
        }
        if (!leaf_cols.empty()) {
            if (ft.species.empty()) for (const auto& kv : leaf_cols) ft.species.push_back(kv.second);
            for (const auto& kv : leaf_cols) ft.counts.push_back(kv.first < (int)tok.size() ? std::atoi(tok[kv.first].c_str()) : 0);
            ft.ids.push_back(strip_cr(tok.back()));
        } else {
            ft.ids.push_back(tok.size() > 1 ? tok[1] : std::string());
            for (size_t i = 0; i < ft.species.size(); ++i) ft.counts.push_back(2 + i < tok.size() ? std::atoi(tok[2 + i].c_str()) : 0);
        }
    }
    if (ft.ids.empty()) throw std::runtime_error("No families found");
    return ft;
}

// max_family_size / max_root_family_size from the largest count (src/user_data.h:26-27 floors, src/user_data.cpp:40-48)
inline void derive_sizes(const std::vector<int32_t>& counts, int& max_family_size, int& max_root_family_size, int floor_family = 120)
{
    int m = floor_family;
    for (int32_t v : counts) m = std::max(m, (int)v);
    max_root_family_size = std::max(30, (int)std::rint(m * 1.25));
    max_family_size = m + std::max(50, m / 5);
}

// the reference's root filter: every child of the root has a descendant leaf with a positive count (src/gene_family.cpp:62-91)
inline std::vector<uint8_t> exists_at_root(const Tree& t, const std::vector<int32_t>& leaf_col, const FamilyTable& ft)
{
    const int n = t.n_nodes(), root = n - 1;
    const size_t S = ft.species.size();
    std::vector<uint8_t> keep(ft.n_families(), 1), present(n);
    for (size_t f = 0; f < ft.n_families(); ++f) {
        std::fill(present.begin(), present.end(), 0);
        for (int i = 0; i < n; ++i) {
            if (t.is_leaf[i]) present[i] = ft.counts[f * S + leaf_col[i]] > 0;
            if (t.parent[i] >= 0 && present[i]) present[t.parent[i]] = 1;
        }
        for (int c = 0; c < n; ++c) if (t.parent[c] == root && !present[c]) keep[f] = 0;
    }
    return keep;
}

// column of the family table for every leaf; the reference's species lookup is case-insensitive (src/gene_family.h:10-33)
inline std::vector<int32_t> leaf_columns(const Tree& t, const std::vector<std::string>& species)
{
    auto lower = [](std::string s) { for (char& ch : s) ch = (char)std::tolower((unsigned char)ch); return s; };
    std::map<std::string, int> col;
    for (size_t j = 0; j < species.size(); ++j) col[lower(species[j])] = (int)j;
    std::vector<int32_t> out(t.n_nodes(), -1);
    for (int i = 0; i < t.n_nodes(); ++i)
        if (t.is_leaf[i]) {
            auto it = col.find(lower(t.name[i]));
            if (it == col.end()) throw std::runtime_error(t.name[i] + " was not found in gene family");   // src/gene_family.cpp:40-42
            out[i] = it->second;
        }
    return out;
}

// ---- error model, root distribution ----------------------------------------------------------------------------------------------
struct ErrorModelTable {
    std::vector<double> probs;   // [rows x 3]: P(deviation -1, 0, +1 | size)
    int rows() const { return (int)probs.size() / 3; }
    int max_count = 0;
};

// "maxcnt:<n>", "cntdiff -1 0 1", then "size p-1 p0 p+1"; missing sizes inherit the previous row (src/io.cpp:228-274,
// src/error_model.cpp:31-50)
inline ErrorModelTable read_error_model(std::istream& in)
{
    ErrorModelTable em;
    std::string line;
    auto nearly = [](double x, double y) { return std::fabs(x - y) <= 0.01 * std::fabs(x); };
    while (std::getline(in, line)) {
        line = strip_cr(line);
        size_t a = line.find_first_not_of(" \t");
        if (a == std::string::npos) continue;
        line = line.substr(a);
        if (line.compare(0, 3, "max") == 0) em.max_count = std::atoi(line.substr(line.find(':') + 1).c_str());
        else if (line.compare(0, 3, "cnt") == 0) {
            std::istringstream ss(line.substr(line.find_first_of(" \t")));
            std::vector<int> dev;
            int d;
            while (ss >> d) dev.push_back(d);
            if (dev != std::vector<int>{-1, 0, 1}) throw std::runtime_error("only the -1 0 1 error classes are supported");
        } else {
            std::istringstream ss(line);
            int sz;
            double v[3];
            if (!(ss >> sz >> v[0] >> v[1] >> v[2])) continue;
            if ((sz == 0 || em.probs.empty()) && !nearly(v[0], 0.0)) throw std::runtime_error("Cannot have a non-zero probability for family size 0 for negative deviation");
            if (!nearly(v[0] + v[1] + v[2], 1.0)) throw std::runtime_error("Sum of probabilities must be equal to one");
            if (em.probs.empty()) em.probs.assign(v, v + 3);
            while (em.rows() <= sz) em.probs.insert(em.probs.end(), em.probs.end() - 3, em.probs.end());
            std::copy(v, v + 3, em.probs.begin() + (size_t)sz * 3);
        }
    }
    return em;
}

// "size<ws>count" per line (src/user_data.cpp:105-117)
inline std::map<int, int> read_rootdist(std::istream& in)
{
    std::map<int, int> out;
    int s, c;
    while (in >> s >> c) out[s] = c;
    return out;
}

// ---- root priors (root_equilibrium_distribution, src/root_equilibrium_distribution.cpp:13-87; chosen by user_data::create_prior,
// src/user_data.cpp:176-206).  A table holds compute(j) for j = 0 .. size-1 as the FLOAT compute() returns (.h:40); beyond it the
// prior is 0.  The index conventions stay in the kernels: inference weights root size j+1 with entry j (base_model.cpp:84,
// gamma_core.cpp:156), Pupko's root weights size j with entry j (gene_family_reconstructor.cpp:65). ----
// default: uniform over sizes 0 .. max_root_family_size-1, each float(1)/float(n) (:30-35, 71-79)
inline std::vector<float> uniform_prior(int max_root_family_size)
{
    return std::vector<float>((size_t)max_root_family_size, float(1) / float(max_root_family_size));
}

// `-f`: a {size: count} root distribution; entry i = float(count_i) / float(total) (:13-28, 71-79)
inline std::vector<float> rootdist_prior(const std::map<int, int>& rootdist)
{
    if (rootdist.empty()) throw std::runtime_error("No root distribution specified");
    size_t total = 0;
    for (const auto& kv : rootdist) total += (size_t)std::max(kv.second, 0);
    std::vector<float> out((size_t)rootdist.rbegin()->first + 1, 0.0f);
    for (const auto& kv : rootdist)
        if (kv.first >= 0 && kv.second > 0) out[(size_t)kv.first] = float((size_t)kv.second) / float(total);
    return out;
}

// `-p<lambda>`: entry i = poisspdf(i, lambda) = exp(i log(lambda) - lgamma(i + 1) - lambda) (src/poisson.cpp:21-24) for every i the
// reference visits while it fills num_values simulated roots (pdf(i) contributes the smallest integer count >= pdf(i) * num_values),
// and five more (:56-68)
inline std::vector<float> poisson_prior(double poisson_lambda, size_t num_values)
{
    auto pdf = [poisson_lambda](size_t x) { return std::exp(double(x) * std::log(poisson_lambda) - std::lgamma(double(x) + 1) - poisson_lambda); };
    std::vector<float> out;
    size_t filled = 0;
    for (size_t i = 0; filled < num_values; ++i) {
        const double pct = pdf(i);
        size_t j = 0;
        while (double(j) < pct * double(num_values)) ++j;
        filled += j;
        out.push_back(float(pct));
        if (i > 100000) throw std::runtime_error("Poisson prior does not fill the requested number of values");
    }
    for (int i = 0; i < 5; ++i) out.push_back(float(pdf(out.size())));
    return out;
}

// ---- result tables -----------------------------------------------------------------------------------------------------------------
// lambda::to_string (src/lambda.cpp:25-30, 47-58): setw(15) << setprecision(14) applied to the first value only by setw
inline std::string lambdas_to_string(const std::vector<double>& lambdas)
{
    std::ostringstream ost;
    ost << std::setw(15) << std::setprecision(14);
    for (size_t i = 0; i < lambdas.size(); ++i) {
        ost << lambdas[i];
        if (i + 1 != lambdas.size()) ost << ", ";
    }
    return ost.str();
}

struct MonitorCounts {     // event_monitor (src/core.h:108-119)
    int attempts = 0, rejects = 0;
    std::map<std::string, int> failure_count;
};

// <Model>_results.txt: model::write_vital_statistics (src/core.cpp:97-112), gamma adds "Alpha:" (src/gamma_core.cpp:46-50),
// event_monitor::log (src/core.cpp:152-174).  epsilon / alpha: pass NaN when absent.
inline void write_vital_statistics(std::ostream& ost, const std::string& model_name, double neg_lnl, const std::vector<double>& lambdas,
                                   double epsilon, double longest_branch, const MonitorCounts& mon, double alpha)
{
    ost << "Model " << model_name << " Final Likelihood (-lnL): " << neg_lnl << std::endl;
    ost << "Lambda: " << lambdas_to_string(lambdas) << std::endl;
    if (!std::isnan(epsilon)) ost << "Epsilon: " << epsilon << std::endl;
    ost << "Maximum possible lambda for this topology: " << 1 / longest_branch << std::endl;
    if (mon.attempts == 0) ost << "No attempts made\n";
    else {
        ost << mon.attempts << " values were attempted (" << std::round(double(mon.rejects) / double(mon.attempts) * 100) << "% rejected)\n";
        if (!mon.failure_count.empty()) {
            int worst = 0;
            for (const auto& kv : mon.failure_count) worst = std::max(worst, kv.second);
            if (worst * 5 > (mon.attempts - mon.rejects)) {
                ost << "The following families had failure rates >20% of the time:\n";
                for (const auto& kv : mon.failure_count)
                    if (kv.second * 5 > (mon.attempts - mon.rejects)) ost << kv.first << " had " << kv.second << " failures\n";
            }
        }
    }
    if (!std::isnan(alpha)) ost << "Alpha: " << alpha << std::endl;
}

// Base_family_likelihoods.txt (src/base_model.cpp:102-109)
inline void write_base_family_likelihoods(std::ostream& ost, const std::vector<std::string>& ids, const double* family_lnl)
{
    ost << "#FamilyID\tLikelihood of Family" << std::endl;
    for (size_t f = 0; f < ids.size(); ++f) ost << ids[f] << "\t" << family_lnl[f] << std::endl;
}

// Gamma_family_likelihoods.txt (src/gamma_core.cpp:52-58, src/core.cpp:53-58): K rows per family
inline void write_gamma_family_likelihoods(std::ostream& ost, const std::vector<std::string>& ids, int K, const double* multipliers,
                                           const double* cat_lk, const double* family_lk, const double* posterior, const uint8_t* significant)
{
    ost << "#FamilyID\tGamma Cat Mean\tLikelihood of Category\tLikelihood of Family\tPosterior Probability\tSignificant" << std::endl;
    for (size_t f = 0; f < ids.size(); ++f)
        for (int k = 0; k < K; ++k)
            ost << ids[f] << "\t" << multipliers[k] << "\t" << cat_lk[f * K + k] << "\t" << family_lk[f] << "\t" << posterior[f * K + k]
                << "\t" << (significant[f * K + k] ? "*" : "N/S") << "\n";
}

// Gamma_category_likelihoods.txt (src/gamma_core.cpp:359-374)
inline void write_category_likelihoods(std::ostream& ost, const std::vector<std::string>& ids, int K, const double* multipliers, const double* cat_lk)
{
    ost << "Family ID\t";
    for (int k = 0; k < K; ++k) ost << multipliers[k] << "\t";
    ost << std::endl;
    for (size_t f = 0; f < ids.size(); ++f) {
        ost << ids[f] << '\t';
        for (int k = 0; k < K; ++k) ost << cat_lk[f * K + k] << "\t";
        ost << std::endl;
    }
}

// ---- reconstruction tables (reconstruction::write_results, src/gene_family_reconstructor.cpp:352-379) ------------------------------
// The reference labels nodes with the numbering R's ape package gives the tree (get_ape_order, src/newick_ape_loader.cpp:213-242):
// tips 1..n in their left-to-right order in the newick text, the root n+1, the other interior nodes n+2.. in preorder.
// Returns id[node] for the Tree's node order; the children of a node, left to right, are its children by DEcreasing index.
inline std::vector<int> ape_ids(const Tree& t)
{
    const int n = t.n_nodes();
    std::vector<std::vector<int>> kids(n);
    for (int i = n - 1; i >= 0; --i) if (t.parent[i] >= 0) kids[t.parent[i]].push_back(i);
    int n_tips = 0;
    for (int i = 0; i < n; ++i) n_tips += t.is_leaf[i];
    std::vector<int> id(n, 0), stack{n - 1};
    int next_tip = 1, next_inner = n_tips + 1;
    while (!stack.empty()) {
        const int v = stack.back();
        stack.pop_back();
        id[v] = t.is_leaf[v] ? next_tip++ : next_inner++;
        for (auto it = kids[v].rbegin(); it != kids[v].rend(); ++it) stack.push_back(*it);   // leftmost child on top
    }
    return id;
}

// clade_index_or_name (src/clade.cpp:225-234)
inline std::string node_label(const Tree& t, const std::vector<int>& ape, int node)
{
    return (t.is_leaf[node] ? t.name[node] : std::string()) + "<" + std::to_string(ape[node]) + ">";
}

inline std::vector<int> nodes_in_ape_order(const std::vector<int>& ape)
{
    std::vector<int> order(ape.size());
    for (size_t i = 0; i < ape.size(); ++i) order[ape[i] - 1] = (int)i;
    return order;
}

// <Model>_count.tab / <Model>_change.tab (print_node_counts :341-350, print_node_change :203-210, print_family_clade_table :270-277).
// states[F x n_nodes]: reconstructed counts in the Tree's node order, leaves = observed counts.
inline void write_node_table(std::ostream& ost, const Tree& t, const std::vector<std::string>& ids, const int32_t* states, bool change)
{
    const std::vector<int> ape = ape_ids(t), order = nodes_in_ape_order(ape);
    const int n = t.n_nodes();
    ost << "FamilyID";
    for (int v : order) ost << "\t" << node_label(t, ape, v);
    ost << std::endl;
    for (size_t f = 0; f < ids.size(); ++f) {
        ost << ids[f];
        for (int v : order) {
            const int here = states[f * n + v];
            ost << "\t" << (change ? (t.parent[v] < 0 ? 0 : here - states[f * n + t.parent[v]]) : here);
        }
        ost << std::endl;
    }
}

// <Model>_asr.tre (print_reconstructed_states :302-339, newick_node :192-201, clade::write_newick src/clade.cpp:206-223); no branch
// probabilities (no node is starred); the gamma model appends its multipliers (write_nexus_extensions, src/gamma_core.cpp:341-349)
// branch_probs[F x n_nodes] (cafe_b200_branch_probabilities; -1 = none): a family "has" branch probabilities when any entry is >= 0,
// and a branch is starred when its probability is below the threshold (is_significant, :315-318)
inline void write_asr_trees(std::ostream& ost, const Tree& t, const std::vector<std::string>& ids, const int32_t* states,
                            const std::vector<double>& gamma_multipliers = std::vector<double>(), const double* branch_probs = nullptr,
                            double threshold = 0.05)
{
    const std::vector<int> ape = ape_ids(t);
    const int n = t.n_nodes();
    std::vector<std::vector<int>> kids(n);
    for (int i = n - 1; i >= 0; --i) if (t.parent[i] >= 0) kids[t.parent[i]].push_back(i);
    ost << "#nexus\nBEGIN TREES;\n";
    for (size_t f = 0; f < ids.size(); ++f) {
        const int32_t* st = states + f * n;
        std::function<void(int)> emit = [&](int v) {
            if (!kids[v].empty()) {
                ost << '(';
                for (size_t i = 0; i < kids[v].size(); ++i) { if (i) ost << ','; emit(kids[v][i]); }
                ost << ')';
            }
            std::ostringstream node;
            const bool star = branch_probs != nullptr && branch_probs[f * n + v] >= 0 && branch_probs[f * n + v] < threshold;
            node << node_label(t, ape, v) << (star ? "*" : "") << "_" << st[v];
            if (t.parent[v] >= 0) node << ':' << t.branch_length[v];
            ost << node.str();
        };
        ost << "  TREE " << ids[f] << " = ";
        emit(n - 1);
        ost << ';' << std::endl;
    }
    ost << "\nEND;\n";
    if (!gamma_multipliers.empty()) {
        ost << "\nBEGIN LAMBDA_MULTIPLIERS;\n";
        for (double m : gamma_multipliers) ost << "  " << m << ";\n";
        ost << "END;\n\n";
    }
}

// <Model>_branch_probabilities.tab (print_branch_probabilities :279-300): one row per family that has branch probabilities
inline void write_branch_probabilities(std::ostream& ost, const Tree& t, const std::vector<std::string>& ids, const double* branch_probs)
{
    const std::vector<int> ape = ape_ids(t), order = nodes_in_ape_order(ape);
    const int n = t.n_nodes();
    ost << "FamilyID";
    for (int v : order) ost << "\t" << node_label(t, ape, v);
    ost << std::endl;
    for (size_t f = 0; f < ids.size(); ++f) {
        bool any = false;
        for (int v = 0; v < n; ++v) any = any || branch_probs[f * n + v] >= 0;
        if (!any) continue;
        ost << ids[f];
        for (int v : order) {
            if (branch_probs[f * n + v] >= 0) ost << "\t" << branch_probs[f * n + v];
            else ost << "\tN/A";
        }
        ost << std::endl;
    }
}

// <Model>_family_results.txt (print_increases_decreases_by_family :213-232)
inline void write_family_results(std::ostream& ost, const std::vector<std::string>& ids, const double* pvalues, double threshold)
{
    if (ids.empty()) { ost << "No increases or decreases recorded\n"; return; }
    ost << "#FamilyID\tpvalue\tSignificant at " << threshold << "\n";
    for (size_t f = 0; f < ids.size(); ++f) ost << ids[f] << '\t' << pvalues[f] << '\t' << (pvalues[f] < threshold ? 'y' : 'n') << std::endl;
}

// <Model>_clade_results.txt (print_increases_decreases_by_clade :234-256): nodes that changed in at least one family.  The reference
// lists them in the order of their addresses in memory; here they come in ape order (compare as a set of lines).
inline void write_clade_results(std::ostream& ost, const Tree& t, size_t n_families, const int32_t* states)
{
    const std::vector<int> ape = ape_ids(t), order = nodes_in_ape_order(ape);
    const int n = t.n_nodes();
    ost << "#Taxon_ID\tIncrease\tDecrease\n";
    for (int v : order) {
        if (t.parent[v] < 0) continue;
        int up = 0, down = 0;
        for (size_t f = 0; f < n_families; ++f) {
            const int d = states[f * n + v] - states[f * n + t.parent[v]];
            up += d > 0;
            down += d < 0;
        }
        if (up || down) ost << node_label(t, ape, v) << "\t" << up << "\t" << down << std::endl;
    }
}

// <Model>_report.cafe (operator<<(ostream&, const Report&), src/report.cpp:45-165; Report::compute_expansion :167-184;
// gene_family2report :186-221; built by estimator::execute, src/execute.cpp:190-197).  lambdas: the fitted values (may be empty);
// lambda_tree: write the tree's lambda indices, else the reference's dummy tree of 1s; branch_probs[F x n_nodes] (-1 = none): only
// families that have branch probabilities get a line (Report::add_line_item :223-229).  The reference writes no line break between
// the "'ID'\t'Newick'" header and the first family; neither do we.
inline void write_report(std::ostream& ost, const Tree& t, const std::vector<double>& lambdas, bool lambda_tree,
                         const std::vector<std::string>& ids, const int32_t* states, const double* pvalues, const double* branch_probs)
{
    const std::vector<int> ape = ape_ids(t);
    const int n = t.n_nodes(), root = n - 1;
    std::vector<std::vector<int>> kids(n);
    for (int i = n - 1; i >= 0; --i) if (t.parent[i] >= 0) kids[t.parent[i]].push_back(i);
    std::function<void(std::ostream&, int, const std::function<std::string(int)>&)> newick =
        [&](std::ostream& o, int v, const std::function<std::string(int)>& text) {        // clade::write_newick, src/clade.cpp:206-223
            if (!kids[v].empty()) {
                o << '(';
                for (size_t i = 0; i < kids[v].size(); ++i) { if (i) o << ','; newick(o, kids[v][i], text); }
                o << ')';
            }
            o << text(v);
        };
    std::vector<int> prefix, stack{root};                                                  // clade::apply_prefix_order, src/clade.cpp:275-291
    while (!stack.empty()) {
        const int v = stack.back();
        stack.pop_back();
        for (auto it = kids[v].rbegin(); it != kids[v].rend(); ++it) stack.push_back(*it);
        prefix.push_back(v);
    }
    auto id_text = [&](int v) { return (t.is_leaf[v] ? t.name[v] : std::string()) + "<" + std::to_string(ape[v]) + ">"; };

    ost << "Tree:";
    newick(ost, root, [&](int v) { std::ostringstream o; o << (t.is_leaf[v] ? t.name[v] : std::string()) << ":" << t.branch_length[v]; return o.str(); });
    ost << "\n";
    if (!lambdas.empty()) {
        ost << "Lambda:\t";
        for (double l : lambdas) ost << l << "\t";
    }
    ost << "\n";
    ost << "Lambda tree:\t";
    newick(ost, root, [&](int v) { return lambda_tree ? std::to_string(t.lambda_class[v] + 1) : std::string("1"); });
    ost << "\n";
    ost << "# IDs of nodes:";
    newick(ost, root, id_text);
    ost << "\n";
    ost << "# Output format for: ' Average Expansion', 'Expansions', 'No Change', 'Contractions', and 'Branch-specific P-values' = (node ID, node ID): ";
    for (int v : prefix)
        if (!t.is_leaf[v]) {
            ost << "(";
            for (size_t i = 0; i < kids[v].size(); ++i) ost << (i ? "," : "") << ape[kids[v][i]];
            ost << ") ";
        }
    ost << "\n";

    // Report::compute_expansion: per non-root node, over all families, the difference from the parent
    const size_t F = ids.size();
    std::vector<float> average(n, 0.0f);
    std::vector<int> expanded(n, 0), same(n, 0), decreased(n, 0);
    for (int v = 0; v < n; ++v) {
        if (t.parent[v] < 0) continue;
        int total = 0;
        for (size_t f = 0; f < F; ++f) {
            const int d = states[f * n + v] - states[f * n + t.parent[v]];
            total += d;
            expanded[v] += d > 0;
            decreased[v] += d < 0;
            same[v] += d == 0;
        }
        average[v] = float(total) / float(F);
    }
    if (n > 1) {
        auto per_child = [&](const char* header, const std::function<float(int)>& value) {
            ost << header;
            for (int v : prefix)
                if (!t.is_leaf[v]) {
                    ost << "\t(";
                    for (size_t i = 0; i < kids[v].size(); ++i) ost << (i ? "," : "") << value(kids[v][i]);
                    ost << ")";
                }
            ost << "\n";
        };
        per_child("Average Expansion:", [&](int v) { return average[v]; });
        per_child("Expansion :", [&](int v) { return (float)expanded[v]; });      // write_delta streams the counts as floats
        per_child("Remain :", [&](int v) { return (float)same[v]; });
        per_child("Decrease :", [&](int v) { return (float)decreased[v]; });
    }
    ost << "'ID'\t'Newick'";
    for (size_t f = 0; f < F; ++f) {
        bool any = false;
        for (int v = 0; branch_probs != nullptr && v < n; ++v) any = any || branch_probs[f * n + v] >= 0;
        if (!any) continue;
        ost << ids[f] << "\t";
        newick(ost, root, [&](int v) {
            std::ostringstream o;
            o << (t.is_leaf[v] ? t.name[v] : std::string()) << "_" << states[f * n + v] << ":" << t.branch_length[v];
            return o.str();
        });
        ost << "\t" << pvalues[f] << "\t";
        newick(ost, root, id_text);
        ost << std::endl;
    }
}

// simulation.txt / simulation_truth.txt (simulator::print_simulations, src/simulator.cpp:135-172; simulator::simulate :97-132 writes
// the leaf table and, with include_internal, the "truth" table whose interior columns are labelled with the node's position in the
// reverse level order).  node_sizes[F x n_nodes] as cafe_b200_simulate returns them; family_lambda[f]: the first lambda of the
// (perturbed / category) lambda family f was simulated with (simulated_family::lambda, create_trial :35).
inline void write_simulations(std::ostream& ost, const Tree& t, size_t n_families, const int32_t* node_sizes, const double* family_lambda,
                              bool include_internal)
{
    const int n = t.n_nodes();
    ost << "DESC\tFID";
    for (int i = 0; i < n; ++i) {
        if (t.is_leaf[i]) ost << '\t' << t.name[i];
        else if (include_internal) ost << '\t' << i;
    }
    ost << std::endl;
    for (size_t f = 0; f < n_families; ++f) {
        ost << "L" << family_lambda[f] << "\tsimfam" << f;
        for (int i = 0; i < n; ++i)
            if (t.is_leaf[i] || include_internal) ost << '\t' << node_sizes[f * n + i];
        ost << std::endl;
    }
}

// The error model file the reference writes after estimating epsilon (write_error_model_file, src/io.cpp:277-297; called from
// src/execute.cpp:37 and src/core.cpp:127): "maxcnt" = the number of table rows - 1 (error_model::get_max_family_size returns the table
// size, src/error_model.h:55-57), the deviations, then one line per family size whose probabilities differ from the previous size's.
inline void write_error_model(std::ostream& ost, const ErrorModelTable& em)
{
    ost << "maxcnt: " << em.rows() - 1 << "\n";
    ost << "cntdiff: -1 0 1\n";
    const double* last = nullptr;
    for (int j = 0; j < em.rows(); ++j) {
        const double* row = em.probs.data() + 3 * (size_t)j;
        if (last != nullptr && row[0] == last[0] && row[1] == last[1] && row[2] == last[2]) continue;
        last = row;
        ost << j << " " << row[0] << " " << row[1] << " " << row[2] << std::endl;
    }
}

}  // namespace cafe_b200_host
