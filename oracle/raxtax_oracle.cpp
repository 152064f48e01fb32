// raxtax_oracle.cpp -- CPU restatement of the raxtax v1.5.0 query-classification path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
// The product (raxtax_b200/) never links, imports or executes anything under oracle/.
//
// Every function cites the reference file:line (relative to /root/reference) it restates.
// The reference is Rust and cannot be compiled in this image (no cargo/rustc), so this is a
// "port" oracle.  It is pinned against every known-answer test the reference holds for this
// path (tests/test_oracle_kats.py):
//   utils.rs:236-263 (2-bit map, k-mer KAT), utils.rs:208-224 (norm / L1-distance KATs),
//   parser.rs:166-299 (ref ordering, k_mer_map content, IUPAC codes),
//   lineage.rs:191-334 (three tree-aggregation KATs), prob.rs:208-235 (PMF properties).
// PARITY UNPINNED for: absolute probability values of prob.rs (the reference only holds
// self-consistency properties, no golden numbers) and everything in raxtax.rs (no tests in
// the reference: hit-count loop, exact-match handling, skip mode, override).  Those parts rest
// on this line-by-line restatement and code review only.
//
// Third-party arithmetic absent from /root/reference: statrs ^0.16 (Cargo.toml:39)
// `function::factorial::ln_binomial`, restated below from its published algorithm
// (ln_factorial via a 171-entry exact factorial cache, else Lanczos ln_gamma with g=10.900511,
// 11 coefficients).  ahash::HashMap iteration order is randomly seeded in the reference; the
// oracle iterates distinct counts in ascending order (a valid instance of that freedom).
//
// Build: make -C oracle   ->  oracle/_build/libraxtax_oracle.so

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace orc {

static const double NEG_INF = -std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------------------------------------
// statrs 0.16 restatement: function::gamma::ln_gamma, function::factorial::{ln_factorial,
// ln_binomial}.  Call sites in the reference: prob.rs:20, prob.rs:117, prob.rs:143.
// ---------------------------------------------------------------------------------------------
static const double GAMMA_R = 10.900511;
static const double GAMMA_DK[11] = {
    2.48574089138753565546e-5, 1.05142378581721974210,    -3.45687097222016235469,
    4.51227709466894823700,    -2.98285225323576655721,   1.05639711577126713077,
    -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
    4.63399473359905636708e-6, -2.71994908488607703910e-9};
static const double LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978;
static const double LN_PI = 1.1447298858494001741434273513530587116472948129153;

static double ln_gamma(double x) {
    if (x < 0.5) {
        double s = GAMMA_DK[0];
        for (int i = 1; i < 11; ++i) s += GAMMA_DK[i] / ((double)i - x);
        return LN_PI - std::log(std::sin(M_PI * x)) - std::log(s) - LN_2_SQRT_E_OVER_PI -
               (0.5 - x) * std::log((0.5 - x + GAMMA_R) / M_E);
    }
    double s = GAMMA_DK[0];
    for (int i = 1; i < 11; ++i) s += GAMMA_DK[i] / (x + (double)i - 1.0);
    return std::log(s) + LN_2_SQRT_E_OVER_PI + (x - 0.5) * std::log((x - 0.5 + GAMMA_R) / M_E);
}

static const double* factorial_cache() {
    static double cache[171];
    static bool init = false;
    if (!init) {
        cache[0] = 1.0;
        for (int i = 1; i <= 170; ++i) cache[i] = cache[i - 1] * (double)i;
        init = true;
    }
    return cache;
}

static double ln_factorial(uint64_t x) {
    if (x <= 170) return std::log(factorial_cache()[x]);
    return ln_gamma((double)x + 1.0);
}

static double ln_binomial(uint64_t n, uint64_t k) {
    if (k > n) return NEG_INF;
    return ln_factorial(n) - ln_factorial(k) - ln_factorial(n - k);
}

// ---------------------------------------------------------------------------------------------
// utils.rs:17-25  map_four_to_two_bit_repr ; utils.rs:27-40 sequence_to_kmers
// ---------------------------------------------------------------------------------------------
static inline int map_four_to_two_bit_repr(uint8_t c) {
    switch (c) {
        case 0b0001: return 0b00;
        case 0b0010: return 0b01;
        case 0b0100: return 0b10;
        case 0b1000: return 0b11;
        default: return -1;  // None
    }
}

// one window of 8 codes -> Some(k-mer) / None   (utils.rs:30-35, tree.rs:115-120)
static inline bool window_to_kmer(const uint8_t* vals, uint16_t* out) {
    uint16_t acc = 0;
    for (int j = 0; j < 8; ++j) {
        int c = map_four_to_two_bit_repr(vals[j]);
        if (c < 0) return false;
        acc |= (uint16_t)(c << (14 - j * 2));
    }
    *out = acc;
    return true;
}

static std::vector<uint16_t> sequence_to_kmers(const uint8_t* seq, size_t len) {
    std::unordered_set<uint16_t> k_mers;  // HashSet (utils.rs:28)
    if (len >= 8) {
        for (size_t i = 0; i + 8 <= len; ++i) {  // sequence.windows(8)
            uint16_t k;
            if (window_to_kmer(seq + i, &k)) k_mers.insert(k);
        }
    }
    std::vector<uint16_t> v(k_mers.begin(), k_mers.end());
    std::sort(v.begin(), v.end());  // .sorted()  (utils.rs:39)
    return v;
}

// utils.rs:91-105
static double euclidean_distance_l1(const double* a, const double* b, size_t n) {
    if (n == 0) return 0.0;
    double a_sum = 0.0, b_sum = 0.0;
    for (size_t i = 0; i < n; ++i) a_sum += a[i];
    for (size_t i = 0; i < n; ++i) b_sum += b[i];
    if (!(a_sum > 0.0) || !(b_sum > 0.0)) throw std::runtime_error("assert sum > 0 (utils.rs:98-99)");
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double d = a[i] / a_sum - b[i] / b_sum;
        s += d * d;  // powi(2)
    }
    return std::sqrt(s);
}

// utils.rs:107-116
static double euclidean_norm(const double* v, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += v[i] * v[i];
    return std::sqrt(s);
}

// utils.rs:70-81
static std::string decompress_sequence(const uint8_t* seq, size_t len) {
    std::string s(len, '-');
    for (size_t i = 0; i < len; ++i) {
        switch (seq[i]) {
            case 1: s[i] = 'A'; break;
            case 2: s[i] = 'C'; break;
            case 4: s[i] = 'G'; break;
            case 8: s[i] = 'T'; break;
            default: s[i] = '-';
        }
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// tree.rs:36-43, 181-228  Tree / Node / NodeType
// ---------------------------------------------------------------------------------------------
enum NodeType { Inner = 0, Taxon = 1, Sequence = 2 };

struct Node {
    std::string label;
    size_t lo, hi;  // confidence_range
    std::vector<Node> children;
    NodeType node_type;
    Node(std::string l, size_t idx, NodeType t) : label(std::move(l)), lo(idx), hi(idx + 1), node_type(t) {}
};

struct BytesHash {
    size_t operator()(const std::string& s) const { return std::hash<std::string>()(s); }
};

struct Tree {
    Node root{"root", 0, Inner};
    std::vector<std::string> lineages;
    std::unordered_map<std::string, std::vector<uint32_t>, BytesHash> sequences;  // key = 4-bit code bytes
    std::vector<std::vector<uint32_t>> k_mer_map;
    size_t num_tips = 0;
    std::vector<std::string> sorted_sequences;  // not in the reference struct; kept for tests only
};

static std::vector<std::string> split_commas(const std::string& s) {
    std::vector<std::string> out;
    size_t start = 0;
    while (true) {
        size_t p = s.find(',', start);
        if (p == std::string::npos) {
            out.emplace_back(s.substr(start));
            break;
        }
        out.emplace_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}

// tree.rs:47-140  Tree::new
static std::unique_ptr<Tree> tree_new(std::vector<std::string> lineages, std::vector<std::string> sequences) {
    if (lineages.size() != sequences.size()) throw std::runtime_error("zip_eq length mismatch (tree.rs:53)");
    if (lineages.size() > 0xFFFFFFFFull) throw std::runtime_error("Too many database sequences (tree.rs:24-31)");
    auto tree = std::make_unique<Tree>();
    for (const auto& s : sequences) tree->sequences.emplace(s, std::vector<uint32_t>());  // tree.rs:50-51
    tree->k_mer_map.assign(2 << 15, std::vector<uint32_t>());                             // tree.rs:52
    std::vector<std::pair<std::string, std::string>> pairs;
    pairs.reserve(lineages.size());
    for (size_t i = 0; i < lineages.size(); ++i) pairs.emplace_back(std::move(lineages[i]), std::move(sequences[i]));
    // tree.rs:54  sort_by(|(l1,_),(l2,_)| l1.cmp(l2)) -- stable, byte-wise
    std::stable_sort(pairs.begin(), pairs.end(),
                     [](const auto& a, const auto& b) { return a.first.compare(b.first) < 0; });
    size_t confidence_idx = 0;
    for (size_t idx = 0; idx < pairs.size(); ++idx) {  // tree.rs:56-126
        const std::string& lineage = pairs[idx].first;
        const std::string& sequence = pairs[idx].second;
        std::vector<std::string> levels = split_commas(lineage);
        size_t last_level_idx = levels.size() - 1;
        Node* current = &tree->root;
        for (size_t level = 0; level < levels.size(); ++level) {
            const std::string& label = levels[level];
            NodeType nt = (level == last_level_idx) ? Taxon : Inner;
            if (!current->children.empty()) {  // get_last_child_label() -> Some(name)
                if (current->children.back().label != label) current->children.emplace_back(label, confidence_idx, nt);
                current->hi = confidence_idx + 1;
            } else {
                current->children.emplace_back(label, confidence_idx, nt);
                current->hi = confidence_idx + 1;
            }
            if (level == last_level_idx) confidence_idx += 1;
            current = &current->children.back();
        }
        current->children.emplace_back(current->label, confidence_idx - 1, Sequence);  // tree.rs:102-106
        current->hi = confidence_idx;                                                  // tree.rs:107
        tree->sequences[sequence].push_back((uint32_t)idx);                            // tree.rs:109-112
        if (sequence.size() >= 8) {
            for (size_t i = 0; i + 8 <= sequence.size(); ++i) {  // tree.rs:114-123
                uint16_t k;
                if (window_to_kmer((const uint8_t*)sequence.data() + i, &k)) tree->k_mer_map[k].push_back((uint32_t)idx);
            }
        }
    }
    tree->root.hi = confidence_idx;  // tree.rs:127
    tree->lineages.reserve(pairs.size());
    for (auto& p : pairs) {
        tree->lineages.push_back(std::move(p.first));
        tree->sorted_sequences.push_back(std::move(p.second));
    }
    for (auto& v : tree->k_mer_map) {  // tree.rs:134-137  unique().sorted()
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    tree->num_tips = confidence_idx;  // tree.rs:138
    return tree;
}

// ---------------------------------------------------------------------------------------------
// parser.rs:11-34  map_dna_char
// ---------------------------------------------------------------------------------------------
static uint8_t map_dna_char(char ch) {
    const uint8_t a = 1, c = 2, g = 4, t = 8;
    char u = (ch >= 'a' && ch <= 'z') ? (char)(ch - 32) : ch;
    switch (u) {
        case 'A': return a;
        case 'C': return c;
        case 'G': return g;
        case 'T': return t;
        case 'W': return a | t;
        case 'S': return c | g;
        case 'M': return a | c;
        case 'K': return g | t;
        case 'R': return a | g;
        case 'Y': return c | t;
        case 'B': return c | g | t;
        case 'D': return a | g | t;
        case 'H': return a | c | t;
        case 'V': return a | c | g;
        case 'N': return a | c | g | t;
        default: throw std::runtime_error(std::string("Unexpected character: ") + ch);  // panic! parser.rs:32
    }
}

static bool is_rust_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }

// str::lines() + trim() + filter(non-empty, not ';')   parser.rs:53-57 / 124-128
static std::vector<std::string> fasta_lines(const std::string& text) {
    std::vector<std::string> out;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        size_t end = (nl == std::string::npos) ? text.size() : nl;
        size_t b = pos, e = end;
        while (b < e && is_rust_ws((unsigned char)text[b])) ++b;
        while (e > b && is_rust_ws((unsigned char)text[e - 1])) --e;
        if (e > b && text[b] != ';') out.emplace_back(text.substr(b, e - b));
        if (nl == std::string::npos) break;
        pos = nl + 1;
    }
    return out;
}

// regex  tax=([^;]+);   first match (parser.rs:50, 78-86)
static bool capture_tax(const std::string& label, std::string* out) {
    size_t from = 0;
    while (true) {
        size_t p = label.find("tax=", from);
        if (p == std::string::npos) return false;
        size_t s = p + 4;
        size_t e = s;
        while (e < label.size() && label[e] != ';') ++e;
        if (e > s && e < label.size()) {  // at least one non-';' char followed by ';'
            *out = label.substr(s, e - s);
            return true;
        }
        from = p + 1;
    }
}

// parser.rs:46-105  parse_reference_fasta_str
static std::unique_ptr<Tree> parse_reference_fasta_str(const std::string& text) {
    if (text.empty()) throw std::runtime_error("File is empty");
    std::vector<std::string> lines = fasta_lines(text);
    if (lines.empty()) throw std::runtime_error("index out of bounds: lines[0] (parser.rs:58)");
    if (lines[0][0] != '>') throw std::runtime_error("Not a valid FASTA file");
    std::vector<std::string> labels, sequences;
    std::string current;
    for (const auto& line : lines) {
        if (line[0] == '>') {
            std::string label = line.substr(1), lineage;
            if (!capture_tax(label, &lineage))
                throw std::runtime_error("Unexpected taxonomical annotation detected in label " + label);
            labels.push_back(lineage);
            if (!current.empty()) {
                sequences.push_back(current);
                current.clear();
            }
        } else {
            for (char ch : line) current.push_back((char)map_dna_char(ch));
        }
    }
    sequences.push_back(current);
    if (labels.size() != sequences.size()) throw std::runtime_error("Number of sequences does not match number of labels");
    return tree_new(std::move(labels), std::move(sequences));
}

// parser.rs:117-154  parse_query_fasta_str (queries_to_skip handled by the caller: empty set here)
static std::vector<std::pair<std::string, std::string>> parse_query_fasta_str(const std::string& text) {
    if (text.empty()) throw std::runtime_error("File is empty");
    std::vector<std::string> lines = fasta_lines(text);
    if (lines.empty()) throw std::runtime_error("index out of bounds: lines[0] (parser.rs:129)");
    if (lines[0][0] != '>') throw std::runtime_error("Not a valid FASTA file");
    std::vector<std::pair<std::string, std::string>> queries;
    std::pair<std::string, std::string> current;
    for (const auto& line : lines) {
        if (line[0] == '>') {
            if (!current.second.empty()) {
                queries.push_back(current);
                current.second.clear();
            }
            current.first = line.substr(1);
        } else {
            for (char ch : line) current.second.push_back((char)map_dna_char(ch));
        }
    }
    queries.push_back(current);
    return queries;
}

// ---------------------------------------------------------------------------------------------
// prob.rs:105-119  only_last_pmf
// ---------------------------------------------------------------------------------------------
static double only_last_pmf(uint64_t total, uint64_t trials, uint64_t m, double num_possible_kmer_sets) {
    if (m == total) return 1.0;
    if (m == 0) return 0.0;
    double num_possible_matches = ln_binomial(m + trials - 1, trials);
    return std::exp(num_possible_matches - num_possible_kmer_sets);
}

// prob.rs:121-170  iterative_pmfs_ln -- one row (one distinct intersection size)
static std::vector<double> iterative_pmf_ln_row(uint64_t total, uint64_t trials, uint64_t m, double T) {
    std::vector<double> res;
    if (m == total) {
        res.assign(trials + 1, NEG_INF);
        res[trials] = 0.0;
        return res;
    }
    if (m == 0) {
        res.assign(trials + 1, NEG_INF);
        res[0] = 0.0;
        return res;
    }
    std::vector<double> poss;  // (1..=trials).scan
    {
        double sum = 0.0;
        for (uint64_t i = 1; i <= trials; ++i) {
            sum += std::log((double)(m + i - 1) / (double)i);
            poss.push_back(sum);
        }
    }
    double impossible_init = ln_binomial(total - m + trials - 1, trials);
    std::vector<double> imp;  // (1..trials).scan(...).chain([0.0])
    {
        double sum = impossible_init;
        for (uint64_t i = 1; i < trials; ++i) {
            sum -= std::log((double)(total - m + trials - i) / (double)(trials - i + 1));
            imp.push_back(sum);
        }
        imp.push_back(0.0);
    }
    if (poss.size() != imp.size()) throw std::runtime_error("zip_eq length mismatch (prob.rs:162)");
    res.push_back(impossible_init - T);
    for (size_t i = 0; i < poss.size(); ++i) res.push_back(poss[i] + imp[i] - T);
    return res;
}

// prob.rs:8-103  highest_hit_prob_per_reference
// The reference keeps the histogram and the per-count probabilities in ahash::HashMaps (prob.rs:13-19,92-95): O(1) per reference,
// iteration order unspecified.  Two equivalent containers here, both iterated in ascending count order (so that the f64 sums of
// prob.rs:62-73 are reproducible): an ordered std::map (the original statement, O(log D) per reference) and -- g_flat_hist, the
// default since round 2 -- a flat table indexed by the u16 count, O(1) per reference like the hash map.  Same values either way;
// the flat table is what the CPU baseline is timed with, so that the port is not slower than the reference's data structure.
static int g_flat_hist = 1;

static std::vector<double> highest_hit_prob_per_reference(uint16_t total_num_k_mers, size_t num_trials,
                                                          const uint16_t* sizes, size_t n) {
    std::vector<std::pair<uint16_t, size_t>> counts;  // HashMap<u16,usize> (prob.rs:13-19) as (count, multiplicity), ascending
    if (g_flat_hist) {
        std::vector<size_t> tab(65536, 0);
        for (size_t i = 0; i < n; ++i) tab[sizes[i]] += 1;
        for (size_t m = 0; m < 65536; ++m)
            if (tab[m]) counts.emplace_back((uint16_t)m, tab[m]);
    } else {
        std::map<uint16_t, size_t> cm;
        for (size_t i = 0; i < n; ++i) cm[sizes[i]] += 1;
        counts.assign(cm.begin(), cm.end());
    }
    // prob.rs:20-23; u64 arithmetic wraps in release builds when K == 0 (0 + 0 - 1)
    uint64_t nn = (uint64_t)total_num_k_mers + (uint64_t)num_trials - 1ull;
    double T = ln_binomial(nn, (uint64_t)num_trials);
    std::vector<std::pair<uint16_t, double>> hp;  // highest_hit_probs: (count, probability), ascending
    bool any_full = false;  // prob.rs:24-26
    for (auto& kv : counts) any_full |= kv.first == total_num_k_mers;
    if (any_full) {
        for (auto& kv : counts) hp.emplace_back(kv.first, only_last_pmf(total_num_k_mers, num_trials, kv.first, T));
    } else {
        std::vector<std::pair<uint16_t, std::vector<double>>> pmfs;
        for (auto& kv : counts) pmfs.emplace_back(kv.first, iterative_pmf_ln_row(total_num_k_mers, num_trials, kv.first, T));
        std::vector<std::vector<double>> cmfs;  // prob.rs:49-61
        for (auto& pv : pmfs) {
            std::vector<double> c;
            double sum = 0.0;
            for (double p : pv.second) {
                if (p != NEG_INF) sum += std::exp(p);
                c.push_back(std::log(sum));
            }
            cmfs.push_back(std::move(c));
        }
        std::vector<double> prod(num_trials + 1);  // prob.rs:62-73
        for (size_t i = 0; i <= num_trials; ++i) {
            double s = 0.0;
            size_t j = 0;
            for (auto& kv : counts) {
                s += (double)kv.second * cmfs[j][i];
                ++j;
            }
            prod[i] = s;
        }
        for (size_t j = 0; j < pmfs.size(); ++j) {  // prob.rs:74-90
            double s = 0.0;
            for (size_t i = 0; i <= num_trials; ++i) {
                double p = pmfs[j].second[i], c = cmfs[j][i], pc = prod[i];
                if (c == NEG_INF || pc == NEG_INF) s += 0.0;
                else s += std::exp(p + pc - c);
            }
            hp.emplace_back(pmfs[j].first, s);
        }
    }
    std::vector<double> out(n);
    if (g_flat_hist) {  // prob.rs:92-95
        std::vector<double> tab(65536, 0.0);
        for (auto& kv : hp) tab[kv.first] = kv.second;
        for (size_t i = 0; i < n; ++i) out[i] = tab[sizes[i]];
    } else {
        std::map<uint16_t, double> hm(hp.begin(), hp.end());
        for (size_t i = 0; i < n; ++i) out[i] = hm[sizes[i]];
    }
    double probs_sum = 0.0;
    for (size_t i = 0; i < n; ++i) probs_sum += out[i];  // prob.rs:97
    if (!(probs_sum > 0.0)) throw std::runtime_error("assert probs_sum > 0.0 (prob.rs:98)");
    for (size_t i = 0; i < n; ++i) out[i] = out[i] / probs_sum;  // prob.rs:99-102
    return out;
}

// ---------------------------------------------------------------------------------------------
// lineage.rs:7-14 EvaluationResult ; 51-180 Lineage
// ---------------------------------------------------------------------------------------------
struct EvaluationResult {
    size_t first_ref_idx;  // index into tree.lineages (lineage.rs:105)
    std::vector<double> confidence_values;
    double local_signal, global_signal;
};

struct Lineage {
    const Tree* tree;
    const std::vector<double>* confidence_values;
    std::vector<double> prefix;
    struct Vec3 {
        size_t idx;
        std::vector<double> conf, expected;
    };
    std::vector<Vec3> confidence_vectors;
    double rounding_factor = 100.0;  // 10^F64_OUTPUT_ACCURACY  (utils.rs:15, lineage.rs:67)

    Lineage(const Tree* t, const std::vector<double>* cv) : tree(t), confidence_values(cv) {  // lineage.rs:61-77
        prefix.reserve(cv->size() + 1);
        prefix.push_back(0.0);
        double sum = 0.0;
        for (double v : *cv) {
            sum += v;
            prefix.push_back(sum);
        }
    }
    double get_confidence(const Node& n) const { return prefix[n.hi] - prefix[n.lo]; }  // lineage.rs:114-117

    // lineage.rs:119-179
    bool eval_recurse(const Node& node, const std::vector<double>& cp, const std::vector<double>& ep) {
        bool no_child_significant = true, pushed_result = false;
        for (const Node& c : node.children) {
            double child_conf = std::round(get_confidence(c) * rounding_factor) / rounding_factor;  // f64::round
            if (child_conf == 0.0) continue;
            no_child_significant = false;
            std::vector<double> conf_prefix = cp, expected_prefix = ep;
            conf_prefix.push_back(child_conf);
            expected_prefix.push_back((double)(c.hi - c.lo) / (double)tree->num_tips);
            bool child_pushed = eval_recurse(c, conf_prefix, expected_prefix);
            if (!child_pushed && c.node_type == Taxon) {
                confidence_vectors.push_back({c.lo, conf_prefix, expected_prefix});
                pushed_result = true;
            }
            pushed_result |= child_pushed;
        }
        if (no_child_significant && node.node_type == Inner) {
            std::vector<double> conf_prefix = cp, expected_prefix = ep;
            const Node* cur = &node;
            while (cur->node_type == Inner) {
                if (cur->children.empty()) throw std::runtime_error("max_by on empty children (lineage.rs:162)");
                // Iterator::max_by returns the LAST maximal element
                const Node* best = &cur->children[0];
                double best_c = get_confidence(*best);
                for (size_t i = 1; i < cur->children.size(); ++i) {
                    double ci = get_confidence(cur->children[i]);
                    if (std::isnan(ci) || std::isnan(best_c)) throw std::runtime_error("partial_cmp unwrap on NaN");
                    if (ci >= best_c) {
                        best = &cur->children[i];
                        best_c = ci;
                    }
                }
                cur = best;
                conf_prefix.push_back(1.0 / rounding_factor);
                expected_prefix.push_back((double)(cur->hi - cur->lo) / (double)tree->num_tips);
            }
            confidence_vectors.push_back({cur->lo, conf_prefix, expected_prefix});
            pushed_result = true;
        }
        return pushed_result;
    }

    // lineage.rs:80-112
    std::vector<EvaluationResult> evaluate() {
        eval_recurse(tree->root, {}, {});
        double inv = 1.0 / (double)tree->num_tips;
        double s = 0.0;
        for (double v : *confidence_values) {
            double d = v - inv;
            s += d * d;
        }
        double leaf_confidence = std::sqrt(s);
        // sorted_by(|a,b| b.1.iter().partial_cmp(a.1.iter())) : stable, descending lexicographic
        std::stable_sort(confidence_vectors.begin(), confidence_vectors.end(), [](const Vec3& a, const Vec3& b) {
            // return true iff a must come before b  <=>  cmp(b.conf, a.conf) == Less
            return std::lexicographical_compare(b.conf.begin(), b.conf.end(), a.conf.begin(), a.conf.end());
        });
        std::vector<EvaluationResult> out;
        for (auto& v : confidence_vectors) {
            size_t start = v.expected.size() - 1;
            for (size_t i = 0; i < v.expected.size(); ++i)
                if (1.0 > v.expected[i]) {
                    start = i;
                    break;
                }
            double local = euclidean_distance_l1(v.conf.data() + start, v.expected.data() + start, v.conf.size() - start);
            out.push_back({v.idx, v.conf, local, leaf_confidence});
        }
        return out;
    }
};

// lineage.rs:17-30  get_output_string
static std::string fmt_fixed(double v, int prec) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.*f", prec, v);
    return buf;
}
static std::string output_string(const std::string& label, const std::string& lineage, const EvaluationResult& r) {
    std::string s = label + "\t" + lineage + "\t";
    for (size_t i = 0; i < r.confidence_values.size(); ++i) {
        if (i) s += ",";
        s += fmt_fixed(r.confidence_values[i], 2);
    }
    s += "\t" + fmt_fixed(r.local_signal, 5) + "\t" + fmt_fixed(r.global_signal, 5);
    return s;
}
// lineage.rs:32-48  get_tsv_string  (itertools interleave: alternate, then drain the longer one)
static std::string tsv_string(const std::string& label, const std::string& lineage, const EvaluationResult& r,
                              const std::string& sequence) {
    std::vector<std::string> a = split_commas(lineage), b;
    for (double v : r.confidence_values) b.push_back(fmt_fixed(v, 2));
    std::vector<std::string> inter;
    size_t ia = 0, ib = 0;
    bool flag = false;
    while (ia < a.size() || ib < b.size()) {
        if (!flag) {
            if (ia < a.size()) inter.push_back(a[ia++]);
            else inter.push_back(b[ib++]);
        } else {
            if (ib < b.size()) inter.push_back(b[ib++]);
            else inter.push_back(a[ia++]);
        }
        flag = !flag;
    }
    std::string s = label + "\t";
    for (size_t i = 0; i < inter.size(); ++i) {
        if (i) s += "\t";
        s += inter[i];
    }
    s += "\t" + fmt_fixed(r.local_signal, 5) + "\t" + fmt_fixed(r.global_signal, 5) + "\t" + sequence;
    return s;
}

// ---------------------------------------------------------------------------------------------
// raxtax.rs:39-88  body of the per-query closure
// ---------------------------------------------------------------------------------------------
struct QueryOut {
    std::vector<uint16_t> k_mers;
    std::vector<uint32_t> exact_matches;
    bool exact_parents_differ = false;  // raxtax.rs:49-52 warning condition
    std::vector<double> probs;          // filled only if keep_probs
    std::vector<EvaluationResult> results;
};

static void classify_one(const Tree& tree, const uint8_t* seq, size_t len, bool skip_exact_matches,
                         bool raw_confidence, std::vector<uint16_t>& intersect_buffer, bool keep_probs, QueryOut& out) {
    std::fill(intersect_buffer.begin(), intersect_buffer.end(), 0);  // raxtax.rs:41
    static const std::vector<uint32_t> empty_vec;
    std::string key((const char*)seq, len);
    auto it = tree.sequences.find(key);  // raxtax.rs:42
    const std::vector<uint32_t>& exact_matches = (it == tree.sequences.end()) ? empty_vec : it->second;
    out.exact_matches = exact_matches;
    if (!skip_exact_matches) {  // raxtax.rs:43-53 (logging only; we record the warning condition)
        bool all_equal = true;
        std::string first;
        bool have = false;
        for (uint32_t idx : exact_matches) {
            const std::string& l = tree.lineages[idx];
            size_t p = l.rfind(',');
            if (p == std::string::npos) throw std::runtime_error("rsplit_once(',').unwrap() on lineage without comma (raxtax.rs:49)");
            std::string parent = l.substr(0, p);
            if (!have) {
                first = parent;
                have = true;
            } else if (parent != first) all_equal = false;
        }
        out.exact_parents_differ = !all_equal;
    }
    out.k_mers = sequence_to_kmers(seq, len);  // raxtax.rs:55
    if (out.k_mers.size() > 0xFFFF) throw std::runtime_error("assert k_mers.len() fits u16 (raxtax.rs:56)");
    size_t num_trials = out.k_mers.size() / 2;  // raxtax.rs:57
    for (uint16_t k : out.k_mers)               // raxtax.rs:58-64
        for (uint32_t id : tree.k_mer_map[k]) intersect_buffer[id] += 1;
    if (skip_exact_matches)  // raxtax.rs:65-68
        for (uint32_t id : exact_matches) intersect_buffer[id] = 0;
    std::vector<double> probs = highest_hit_prob_per_reference((uint16_t)out.k_mers.size(), num_trials,
                                                               intersect_buffer.data(), intersect_buffer.size());
    Lineage lin(&tree, &probs);
    std::vector<EvaluationResult> eval_res = lin.evaluate();  // raxtax.rs:71
    if (eval_res.empty()) throw std::runtime_error("assert !eval_res.is_empty() (raxtax.rs:72)");
    if (!raw_confidence && !skip_exact_matches) {  // raxtax.rs:73-84
        if (exact_matches.size() == 1) {
            uint32_t idx = exact_matches[0];
            const std::string& l = tree.lineages[idx];
            size_t commas = std::count(l.begin(), l.end(), ',');
            EvaluationResult r{idx, std::vector<double>(commas + 1, 1.0), eval_res[0].local_signal, eval_res[0].global_signal};
            eval_res.clear();
            eval_res.push_back(r);
        }
    }
    out.results = std::move(eval_res);
    if (keep_probs) out.probs = std::move(probs);
}

}  // namespace orc

// =============================================================================================
// C interface for ctypes (tests / smoke / cpu_baseline only)
// =============================================================================================
using namespace orc;

struct orc_tree {
    std::unique_ptr<Tree> t;
};

static thread_local std::string g_err;
static int fail(const std::exception& e) {
    g_err = e.what();
    return -1;
}

extern "C" {

// 1 (default): flat count-indexed tables in highest_hit_prob_per_reference; 0: ordered std::map (the round-1 statement)
void orc_set_flat_hist(int on) { g_flat_hist = on ? 1 : 0; }
int orc_get_flat_hist(void) { return g_flat_hist; }

const char* orc_last_error() { return g_err.c_str(); }

double orc_ln_binomial(uint64_t n, uint64_t k) { return ln_binomial(n, k); }
double orc_ln_gamma(double x) { return ln_gamma(x); }
double orc_euclidean_norm(const double* v, size_t n) { return euclidean_norm(v, n); }
int orc_euclidean_distance_l1(const double* a, const double* b, size_t n, double* out) {
    try {
        *out = euclidean_distance_l1(a, b, n);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
int orc_map_four_to_two_bit_repr(uint8_t c) { return map_four_to_two_bit_repr(c); }

// returns K; out must hold max(len-7,0) entries
size_t orc_sequence_to_kmers(const uint8_t* seq, size_t len, uint16_t* out) {
    auto v = sequence_to_kmers(seq, len);
    std::copy(v.begin(), v.end(), out);
    return v.size();
}

// codes of a DNA string via map_dna_char; returns 0 / -1
int orc_map_dna(const char* s, size_t len, uint8_t* out) {
    try {
        for (size_t i = 0; i < len; ++i) out[i] = map_dna_char(s[i]);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

orc_tree* orc_tree_from_fasta(const char* text, size_t len) {
    try {
        auto h = new orc_tree();
        h->t = parse_reference_fasta_str(std::string(text, len));
        return h;
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

// lineages: '\n'-joined blob; sequences: 4-bit codes concatenated with offsets[n+1]
orc_tree* orc_tree_new(size_t n, const char* lineage_blob, size_t blob_len, const uint64_t* seq_off, const uint8_t* codes) {
    try {
        std::vector<std::string> lin, seqs;
        std::string blob(lineage_blob, blob_len);
        size_t pos = 0;
        for (size_t i = 0; i < n; ++i) {
            size_t nl = blob.find('\n', pos);
            if (nl == std::string::npos) nl = blob.size();
            lin.emplace_back(blob.substr(pos, nl - pos));
            pos = nl + 1;
            seqs.emplace_back((const char*)codes + seq_off[i], seq_off[i + 1] - seq_off[i]);
        }
        auto h = new orc_tree();
        h->t = tree_new(std::move(lin), std::move(seqs));
        return h;
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}

void orc_tree_free(orc_tree* h) { delete h; }
size_t orc_tree_num_tips(const orc_tree* h) { return h->t->num_tips; }
const char* orc_tree_lineage(const orc_tree* h, size_t i) { return h->t->lineages[i].c_str(); }
size_t orc_tree_kmer_list_len(const orc_tree* h, uint32_t kmer) { return h->t->k_mer_map[kmer].size(); }
void orc_tree_kmer_list(const orc_tree* h, uint32_t kmer, uint32_t* out) {
    const auto& v = h->t->k_mer_map[kmer];
    std::copy(v.begin(), v.end(), out);
}
uint64_t orc_tree_nnz(const orc_tree* h) {
    uint64_t s = 0;
    for (auto& v : h->t->k_mer_map) s += v.size();
    return s;
}
// CSR export of k_mer_map: offsets[65537], ids[nnz]
void orc_tree_csr(const orc_tree* h, uint64_t* offsets, uint32_t* ids) {
    uint64_t o = 0;
    for (size_t k = 0; k < h->t->k_mer_map.size(); ++k) {
        offsets[k] = o;
        for (uint32_t id : h->t->k_mer_map[k]) ids[o++] = id;
    }
    offsets[h->t->k_mer_map.size()] = o;
}
// sorted sequence i (codes) -- test helper
size_t orc_tree_sequence_len(const orc_tree* h, size_t i) { return h->t->sorted_sequences[i].size(); }
void orc_tree_sequence(const orc_tree* h, size_t i, uint8_t* out) {
    memcpy(out, h->t->sorted_sequences[i].data(), h->t->sorted_sequences[i].size());
}
// exact-match lookup (tree.sequences.get): returns count, writes up to cap ids
size_t orc_tree_exact(const orc_tree* h, const uint8_t* seq, size_t len, uint32_t* out, size_t cap) {
    auto it = h->t->sequences.find(std::string((const char*)seq, len));
    if (it == h->t->sequences.end()) return 0;
    for (size_t i = 0; i < it->second.size() && i < cap; ++i) out[i] = it->second[i];
    return it->second.size();
}

// pre-order flattening of the Node tree (all node types) for structural cross-checks:
// per node: lo, hi, type, depth, parent(-1 root), n_children.  returns node count; pass NULLs to size.
static void flatten(const Node& n, int depth, int64_t parent, std::vector<uint64_t>* lo, std::vector<uint64_t>* hi,
                    std::vector<uint8_t>* type, std::vector<int32_t>* dep, std::vector<int64_t>* par,
                    std::vector<uint32_t>* nch, std::string* labels) {
    int64_t me = (int64_t)lo->size();
    lo->push_back(n.lo);
    hi->push_back(n.hi);
    type->push_back((uint8_t)n.node_type);
    dep->push_back(depth);
    par->push_back(parent);
    nch->push_back((uint32_t)n.children.size());
    labels->append(n.label);
    labels->push_back('\n');
    for (const Node& c : n.children) flatten(c, depth + 1, me, lo, hi, type, dep, par, nch, labels);
}
size_t orc_tree_flatten(const orc_tree* h, uint64_t* lo, uint64_t* hi, uint8_t* type, int32_t* depth, int64_t* parent,
                        uint32_t* nchildren, char* labels, size_t labels_cap, size_t* labels_len) {
    std::vector<uint64_t> vlo, vhi;
    std::vector<uint8_t> vt;
    std::vector<int32_t> vd;
    std::vector<int64_t> vp;
    std::vector<uint32_t> vn;
    std::string lab;
    flatten(h->t->root, 0, -1, &vlo, &vhi, &vt, &vd, &vp, &vn, &lab);
    if (lo) {
        std::copy(vlo.begin(), vlo.end(), lo);
        std::copy(vhi.begin(), vhi.end(), hi);
        std::copy(vt.begin(), vt.end(), type);
        std::copy(vd.begin(), vd.end(), depth);
        std::copy(vp.begin(), vp.end(), parent);
        std::copy(vn.begin(), vn.end(), nchildren);
    }
    if (labels && labels_cap >= lab.size()) memcpy(labels, lab.data(), lab.size());
    if (labels_len) *labels_len = lab.size();
    return vlo.size();
}

// prob.rs:8  -- out[n]
int orc_highest_hit_prob(uint16_t K, size_t t, const uint16_t* sizes, size_t n, double* out) {
    try {
        auto v = highest_hit_prob_per_reference(K, t, sizes, n);
        std::copy(v.begin(), v.end(), out);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// prob.rs:121 -- one pmf row, out[t+1]
int orc_iterative_pmf_ln(uint64_t K, uint64_t t, uint64_t m, double* out) {
    try {
        double T = ln_binomial(K + t - 1, t);
        auto v = iterative_pmf_ln_row(K, t, m, T);
        std::copy(v.begin(), v.end(), out);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// Result buffers shared by orc_lineage_evaluate / orc_classify
struct orc_results {
    std::vector<uint32_t> query, first_ref;
    std::vector<uint8_t> nlev;
    std::vector<double> conf;  // [n * max_lev]
    std::vector<double> local, global;
    int max_lev = 0;
};

static void push_results(orc_results& R, uint32_t q, const std::vector<EvaluationResult>& res) {
    for (auto& r : res) {
        if ((int)r.confidence_values.size() > R.max_lev) throw std::runtime_error("result deeper than max_lev");
        R.query.push_back(q);
        R.first_ref.push_back((uint32_t)r.first_ref_idx);
        R.nlev.push_back((uint8_t)r.confidence_values.size());
        size_t base = R.conf.size();
        R.conf.resize(base + R.max_lev, 0.0);
        std::copy(r.confidence_values.begin(), r.confidence_values.end(), R.conf.begin() + base);
        R.local.push_back(r.local_signal);
        R.global.push_back(r.global_signal);
    }
}

orc_results* orc_results_new(int max_lev) {
    auto r = new orc_results();
    r->max_lev = max_lev;
    return r;
}
void orc_results_free(orc_results* r) { delete r; }
size_t orc_results_len(const orc_results* r) { return r->query.size(); }
void orc_results_copy(const orc_results* r, uint32_t* query, uint32_t* first_ref, uint8_t* nlev, double* conf, double* local,
                      double* global) {
    std::copy(r->query.begin(), r->query.end(), query);
    std::copy(r->first_ref.begin(), r->first_ref.end(), first_ref);
    std::copy(r->nlev.begin(), r->nlev.end(), nlev);
    std::copy(r->conf.begin(), r->conf.end(), conf);
    std::copy(r->local.begin(), r->local.end(), local);
    std::copy(r->global.begin(), r->global.end(), global);
}

// lineage.rs:61-112 with caller-supplied per-reference confidences (the lineage.rs KATs)
int orc_lineage_evaluate(const orc_tree* h, const double* confidences, size_t n, orc_results* out) {
    try {
        std::vector<double> cv(confidences, confidences + n);
        Lineage lin(h->t.get(), &cv);
        push_results(*out, 0, lin.evaluate());
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// raxtax.rs:14-97 over a batch.  Threads: chunks of chunk_size handed out dynamically (rayon par_chunks).
// Optional taps: out_K[n_q], out_counts[n_q*N] (post skip-zeroing buffer), out_probs[n_q*N],
// out_kmers[n_q*kmer_stride], out_nexact[n_q], out_warn[n_q]; strings: primary/tsv joined per query with '\n'.
int orc_classify(const orc_tree* h, size_t n_q, const uint64_t* seq_off, const uint8_t* codes, int skip_exact,
                 int raw_conf, int n_threads, size_t chunk_size, uint16_t* out_K, uint16_t* out_counts, double* out_probs,
                 uint16_t* out_kmers, size_t kmer_stride, uint32_t* out_nexact, uint8_t* out_warn, orc_results* out_results,
                 double* out_seconds) {
    const Tree& tree = *h->t;
    const size_t N = tree.num_tips;
    if (n_threads < 1) n_threads = 1;
    if (chunk_size == 0) chunk_size = n_q ? n_q : 1;
    std::vector<std::vector<EvaluationResult>> per_query(n_q);
    std::atomic<size_t> next_chunk{0};
    std::atomic<bool> failed{false};
    std::string err;
    std::mutex err_mtx;
    size_t n_chunks = (n_q + chunk_size - 1) / chunk_size;
    auto worker = [&]() {
        try {
            while (true) {
                size_t c = next_chunk.fetch_add(1);
                if (c >= n_chunks || failed.load()) break;
                std::vector<uint16_t> buffer(N, 0);  // raxtax.rs:38
                size_t q0 = c * chunk_size, q1 = std::min(n_q, q0 + chunk_size);
                for (size_t q = q0; q < q1; ++q) {
                    QueryOut qo;
                    classify_one(tree, codes + seq_off[q], seq_off[q + 1] - seq_off[q], skip_exact != 0, raw_conf != 0, buffer,
                                 out_probs != nullptr, qo);
                    if (out_K) out_K[q] = (uint16_t)qo.k_mers.size();
                    if (out_counts) memcpy(out_counts + q * N, buffer.data(), N * sizeof(uint16_t));
                    if (out_probs) memcpy(out_probs + q * N, qo.probs.data(), N * sizeof(double));
                    if (out_kmers)
                        for (size_t i = 0; i < qo.k_mers.size() && i < kmer_stride; ++i) out_kmers[q * kmer_stride + i] = qo.k_mers[i];
                    if (out_nexact) out_nexact[q] = (uint32_t)qo.exact_matches.size();
                    if (out_warn) out_warn[q] = qo.exact_parents_differ ? 1 : 0;
                    per_query[q] = std::move(qo.results);
                }
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> g(err_mtx);
            err = e.what();
            failed.store(true);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads == 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (out_seconds) *out_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (failed.load()) {
        g_err = err;
        return -1;
    }
    try {
        if (out_results)
            for (size_t q = 0; q < n_q; ++q) push_results(*out_results, (uint32_t)q, per_query[q]);
    } catch (const std::exception& e) { return fail(e); }
    return 0;
}

// utils.rs:62-68 get_results / 83-89 get_results_tsv for results already in an orc_results.
// labels: '\n'-joined blob of query labels.  Returns malloc'ed NUL-terminated text (one block per query, in
// query order, lines joined by '\n', blocks joined by '\n'); caller frees with orc_free.
char* orc_format(const orc_tree* h, const orc_results* R, size_t n_q, const char* label_blob, size_t blob_len,
                 const uint64_t* seq_off, const uint8_t* codes, int tsv) {
    std::vector<std::string> labels;
    std::string blob(label_blob, blob_len);
    size_t pos = 0;
    for (size_t i = 0; i < n_q; ++i) {
        size_t nl = blob.find('\n', pos);
        if (nl == std::string::npos) nl = blob.size();
        labels.emplace_back(blob.substr(pos, nl - pos));
        pos = nl + 1;
    }
    std::string out;
    for (size_t i = 0; i < R->query.size(); ++i) {
        uint32_t q = R->query[i];
        EvaluationResult r{R->first_ref[i],
                           std::vector<double>(R->conf.begin() + i * R->max_lev, R->conf.begin() + i * R->max_lev + R->nlev[i]),
                           R->local[i], R->global[i]};
        const std::string& lineage = h->t->lineages[r.first_ref_idx];
        if (!out.empty()) out.push_back('\n');
        if (tsv) out += tsv_string(labels[q], lineage, r, decompress_sequence(codes + seq_off[q], seq_off[q + 1] - seq_off[q]));
        else out += output_string(labels[q], lineage, r);
    }
    char* p = (char*)malloc(out.size() + 1);
    memcpy(p, out.data(), out.size() + 1);
    return p;
}
void orc_free(void* p) { free(p); }

// query FASTA parsing (parser.rs:117-154).  Returns number of queries, fills caller buffers when non-NULL.
// label blob '\n'-joined.
int64_t orc_parse_queries(const char* text, size_t len, char* label_blob, size_t label_cap, size_t* label_len,
                          uint64_t* seq_off, uint8_t* codes, size_t codes_cap, size_t* codes_len) {
    try {
        auto q = parse_query_fasta_str(std::string(text, len));
        std::string lab;
        size_t total = 0;
        for (auto& p : q) {
            lab += p.first;
            lab.push_back('\n');
            total += p.second.size();
        }
        if (label_len) *label_len = lab.size();
        if (codes_len) *codes_len = total;
        if (label_blob && label_cap >= lab.size()) memcpy(label_blob, lab.data(), lab.size());
        if (seq_off && codes && codes_cap >= total) {
            size_t o = 0;
            for (size_t i = 0; i < q.size(); ++i) {
                seq_off[i] = o;
                memcpy(codes + o, q[i].second.data(), q[i].second.size());
                o += q[i].second.size();
            }
            seq_off[q.size()] = o;
        }
        return (int64_t)q.size();
    } catch (const std::exception& e) {
        fail(e);
        return -1;
    }
}

}  // extern "C"
