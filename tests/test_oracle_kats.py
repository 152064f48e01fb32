"""Pin the CPU oracle against every known-answer test the reference holds for the hot path.

Each test names the reference test it restates (file:line under /root/reference).  Nothing here reads
/root/reference at run time: the fixtures are the literals of the reference's own unit tests.
"""
import math

import numpy as np
import pytest

REF_FASTA_STR_PARSER = """>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species1;
AAACCCTTTGGGA
>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species2;
ATACGCTTTGGGA
>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order4,f:Family5,g:Genus2,s:Species3;
ATCCGCTATGGGA
>Badabing|Badabum;tax=p:Phylum1,c:Class2,o:Order2,f:Family3,g:Genus3,s:Species6;
ATACGCTTTGCGT
>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species2;
GTGCGCTATGCGA
>Badabing|Badabum;tax=p:Phylum2,c:Class3,o:Order3,f:Family4,g:Genus4,s:Species5;
ATACGCTTTGCGT"""

REF_FASTA_KMERS = """>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species1;
AAACCCCGT
>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species1;
TAACCCCGG
>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus2,s:Species3;
TTTAAAACC
>Badabing|Badabum;tax=p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus2,s:Species3;
TTTAAAACA
>Badabing|Badabum;tax=p:Phylum1,c:Class2,o:Order2,f:Family2,g:Genus3,s:Species4;
AAACCCCGG"""


# ---- utils.rs ---------------------------------------------------------------------------------
def test_euclidean_norm(oracle):  # utils.rs:208-214
    assert oracle.euclidean_norm([1.0, 2.0, 3.0, 4.0]) == pytest.approx(math.sqrt(30.0), abs=1e-7)
    assert oracle.euclidean_norm([0.5, 0.5, 0.25, 0.2]) == pytest.approx(math.sqrt(0.6025), abs=1e-7)


def test_euclidean_distance(oracle):  # utils.rs:216-224
    assert oracle.euclidean_distance_l1([1.0, 0.0, 0.0], [0.0, 1.0, 0.0]) == pytest.approx(math.sqrt(2.0), abs=1e-7)
    assert oracle.euclidean_distance_l1([0.5, 0.1, 0.1], [1.0, 1.0, 0.5]) == pytest.approx(0.4100771455544949, abs=1e-7)


def test_map(oracle):  # utils.rs:236-243
    assert oracle.map_four_to_two_bit_repr(1) == 0
    assert oracle.map_four_to_two_bit_repr(2) == 1
    assert oracle.map_four_to_two_bit_repr(4) == 2
    assert oracle.map_four_to_two_bit_repr(8) == 3
    assert oracle.map_four_to_two_bit_repr(10) is None


KMER_KAT_CODES = [1, 2, 1, 4, 8, 2, 8, 4, 1, 4, 8, 2, 8, 4, 1, 4]
KMER_KAT_EXPECTED = [
    0b0001_0010_1101_1110,
    0b0010_1101_1110_0010,
    0b0100_1011_0111_1000,
    0b0111_1000_1011_0111,
    0b1000_1011_0111_1000,
    0b1011_0111_1000_1011,
    0b1101_1110_0010_1101,
    0b1110_0010_1101_1110,
]


def test_sequence_to_kmers(oracle):  # utils.rs:245-263
    kmers = oracle.sequence_to_kmers(KMER_KAT_CODES)
    assert list(kmers) == KMER_KAT_EXPECTED


def test_sequence_to_kmers_edges(oracle):
    assert len(oracle.sequence_to_kmers([1] * 7)) == 0  # shorter than one window
    assert len(oracle.sequence_to_kmers([])) == 0
    assert list(oracle.sequence_to_kmers([1] * 8)) == [0]
    assert list(oracle.sequence_to_kmers([8] * 30)) == [0xFFFF]  # dedup
    assert len(oracle.sequence_to_kmers([1] * 7 + [15] + [1] * 7)) == 0  # every window holds the ambiguous base
    assert len(oracle.sequence_to_kmers([0] * 9)) == 0  # lineage.rs tests use code 0 sequences


# ---- parser.rs --------------------------------------------------------------------------------
def test_str_parser(oracle):  # parser.rs:166-217
    tree = oracle.Tree.from_fasta(REF_FASTA_STR_PARSER)
    assert list(tree.k_mer_map(0b1_0101_1111_1110)) == [0]
    assert list(tree.k_mer_map(0b11_0001_1001_1111)) == [1, 4, 5]
    assert list(tree.k_mer_map(0b110_0111_0011_1010)) == [3]
    assert tree.num_tips == 6
    assert tree.lineages == [
        "p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species1",
        "p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species2",
        "p:Phylum1,c:Class1,o:Order1,f:Family1,g:Genus1,s:Species2",
        "p:Phylum1,c:Class1,o:Order4,f:Family5,g:Genus2,s:Species3",
        "p:Phylum1,c:Class2,o:Order2,f:Family3,g:Genus3,s:Species6",
        "p:Phylum2,c:Class3,o:Order3,f:Family4,g:Genus4,s:Species5",
    ]


def test_query_parser(oracle):  # parser.rs:219-233
    _, off, codes = oracle.parse_queries(">label1\nAAACCCTTTGGGA")
    assert list(codes[off[0]:off[1]]) == [1, 1, 1, 2, 2, 2, 8, 8, 8, 4, 4, 4, 1]
    labels, off, codes = oracle.parse_queries(">label1\nACGTWSMKRYBDHVN")
    assert labels == ["label1"]
    assert list(codes[off[0]:off[1]]) == [1, 2, 4, 8, 9, 6, 3, 12, 5, 10, 14, 13, 11, 7, 15]


def test_kmers(oracle):  # parser.rs:235-299
    tree = oracle.Tree.from_fasta(REF_FASTA_KMERS)
    assert list(tree.k_mer_map(0b1_0101_0110)) == [0, 4]
    assert list(tree.k_mer_map(0b101_0101_1010)) == [1, 4]
    assert list(tree.k_mer_map(0b101_0101_1011)) == [0]
    assert list(tree.k_mer_map(0b1100_0001_0101_0110)) == [1]
    assert list(tree.k_mer_map(0b1111_0000_0000_0101)) == [2]
    assert list(tree.k_mer_map(0b1111_1100_0000_0001)) == [2, 3]


def test_parser_errors(oracle):  # parser.rs:47-49,58-60,32,79-83,96-98
    with pytest.raises(oracle.OracleError, match="File is empty"):
        oracle.Tree.from_fasta("")
    with pytest.raises(oracle.OracleError, match="Not a valid FASTA"):
        oracle.Tree.from_fasta("ACGT\n>x;tax=a,b;\nACGT")
    with pytest.raises(oracle.OracleError, match="Unexpected character"):
        oracle.Tree.from_fasta(">x;tax=a,b;\nACGU")
    with pytest.raises(oracle.OracleError, match="taxonomical annotation"):
        oracle.Tree.from_fasta(">x;taxon=a,b\nACGT")
    with pytest.raises(oracle.OracleError, match="does not match"):
        oracle.Tree.from_fasta(">x;tax=a,b;\n>y;tax=a,c;\nACGT")


def test_parser_multiline_comments_case(oracle):  # parser.rs:53-57: trim, ';' comment lines, multi-line records, lower case
    text = ";comment\n>r1;tax=a,b;\nacgt\n  ACGTAC  \n\n;another\n>r2;tax=a,c;\r\nTTTTTTTTT\r\n"
    tree = oracle.Tree.from_fasta(text)
    assert tree.num_tips == 2
    assert list(tree.sequence(0)) == [1, 2, 4, 8, 1, 2, 4, 8, 1, 2]
    assert list(tree.k_mer_map(0xFFFF)) == [1]


# ---- lineage.rs -------------------------------------------------------------------------------
def _eval(oracle, lineages, confidences):
    tree = oracle.Tree.new(lineages, [[0] * 9 for _ in lineages])
    res = tree.evaluate(confidences)
    return [(tree.lineage(int(res.first_ref[i])), [float(x) for x in res.conf[i, : res.nlev[i]]]) for i in range(len(res.query))]


def test_tree_construction(oracle):  # lineage.rs:191-239
    lineages = [
        "Animalia,Chordata,Mammalia,Primates,Hominidae,Homo",
        "Animalia,Chordata,Mammalia,Primates,Hominidae,Pan",
        "Animalia,Chordata,Mammalia,Carnivora,Canidae,Canis",
        "Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis",
        "Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis",
    ]
    # NB: confidences are given in SORTED reference order (Tree::new sorts the lineages)
    got = _eval(oracle, lineages, [0.1, 0.3, 0.4, 0.004, 0.004])
    assert got == [
        ("Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis", [0.81, 0.81, 0.81, 0.8, 0.7, 0.7]),
        ("Animalia,Chordata,Mammalia,Carnivora,Canidae,Canis", [0.81, 0.81, 0.81, 0.8, 0.1, 0.1]),
        ("Animalia,Chordata,Mammalia,Primates,Hominidae,Pan", [0.81, 0.81, 0.81, 0.01, 0.01, 0.01]),
    ]


def test_variable_lineage_length(oracle):  # lineage.rs:241-302
    lineages = [
        "Animalia,Chordata,Mammalia,Primates,Hominidae,Homo,Homo_sapiens",
        "Animalia,Chordata,Mammalia,Primates,Hominidae,Pan",
        "Animalia,Chordata,Mammalia,Carnivora,Canidae,Canis",
        "Animalia,Chordata,Mammalia,Carnivora,Doggo",
        "Animalia,Chordata,Mammalia,Mouse",
        "Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis",
        "Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis",
    ]
    got = _eval(oracle, lineages, [0.05, 0.1, 0.3, 0.4, 0.1, 0.004, 0.004])
    assert got == [
        ("Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis", [0.96, 0.96, 0.96, 0.85, 0.7, 0.7]),
        ("Animalia,Chordata,Mammalia,Carnivora,Doggo", [0.96, 0.96, 0.96, 0.85, 0.1]),
        ("Animalia,Chordata,Mammalia,Carnivora,Canidae,Canis", [0.96, 0.96, 0.96, 0.85, 0.05, 0.05]),
        ("Animalia,Chordata,Mammalia,Mouse", [0.96, 0.96, 0.96, 0.1]),
        ("Animalia,Chordata,Mammalia,Primates,Hominidae,Pan", [0.96, 0.96, 0.96, 0.01, 0.01, 0.01]),
    ]


def test_likelihood_edge_case(oracle):  # lineage.rs:304-334
    lineages = [
        "Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis",
        "Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis_ferrocius",
        "Animalia,Chordata,Mammalia,Carnivora,Canidae,Canis",
    ]
    got = _eval(oracle, lineages, [0.004, 0.004, 0.004])
    assert got == [("Animalia,Chordata,Mammalia,Carnivora,Felidae,Felis_ferrocius", [0.01] * 6)]


# ---- prob.rs ----------------------------------------------------------------------------------
def _closed_form_pmf(oracle, K, i, t, m, T):  # prob.rs:183-206 (test helper `pmf`)
    if m == K:
        return 1.0 if i == t else 0.0
    if m == 0:
        return 1.0 if i == 0 else 0.0
    a = oracle.ln_binomial(m + i - 1, i)
    b = oracle.ln_binomial((K - m) + (t - i) - 1, t - i)
    return math.exp(a + b - T)


def test_pmf(oracle):  # prob.rs:208-227
    T = oracle.ln_binomial(200 + 32 - 1, 32)
    p = oracle.iterative_pmf_ln(200, 32, 50)
    p2 = [_closed_form_pmf(oracle, 200, i, 32, 50, T) for i in range(33)]
    assert np.exp(p).sum() == pytest.approx(1.0, abs=1e-7)
    assert sum(p2) == pytest.approx(1.0, abs=1e-7)
    for a, b in zip(p, p2):
        assert math.exp(a) == pytest.approx(b, abs=1e-7)


def test_hit_prob(oracle):  # prob.rs:229-235
    probs = oracle.highest_hit_prob_per_reference(400, 200, np.arange(401, dtype=np.uint16))
    assert probs.sum() == pytest.approx(1.0, abs=1e-7)
    assert np.all(probs[:-1] <= probs[1:])
    # SURVEY.md Appendix B (derived from an independent scratch restatement, not from the Rust binary)
    assert probs[-3:] == pytest.approx([0.148515837432, 0.223146911519, 0.335], abs=1e-9)


def test_ln_binomial_against_lgamma(oracle):
    """statrs restatement (Lanczos g=10.900511) must agree with libm lgamma: catches a mistyped coefficient."""
    for n, k in [(10, 3), (170, 85), (171, 1), (964, 321), (2238, 746), (98301, 32767), (5, 6)]:
        got = oracle.ln_binomial(n, k)
        if k > n:
            assert got == -math.inf
            continue
        exp = math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1)
        assert got == pytest.approx(exp, rel=1e-12, abs=1e-10)
    for x in [0.1, 0.5, 1.0, 2.5, 171.0, 1e3, 1e5]:
        assert oracle.ln_gamma(x) == pytest.approx(math.lgamma(x), rel=1e-13, abs=1e-13)


def test_hit_prob_edge_cases(oracle):
    # K == 0 (query shorter than 8 / all ambiguous): u64 wrap in prob.rs:20-23, fast branch, uniform output
    p = oracle.highest_hit_prob_per_reference(0, 0, np.zeros(5, np.uint16))
    assert p == pytest.approx([0.2] * 5)
    # K == 1, no hit anywhere: slow branch with the m == 0 row only
    p = oracle.highest_hit_prob_per_reference(1, 0, np.zeros(4, np.uint16))
    assert p == pytest.approx([0.25] * 4)
    # a full hit forces the fast branch: zero-count references get exactly 0
    p = oracle.highest_hit_prob_per_reference(10, 5, np.array([10, 0, 3], np.uint16))
    assert p[1] == 0.0 and p[0] > p[2] > 0.0 and p.sum() == pytest.approx(1.0)
    # all references hit zero k-mers while K > 0: slow branch, uniform
    p = oracle.highest_hit_prob_per_reference(20, 10, np.zeros(3, np.uint16))
    assert p == pytest.approx([1 / 3] * 3)


# ---- golden end-to-end text of the oracle on real data (tests/golden/make_golden_outputs.py) ----------------------
def test_oracle_reproduces_golden_outputs(oracle):
    import os

    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    text = open(os.path.join(g, "diptera_sample.fasta")).read()
    tree = oracle.Tree.from_fasta(text)
    labels, off, codes = oracle.parse_queries(text)
    assert tree.num_tips == 400 and len(labels) == 400
    for skip, name in ((False, "default"), (True, "skip")):
        out = tree.classify(off, codes, skip_exact=skip, threads=2, chunk_size=32)
        txt = oracle.format_results(tree, out["results"], labels) + "\n"
        assert txt == open(os.path.join(g, f"diptera_sample.{name}.out")).read()
