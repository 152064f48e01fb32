"""The binary database (tree.rs:146-164): the C++ host library's reader and writer against an independent
Python model of Tree::new + bincode 1.3 (tests/bincode_model.py).  No GPU."""
import numpy as np
import pytest

from raxtax_b200 import _build, capi, synth
from tests import bincode_model as bm
from tests.test_oracle_kats import REF_FASTA_STR_PARSER


@pytest.fixture(scope="module", autouse=True)
def _built():
    _build.build_all()


def _parse_fasta(text):
    lin, seqs = [], []
    for line in text.strip().splitlines():
        line = line.strip()
        if line.startswith(">"):
            lin.append(line.split("tax=")[1].split(";")[0])
            seqs.append([])
        elif line:
            seqs[-1].extend(int(c) for c in capi.map_dna(line)) if hasattr(capi, "map_dna") else seqs[-1].extend(
                {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15, "W": 9, "S": 6, "M": 3, "K": 12, "R": 5, "Y": 10, "B": 14, "D": 13, "H": 11, "V": 7}[c]
                for c in line.upper())
    return lin, seqs


def _datasets():
    lin, seqs = _parse_fasta(REF_FASTA_STR_PARSER)
    yield "parser.rs KAT", lin, seqs
    ds = synth.generate("tiny", measure=False)
    yield "tiny", list(ds.ref_lineages), [list(map(int, ds.ref_seq(i))) for i in range(ds.n_refs)]
    # variable depth, identical lineages, duplicate sequences across taxa, a rank that repeats its parent's label
    lin = ["a,b,c", "a,b", "a,b,c", "a,d", "a,d,d", "e", "a,b,c"]
    seqs = [[1, 2, 4, 8, 1, 2, 4, 8, 1], [1, 2, 4, 8, 1, 2, 4, 8, 1], [2] * 12, [15] * 9, [1, 2, 4, 8, 1, 2, 4, 8, 1], [], [8] * 8]
    yield "ragged", lin, seqs


def _host_tree(lin, seqs):
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    codes = np.array([c for s in seqs for c in s], np.uint8)
    return capi.Tree.new(lin, off, codes)


def _same_tree(a: capi.Tree, b: capi.Tree, seqs):
    assert a.num_tips == b.num_tips and a.lineages == b.lineages
    ao, ai = a.csr()
    bo, bi = b.csr()
    assert np.array_equal(ao, bo) and np.array_equal(ai, bi)
    ia, ib = a.index_arrays(), b.index_arrays()
    for k in ("node_lo", "node_hi", "node_type", "child_first", "child_count", "ref_levels"):
        assert np.array_equal(ia[k], ib[k]), k
    for s in seqs:
        s = np.array(s, np.uint8)
        assert np.array_equal(a.exact(s), b.exact(s))


@pytest.mark.parametrize("name,lin,seqs", list(_datasets()), ids=lambda x: x if isinstance(x, str) else "")
def test_reader_accepts_the_reference_layout(name, lin, seqs):
    """bytes produced by the Python model of Tree::new + bincode -> rxh_tree_from_bin == rxh_tree_new on the same input"""
    model = bm.tree_new(lin, seqs)
    loaded = capi.Tree.from_bin(bm.serialize(model))
    assert loaded is not None and loaded.has_kmer_map
    _same_tree(loaded, _host_tree(lin, seqs), seqs)


@pytest.mark.parametrize("name,lin,seqs", list(_datasets()), ids=lambda x: x if isinstance(x, str) else "")
def test_writer_emits_the_reference_layout(name, lin, seqs, tmp_path):
    """rxh_tree_save_bin -> independent parse == the Python model (node tree incl. Sequence leaves, lineages, sequence map, k_mer_map)"""
    p = str(tmp_path / "db.bin")
    _host_tree(lin, seqs).save_bin(p)
    got = bm.deserialize(open(p, "rb").read())
    want = bm.tree_new(lin, seqs)
    assert got["root"] == want["root"]
    assert got["lineages"] == want["lineages"] and got["num_tips"] == want["num_tips"]
    assert got["k_mer_map"] == want["k_mer_map"]
    assert {k: sorted(v) for k, v in got["sequences"].items()} == {k: sorted(v) for k, v in want["sequences"].items()}
    # and it loads back
    _same_tree(capi.Tree.from_bin(open(p, "rb").read()), _host_tree(lin, seqs), seqs)


def test_fasta_and_garbage_are_not_databases():
    assert capi.Tree.from_bin(REF_FASTA_STR_PARSER.encode()) is None
    assert capi.Tree.from_bin(b"") is None
    good = bm.serialize(bm.tree_new(*_parse_fasta(REF_FASTA_STR_PARSER)))
    assert capi.Tree.from_bin(good[:-9]) is None  # truncated
    rng = np.random.default_rng(5)
    for _ in range(50):  # corrupted lengths must fail cleanly, never crash
        b = bytearray(good)
        pos = int(rng.integers(0, len(b) - 8))
        b[pos:pos + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
        t = capi.Tree.from_bin(bytes(b))
        assert t is None or t.num_tips == 6


def test_queries_skip():
    q = capi.Queries.from_fasta(">q1\nACGT\n>q2 x\nAAAA\n>q3\nCCCC\n")
    q.skip(["q2 x", "nope"])
    assert q.labels == ["q1", "q3"]
    off, codes = q.arrays()
    assert list(off) == [0, 4, 8] and list(codes) == [1, 2, 4, 8, 2, 2, 2, 2]


def _random_tree(seed):
    """Random lineages of 1-7 ranks over a 5-letter alphabet (repeated labels, strict prefixes, duplicates) with short random sequences."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 60))
    alphabet = ["a", "b", "c", "dd", "e"]
    lin = [",".join(alphabet[int(rng.integers(0, 5))] for _ in range(int(rng.integers(1, 8)))) for _ in range(n)]
    pool = [list(map(int, rng.choice([1, 2, 4, 8, 15], int(rng.integers(0, 30)), p=[0.24, 0.24, 0.24, 0.24, 0.04]))) for _ in range(max(1, n // 2))]
    seqs = [pool[int(rng.integers(0, len(pool)))] for _ in range(n)]
    return lin, seqs


@pytest.mark.parametrize("seed", range(40))
def test_random_trees_round_trip(seed, tmp_path):
    lin, seqs = _random_tree(seed)
    model = bm.tree_new(lin, seqs)
    host = _host_tree(lin, seqs)
    _same_tree(capi.Tree.from_bin(bm.serialize(model)), host, seqs)
    p = str(tmp_path / "db.bin")
    host.save_bin(p)
    got = bm.deserialize(open(p, "rb").read())
    assert got["root"] == model["root"] and got["lineages"] == model["lineages"] and got["k_mer_map"] == model["k_mer_map"]
    assert {k: sorted(v) for k, v in got["sequences"].items()} == {k: sorted(v) for k, v in model["sequences"].items()}
