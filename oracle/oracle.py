"""ctypes binding of the CPU oracle (oracle/raxtax_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under raxtax_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libraxtax_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raxtax_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u8p, u16p, u32p, u64p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_double))
    L.orc_last_error.restype = C.c_char_p
    L.orc_ln_binomial.restype = C.c_double
    L.orc_ln_binomial.argtypes = [C.c_uint64, C.c_uint64]
    L.orc_ln_gamma.restype = C.c_double
    L.orc_ln_gamma.argtypes = [C.c_double]
    L.orc_euclidean_norm.restype = C.c_double
    L.orc_euclidean_norm.argtypes = [f64p, C.c_size_t]
    L.orc_euclidean_distance_l1.argtypes = [f64p, f64p, C.c_size_t, f64p]
    L.orc_map_four_to_two_bit_repr.argtypes = [C.c_uint8]
    L.orc_sequence_to_kmers.restype = C.c_size_t
    L.orc_sequence_to_kmers.argtypes = [u8p, C.c_size_t, u16p]
    L.orc_map_dna.argtypes = [C.c_char_p, C.c_size_t, u8p]
    L.orc_tree_from_fasta.restype = C.c_void_p
    L.orc_tree_from_fasta.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_tree_new.restype = C.c_void_p
    L.orc_tree_new.argtypes = [C.c_size_t, C.c_char_p, C.c_size_t, u64p, u8p]
    L.orc_tree_free.argtypes = [C.c_void_p]
    L.orc_tree_num_tips.restype = C.c_size_t
    L.orc_tree_num_tips.argtypes = [C.c_void_p]
    L.orc_tree_lineage.restype = C.c_char_p
    L.orc_tree_lineage.argtypes = [C.c_void_p, C.c_size_t]
    L.orc_tree_kmer_list_len.restype = C.c_size_t
    L.orc_tree_kmer_list_len.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_tree_kmer_list.argtypes = [C.c_void_p, C.c_uint32, u32p]
    L.orc_tree_nnz.restype = C.c_uint64
    L.orc_tree_nnz.argtypes = [C.c_void_p]
    L.orc_tree_csr.argtypes = [C.c_void_p, u64p, u32p]
    L.orc_tree_sequence_len.restype = C.c_size_t
    L.orc_tree_sequence_len.argtypes = [C.c_void_p, C.c_size_t]
    L.orc_tree_sequence.argtypes = [C.c_void_p, C.c_size_t, u8p]
    L.orc_tree_exact.restype = C.c_size_t
    L.orc_tree_exact.argtypes = [C.c_void_p, u8p, C.c_size_t, u32p, C.c_size_t]
    L.orc_tree_flatten.restype = C.c_size_t
    L.orc_tree_flatten.argtypes = [C.c_void_p, u64p, u64p, u8p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), u32p,
                                   C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.orc_highest_hit_prob.argtypes = [C.c_uint16, C.c_size_t, u16p, C.c_size_t, f64p]
    L.orc_iterative_pmf_ln.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, f64p]
    L.orc_results_new.restype = C.c_void_p
    L.orc_results_new.argtypes = [C.c_int]
    L.orc_results_free.argtypes = [C.c_void_p]
    L.orc_results_len.restype = C.c_size_t
    L.orc_results_len.argtypes = [C.c_void_p]
    L.orc_results_copy.argtypes = [C.c_void_p, u32p, u32p, u8p, f64p, f64p, f64p]
    L.orc_lineage_evaluate.argtypes = [C.c_void_p, f64p, C.c_size_t, C.c_void_p]
    L.orc_classify.argtypes = [C.c_void_p, C.c_size_t, u64p, u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, u16p, u16p, f64p,
                               u16p, C.c_size_t, u32p, u8p, C.c_void_p, f64p]
    L.orc_format.restype = C.c_void_p
    L.orc_format.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, u64p, u8p, C.c_int]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_parse_queries.restype = C.c_int64
    L.orc_parse_queries.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), u64p, u8p,
                                    C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class OracleError(RuntimeError):
    pass


def _err():
    return OracleError(lib().orc_last_error().decode())


def set_flat_hist(on: bool):
    """histogram / per-count probability containers of highest_hit_prob_per_reference: flat tables (default, O(1) per reference like
    the reference's ahash maps) or the ordered std::map of the round-1 statement.  Same results either way."""
    lib().orc_set_flat_hist(int(bool(on)))


def ln_binomial(n, k):
    return lib().orc_ln_binomial(n, k)


def ln_gamma(x):
    return lib().orc_ln_gamma(x)


def map_four_to_two_bit_repr(c):
    r = lib().orc_map_four_to_two_bit_repr(c)
    return None if r < 0 else r


def map_dna(s: str) -> np.ndarray:
    b = s.encode()
    out = np.zeros(len(b), np.uint8)
    if lib().orc_map_dna(b, len(b), _p(out, C.c_uint8)) != 0:
        raise _err()
    return out


def sequence_to_kmers(codes) -> np.ndarray:
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.zeros(max(len(codes), 1), np.uint16)
    k = lib().orc_sequence_to_kmers(_p(codes, C.c_uint8), len(codes), _p(out, C.c_uint16))
    return out[:k].copy()


def euclidean_norm(v):
    v = np.ascontiguousarray(v, np.float64)
    return lib().orc_euclidean_norm(_p(v, C.c_double), len(v))


def euclidean_distance_l1(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    out = C.c_double()
    if lib().orc_euclidean_distance_l1(_p(a, C.c_double), _p(b, C.c_double), len(a), C.byref(out)) != 0:
        raise _err()
    return out.value


def highest_hit_prob_per_reference(K, t, sizes) -> np.ndarray:
    sizes = np.ascontiguousarray(sizes, np.uint16)
    out = np.zeros(len(sizes), np.float64)
    if lib().orc_highest_hit_prob(K, t, _p(sizes, C.c_uint16), len(sizes), _p(out, C.c_double)) != 0:
        raise _err()
    return out


def iterative_pmf_ln(K, t, m) -> np.ndarray:
    out = np.zeros(t + 1, np.float64)
    if lib().orc_iterative_pmf_ln(K, t, m, _p(out, C.c_double)) != 0:
        raise _err()
    return out


@dataclass
class Results:
    query: np.ndarray
    first_ref: np.ndarray
    nlev: np.ndarray
    conf: np.ndarray  # [n, max_lev]
    local: np.ndarray
    glob: np.ndarray

    def for_query(self, q):
        idx = np.nonzero(self.query == q)[0]
        return [(int(self.first_ref[i]), self.conf[i, : self.nlev[i]].copy(), float(self.local[i]), float(self.glob[i])) for i in idx]


def _take_results(h, max_lev) -> Results:
    L = lib()
    n = L.orc_results_len(h)
    r = Results(np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros((n, max_lev), np.float64),
                np.zeros(n, np.float64), np.zeros(n, np.float64))
    if n:
        L.orc_results_copy(h, _p(r.query, C.c_uint32), _p(r.first_ref, C.c_uint32), _p(r.nlev, C.c_uint8),
                           _p(r.conf, C.c_double), _p(r.local, C.c_double), _p(r.glob, C.c_double))
    return r


def pack_sequences(seqs):
    """list of uint8 code arrays -> (offsets u64[n+1], codes u8[total])"""
    off = np.zeros(len(seqs) + 1, np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs])
    codes = np.concatenate([np.asarray(s, np.uint8) for s in seqs]) if len(seqs) and off[-1] > 0 else np.zeros(0, np.uint8)
    return off, np.ascontiguousarray(codes, np.uint8)


class Tree:
    MAX_LEV = 32

    def __init__(self, handle):
        if not handle:
            raise _err()
        self._h = C.c_void_p(handle)

    @classmethod
    def from_fasta(cls, text: str) -> "Tree":
        b = text.encode()
        return cls(lib().orc_tree_from_fasta(b, len(b)))

    @classmethod
    def new(cls, lineages, sequences) -> "Tree":
        blob = "\n".join(lineages).encode()
        off, codes = pack_sequences(sequences)
        if codes.size == 0:
            codes = np.zeros(1, np.uint8)
        return cls(lib().orc_tree_new(len(lineages), blob, len(blob), _p(off, C.c_uint64), _p(codes, C.c_uint8)))

    def __del__(self):
        try:
            if self._h:
                lib().orc_tree_free(self._h)
        except Exception:
            pass

    @property
    def num_tips(self):
        return lib().orc_tree_num_tips(self._h)

    @property
    def lineages(self):
        return [lib().orc_tree_lineage(self._h, i).decode() for i in range(self.num_tips)]

    def lineage(self, i):
        return lib().orc_tree_lineage(self._h, i).decode()

    def k_mer_map(self, kmer) -> np.ndarray:
        n = lib().orc_tree_kmer_list_len(self._h, kmer)
        out = np.zeros(max(n, 1), np.uint32)
        lib().orc_tree_kmer_list(self._h, kmer, _p(out, C.c_uint32))
        return out[:n]

    def csr(self):
        nnz = lib().orc_tree_nnz(self._h)
        off = np.zeros(65537, np.uint64)
        ids = np.zeros(max(nnz, 1), np.uint32)
        lib().orc_tree_csr(self._h, _p(off, C.c_uint64), _p(ids, C.c_uint32))
        return off, ids[:nnz]

    def sequence(self, i) -> np.ndarray:
        n = lib().orc_tree_sequence_len(self._h, i)
        out = np.zeros(max(n, 1), np.uint8)
        lib().orc_tree_sequence(self._h, i, _p(out, C.c_uint8))
        return out[:n]

    def exact(self, codes) -> np.ndarray:
        codes = np.ascontiguousarray(codes, np.uint8)
        buf = np.zeros(64, np.uint32)
        n = lib().orc_tree_exact(self._h, _p(codes, C.c_uint8), len(codes), _p(buf, C.c_uint32), len(buf))
        if n > len(buf):
            buf = np.zeros(n, np.uint32)
            lib().orc_tree_exact(self._h, _p(codes, C.c_uint8), len(codes), _p(buf, C.c_uint32), len(buf))
        return buf[:n].copy()

    def flatten(self):
        """pre-order node arrays of the full Node tree (incl. Sequence nodes)"""
        L = lib()
        ll = C.c_size_t()
        n = L.orc_tree_flatten(self._h, None, None, None, None, None, None, None, 0, C.byref(ll))
        lo, hi = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        ty, dep, par, nch = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int64), np.zeros(n, np.uint32)
        lab = C.create_string_buffer(ll.value + 1)
        L.orc_tree_flatten(self._h, _p(lo, C.c_uint64), _p(hi, C.c_uint64), _p(ty, C.c_uint8), _p(dep, C.c_int32),
                           _p(par, C.c_int64), _p(nch, C.c_uint32), lab, ll.value, C.byref(ll))
        labels = lab.raw[: ll.value].decode().split("\n")[:-1]
        return dict(lo=lo, hi=hi, type=ty, depth=dep, parent=par, nchildren=nch, labels=labels)

    def evaluate(self, confidences) -> Results:
        cv = np.ascontiguousarray(confidences, np.float64)
        h = C.c_void_p(lib().orc_results_new(self.MAX_LEV))
        try:
            if lib().orc_lineage_evaluate(self._h, _p(cv, C.c_double), len(cv), h) != 0:
                raise _err()
            return _take_results(h, self.MAX_LEV)
        finally:
            lib().orc_results_free(h)

    def classify(self, seq_off, codes, skip_exact=False, raw_conf=False, threads=1, chunk_size=0, want_counts=False,
                 want_probs=False, want_kmers=False):
        """raxtax.rs:14-97 over a batch.  Returns dict with K, results, seconds and the requested taps."""
        L = lib()
        seq_off = np.ascontiguousarray(seq_off, np.uint64)
        codes = np.ascontiguousarray(codes, np.uint8)
        if codes.size == 0:
            codes = np.zeros(1, np.uint8)
        nq = len(seq_off) - 1
        N = self.num_tips
        K = np.zeros(nq, np.uint16)
        counts = np.zeros((nq, N), np.uint16) if want_counts else None
        probs = np.zeros((nq, N), np.float64) if want_probs else None
        maxlen = int((seq_off[1:] - seq_off[:-1]).max()) if nq else 0
        kstride = max(maxlen - 7, 1)
        kmers = np.zeros((nq, kstride), np.uint16) if want_kmers else None
        nexact = np.zeros(nq, np.uint32)
        warn = np.zeros(nq, np.uint8)
        secs = C.c_double()
        h = C.c_void_p(L.orc_results_new(self.MAX_LEV))
        try:
            rc = L.orc_classify(self._h, nq, _p(seq_off, C.c_uint64), _p(codes, C.c_uint8), int(skip_exact), int(raw_conf),
                                int(threads), int(chunk_size), _p(K, C.c_uint16), _p(counts, C.c_uint16),
                                _p(probs, C.c_double), _p(kmers, C.c_uint16), kstride, _p(nexact, C.c_uint32),
                                _p(warn, C.c_uint8), h, C.byref(secs))
            if rc != 0:
                raise _err()
            res = _take_results(h, self.MAX_LEV)
        finally:
            L.orc_results_free(h)
        return dict(K=K, counts=counts, probs=probs, kmers=kmers, nexact=nexact, warn=warn, results=res, seconds=secs.value)


def format_results(tree: Tree, results: Results, labels, seq_off=None, codes=None, tsv=False):
    """lineage.rs:17-48 formatting, in Python (mirrors orc_format; used to produce expected text in tests)."""
    lines = []
    for i in range(len(results.query)):
        q = int(results.query[i])
        lin = tree.lineage(int(results.first_ref[i]))
        conf = ["%.2f" % v for v in results.conf[i, : results.nlev[i]]]
        if not tsv:
            lines.append("%s\t%s\t%s\t%.5f\t%.5f" % (labels[q], lin, ",".join(conf), results.local[i], results.glob[i]))
        else:
            a, b = lin.split(","), conf
            inter, ia, ib, flag = [], 0, 0, False
            while ia < len(a) or ib < len(b):
                if not flag:
                    if ia < len(a):
                        inter.append(a[ia]); ia += 1
                    else:
                        inter.append(b[ib]); ib += 1
                else:
                    if ib < len(b):
                        inter.append(b[ib]); ib += 1
                    else:
                        inter.append(a[ia]); ia += 1
                flag = not flag
            seq = "".join({1: "A", 2: "C", 4: "G", 8: "T"}.get(int(c), "-") for c in codes[int(seq_off[q]): int(seq_off[q + 1])])
            lines.append("%s\t%s\t%.5f\t%.5f\t%s" % (labels[q], "\t".join(inter), results.local[i], results.glob[i], seq))
    return "\n".join(lines)


def parse_queries(text: str):
    L = lib()
    b = text.encode()
    ll, cl = C.c_size_t(), C.c_size_t()
    n = L.orc_parse_queries(b, len(b), None, 0, C.byref(ll), None, None, 0, C.byref(cl))
    if n < 0:
        raise _err()
    lab = C.create_string_buffer(ll.value + 1)
    off = np.zeros(n + 1, np.uint64)
    codes = np.zeros(max(cl.value, 1), np.uint8)
    L.orc_parse_queries(b, len(b), lab, ll.value, C.byref(ll), _p(off, C.c_uint64), _p(codes, C.c_uint8), cl.value, C.byref(cl))
    labels = lab.raw[: ll.value].decode().split("\n")[:-1]
    return labels, off, codes[: cl.value]
