"""Independent Python model of the reference's binary database: Tree::new (tree.rs:47-140) restated with plain dicts/lists and
serialised / parsed in bincode 1.3's default layout (little-endian, fixed-width ints, usize and lengths as
u64, enum variant as u32, fields in declaration order tree.rs:36-43,181-194).  Test infrastructure only."""
import struct

INNER, TAXON, SEQUENCE = 0, 1, 2
TWO_BIT = {1: 0, 2: 1, 4: 2, 8: 3}


def tree_new(lineages, sequences):
    """-> dict(root, lineages, sequences, k_mer_map, num_tips); a node is [label, lo, hi, children, type]"""
    pairs = sorted(zip(lineages, sequences), key=lambda p: p[0].encode())  # stable, byte-wise (tree.rs:54)
    root = ["root", 0, 1, [], INNER]
    seq_map = {}
    kmap = [[] for _ in range(65536)]
    ci = 0
    for idx, (lineage, seq) in enumerate(pairs):
        levels = lineage.split(",")
        cur = root
        for level, label in enumerate(levels):
            nt = TAXON if level == len(levels) - 1 else INNER
            if not cur[3] or cur[3][-1][0] != label:
                cur[3].append([label, ci, ci + 1, [], nt])
            cur[2] = ci + 1
            if level == len(levels) - 1:
                ci += 1
            cur = cur[3][-1]
        cur[3].append([cur[0], ci - 1, ci, [], SEQUENCE])
        cur[2] = ci
        seq_map.setdefault(bytes(seq), []).append(idx)
        for i in range(len(seq) - 7):
            w = seq[i:i + 8]
            if all(c in TWO_BIT for c in w):
                k = 0
                for c in w:
                    k = (k << 2) | TWO_BIT[c]
                kmap[k].append(idx)
    root[2] = ci
    return dict(root=root, lineages=[p[0] for p in pairs], sequences=seq_map, k_mer_map=[sorted(set(l)) for l in kmap], num_tips=ci)


def _u64(v):
    return struct.pack("<Q", v)


def _node_bytes(n, out):
    lb = n[0].encode()
    out.append(_u64(len(lb)) + lb + _u64(n[1]) + _u64(n[2]) + _u64(len(n[3])))
    for c in n[3]:
        _node_bytes(c, out)
    out.append(struct.pack("<I", n[4]))


def serialize(t) -> bytes:
    out = []
    _node_bytes(t["root"], out)
    out.append(_u64(len(t["lineages"])))
    for l in t["lineages"]:
        b = l.encode()
        out.append(_u64(len(b)) + b)
    out.append(_u64(len(t["sequences"])))
    for k, v in t["sequences"].items():
        out.append(_u64(len(k)) + k + _u64(len(v)) + struct.pack(f"<{len(v)}I", *v))
    out.append(_u64(len(t["k_mer_map"])))
    for l in t["k_mer_map"]:
        out.append(_u64(len(l)) + struct.pack(f"<{len(l)}I", *l))
    out.append(_u64(t["num_tips"]))
    return b"".join(out)


class _R:
    def __init__(self, b):
        self.b, self.p = b, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.p)[0]
        self.p += 8
        return v

    def u32(self):
        v = struct.unpack_from("<I", self.b, self.p)[0]
        self.p += 4
        return v

    def take(self, n):
        v = self.b[self.p:self.p + n]
        assert len(v) == n
        self.p += n
        return v


def _node_parse(r):
    label = r.take(r.u64()).decode()
    lo, hi = r.u64(), r.u64()
    kids = [_node_parse(r) for _ in range(r.u64())]
    return [label, lo, hi, kids, r.u32()]


def deserialize(b: bytes):
    r = _R(b)
    root = _node_parse(r)
    lineages = [r.take(r.u64()).decode() for _ in range(r.u64())]
    seqs = {}
    for _ in range(r.u64()):
        k = r.take(r.u64())
        n = r.u64()
        seqs[k] = list(struct.unpack_from(f"<{n}I", r.take(4 * n)))
    kmap = []
    for _ in range(r.u64()):
        n = r.u64()
        kmap.append(list(struct.unpack_from(f"<{n}I", r.take(4 * n))))
    num_tips = r.u64()
    assert r.p == len(b), "trailing bytes"
    return dict(root=root, lineages=lineages, sequences=seqs, k_mer_map=kmap, num_tips=num_tips)
