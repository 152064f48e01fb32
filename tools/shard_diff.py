"""Why a reference-sharded run can differ from the unsharded one: the same queries of a workload through N shards on ONE GPU (staged
exchanges, the same kernels as the NCCL path) and through an unsharded context; every differing query is printed and judged against the
CPU oracle with the tolerant checker (fallback ties / rounding boundaries, tests/parity.py).
usage: python tools/shard_diff.py c3 8 16384   |   python tools/shard_diff.py c2:12000:1500 8 1500"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from raxtax_b200 import capi
from raxtax_b200 import dist as rdist

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_shards = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
if ":" in name:  # name:n_refs:n_queries -- a reduced data set, e.g. c2:12000:1500 (tests/test_gpu_nccl.py)
    from raxtax_b200 import synth

    name, n_refs, n_q = name.split(":")
    ds = synth.generate(name, n_refs=int(n_refs), n_queries=int(n_q), measure=False)
    nq = min(nq, ds.n_queries)
else:
    q_total, _, _ = bench.workload_queries(name, 1)
    ds = bench.load_workload(name, q_total)
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
off = np.ascontiguousarray(ds.query_off[: nq + 1], np.uint64)
codes = ds.query_codes[: int(off[-1])]
eo, eids = tree.exact_batch(off, codes)
c0 = capi.Context(0)
c0.upload_tree(tree)
want = c0.classify(off, codes, eo, eids, taps=())
c0.close()
N = tree.num_tips
cuts = np.array([N * r // n_shards for r in range(n_shards)] + [N], np.uint64)
ctxs = [capi.Context(0) for _ in range(n_shards)]
for r, c in enumerate(ctxs):
    c.upload_tree_sharded(tree, n_shards, r, cuts)
    c.set_option(capi.RTX_OPT_SUB_BATCH, 2048)
ref_levels = tree.index_arrays()["ref_levels"] if False else None
import ctypes as C
d = capi.IndexDesc()
capi.host_lib().rxh_tree_index_desc(tree._h, C.byref(d))
ref_levels = np.ctypeslib.as_array(d.ref_levels, (N,)).copy()
diff = []
CH = 2048
got_lines = {}
for c0q in range(0, nq, CH):
    c1q = min(nq, c0q + CH)
    o2 = (off[c0q: c1q + 1] - off[c0q]).astype(np.uint64)
    cd = codes[int(off[c0q]): int(off[c1q])]
    e2 = (eo[c0q: c1q + 1] - eo[c0q]).astype(np.uint32)
    ei = eids[int(eo[c0q]): int(eo[c1q])]
    merged, _ = rdist.classify_sharded_local(ctxs, o2, cd, e2, ei, ref_levels)
    for q in range(c1q - c0q):
        a, b = merged.for_query(q), want.for_query(c0q + q)
        same = len(a) == len(b) and all(x[0] == y[0] and len(x[1]) == len(y[1]) and np.array_equal(x[1], y[1]) and abs(x[2] - y[2]) <= 1e-9 and abs(x[3] - y[3]) <= 1e-9
                                        for x, y in zip(a, b))
        if not same:
            diff.append(c0q + q)
            got_lines[c0q + q] = (a, b)
print(f"{name}: {n_shards} shards vs unsharded on {nq} queries: {len(diff)} queries differ: {diff[:20]}", flush=True)
for q in diff[:6]:
    a, b = got_lines[q]
    print(" query", q, ds.query_labels[q])
    print("   sharded  :", [(x[0], list(np.round(x[1], 2)), round(x[2], 6)) for x in a][:6])
    print("   unsharded:", [(x[0], list(np.round(x[1], 2)), round(x[2], 6)) for x in b][:6])
if diff and os.environ.get("SHARD_DIFF_ORACLE", "1") != "0":
    from oracle import oracle as orc
    from tests import parity

    ot = parity.oracle_tree_from_ds(orc, ds)
    checker = parity.TolerantChecker(ot.flatten(), ot.num_tips)
    for q in diff[:12]:
        o = ot.classify(off[q: q + 2] - off[q], codes[int(off[q]): int(off[q + 1])], threads=1, want_probs=True)
        a, b = got_lines[q]
        ora = o["results"].for_query(0)
        print(f" query {q}: oracle == unsharded {parity.results_equal(ora, b)}, oracle == sharded {parity.results_equal(ora, a)}, "
              f"sharded acceptable {checker.acceptable(o['probs'][0], a, ora)}, unsharded acceptable {checker.acceptable(o['probs'][0], b, ora)}", flush=True)
