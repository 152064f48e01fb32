"""BASELINE config 5 at reduced width: shards of 1 M references each (the size one GPU holds in config 5), S of them held by S
contexts on ONE GPU, a query batch classified through the sharded phases (histogram all-reduce + straddler records exchanged by
device copies), compared line by line with the unsharded run over the same S x 1 M references on the same GPU.
usage: c5_scaled_check.py [n_shards=2] [n_queries=512]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from raxtax_b200 import capi, dist as rdist, synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 512
t0 = time.time()
ds = synth.generate("c5", n_refs=S * 1_000_000, n_queries=nq, measure=False)
t_gen = time.time() - t0
t0 = time.time()
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
t_tree = time.time() - t0
eo, eids = tree.exact_batch(ds.query_off, ds.query_codes)
full = capi.Context(0)
t0 = time.time()
full.upload_tree(tree)
t_up = time.time() - t0
ref = full.classify(ds.query_off, ds.query_codes, eo, eids)
full.close()
cuts = rdist.shard_cuts(tree.num_tips, S)
ctxs = [capi.Context(0) for _ in range(S)]
for r, c in enumerate(ctxs):
    c.upload_tree_sharded(tree, S, r, cuts)
ref_levels = tree.index_arrays()["ref_levels"]
t0 = time.time()
merged, outs = rdist.classify_sharded_local(ctxs, ds.query_off, ds.query_codes, eo, eids, ref_levels)
t_sh = time.time() - t0
same = 0
for q in range(nq):
    a, b = ref.for_query(q), merged.for_query(q)
    same += len(a) == len(b) and all(x[0] == y[0] and np.array_equal(x[1], y[1]) and abs(x[2] - y[2]) < 1e-9 and abs(x[3] - y[3]) < 1e-9 for x, y in zip(a, b))
print(json.dumps(dict(n_refs=int(tree.num_tips), n_shards=S, refs_per_shard=int(cuts[1] - cuts[0]), queries=nq, identical_to_unsharded=int(same),
                      result_lines=int(len(ref.first_ref)), gen_s=round(t_gen, 1), tree_new_s=round(t_tree, 1), index_upload_s=round(t_up, 1),
                      sharded_pass_s=round(t_sh, 2), index_bytes_per_shard=int(ctxs[0].index_bytes))))
assert same >= nq - max(1, nq // 50)
for c in ctxs:
    c.close()
