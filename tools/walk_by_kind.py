"""Per-kernel times by query kind (synthetic query mix: kind0 exact copies, kind1 near copies, kind2 unseen species, kind3 junk) and walk
variant.  usage: python tools/walk_by_kind.py [workload] [n_queries] [variants, e.g. 0,3]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from raxtax_b200 import capi, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
variants = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,3").split(",")]
ds = synth.generate(name, n_queries=nq, measure=False)
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
ctx = capi.Context(0)
ctx.upload_tree(tree)
kinds = np.array([int(l.rsplit("kind", 1)[1]) for l in ds.query_labels])
lens = (ds.query_off[1:] - ds.query_off[:-1]).astype(np.int64)
for kind in (None, 0, 1, 2, 3):
    sel = np.arange(ds.n_queries) if kind is None else np.nonzero(kinds == kind)[0]
    off = np.zeros(len(sel) + 1, np.uint64)
    off[1:] = np.cumsum(lens[sel])
    codes = np.concatenate([ds.query_seq(int(i)) for i in sel]) if len(sel) else np.zeros(0, np.uint8)
    eo, eids = tree.exact_batch(off, codes)
    for wv in variants:
        ctx.set_option(capi.RTX_OPT_WALK_VARIANT, wv)
        ctx.batch_upload(off, codes, eo, eids)
        ctx.set_option(capi.RTX_OPT_PROFILE, 0)
        ctx.batch_run()
        ctx.synchronize()
        ctx.set_option(capi.RTX_OPT_PROFILE, 1)
        ctx.profile_reset()
        for _ in range(2):
            ctx.batch_run()
        out = ctx.batch_download()
        p = ctx.profile()
        print(json.dumps(dict(kind="all" if kind is None else kind, n=len(sel), walk_variant=wv, lines=int(len(out.first_ref)),
                              **{k: round(p[k]["total_ms"] / 2, 3) for k in ("hitcount", "prob", "prefix", "walk")})), flush=True)
