"""GPU micro-benchmark: hit-count kernel time on the C2 workload for each geometry (RTX_OPT_HITCOUNT_TUNE)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from raxtax_b200 import capi, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
tunes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 804, 504, 404, 802, 502, 602, 702, 402, 812, 512, 412]
ds = synth.generate(name, measure=False)
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
ctx = capi.Context(0)
ctx.upload_tree(tree)
eo, eids = tree.exact_batch(ds.query_off, ds.query_codes)
ctx.batch_upload(ds.query_off, ds.query_codes, eo, eids)
ref = None
for t in tunes:
    ctx.set_option(capi.RTX_OPT_HITCOUNT_TUNE, t % 1000)
    ctx.set_option(capi.RTX_OPT_HITCOUNT_MAX_TILES, t // 1000)
    ctx.set_option(capi.RTX_OPT_PROFILE, 0)
    ctx.batch_run()
    ctx.synchronize()
    ctx.set_option(capi.RTX_OPT_PROFILE, 1)
    ctx.profile_reset()
    for _ in range(3):
        ctx.batch_run()
    out = ctx.batch_download(taps=("hist",))
    p = ctx.profile()
    chk = int(out.hist.astype(np.int64).sum()), int((out.hist.astype(np.int64) * np.arange(out.hist.shape[1])).sum())
    if ref is None:
        ref = chk
    ms = p["hitcount"]["total_ms"] / p["hitcount"]["launches"]
    gbs = p["bitrow_bytes"] / p["hitcount"]["launches"] / ms / 1e6
    print(json.dumps(dict(tune=t, hitcount_ms=round(ms, 3), bitrow_GBps=round(gbs, 1), prob_ms=round(p["prob"]["total_ms"] / p["prob"]["launches"], 3), walk_ms=round(p["walk"]["total_ms"] / max(p["walk"]["launches"], 1), 3),
                          checksum_ok=chk == ref)), flush=True)
