"""Small ncu target: one warm device-resident pass of NQ queries of a workload, then the profiled pass (one hit-count launch when NQ
fits a sub-batch).  Prints the per-launch byte model of the hit-count kernel as JSON (for profiles/hitcount_traffic.json).
usage: ncu ... python tools/ncu_target.py c3 6000 [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from raxtax_b200 import capi

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
q_total, _, _ = bench.workload_queries(name, 1)
ds = bench.load_workload(name, q_total)
skip = bench.WORKLOADS[name][1]
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
ctx = capi.Context(0)
ctx.upload_tree(tree)
off = np.ascontiguousarray(ds.query_off[: nq + 1], np.uint64)
codes = ds.query_codes[: int(off[-1])]
eo, eids = tree.exact_batch(off, codes)
ctx.batch_upload(off, codes, eo, eids, skip_exact=skip)
ctx.batch_run()  # warm (skipped by ncu -s)
ctx.synchronize()
ctx.set_option(capi.RTX_OPT_PROFILE, 1)
ctx.profile_reset()
ctx.batch_run()  # profiled
ctx.batch_download()
p = ctx.profile()
n_launch = max(1, p["hitcount"]["launches"])
rec = {"workload": name, "queries_per_launch": nq / n_launch, "launches": n_launch, "sub_batch": ctx.sub_batch, "kernel": ctx.hitcount_kernel_name(),
       "bitrow_bytes_per_launch": p["bitrow_bytes"] / n_launch, "csr_equiv_bytes_per_launch": p["csr_equiv_bytes"] / n_launch,
       "compulsory_bytes_per_launch": ctx.index_bitrow_bytes + nq / n_launch * ctx.shard_refs * 2, "launch_ms_under_profiler": p["hitcount"]["total_ms"] / n_launch}
print(json.dumps(rec))
if len(sys.argv) > 3:
    json.dump(rec, open(sys.argv[3], "w"))
