"""Where the time of rxh_raxtax goes (RXH_TIMING stage clocks) next to the device-resident time of the same queries.
usage: python tools/e2e_probe.py c2|c3 [chunk ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from raxtax_b200 import capi

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
chunks = [int(x) for x in sys.argv[2:]] or [0]
q_total, nq, _ = bench.workload_queries(name, 1)
ds = bench.load_workload(name, q_total)
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
ctx = capi.Context(0)
ctx.upload_tree(tree)
off, codes = ds.query_off, ds.query_codes
eo, eids = tree.exact_batch(off, codes)
ctx.batch_upload(off, codes, eo, eids)
for _ in range(2):
    ctx.batch_run()
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ctx.batch_run()
ctx.synchronize()
dev = (time.perf_counter() - t0) / 3
print(f"{name}: device-resident {dev * 1e3:.2f} ms / {nq} queries = {nq / dev:.0f} q/s", flush=True)
qs = capi.Queries.new(ds.query_labels, off, codes)
os.environ["RXH_TIMING"] = "1"
for chunk in chunks:
    for fmt in (None, "1"):
        if fmt:
            os.environ["RXH_FORMAT_THREADS"] = fmt
        else:
            os.environ.pop("RXH_FORMAT_THREADS", None)
        capi.raxtax_counted(ctx, qs, tree, chunk_size=chunk)
        t0 = time.perf_counter()
        for _ in range(3):
            capi.raxtax_counted(ctx, qs, tree, chunk_size=chunk)
        e = (time.perf_counter() - t0) / 3
        print(f"  chunk {chunk} fmt {fmt}: e2e {e * 1e3:.2f} ms = {nq / e:.0f} q/s ({dev / e:.3f} of device)", flush=True)
