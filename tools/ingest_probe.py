"""Host-side ingest times on the box's cores: FASTA (plain / gz) -> Tree through rxh_tree_from_file, with the stage clocks of
RXH_TIMING.  usage: python tools/ingest_probe.py c3 [copies]   (copies > 1 concatenates relabelled copies of the set: 8 = 8 M references)"""
import gzip
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raxtax_b200 import capi

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 1
q_total, _, _ = bench.workload_queries(name, 1)
ds = bench.load_workload(name, q_total)
t0 = time.time()
txt = ds.ref_fasta()
path = f"/tmp/ingest_{name}_{copies}.fasta"
with open(path, "w") as f:
    for c in range(copies):
        f.write(txt if c == 0 else txt.replace("tax=p:P", f"tax=p:Z{c}P"))
print(f"wrote {path}: {os.path.getsize(path) / 1e6:.0f} MB in {time.time() - t0:.1f} s, {ds.n_refs * copies} references", flush=True)
del txt
os.environ["RXH_TIMING"] = "1"
for threads in ("1", None):
    if threads:
        os.environ["RXH_THREADS"] = threads
    else:
        os.environ.pop("RXH_THREADS", None)
    t0 = time.time()
    tree, was = capi.Tree.from_file(path)
    print(f"RXH_THREADS={threads or 'default'} ({os.cpu_count()} cores): rxh_tree_from_file {time.time() - t0:.2f} s, {tree.num_tips} references", flush=True)
    del tree
if copies == 1:
    subprocess.run(["gzip", "-1", "-k", "-f", path], check=True)
    t0 = time.time()
    tree, was = capi.Tree.from_file(path + ".gz")
    print(f"gz ({os.path.getsize(path + '.gz') / 1e6:.0f} MB): rxh_tree_from_file {time.time() - t0:.2f} s", flush=True)
