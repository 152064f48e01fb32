"""Wall time of the `raxtax` binary on a synthetic workload (FASTA written to a temp dir first).
usage: cli_wall.py [workload] [extra CLI flags ...]"""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raxtax_b200 import _build

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
extra = sys.argv[2:]
ds = bench.load_workload(name, bench.workload_queries(name, 1)[0])
only_first = os.environ.get("CLI_WALL_ONLY_SKIP_DB") == "1"  # large sets: skip the two runs that write / read the multi-GB .bin
with tempfile.TemporaryDirectory() as d:
    refs, qs = os.path.join(d, "refs.fasta"), os.path.join(d, "queries.fasta")
    open(refs, "w").write(ds.ref_fasta())
    open(qs, "w").write(ds.query_fasta())
    rows = []
    for label, flags in (("fasta database, --skip-db", ["--skip-db"]), ("fasta database, writes the .bin", []), ("the .bin as database", None))[: 1 if only_first else 3]:
        prefix = os.path.join(d, "out_" + str(len(rows)))
        db = refs if flags is not None else os.path.join(d, "out_1", "refs.bin")
        cmd = [_build.CLI_BIN, "-d", db, "-i", qs, "-o", prefix, "--tsv"] + (flags or []) + extra
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RXH_TIMING="1"))
        wall = time.time() - t0
        log = open(os.path.join(prefix, "raxtax.log")).read() if os.path.exists(os.path.join(prefix, "raxtax.log")) else ""
        el = [l for l in log.splitlines() if "Elapsed" in l]
        n_lines = sum(1 for _ in open(os.path.join(prefix, "raxtax.out"))) if r.returncode == 0 else 0
        rows.append(dict(run=label, rc=r.returncode, wall_s=round(wall, 2), result_lines=n_lines, log=el, stderr_tail=r.stderr.strip().splitlines()[-6:]))
    print(json.dumps(dict(workload=name, n_refs=ds.n_refs, n_queries=ds.n_queries, ref_fasta_mb=round(os.path.getsize(refs) / 1e6, 1), runs=rows), indent=1))
