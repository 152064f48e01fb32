"""Generates the golden raxtax.out texts for tests/golden/diptera_sample.fasta (used as its own database, default and
--skip-exact-matches) with the CPU oracle.  The oracle is pinned on the reference's own KATs (tests/test_oracle_kats.py);
the Rust binary itself cannot be built in this image, so these files pin the ORACLE's end-to-end output, and the CUDA
path is compared against them on the GPU box."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

if __name__ == "__main__":
    text = open(os.path.join(HERE, "diptera_sample.fasta")).read()
    tree = orc.Tree.from_fasta(text)
    labels, off, codes = orc.parse_queries(text)
    for skip, name in ((False, "default"), (True, "skip")):
        out = tree.classify(off, codes, skip_exact=skip, threads=4, chunk_size=16)
        txt = orc.format_results(tree, out["results"], labels) + "\n"
        open(os.path.join(HERE, f"diptera_sample.{name}.out"), "w").write(txt)
        print(name, len(txt.splitlines()), "lines")
