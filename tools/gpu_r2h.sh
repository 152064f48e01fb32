# sparse frontier expansion in the level-synchronous walk: GPU suite, per-kernel times on c3 (20 k queries) and c2, sub-batch pipeline on / off
tag=${1:-r2h}
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/${tag}_tests.txt
python tools/sweep_hitcount.py c2 P0,P1 2>&1 | tail -2 | tee gpurun_out/${tag}_sweep_c2.txt
python tools/sweep_hitcount.py c3 P0,P1 20000 2>&1 | tail -2 | tee gpurun_out/${tag}_sweep_c3.txt
