# full round evidence (usage: bash tools/gpu_round.sh <tag>): gpu tests, bench line (+cpu baseline), reference arm, launch list,
# ncu --set full of the hot kernels
tag=${1:-r02j}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_tests.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -3 gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_b.log 2>&1
tail -2 gpurun_out/${tag}_ncu_b.log
bash tools/gpu_ncu_full.sh ${tag}_full "hitcount_group|prob_table|prefix_kernel|lineage_bfs" 4 4
