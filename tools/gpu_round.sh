set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/r01e_bench_n1.json 2> gpurun_out/r01e_bench_n1.err; tail -3 gpurun_out/r01e_bench_n1.err; cat gpurun_out/r01e_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r01e_ncu_b.log 2>&1
tail -2 gpurun_out/r01e_ncu_b.log
