# final check of a round: full GPU suite, default bench line, config 4 bench line
tag=${1:-r02g}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_tests.txt
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -2 gpurun_out/${tag}_bench_n1.err; cut -c1-400 gpurun_out/${tag}_bench_n1.json
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; tail -2 gpurun_out/${tag}_bench_c4.err; cut -c1-300 gpurun_out/${tag}_bench_c4.json
