# round 2, first check of the pipelined driver + new bench (usage: bash tools/gpu_r2a.sh <tag>)
tag=${1:-r2a}
python -m pytest tests/test_gpu_driver.py -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_driver_tests.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.txt
for w in c2 c3; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -3 gpurun_out/${tag}_bench_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_$w.json"))
print("$w value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["e2e"]["frac_of_device_resident"],3), "abi", round(d["e2e_device_abi"]["value"]), "ms/step", round(d["ms_per_step"],3), d["roofline"]["kernel_ms_per_step"])
PY
done
