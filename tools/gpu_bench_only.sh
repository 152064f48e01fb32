# bench line only (usage: bash tools/gpu_bench_only.sh <tag> [extra bench args])
tag=${1:-quick}; shift
python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), d["roofline"]["kernel_ms_per_step"])
PY
