# c3 device-resident rate with the two-stream sub-batch pipeline on / off (usage: bash tools/gpu_pipe_c3.sh <tag>)
tag=${1:-pipe}
for p in 0 1; do
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --pipeline $p > gpurun_out/${tag}_p$p.json 2> gpurun_out/${tag}_p$p.err; tail -2 gpurun_out/${tag}_p$p.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_p$p.json"))
print("pipeline $p: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "sub_batch", d["config"]["sub_batch"], "ms/step", round(d["ms_per_step"],2), "roofline frac", d["roofline"]["frac"])
PY
done
