# round-2 evidence on one B200 (usage: bash tools/gpu_r2_final.sh <tag>): GPU suite, smoke, default bench line (c3) with CPU baseline, reference arm,
# c2 / c4 bench lines, ncu launch list of the default bench command
tag=${1:-r2f}
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/${tag}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-260 gpurun_out/${tag}_bench_ref.json
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; tail -3 gpurun_out/${tag}_bench_c3.err
for w in c2 c4; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -3 gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json
for w in ("c3","c2","c4"):
    d=json.load(open("gpurun_out/${tag}_bench_%s.json" % w))
    r=d["roofline"]
    print(w, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["e2e"]["frac_of_device_resident"],3), "abi", round(d["e2e_device_abi"]["value"]),
          "ms/step", round(d["ms_per_step"],2), "roofline", round(r["achieved"]), "/", round(r["peak"]), "=", round(r["frac"],3), "hbm", round(r["hbm"]["frac"],3),
          {k: round(v,2) for k,v in r["kernel_ms_per_step"].items()}, d.get("cpu_baseline",{}).get("value"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_b.log 2>&1
tail -2 gpurun_out/${tag}_ncu_b.log | cut -c1-200
