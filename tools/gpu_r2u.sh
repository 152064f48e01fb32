# final evidence of round 2 on one B200 (usage: bash tools/gpu_r2u.sh <tag> [notests]): GPU suite, smoke, reference arm, c3 (default) / c2 / c4 bench
# lines, ncu launch list of the default bench command, ncu --set full of the hot kernels on c2 / c3 / c4 (summarised on the box: the reports
# themselves are larger than what gpurun copies back; c3's is kept)
tag=${1:-r2u}
if [ "$2" != "notests" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -6 | tee gpurun_out/${tag}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
fi
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
for w in c2:10000 c3:6000 c4:6000; do
  name=${w%%:*}; nq=${w##*:}
  ncu --set full --clock-control none --import-source on -k regex:"hitcount_group|prob_table|prefix_kernel|lineage_bfs" -s 4 -c 4 -o /tmp/${tag}_ncu_${name} -f \
      python tools/ncu_target.py $name $nq gpurun_out/${tag}_ncu_${name}_model.json > gpurun_out/${tag}_ncu_${name}.log 2>&1
  tail -1 gpurun_out/${tag}_ncu_${name}.log | cut -c1-200
  python tools/ncu_summary.py /tmp/${tag}_ncu_${name}.ncu-rep > gpurun_out/${tag}_ncu_full_summary_${name}.json
  for k in hitcount_group prob_table prefix_kernel lineage_bfs; do
    echo "== $k ($name): top source lines by stall samples" >> gpurun_out/${tag}_ncu_source_hotlines_${name}.txt
    python tools/ncu_lines.py /tmp/${tag}_ncu_${name}.ncu-rep $k 12 >> gpurun_out/${tag}_ncu_source_hotlines_${name}.txt 2>&1
  done
done
cp /tmp/${tag}_ncu_c3.ncu-rep gpurun_out/
du -sh gpurun_out
