# round 2 evidence run on one B200 (usage: bash tools/gpu_r2c.sh <tag>): GPU suite, e2e stage clocks, ncu --set full of the hot kernels on c2 and c3
tag=${1:-r2c}
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/${tag}_tests.txt
python tools/e2e_probe.py c2 0 1250 2>&1 | grep -v "fmt 1\|1 format threads" | tail -12 | tee gpurun_out/${tag}_e2e_probe_c2.txt
for w in c2:10000 c3:6000; do
  name=${w%%:*}; nq=${w##*:}
  ncu --set full --clock-control none --import-source on -k regex:"hitcount_group|prob_table|prefix_kernel|lineage_bfs" -s 4 -c 4 -o gpurun_out/${tag}_ncu_${name} -f \
      python tools/ncu_target.py $name $nq gpurun_out/${tag}_ncu_${name}_model.json > gpurun_out/${tag}_ncu_${name}.log 2>&1
  tail -2 gpurun_out/${tag}_ncu_${name}.log
done
ls -la gpurun_out/${tag}_ncu_*.ncu-rep
