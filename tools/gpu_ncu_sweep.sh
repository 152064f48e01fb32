# ncu --set full of one hit-count launch per sweep configuration (usage: bash tools/gpu_ncu_sweep.sh <tag> <workload> <cfg> [<cfg> ...])
tag=$1; wl=$2; shift 2
for cfg in "$@"; do
  ncu --set full --clock-control none -k regex:"hitcount_group" -s 2 -c 1 -o gpurun_out/${tag}_${cfg} -f python tools/sweep_hitcount.py $wl $cfg > gpurun_out/${tag}_${cfg}.log 2>&1
  tail -1 gpurun_out/${tag}_${cfg}.log
done
