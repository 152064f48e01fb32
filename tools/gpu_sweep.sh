# usage: bash tools/gpu_sweep.sh <tag> <workload> <configs>
tag=$1; wl=$2; cfgs=$3
python tools/sweep_hitcount.py $wl $cfgs 2>&1 | tee gpurun_out/${tag}_sweep.txt
