# ncu --set full of one launch of each hot kernel in a warm bench step (usage: bash tools/gpu_ncu_full.sh <tag> [kernel-regex])
tag=${1:-ncu}
pat=${2:-"hitcount_bitrows|prob_table|prefix_kernel|lineage_walk"}
n=${3:-4}
skip=${4:-8}
ncu --set full --clock-control none --import-source on -k regex:"$pat" -s $skip -c $n -o gpurun_out/${tag} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log
ls -la gpurun_out/${tag}.ncu-rep
