# gpu tests + ncu --set full of the four hot kernels in a warm bench step (usage: bash tools/gpu_ncu_round.sh <tag>)
tag=${1:-r01k}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.txt
bash tools/gpu_ncu_full.sh ${tag}_full "hitcount_group|prob_table|prefix_kernel|lineage_bfs" 4 4
