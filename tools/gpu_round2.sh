# 2-GPU evidence (usage: bash tools/gpu_round2.sh <tag>): weak-scaling bench line at N = 2, reference-sharded path over NCCL (check + timing),
# smoke(), and the wall time of the raxtax binary on C2
tag=${1:-r01q}
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; tail -2 gpurun_out/${tag}_bench_n2.err; cat gpurun_out/${tag}_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/shard_nccl_check.py c2 2000 time > gpurun_out/${tag}_shard_nccl_n2.log 2>&1; grep -v "^\*\|OMP_NUM" gpurun_out/${tag}_shard_nccl_n2.log | tail -5
python tools/cli_wall.py c2 > gpurun_out/${tag}_cli_wall_c2.json 2> gpurun_out/${tag}_cli_wall.err; cat gpurun_out/${tag}_cli_wall_c2.json | head -60
python tools/cli_wall.py c2 --gpus 2 > gpurun_out/${tag}_cli_wall_c2_g2.json 2>> gpurun_out/${tag}_cli_wall.err; grep wall_s gpurun_out/${tag}_cli_wall_c2_g2.json
