# multi-GPU evidence of the final kernels (usage: bash tools/gpu_r2v.sh <tag> <ngpus> [c3]): NCCL tests, BASELINE config 5 (8 M references sharded over the ranks),
# optionally the default bench line (c3, strong scaling, sharded check leg)
tag=${1:-r2v}; n=${2:-8}; c3=${3:-}
python -m pytest tests/test_gpu_nccl.py -x -q 2>&1 | grep -v "^$" | tail -4 | tee gpurun_out/${tag}_nccl_tests_n$n.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --gpus $n --steps 3 --warmup 2 > gpurun_out/${tag}_bench_c5_n$n.json 2> gpurun_out/${tag}_bench_c5_n$n.err
tail -3 gpurun_out/${tag}_bench_c5_n$n.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_c5_n$n.json"))
s=d.get("sharded") or {}
print("c5 N=$n value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), d["config"])
print(json.dumps({k:s.get(k) for k in ("queries","references","references_per_shard","sub_batch","index_upload_s","result_lines","phase_ms_per_step_rank0","collective_bytes_per_step_rank0")}))
PY
if [ -n "$c3" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c3_n$n.json 2> gpurun_out/${tag}_bench_c3_n$n.err
tail -3 gpurun_out/${tag}_bench_c3_n$n.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_c3_n$n.json"))
s=d.get("sharded") or {}
print("c3 N=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["e2e"]["frac_of_device_resident"],3), "ms/step", round(d["ms_per_step"],2))
print("sharded leg:", round(s.get("value",0)), "q/s", s.get("n_ranks"), "ranks", s.get("queries"), "queries differing from unsharded:", s.get("queries_differing_from_unsharded"), s.get("phase_ms_per_step_rank0"))
PY
fi
