# 2-GPU check: NCCL tests, the driver suite, the N=2 bench line (usage: bash tools/gpu_r2b.sh <tag> <ngpus>)
tag=${1:-r2b}; n=${2:-2}
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_driver.py -x -q 2>&1 | tail -25 | tee gpurun_out/${tag}_nccl_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
tail -5 gpurun_out/${tag}_bench_n$n.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_n$n.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["e2e"]["frac_of_device_resident"],3), "ms/step", round(d["ms_per_step"],3))
print(json.dumps(d.get("sharded"), indent=1))
PY
