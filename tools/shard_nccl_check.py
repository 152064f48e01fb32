"""Multi-GPU check of the reference-sharded path over NCCL (run under torchrun, one rank per GPU):
every rank holds one shard, histograms are all-reduced, straddler records all-gathered; rank 0 merges the lines and
compares them with an unsharded run on its own GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from raxtax_b200 import capi, dist as rdist, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "small"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ds = synth.generate(name, n_queries=nq, measure=False)
tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
cuts = rdist.shard_cuts(tree.num_tips, world)
ctx = capi.Context(local)
ctx.upload_tree_sharded(tree, world, rank, cuts)
eo, eids = tree.exact_batch(ds.query_off, ds.query_codes)
group = rdist.TorchShardGroup(ctx)
for skip in (False, True):
    mine = rdist.classify_sharded_rank(ctx, group, ds.query_off, ds.query_codes, eo, eids, skip_exact=skip)
    outs = rdist.gather_outputs(mine)
    if rank == 0:
        merged = rdist.merge_shard_results(outs, eo, eids, tree.index_arrays()["ref_levels"], skip_exact=skip)
        full = capi.Context(local)
        full.upload_tree(tree)
        ref = full.classify(ds.query_off, ds.query_codes, eo, eids, skip_exact=skip)
        full.close()
        same = 0
        for q in range(nq):
            a, b = ref.for_query(q), merged.for_query(q)
            same += len(a) == len(b) and all(x[0] == y[0] and np.array_equal(x[1], y[1]) and abs(x[2] - y[2]) < 1e-9 and abs(x[3] - y[3]) < 1e-9
                                             for x, y in zip(a, b))
        print(f"sharded over {world} GPUs (NCCL), skip={skip}: {same}/{nq} queries identical to the unsharded run", flush=True)
        assert same >= nq - max(1, nq // 50)
# ---- timing of the sharded data path (batch resident; phases + the two collectives; the per-query merge on rank 0 is host work) ----
if len(sys.argv) > 3 and sys.argv[3] == "time":
    import json
    import time

    ctx.batch_upload(ds.query_off, ds.query_codes, eo, eids)

    def one_pass():
        ctx.shard_phase(1)
        group.allreduce_hist()
        ctx.shard_phase(2)
        group.allgather_records()
        ctx.shard_phase(3)
        ctx.synchronize()

    for _ in range(2):
        one_pass()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 5
    for _ in range(steps):
        one_pass()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(dt.item()) * 1e3 / steps
        print(json.dumps({"mode": "reference-sharded", "workload": name, "n_gpus": world, "refs_per_shard": int(cuts[1] - cuts[0]), "queries": nq,
                          "ms_per_batch": round(ms, 3), "queries_per_s": round(nq / ms * 1e3), "hist_allreduce_bytes": int(ctx.shard_hist_buffer()[1]) * 4,
                          "timed": "phase1 + all_reduce + phase2 + all_gather + phase3, max over ranks, batch resident"}), flush=True)
dist.barrier()
dist.destroy_process_group()
