#!/usr/bin/env python
"""bench.py -- queries/sec of the raxtax query-classification hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                 our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference algorithm on the host cores (CPU oracle port)

A "step" is one pass of the hot path (k-mers -> hit counts -> probabilities -> lineage results) over one batch of
synthetic queries of the workload.  Workload at every N: BASELINE config 2 (100k COI-like 650 bp references x 10k
queries per GPU; weak scaling: the index is replicated, every rank classifies its own 10k queries, no collective).
`value` = queries of all ranks / max-over-ranks device time with the batch resident in HBM; `e2e` = the same through
rtx_classify_batch with pinned HOST buffers (H2D + kernels + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (box, device-timed) at 1/2/4/8 B200; hit-count HBM GB/s vs peak"
UNIT = "queries/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread (the timed region of the
    default run is ~40 ms, too short for `nvidia-smi -lms`); falls back to an nvidia-smi loop when pynvml is missing."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain list of ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu_index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)), int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                try:
                    self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)), int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))))
                except Exception:
                    break
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = [x[0] for x in self.samples]
            reasons = sorted(k for k, bit in names.items() if any(x[1] & bit for x in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm), "source": "nvml, 2 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 200"}


def make_workload(name, world):
    from raxtax_b200 import synth

    cfg = synth.CONFIGS[name]
    ds = synth.generate(name, n_queries=cfg[1] * world, measure=False)
    return ds, cfg[1]


def cpu_reference_rate(ds, q0, n_sample, threads, full_chunk, skip=False):
    """The reference algorithm (oracle port of raxtax.rs:14-97, --threads 0 chunking of main.rs:119-124) on host cores."""
    from oracle import oracle as orc

    ot = orc.Tree.new(ds.ref_lineages, [ds.ref_seq(i) for i in range(ds.n_refs)])
    off = ds.query_off[q0: q0 + n_sample + 1]
    codes = ds.query_codes
    chunk = n_sample if threads == 1 else full_chunk

    def run():
        o = ot.classify(off - off[0], codes[int(off[0]): int(off[-1])], skip_exact=skip, threads=threads, chunk_size=chunk)
        return o["seconds"]

    return ot, run


def cpu_sample_plan(q_total, threads, override=0):
    """chunk size of the reference for the FULL job (main.rs:119-124) and a bounded sample that keeps every thread busy."""
    chunk = q_total if threads == 1 else max(100, q_total // (threads * 10) + 1)
    n = override or min(q_total, 2 * chunk * threads)
    return n, chunk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub-batch", type=int, default=0, help="RTX_OPT_SUB_BATCH for the timed legs (0 = library default)")
    ap.add_argument("--pipeline", type=int, default=-1, help="RTX_OPT_PIPELINE for the timed legs (-1 = library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    # stdout carries the ONE JSON line and nothing else: whatever libraries print (NCCL writes its version banner to stdout) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        ds, q_per_gpu = make_workload(args.workload, 1)
        n_sample, chunk = cpu_sample_plan(q_per_gpu, cores, args.cpu_sample)
        from raxtax_b200 import synth

        _, run = cpu_reference_rate(ds, 0, n_sample, cores, chunk, skip=args.workload == "c4")
        for _ in range(args.warmup):
            run()
        secs = [run() for _ in range(args.steps)]
        total = float(sum(secs))
        v = n_sample * args.steps / total
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/f64",
                "data": "synthetic", "config": {"workload": f"{args.workload}: {ds.n_refs} refs x {synth.CONFIGS[args.workload][2]} bp, bounded sample of {n_sample} queries per step"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{n_sample} queries of the {args.workload} workload per step, {args.steps} steps, all {cores} host threads"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    from raxtax_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ds, q_per_gpu = make_workload(args.workload, world)
    from raxtax_b200 import synth

    ref_len, kind = synth.CONFIGS[args.workload][2], synth.CONFIGS[args.workload][3]
    kind_name = "16S-like" if kind == "16s" else "COI-like"
    skip = args.workload == "c4"  # BASELINE config 4 is the mislabel mode (--skip-exact-matches)
    t0 = time.time()
    tree = capi.Tree.new(ds.ref_lineages, ds.ref_off, ds.ref_codes)
    t_tree = time.time() - t0
    ctx = capi.Context(local_rank)
    t0 = time.time()
    ctx.upload_tree(tree)
    t_upload = time.time() - t0
    q0 = rank * q_per_gpu
    off = (ds.query_off[q0: q0 + q_per_gpu + 1] - ds.query_off[q0]).astype(np.uint64)
    codes = ds.query_codes[int(ds.query_off[q0]): int(ds.query_off[q0 + q_per_gpu])]
    eo, eids = tree.exact_batch(off, codes)
    # pinned host buffers for the end-to-end leg
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    off_p, codes_p, eo_p, eids_p = pin(off), pin(codes), pin(eo), pin(eids if len(eids) else np.zeros(1, np.uint32))

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    if args.sub_batch:
        ctx.set_option(capi.RTX_OPT_SUB_BATCH, args.sub_batch)
    if args.pipeline >= 0:
        ctx.set_option(capi.RTX_OPT_PIPELINE, args.pipeline)
    # ---- device-resident leg ------------------------------------------------------------------------------
    ctx.batch_upload(off_p, codes_p, eo_p, eids_p, skip_exact=skip)
    for _ in range(args.warmup):
        ctx.batch_run()
    ctx.synchronize()
    ctx.set_option(capi.RTX_OPT_PROFILE, 1)
    ctx.profile_reset()
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.batch_run()
    ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    res = ctx.batch_download()
    prof_main = ctx.profile()
    n_results = len(res.first_ref)
    value = q_per_gpu * world * args.steps / (dev_ms / 1e3)

    # ---- per-kernel leg: the same batch with the sub-batch pipeline off, so that every kernel runs alone on the GPU and its
    # CUDA-event duration is its own (in the pipelined run above hit counting shares the SMs with the other stream's kernels)
    ctx.set_option(capi.RTX_OPT_PIPELINE, 0)
    ctx.set_option(capi.RTX_OPT_SUB_BATCH, 0)
    ctx.batch_upload(off_p, codes_p, eo_p, eids_p, skip_exact=skip)
    for _ in range(2):
        ctx.batch_run()
    ctx.profile_reset()
    ser0, ser1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ser0.record(stream)
    for _ in range(args.steps):
        ctx.batch_run()
    ser1.record(stream)
    ctx.synchronize()
    serial_ms = ser0.elapsed_time(ser1) / args.steps
    ctx.batch_download()
    prof = ctx.profile()
    ctx.set_option(capi.RTX_OPT_PROFILE, 0)
    ctx.set_option(capi.RTX_OPT_PIPELINE, args.pipeline if args.pipeline >= 0 else 0)
    ctx.set_option(capi.RTX_OPT_SUB_BATCH, args.sub_batch)

    # ---- end-to-end leg (host buffers, H2D + kernels + D2H per step) ---------------------------------------------
    # inputs and result arrays page-locked (rtx_host_alloc), reused from step to step as a long-running caller would
    res_buf = ctx.pinned_results(q_per_gpu, max(q_per_gpu * 8 + 64, n_results + n_results // 4 + 64))
    for _ in range(2):
        ctx.classify(off_p, codes_p, eo_p, eids_p, skip_exact=skip, out=res_buf)
    ctx.profile_reset()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = ctx.classify(off_p, codes_p, eo_p, eids_p, skip_exact=skip, out=res_buf)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()  # polled from the start of the device-resident leg to the end of the end-to-end leg (all three timed legs)
    prof_e2e = ctx.profile()
    e2e_value = q_per_gpu * world * args.steps / e2e_s
    assert len(out.first_ref) == n_results

    # ---- roofline of the dominant kernel (hit count) ----------------------------------------------------------
    peak, peak_src = measured_peaks()
    hc = prof["hitcount"]
    hc_ms = hc["total_ms"] / max(hc["launches"], 1)
    bytes_per_launch = prof["bitrow_bytes"] / max(hc["launches"], 1)
    csr_bytes_per_launch = prof["csr_equiv_bytes"] / max(hc["launches"], 1)
    achieved = bytes_per_launch / (hc_ms * 1e-3) / 1e9 if hc_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "hitcount_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch") if str(tj.get("workload", "")).startswith(args.workload + ",") else None  # measured on that workload only
        except Exception:
            traffic = None
    kernel_ms = {k: prof[k]["total_ms"] / args.steps for k in ("kmers", "hitcount", "fixup", "prob", "prefix", "walk")}
    roofline = {"bound": "hbm", "kernel": "hitcount_bitrows_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                "launch_ms": hc_ms, "timed": "kernels serialized (RTX_OPT_PIPELINE=0), CUDA events around each launch on the launching stream",
                "serial_ms_per_step": serial_ms, "csr_equivalent": {"bytes_per_launch": csr_bytes_per_launch,
                                                       "achieved": csr_bytes_per_launch / (hc_ms * 1e-3) / 1e9 if hc_ms > 0 else 0.0,
                                                       "note": "4*hits+2*N per query: what the reference's CSR walk would move (SURVEY 8d primary figure)"},
                "kernel_ms_per_step": kernel_ms}
    if traffic and hc_ms > 0:  # what actually crossed the HBM interface (ncu), next to the algorithmic figure above
        roofline["dram"] = {"bytes_per_launch": traffic, "gbs": traffic / (hc_ms * 1e-3) / 1e9, "frac_of_peak": traffic / (hc_ms * 1e-3) / 1e9 / peak,
                            "note": "L2 blocking keeps the bit matrix on chip: the kernel is bound by the SMs' L1 data pipe (80 % of peak, ncu), not by HBM"}

    launches = sum(prof_main[k]["launches"] for k in capi.KERNEL_NAMES)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bit-planes/u16 counts/f64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {ds.n_refs} {kind_name} refs x {ref_len} bp (6-rank lineages) x {q_per_gpu} queries per GPU, index replicated, queries partitioned"
                                   + (", --skip-exact-matches" if skip else ""),
                       "l2": "no flush needed: bit rows %.0f MB + count vectors %.0f MB per step >> 126 MB L2" % (
                           ctx.index_bytes / 1e6, q_per_gpu * ctx.shard_refs * 2 / 1e6),
                       "sub_batch": ctx.sub_batch, "pipeline": bool(args.pipeline > 0),
                       "tree_build_s": round(t_tree, 2), "index_upload_s": round(t_upload, 2), "results_per_step": n_results},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": prof_e2e["h2d_bytes"] // args.steps,
                    "d2h_bytes_per_step": prof_e2e["d2h_bytes"] // args.steps, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "roofline": roofline}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        n_sample, chunk = cpu_sample_plan(q_per_gpu, cores, args.cpu_sample)
        _, run = cpu_reference_rate(ds, 0, n_sample, cores, chunk, skip=skip)
        run()
        secs = run()
        line["cpu_baseline"] = {"value": n_sample / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {n_sample} queries of the same workload, all {cores} host threads, oracle port of raxtax.rs (Rust reference not buildable here)"}
    if rank == 0:
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
