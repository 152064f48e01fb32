"""profiles/hitcount_traffic.json from ncu captures: for every workload W with gpurun_out/<tag>_ncu_full_summary_W.json (tools/ncu_summary.py of
an ncu --set full capture of one pass, tools/gpu_r2u.sh) and gpurun_out/<tag>_ncu_W_model.json (tools/ncu_target.py), the hit-count launch's measured DRAM bytes and pipe figures
next to its byte model.  usage: python tools/traffic_from_ncu.py <tag> [workloads...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
workloads = sys.argv[2:] or ["c2", "c3", "c4"]
path = os.path.join(ROOT, "profiles", "hitcount_traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def val(x):
    return float(x.split()[0].replace(",", ""))


def scaled(x, table):
    v, u = x.split()[0], x.split()[1]
    return float(v.replace(",", "")) * table[u]


for w in workloads:
    summ = os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_full_summary_{w}.json")
    model = os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_{w}_model.json")
    if not (os.path.exists(summ) and os.path.exists(model)):
        print("missing capture for", w)
        continue
    m = json.load(open(model))
    hit = [k for k in json.load(open(summ)) if "hitcount" in k.get("Kernel Name", "")]
    if not hit:
        print("no hit-count launch in", summ)
        continue
    k = hit[0]
    rd, wr = scaled(k["dram__bytes_read.sum"], UNIT), scaled(k["dram__bytes_write.sum"], UNIT)
    out[w] = {
        "kernel": m["kernel"], "kernel_symbol": k["Kernel Name"].replace("void ", "").split("(")[0], "queries_per_launch": m["queries_per_launch"],
        "bitrow_bytes_per_launch": m["bitrow_bytes_per_launch"], "compulsory_bytes_per_launch": m["compulsory_bytes_per_launch"],
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "launch_ms_ncu": scaled(k["gpu__time_duration.sum"], TIME),
        "alu_pipe_pct": val(k["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]),
        "l1_data_pipe_pct": val(k["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
        "lts_pct": val(k["lts__throughput.avg.pct_of_peak_sustained_elapsed"]),
        "l1_hit_pct": val(k["l1tex__t_sector_hit_rate.pct"]), "l2_hit_pct": val(k["lts__t_sector_hit_rate.pct"]),
        "inst_executed": val(k["smsp__inst_executed.sum"]), "registers": k["launch__registers_per_thread"],
        "warps_active_pct": val(k["sm__warps_active.avg.pct_of_peak_sustained_active"]),
        "source": f"ncu --set full --clock-control none of one launch (tools/gpu_r2u.sh, tools/ncu_target.py {w} {int(m['queries_per_launch'] * m['launches'])}): "
                  f"profiles/{tag}_ncu_full_summary_{w}.json",
    }
    print(w, json.dumps(out[w])[:300])
json.dump(out, open(path, "w"), indent=1)
