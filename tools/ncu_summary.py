"""Summarise an .ncu-rep (raw page + source page) into JSON/text for profiles/.  Usage: ncu_summary.py rep [--src N]"""
import csv, io, json, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_bytes.sum', 'sm__cycles_active.avg', 'smsp__cycles_active.avg',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {h: f"{vals[i]} {units[i]}".strip() for i, h in enumerate(hdr) if h in KEYS or h == 'Kernel Name'}
        res.append(d)
    return res


def src(rep, top):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
    cur, agg = None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif len(r) >= 8 and r[0].isdigit():
            try:
                agg.append((int(r[7]), int(r[4]) if r[4].isdigit() else 0, cur, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    ti, ts = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    lines = [f"{a[0] / ti * 100:5.1f}% inst {a[1] / ts * 100:5.1f}% samples  {a[2]}:{a[3]}  {a[4]}" for a in sorted(agg, key=lambda a: -a[1])[:top]]
    return lines


if __name__ == '__main__':
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--src') + 1]) if '--src' in sys.argv else 0
    print(json.dumps(raw(rep), indent=1))
    if top:
        print("\n".join(src(rep, top)))
