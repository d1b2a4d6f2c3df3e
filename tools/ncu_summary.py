"""Condenses an .ncu-rep (ncu --set full) into the handful of per-kernel numbers the design decisions rest on."""
import csv
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct']


def summarize(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        e = {'kernel': r[idx['Kernel Name']].split('(')[0], 'grid': r[idx['Grid Size']], 'block': r[idx['Block Size']]}
        for w in WANT:
            if w in idx:
                e[w] = f"{r[idx[w]]} {units[idx[w]]}".strip()
        st = {h: float(r[i]) for h, i in idx.items()
              if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]}
        e['top_stalls_per_issue'] = {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(v, 2)
                                     for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6]}
        res.append(e)
    return res


if __name__ == '__main__':
    print(json.dumps(summarize(sys.argv[1]), indent=1))
