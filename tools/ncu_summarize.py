"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/r1_launches_summary.txt
    python tools/ncu_summarize.py report   gpurun_out/prof_x.ncu-rep profiles/r1_ncu_x.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__warps_active.avg.per_cycle_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']


def short(name):
    name = re.sub(r'void |yunet::|\(anonymous namespace\)::|<unnamed>::|unnamed>::', '', name)
    return re.sub(r'\(.*', '', name).strip()


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        k = short(r[4])
        ns = float(r[14].replace(',', ''))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(dst, 'w') as f:
        f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised:\n'
                f'# compare SHARES, not absolutes); {len(rows)} launches, {total / 1e6:.3f} ms total\n')
        f.write(f'{"kernel":70s} {"launches":>8s} {"total_us":>10s} {"share":>7s}\n')
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k:70s} {n:8d} {ns / 1e3:10.1f} {100 * ns / total:6.2f}%\n')


def report(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full --clock-control none --import-source on; source: {src}\n')
        for r in rows[2:]:
            f.write(f'\n== {short(r[idx["Kernel Name"]])}  grid {r[idx["Grid Size"]]} block {r[idx["Block Size"]]}\n')
            for k in KEYS:
                if k in idx:
                    f.write(f'  {k:72s} {r[idx[k]]} {units[idx[k]]}\n')
            vals = sorted(((float(r[idx[h]].replace(',', '') or 0), h) for h in stalls), reverse=True)
            f.write('  top stall reasons (warps per issue-active cycle): ' + ', '.join(
                f'{h.split("stalled_")[1].split("_per_issue")[0]} {v:.2f}' for v, h in vals[:5]) + '\n')


if __name__ == '__main__':
    {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2], sys.argv[3])
