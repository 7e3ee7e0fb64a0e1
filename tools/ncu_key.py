"""Key metrics of one .ncu-rep (first profiled launch): python tools/ncu_key.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct', 'sm__warps_active.avg.pct', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum ',
        'launch__registers_per_thread ', 'launch__block_size', 'launch__grid_size', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_fma.avg.pct',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg', 'sm__cycles_elapsed.max ', 'derived__memory_l1_wavefronts_shared_excessive', 'smsp__inst_executed_op_local',
        'local_load', 'local_store', 'lsu_mem_local']
for h, u, v in zip(hdr, units, vals):
    hh = h + ' '
    if any(w in hh for w in want) and 'Triage' not in h:
        try:
            if 'stalled' in h and float(v) < 0.3:
                continue
        except ValueError:
            pass
        print(f'{h:88s} {u:16s} {v}')
