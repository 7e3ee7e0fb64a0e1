#!/bin/bash
# round 2, session 2: ncu evidence of the final build for profiles/ (one GPU; numbers printed under ncu
# are never bench values).  Every step under its own hard timeout.
mkdir -p gpurun_out /tmp/rep
P="python tools/profile_step.py 2"
N="ncu --set full --clock-control none --import-source on"
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $P > gpurun_out/ncu_launch.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/r2_launches.csv gpurun_out/r2_launches_summary.txt; head -12 gpurun_out/r2_launches_summary.txt
# second step only.  Per step: 6 unit_bwd_kernel (3 head-branch units, then 16->64, pooled 16->16, 16->16 @160^2),
# 6 unit_bwd_st (the last two are the 80x80 64->64 units)
timeout -s KILL 240 $N -k regex:unit_bwd_kernel -s 9 -c 3 -o /tmp/rep/bwd_fp32 $P > gpurun_out/ncu_bwd_fp32.log 2>&1
python tools/ncu_summarize.py report /tmp/rep/bwd_fp32.ncu-rep gpurun_out/r2_ncu_bwd_fp32.txt
python tools/ncu_hotlines.py /tmp/rep/bwd_fp32.ncu-rep 40 2 > gpurun_out/r2_hotlines_bwd_fp32.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes|l1tex__throughput|local_op|registers_per_thread|top stall" gpurun_out/r2_ncu_bwd_fp32.txt
if [ "$1" = "more" ]; then
  timeout -s KILL 240 $N -k regex:unit_bwd_st -s 10 -c 2 -o /tmp/rep/bwd_st $P > gpurun_out/ncu_bwd_st.log 2>&1
  python tools/ncu_summarize.py report /tmp/rep/bwd_st.ncu-rep gpurun_out/r2_ncu_bwd_st.txt
  grep -E "^==|gpu__time_duration|dram__bytes|registers_per_thread|top stall" gpurun_out/r2_ncu_bwd_st.txt
fi
