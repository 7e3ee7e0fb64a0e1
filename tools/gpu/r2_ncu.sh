#!/bin/bash
# round-2 ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
P="python tools/profile_step.py 2"
N="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $P > gpurun_out/ncu_launch.log 2>&1
# second step only (the first is warm-up).  Per step: 6 unit_bwd_st, 6 unit_bwd_kernel (3 head-branch
# units, then 16->64, pooled 16->16, 16->16 @160^2), 2 unit_fwd_kernel, 18 unit_fwd_ws, stem fwd + bwd
timeout 900 $N -k regex:unit_bwd_st_kernel -s 6 -c 6 -o gpurun_out/r2_prof_bwd_st $P > gpurun_out/ncu_bwd_st.log 2>&1
timeout 600 $N -k regex:unit_bwd_kernel -s 9 -c 3 -o gpurun_out/r2_prof_bwd_fp32 $P > gpurun_out/ncu_bwd_fp32.log 2>&1
timeout 600 $N -k regex:unit_fwd_kernel -s 2 -c 2 -o gpurun_out/r2_prof_fwd_fp32 $P > gpurun_out/ncu_fwd_fp32.log 2>&1
timeout 600 $N -k regex:unit_fwd_ws_kernel -s 18 -c 3 -o gpurun_out/r2_prof_fwd_ws $P > gpurun_out/ncu_fwd_ws.log 2>&1
timeout 600 $N -k regex:stem_ -s 2 -c 2 -o gpurun_out/r2_prof_stem $P > gpurun_out/ncu_stem.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_*.log | cut -c1-160
