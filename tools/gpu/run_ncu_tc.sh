#!/bin/bash
mkdir -p gpurun_out
P="python tools/profile_step.py 2"
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:unit_bwd_tc_kernel -s 26 -c 1 -o gpurun_out/prof_bwd_tc $P > gpurun_out/ncu_bwd_tc.log 2>&1
timeout 600 $N -k regex:unit_fwd_tc_kernel -s 17 -c 1 -o gpurun_out/prof_fwd_tc $P > gpurun_out/ncu_fwd_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
