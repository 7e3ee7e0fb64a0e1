#!/bin/bash
# ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
P="python tools/profile_step.py 2"
N="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launch.log 2>&1
# second step only (first step = warm-up): the two 80x80 64->64 units are launches 13,14 of the
# 14 unit_bwd_tc launches of a step and launches 1,2 of the 17 unit_fwd_tc launches
timeout 600 $N -k regex:unit_bwd_tc_kernel -s 26 -c 2 -o gpurun_out/prof_bwd_tc $P > gpurun_out/ncu_bwd_tc.log 2>&1
timeout 600 $N -k regex:unit_fwd_tc_kernel -s 17 -c 2 -o gpurun_out/prof_fwd_tc $P > gpurun_out/ncu_fwd_tc.log 2>&1
# fp32 units of the 160x160 / 80x80 layers: backward launches 4..6 of 6, forward launches 1..3 of 3
timeout 600 $N -k regex:unit_bwd_kernel -s 9 -c 3 -o gpurun_out/prof_bwd_fp32 $P > gpurun_out/ncu_bwd_fp32.log 2>&1
timeout 600 $N -k regex:unit_fwd_kernel -s 3 -c 3 -o gpurun_out/prof_fwd_fp32 $P > gpurun_out/ncu_fwd_fp32.log 2>&1
timeout 600 $N -k regex:stem_ -s 2 -c 2 -o gpurun_out/prof_stem $P > gpurun_out/ncu_stem.log 2>&1
timeout 600 $N -k regex:simota_assign -s 1 -c 1 -o gpurun_out/prof_simota $P > gpurun_out/ncu_simota.log 2>&1
ls -la gpurun_out/*.ncu-rep
