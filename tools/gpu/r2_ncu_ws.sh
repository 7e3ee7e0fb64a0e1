#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
# 17 ws launches per forward (yunet_n); index 17 = forward #2's backbone.model2.conv1 (80x80 64->64)
timeout 600 $N -k regex:unit_fwd_ws_kernel -s 17 -c 1 -o gpurun_out/prof_fwd_ws python tools/profile_fwd.py 2 > gpurun_out/ncu_fwd_ws.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3; tail -3 gpurun_out/ncu_fwd_ws.log
