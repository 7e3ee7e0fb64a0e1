#!/bin/bash
mkdir -p gpurun_out
cd libfacedetection/train_b200/csrc && touch unit_bwd_st.cu && make EXTRA=-DYUNET_PHASE_TIMING > /root/repo/gpurun_out/make_timing.log 2>&1; cd /root/repo
timeout 300 python tools/phase_timing.py > gpurun_out/phase_timing.log 2>&1
tail -20 gpurun_out/phase_timing.log
