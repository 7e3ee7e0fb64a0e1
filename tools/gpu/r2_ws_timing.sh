#!/bin/bash
mkdir -p gpurun_out
cd libfacedetection/train_b200/csrc && touch unit_fwd_ws.cu && make EXTRA=-DYUNET_WS_TIMING > /root/repo/gpurun_out/make_timing.log 2>&1; cd /root/repo
timeout 300 python tools/ws_timing.py > gpurun_out/ws_timing.log 2>&1
tail -20 gpurun_out/ws_timing.log
