#!/bin/bash
mkdir -p gpurun_out
cd libfacedetection/train_b200/csrc && touch unit_fwd_ws.cu && make EXTRA=-DYUNET_WS_TIMING > /root/repo/gpurun_out/make_timing.log 2>&1; cd /root/repo
for d in 0 30 6 7; do
echo "== dbg=$d"
YUNET_WS_DBG=$d timeout 300 python tools/ws_timing.py 2>&1 | tail -11
done
cd libfacedetection/train_b200/csrc && touch unit_fwd_ws.cu && make > /root/repo/gpurun_out/make_plain.log 2>&1; cd /root/repo
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:unit_fwd_ws_kernel -s 17 -c 1 -o gpurun_out/prof_fwd_ws python tools/profile_fwd.py 2 > gpurun_out/ncu_fwd_ws.log 2>&1
tail -2 gpurun_out/ncu_fwd_ws.log
