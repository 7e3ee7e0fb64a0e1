#!/bin/bash
# round-2 ncu evidence for profiles/ (one GPU; numbers printed under ncu are never bench values).
# The reports stay on the box (gpurun_out/ is capped at 64 MiB): only the text summaries come back.
mkdir -p gpurun_out /tmp/rep
P="python tools/profile_step.py 2"
N="ncu --set full --clock-control none --import-source on"
if [ "$1" = "launches" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $P > gpurun_out/ncu_launch.log 2>&1
fi
# second step only (the first is warm-up).  Per step: 6 unit_bwd_st, 6 unit_bwd_kernel (3 head-branch
# units, then 16->64, pooled 16->16, 16->16 @160^2), 2 unit_fwd_kernel, 18 unit_fwd_ws, stem fwd + bwd
timeout 900 $N -k regex:unit_bwd_st_kernel -s 6 -c 6 -o /tmp/rep/bwd_st $P > gpurun_out/ncu_bwd_st.log 2>&1
timeout 600 $N -k regex:unit_bwd_kernel -s 9 -c 3 -o /tmp/rep/bwd_fp32 $P > gpurun_out/ncu_bwd_fp32.log 2>&1
timeout 600 $N -k regex:unit_fwd_kernel -s 2 -c 2 -o /tmp/rep/fwd_fp32 $P > gpurun_out/ncu_fwd_fp32.log 2>&1
timeout 600 $N -k regex:unit_fwd_ws_kernel -s 18 -c 3 -o /tmp/rep/fwd_ws $P > gpurun_out/ncu_fwd_ws.log 2>&1
timeout 600 $N -k regex:stem_ -s 2 -c 2 -o /tmp/rep/stem $P > gpurun_out/ncu_stem.log 2>&1
for r in bwd_st bwd_fp32 fwd_fp32 fwd_ws stem; do
  python tools/ncu_summarize.py report /tmp/rep/$r.ncu-rep gpurun_out/r2_ncu_$r.txt
done
python tools/ncu_hotlines.py /tmp/rep/bwd_st.ncu-rep 40 4 > gpurun_out/r2_hot_bwd_st_4.txt 2>&1
python tools/ncu_hotlines.py /tmp/rep/bwd_st.ncu-rep 40 5 > gpurun_out/r2_hot_bwd_st_5.txt 2>&1
python tools/ncu_hotlines.py /tmp/rep/bwd_fp32.ncu-rep 40 2 > gpurun_out/r2_hot_bwd_fp32.txt 2>&1
python tools/ncu_hotlines.py /tmp/rep/fwd_fp32.ncu-rep 30 0 > gpurun_out/r2_hot_fwd_fp32.txt 2>&1
python tools/ncu_hotlines.py /tmp/rep/stem.ncu-rep 30 0 > gpurun_out/r2_hot_stem0.txt 2>&1
python tools/ncu_hotlines.py /tmp/rep/stem.ncu-rep 30 1 > gpurun_out/r2_hot_stem1.txt 2>&1
python tools/ncu_hotlines.py /tmp/rep/fwd_ws.ncu-rep 30 1 > gpurun_out/r2_hot_fwd_ws.txt 2>&1
ls -la /tmp/rep gpurun_out | tail -30
