#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 40" "32 40" "38 40" "8 40" "40 40"; do
  set -- $cfg
  YUNET_WS_DBG=$1 YUNET_WS_SLEEP=$2 timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --kernel-table gpurun_out/k.json > gpurun_out/bench_dbg.log 2>&1
  python - <<P
import json
rows=json.load(open('gpurun_out/k.json'))
for r in rows:
    if r['kernel'] in ('fwd_ws:backbone.model2.conv1','fwd_ws:backbone.model3.conv2'):
        print('dbg=$1 %-40s %7.3f ms' % (r['kernel'], r['ms']))
P
done
