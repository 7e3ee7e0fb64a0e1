#!/bin/bash
# role-ablation experiment of the ws forward kernel (timings only; outputs are wrong with dbg != 0)
mkdir -p gpurun_out
for d in 6 14 22 30; do
  YUNET_WS_DBG=$d timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --kernel-table gpurun_out/k_$d.json > gpurun_out/bench_dbg.log 2>&1
  python - <<P
import json
rows=json.load(open('gpurun_out/k_$d.json'))
for r in rows:
    if r['kernel'] in ('fwd_ws:backbone.model2.conv1','fwd_ws:backbone.model3.conv1','fwd_ws:neck.lateral_convs.0','fwd_ws:backbone.model3.conv2'):
        print('dbg=$d %-40s %7.3f ms' % (r['kernel'], r['ms']))
P
done
