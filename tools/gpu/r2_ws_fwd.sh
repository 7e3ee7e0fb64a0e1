#!/bin/bash
# round 2: warp-specialised forward kernel bring-up: parity vs the fp32 path, then a short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "forward_equals" -x > gpurun_out/pytest_ws.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ws.log
grep -E "^(PASSED|FAILED|ERROR)|passed|failed|tc vs fp32|kernel reported|Error|error" gpurun_out/pytest_ws.log | tail -40
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernels_ws.json > gpurun_out/bench_ws.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_ws.log
tail -2 gpurun_out/bench_ws.log | cut -c1-400
python - <<'P'
import json
try:
    k=json.load(open('gpurun_out/kernels_ws.json'))
    rows=k if isinstance(k,list) else k.get('kernels',k)
    for r in rows[:60]:
        print(r)
except Exception as e:
    print('no table',e)
P
