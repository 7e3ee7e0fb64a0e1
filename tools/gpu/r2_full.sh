#!/bin/bash
# what the driver runs at round end: the gpu-marked tests, then the default bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|kernel reported|Error" gpurun_out/pytest.log | tail -20
SECONDS=0; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $? after ${SECONDS}s"; tail -3 gpurun_out/bench_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','gpu_launches','cuda_graph','e2e','e2e_plugin','cpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:300])
print('roofline', json.dumps({k:v for k,v in d['roofline'].items() if k not in ('forward_total','step_total')}))
print('forward_total', d['roofline']['forward_total']); print('step_total', d['roofline']['step_total'])
for k,v in (d.get('extra_configs') or {}).items():
    print(k, json.dumps(v)[:900])
P
