#!/bin/bash
# 32 -> 64 fp32 backward on 4-row tiles / two CTAs per SM (yunet_s): parity subset, plugin + trainer tests, yunet_s bench
mkdir -p gpurun_out
PT="python -m pytest -q -rA -p no:cacheprovider --timeout 120 --timeout-method=thread -m gpu"
timeout -s KILL 240 $PT tests/test_gpu_backward_units.py tests/test_gpu_plugins.py tests/test_gpu_trainer.py tests/test_gpu_parity.py -k "yunet_s or plugins or trainer or backward or every_unit" > gpurun_out/pytest_mid.log 2>&1; echo "exit $?" >> gpurun_out/pytest_mid.log
echo "== mid: $(grep -E 'passed|failed|error' gpurun_out/pytest_mid.log | tail -1) $(tail -1 gpurun_out/pytest_mid.log)"
grep -E "^(FAILED|ERROR)|Timeout|^E  " gpurun_out/pytest_mid.log | head -12
timeout -s KILL 200 python bench.py --arch yunet_s --steps 10 --warmup 3 --no-cpu-baseline --no-extra --kernel-table gpurun_out/k_s.json > gpurun_out/bench_s.log 2>&1
python - <<'P'
import json
d = json.loads(open('gpurun_out/bench_s.log').read().strip().splitlines()[-1])
k = json.load(open('gpurun_out/k_s.json'))
print('yunet_s ms/step', round(d['ms_per_step'], 3), 'img/s', round(d['value']), 'roofline', d['roofline'].get('kernel'), round(d['roofline']['frac'], 3))
for r in k[:10]: print('    %-46s %7.3f ms %6.0f GB/s' % (r['kernel'], r['ms'], r['gbs'] or 0))
P
