#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider -k "backward" -x -rA > gpurun_out/pytest_st.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_st.log
grep -E "^(PASSED|FAILED|ERROR)|passed|failed|kernel reported|Error|assert|worst normalised" gpurun_out/pytest_st.log | tail -30
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-200
python - <<'P'
import json
k=json.load(open('gpurun_out/kernels.json'))
tot={}
for r in k:
    fam=r['kernel'].split(':')[0]
    tot[fam]=tot.get(fam,0)+r['ms']
print({a:round(b,3) for a,b in tot.items()})
for r in k:
    if r['kernel'].startswith('bwd'):
        print('%-52s %7.3f ms %7.0f GB/s' % (r['kernel'], r['ms'], r['gbs'] or 0))
P
