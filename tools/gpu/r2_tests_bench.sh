#!/bin/bash
# full GPU test suite + a short bench with the per-kernel table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|kernel reported|Error" gpurun_out/pytest.log | tail -20
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-300
python - <<'P'
import json
k=json.load(open('gpurun_out/kernels.json'))
rows=k if isinstance(k,list) else k.get('kernels',k)
tot={}
for r in rows:
    fam=r['kernel'].split(':')[0]
    tot[fam]=tot.get(fam,0)+r['ms']
print({a:round(b,3) for a,b in tot.items()})
for r in rows[:45]:
    print('%-52s %7.3f ms %7.0f GB/s' % (r['kernel'], r['ms'], r['gbs'] or 0))
P
