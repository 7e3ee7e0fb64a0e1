#!/bin/bash
# round 2, session 2: full GPU suite on the current build, then the bench with the per-kernel table for
# the default build and for the development knobs given as arguments ("TAG:ENV=VAL" ...)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|kernel reported|Error|pytest exit" gpurun_out/pytest.log | tail -12
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --kernel-table gpurun_out/k_$tag.json > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_{tag}.log').read().strip().splitlines()[-1])
    k = json.load(open(f'gpurun_out/k_{tag}.json'))
    print(tag, 'ms/step', round(d['ms_per_step'], 3), 'eager', round(d['cuda_graph']['eager_ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'roofline', d['roofline'].get('kernel'), round(d['roofline']['frac'], 3))
    fam = {}
    for r in k: fam[r['kernel'].split(':')[0]] = fam.get(r['kernel'].split(':')[0], 0) + r['ms']
    print('   ', {a: round(b, 3) for a, b in fam.items()})
    for r in k[:40]: print('    %-46s %7.3f ms %6.0f GB/s' % (r['kernel'], r['ms'], r['gbs'] or 0))
except Exception as e:
    print(tag, 'failed', e); print(open(f'gpurun_out/bench_{tag}.log').read()[-800:])
P
}
run default A=1
for spec in "$@"; do run "${spec%%:*}" "${spec#*:}"; done
