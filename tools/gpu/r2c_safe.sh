#!/bin/bash
# round 2, session 2: hang-proof run -- every pytest file and every bench under its own hard timeout
# (a dead-locked kernel must cost seconds, not the budget), default build first, knobs afterwards
mkdir -p gpurun_out
PT="python -m pytest -q -rA -p no:cacheprovider --timeout 120 --timeout-method=thread -m gpu"
for f in backward_units tc parity plugins trainer pipeline; do
  timeout -s KILL 330 $PT tests/test_gpu_$f.py > gpurun_out/pytest_$f.log 2>&1; echo "exit $?" >> gpurun_out/pytest_$f.log
  echo "== $f: $(grep -E 'passed|failed|error' gpurun_out/pytest_$f.log | tail -1) $(tail -1 gpurun_out/pytest_$f.log)"
  grep -E "^(FAILED|ERROR)|Timeout|kernel reported" gpurun_out/pytest_$f.log | head -8
done
run() { tag=$1; extra=$2; shift 2; env "$@" timeout -s KILL 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $extra --kernel-table gpurun_out/k_$tag.json > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_{tag}.log').read().strip().splitlines()[-1])
    k = json.load(open(f'gpurun_out/k_{tag}.json'))
    print(tag, 'ms/step', round(d['ms_per_step'], 3), 'eager', round(d['cuda_graph']['eager_ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'roofline', d['roofline'].get('kernel'), round(d['roofline']['frac'], 3))
    if d.get('e2e_plugin'): print('    e2e_plugin', json.dumps(d['e2e_plugin'])[:700])
    if d.get('extra_configs'): print('    extra', json.dumps(d['extra_configs'])[:900])
    fam = {}
    for r in k: fam[r['kernel'].split(':')[0]] = fam.get(r['kernel'].split(':')[0], 0) + r['ms']
    print('   ', {a: round(b, 3) for a, b in fam.items()})
    for r in k[:32]: print('    %-46s %7.3f ms %6.0f GB/s' % (r['kernel'], r['ms'], r['gbs'] or 0))
except Exception as e:
    print(tag, 'failed', e); print(open(f'gpurun_out/bench_{tag}.log').read()[-800:])
P
}
run default "" A=1
run quad --no-extra YUNET_BWD_QUAD=1
# experimental strip backward (g pass on dedicated warps): bounded tests, then a bench
YUNET_ST_GW=1 timeout -s KILL 200 $PT tests/test_gpu_backward_units.py tests/test_gpu_tc.py "tests/test_gpu_parity.py::test_train_step_matches_reference_golden" "tests/test_gpu_parity.py::test_loss_and_grads_match_oracle" tests/test_gpu_parity.py::test_full_size_properties_bs256 > gpurun_out/pytest_gw.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gw.log
echo "== gw: $(grep -E 'passed|failed|error' gpurun_out/pytest_gw.log | tail -1) $(tail -1 gpurun_out/pytest_gw.log)"
grep -E "^(FAILED|ERROR)|Timeout|kernel reported" gpurun_out/pytest_gw.log | head -8
run gw --no-extra YUNET_ST_GW=1
