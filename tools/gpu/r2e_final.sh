#!/bin/bash
# plugin-surface tests of the current build, ncu evidence, then the default bench (as the driver runs it, minus the CPU leg)
mkdir -p gpurun_out
PT="python -m pytest -q -rA -p no:cacheprovider --timeout 120 --timeout-method=thread -m gpu"
timeout -s KILL 200 $PT tests/test_gpu_plugins.py > gpurun_out/pytest_plugins.log 2>&1; echo "exit $?" >> gpurun_out/pytest_plugins.log
echo "== plugins: $(grep -E 'passed|failed|error' gpurun_out/pytest_plugins.log | tail -1) $(tail -1 gpurun_out/pytest_plugins.log)"
grep -E "^(FAILED|ERROR)|Timeout|^E  " gpurun_out/pytest_plugins.log | head -12
bash tools/gpu/r2d_ncu.sh "$1"
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-table gpurun_out/k_final.json > gpurun_out/bench_final.log 2>&1
tail -1 gpurun_out/bench_final.log | cut -c1-6000
