#!/bin/bash
# first GPU run: smoke, sanitizer on the tiny smoke, full gpu tests, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitizer.log
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/smoke.log; tail -3 gpurun_out/sanitizer.log; tail -15 gpurun_out/pytest.log; tail -3 gpurun_out/bench.log
