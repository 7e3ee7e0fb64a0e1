#!/bin/bash
# what the driver runs at round end, each step under a hard timeout: the GPU suite, smoke(), the default bench
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --timeout 150 --timeout-method=thread > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log | cut -c1-300
timeout -s KILL 400 python bench.py --gpus 1 --steps 20 --warmup 5 --kernel-table gpurun_out/k_validate.json > gpurun_out/bench_validate.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/bench_validate.log | cut -c1-1500
