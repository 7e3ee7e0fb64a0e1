#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "^(PASSED|FAILED|ERROR)|passed|failed|tc vs fp32|tensor-core kernel reported|oracle\(|worst normalised|tc forward vs" gpurun_out/pytest.log | tail -60
timeout 600 python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-700
