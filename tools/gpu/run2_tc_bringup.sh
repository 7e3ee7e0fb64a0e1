#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/tc_gemm_test > gpurun_out/tc_gemm_test.log 2>&1; echo "tc exit $?" >> gpurun_out/tc_gemm_test.log
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
cat gpurun_out/tc_gemm_test.log; grep -E "^(PASSED|FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -30; tail -2 gpurun_out/bench.log | cut -c1-600
