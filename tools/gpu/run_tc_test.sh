#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/tc_gemm_test > gpurun_out/tc_gemm_test.log 2>&1; echo "tc exit $?" >> gpurun_out/tc_gemm_test.log
cat gpurun_out/tc_gemm_test.log
timeout 300 python bench.py --arch yunet_s --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.log 2>&1; tail -1 gpurun_out/bench_s.log | cut -c1-400
timeout 300 python bench.py --workload infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.log 2>&1; tail -1 gpurun_out/bench_infer.log | cut -c1-400
