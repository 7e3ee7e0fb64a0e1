#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "tc or golden" > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
grep -E "^(FAILED|ERROR)|passed|failed|tensor-core kernel reported" gpurun_out/pytest_tc.log | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-200
# launch list of two plain steps (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/ncu_launch.log 2>&1
# full capture: tensor-core forward + backward unit kernels of the second step (largest units come first in bwd)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_bwd_tc_kernel -s 9 -c 3 -o gpurun_out/prof_bwd_tc python tools/profile_step.py 2 > gpurun_out/ncu_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:unit_fwd_tc_kernel -s 14 -c 4 -o gpurun_out/prof_fwd_tc python tools/profile_step.py 2 > gpurun_out/ncu_fwd.log 2>&1
ls -la gpurun_out/
