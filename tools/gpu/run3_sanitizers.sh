#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward_units.py -q -rA -p no:cacheprovider -k yunet_n > gpurun_out/bwd_units.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_backward_units.py -q -p no:cacheprovider -k yunet_n > gpurun_out/racecheck.log 2>&1
timeout 900 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_backward_units.py -q -p no:cacheprovider -k yunet_n > gpurun_out/initcheck.log 2>&1
grep -B2 -A60 "first bad tensor" gpurun_out/bwd_units.log | head -90
grep -E "Race|hazard|ERROR SUMMARY|Uninit" gpurun_out/racecheck.log | sort | uniq -c | head -20
grep -E "Uninitialized|ERROR SUMMARY" gpurun_out/initcheck.log | sort | uniq -c | head
