#!/bin/bash
bash tools/gpu/run_tests_bench.sh
bash tools/gpu/run_phase_timing.sh
