#!/bin/bash
# last seconds of the round's GPU budget: smoke() and the plugin / trainer tests on the final build
mkdir -p gpurun_out
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200
timeout -s KILL 90 python -m pytest -q -p no:cacheprovider --timeout 60 --timeout-method=thread -m gpu tests/test_gpu_plugins.py tests/test_gpu_trainer.py 2>&1 | tail -3
