#!/bin/bash
# gpurun --gpus 2 -- 'bash tools/gpu/run_2gpu.sh'
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/bench_2gpu.log
tail -3 gpurun_out/bench_2gpu.log | cut -c1-1500
