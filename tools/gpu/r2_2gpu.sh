#!/bin/bash
# gpurun --gpus 2 -- 'bash tools/gpu/r2_2gpu.sh' : both arms as the driver launches them at N = 2
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/bench_2gpu.log
timeout 400 $T --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_2gpu_nograph.log 2>&1; echo "exit $?" >> gpurun_out/bench_2gpu_nograph.log
timeout 400 $T --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "exit $?" >> gpurun_out/bench_2gpu_ref.log
python - <<'P'
import json
for f in ('bench_2gpu', 'bench_2gpu_nograph', 'bench_2gpu_ref'):
    t = open(f'gpurun_out/{f}.log').read().strip().splitlines()
    print(f, t[-1])
    try:
        d = json.loads(t[-2])
        print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'cuda_graph', 'comm_ms', 'e2e', 'impl')})
    except Exception as e:
        print('  no json:', e, '|', '\n'.join(t[-6:])[-1200:])
P
