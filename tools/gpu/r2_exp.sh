#!/bin/bash
# experiments: default build, then the development knobs, one short bench each
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_plugins.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_exp.log 2>&1; tail -3 gpurun_out/pytest_exp.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --kernel-table gpurun_out/k_$tag.json > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_{tag}.log').read().strip().splitlines()[-1])
    k = json.load(open(f'gpurun_out/k_{tag}.json'))
    print(tag, 'ms/step', round(d['ms_per_step'], 3), 'eager', round(d['cuda_graph']['eager_ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3))
    print('   ', ' | '.join('%s %.3f' % (r['kernel'].replace('backbone.', ''), r['ms']) for r in k[:16]))
except Exception as e:
    print(tag, 'failed', e); print(open(f'gpurun_out/bench_{tag}.log').read()[-600:])
P
}
run default A=1
python - <<'P'
import torch, time
h = torch.empty(314572800 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device='cuda')
s = torch.cuda.Stream()
for n in range(3):
    with torch.cuda.stream(s):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            d.copy_(h, non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    print('H2D pinned 314.6 MB x5: %.1f GB/s' % (5 * 0.3145728 / (e0.elapsed_time(e1) / 1e3)))
P
