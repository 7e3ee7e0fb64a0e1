#!/bin/bash
# iteration loop of the ws forward kernel: parity (forward tests) + bench table + ncu + per-role timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "forward or every_unit" -x > gpurun_out/pytest_ws.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ws.log
grep -E "^(FAILED|ERROR)|passed|failed|kernel reported|Error|assert" gpurun_out/pytest_ws.log | tail -12
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernels.json > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-200
python - <<'P'
import json
k=json.load(open('gpurun_out/kernels.json'))
rows=k if isinstance(k,list) else k.get('kernels',k)
tot={}
for r in rows:
    fam=r['kernel'].split(':')[0]
    tot[fam]=tot.get(fam,0)+r['ms']
print({a:round(b,3) for a,b in tot.items()})
for r in rows:
    if r['kernel'].startswith('fwd'):
        print('%-52s %7.3f ms %7.0f GB/s' % (r['kernel'], r['ms'], r['gbs'] or 0))
P
if [ "$1" = "ncu" ]; then
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:unit_fwd_ws_kernel -s 17 -c 1 -o gpurun_out/prof_fwd_ws python tools/profile_fwd.py 2 > gpurun_out/ncu_fwd_ws.log 2>&1
tail -2 gpurun_out/ncu_fwd_ws.log
fi
cd libfacedetection/train_b200/csrc && touch unit_fwd_ws.cu && make EXTRA=-DYUNET_WS_TIMING > /root/repo/gpurun_out/make_timing.log 2>&1; cd /root/repo
timeout 300 python tools/ws_timing.py > gpurun_out/ws_timing.log 2>&1
tail -12 gpurun_out/ws_timing.log
