"""The reference arm of bench.py (``--impl reference``) is CPU-only: it must print ONE JSON line with the
contract's keys and must not load the product library (it imports oracle/ and the stand-alone synthetic
generator only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line_without_the_product():
    code = (
        "import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', "
        "'--cpu-sample', '2', '--size', '64']\n"
        "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('LOADED_PRODUCT_SO', 'libyunet_b200' in maps)\n")
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'train_images_per_sec_320' and d['unit'] == 'images/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['value'] > 0 and d['gpu_launches'] == 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value']
    assert set(cb['split_ms']) == {'forward', 'assign_loss', 'backward', 'sgd'}
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'LOADED_PRODUCT_SO False' in out.stdout
