"""Static properties of the built library read with ``cuobjdump`` (no GPU): it holds sm_100a code only,
the kernels the plan launches exist, and the register / stack budgets the occupancy design relies on hold —
the two-CTAs-per-SM fp32 backward of the 16-channel units (128 registers) keeps its gradient accumulators
in registers (no stack frame) since the channel-pair depthwise mapping."""
import re
import shutil
import subprocess

import pytest

from libfacedetection.train_b200 import _capi

pytestmark = pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='cuobjdump not on PATH')


def _resources():
    out = subprocess.run(['cuobjdump', '-res-usage', _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'arch = (sm_\w+)', out))
    res, cur = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'REG:(\d+) STACK:(\d+)', line)
        if m and cur:
            res[cur] = (int(m.group(1)), int(m.group(2)))
            cur = None
    return archs, res


def test_library_is_sm100a_only_and_holds_the_planned_kernels():
    archs, res = _resources()
    assert archs == {'sm_100a'}, archs
    names = ' '.join(res)
    for k in ('unit_fwd_ws_kernel', 'unit_fwd_tc_kernel', 'unit_bwd_st_kernel', 'unit_bwd_tc_kernel',
              'unit_fwd_kernel', 'unit_bwd_kernel', 'stem_fwd_kernel', 'stem_bwd_kernel', 'simota_assign_kernel',
              'loss_grad_kernel', 'decode_nms_kernel', 'sgd_kernel', 'reduce_partials_kernel',
              'preprocess'):
        assert k in names, k


def test_two_cta_fp32_backward_of_the_16_channel_units_does_not_spill():
    _, res = _resources()
    # unit_bwd_kernel<CIN, COUT, MODE, HAS_BN, OCC = 2, PAIR = 1>: Li<CIN>ELi<COUT>ELi<MODE>ELi<BN>ELi2ELi1E
    checked = 0
    for name, (reg, stack) in res.items():
        m = re.search(r'unit_bwd_kernelILi(\d+)ELi(\d+)ELi(\d)ELi(\d)ELi2ELi1E', name)
        if not m:
            continue
        cin, cout, mode = int(m.group(1)), int(m.group(2)), int(m.group(3))
        assert reg <= 128, (name, reg)            # two CTAs of 256 threads per SM
        if cin == 16 and cout == 16 and mode in (0, 1):
            assert stack == 0, (name, stack)      # plain and pooled 16 -> 16 units: no local memory at all
            checked += 1
        else:
            assert stack <= 64, (name, stack)     # the others: a few spilled words at most (quad mapping: 190..270 B)
    assert checked >= 3
