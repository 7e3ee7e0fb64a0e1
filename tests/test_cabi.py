"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and its host-only entry points (plan queries, argument validation) behave."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from libfacedetection.train_b200 import _capi


def _declared():
    src = open(os.path.join(ROOT, 'include', 'yunet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(yunet_[a-z_0-9]+)\s*\(', src)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/yunet_b200.h but not exported'
    # and the ctypes mirror binds exactly the declared set
    assert set(_capi.EXPORTED) == set(names)


def test_library_is_sm100a_only():
    out = os.popen(f'cuobjdump -lelf {_capi.LIB_PATH} 2>/dev/null').read()
    if not out:
        pytest.skip('cuobjdump not available')
    assert 'sm_100a' in out
    assert not re.search(r'sm_(?!100a)\d+', out), out


def test_bad_arch_is_rejected_with_message():
    cfg = _capi.make_arch_cfg([[3, 16, 16], [16, 48], [48, 64], [64, 64], [64, 64], [64, 64]],
                              [0, 2, 3, 4], [3, 4, 5], 1)
    with pytest.raises(_capi.YuNetError, match='unsupported channels'):
        _capi.Ctx(cfg)


def test_geometry_queries():
    from libfacedetection.train_b200.engine import ARCHS
    a = ARCHS['yunet_n']
    ctx = _capi.Ctx(_capi.make_arch_cfg(a['stage_channels'], a['downsample_idx'], a['out_idx'],
                                        a['shared_stacked_convs']))
    assert ctx.num_priors(320, 320) == 2100
    assert ctx.num_priors(321, 320) == -1
    assert ctx.workspace_bytes(1, 300, 320, False) == 0          # not a multiple of 32
    small = ctx.workspace_bytes(2, 64, 64, False)
    assert ctx.workspace_bytes(2, 64, 64, True) > 2 * small * 0.9
    units = ctx.units()
    assert len(units) == 11 + 3 + 3 + 3     # backbone, neck, shared, fused branch units
    assert sum(1 for u in units if u.mode == 1) == 4 and sum(1 for u in units if u.mode == 2) == 2
    # null pointers are refused before any CUDA call
    assert _capi.lib.yunet_forward(ctx.handle, None, None, None, 1, 64, 64, 0, 0.1, None, None, 0,
                                   None) == -1
    assert b'null pointer' in _capi.lib.yunet_last_error(ctx.handle)


def test_engine_refuses_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from libfacedetection.train_b200 import YuNetEngine
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        YuNetEngine('yunet_n')
