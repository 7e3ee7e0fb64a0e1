"""Deployment export from the flat parameter store: the reference's two hand-off
formats, produced from a reference-format ``state_dict`` (``engine.state_dict()`` or a loaded
``weights/yunet_*.pth``) without mmdet / mmcv / onnx:

  * ``cpp_data(state_dict, arch)``: the ``facedetectcnn-data.cpp`` weight file of libfacedetection
    (``tools/yunet2cpp.py:24-150``): BatchNorm folded into the preceding convolution
    (``combine_conv_bn``, ``yunet2cpp.py:42-51``), the stem's 3x3x3 kernel re-ordered to
    (tap, channel) and padded to 32 (``yunet2cpp.py:60-67``), depthwise kernels tap-major
    (``yunet2cpp.py:68-69``), ``.3g`` floats with the ``f`` suffix (``yunet2cpp.py:18-23``), followed
    by the ``ConvInfoStruct`` table (``yunet2cpp.py:135-149``).  Byte-identical to the reference tool
    (pinned by ``tests/golden/export_golden.json``).
  * ``onnx_model(state_dict, arch, height, width)``: the 12-output ONNX graph of
    ``tools/yunet2onnx.py:86-108`` / ``yunet_head.py:227-245`` (``cls_*``, ``obj_*`` after sigmoid,
    ``bbox_*``, ``kps_*`` as ``(1, H*W, C)`` for strides 8/16/32), BatchNorm folded, serialised with
    a minimal protobuf writer (the ``onnx`` package is not needed).

Pure host code (torch CPU for the fold so that the arithmetic equals the reference's).
"""
import struct

import numpy as np
import torch

from .engine import ARCHS

BN_EPS = 1e-5


# --------------------------------------------------------------------------- module walk
def _t(sd, key):
    v = sd[key]
    return v.detach().float().cpu() if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)).float()


def _fold(w, b, sd, bn):
    """conv (w, b) followed by BatchNorm ``bn`` (eval) -> one conv; fp32, reference operation order."""
    scales = _t(sd, bn + '.weight') / torch.sqrt(_t(sd, bn + '.running_var') + BN_EPS)
    bias = (b - _t(sd, bn + '.running_mean')) * scales + _t(sd, bn + '.bias')
    weight = w * scales.reshape(-1, 1, 1, 1)
    return weight, bias


def walk_units(sd, arch):
    """The detector's convolutions in the reference's module order (``named_children`` recursion of
    ``yunet2cpp.py:117-125``): list of dicts ``name, weight (OIHW), bias, is_dw, with_bn, first``
    with BatchNorm already folded where the module has one."""
    a = ARCHS[arch] if isinstance(arch, str) else arch
    out = []

    def dp_unit(prefix, name, with_bn):
        out.append(dict(name=name + '_pw', weight=_t(sd, prefix + '.conv1.weight'),
                        bias=_t(sd, prefix + '.conv1.bias'), is_dw=False, with_bn=False, first=False))
        w, b = _t(sd, prefix + '.conv2.weight'), _t(sd, prefix + '.conv2.bias')
        if with_bn:
            w, b = _fold(w, b, sd, prefix + '.bn')
        out.append(dict(name=name + '_dw', weight=w, bias=b, is_dw=True, with_bn=with_bn, first=False))

    nstage = len(a['stage_channels'])
    w, b = _fold(_t(sd, 'backbone.model0.conv1.weight'), _t(sd, 'backbone.model0.conv1.bias'), sd,
                 'backbone.model0.bn1')
    out.append(dict(name='backbone__model0_pw', weight=w, bias=b, is_dw=False, with_bn=True, first=True))
    dp_unit('backbone.model0.conv2', 'backbone__model0_dp', True)
    for s in range(1, nstage):
        dp_unit(f'backbone.model{s}.conv1', f'backbone__model{s}_dp1', True)
        dp_unit(f'backbone.model{s}.conv2', f'backbone__model{s}_dp2', True)
    for i in range(3):
        dp_unit(f'neck.lateral_convs.{i}', f'neck__lateral_convs__{i}', True)
    for i in range(3):
        for j in range(a.get('shared_stacked_convs', 0)):
            dp_unit(f'bbox_head.multi_level_share_convs.{i}.{j}',
                    f'bbox_head__multi_level_share_convs__{i}__{j}', True)
    for branch in ('cls', 'bbox', 'obj', 'kps'):
        for i in range(3):
            dp_unit(f'bbox_head.multi_level_{branch}.{i}', f'bbox_head__multi_level_{branch}__{i}', False)
    return out


# --------------------------------------------------------------------------- facedetectcnn-data.cpp
def _fmt(x, precision='.3g'):
    s = format(x, precision)
    return s + '.f' if (s.count('.') == 0 and s.count('e') == 0) else s + 'f'


def _cbool(v):
    return 'true' if v else 'false'


def cpp_data(sd, arch='yunet_n'):
    """Text of ``facedetectcnn-data.cpp`` for the given weights (see module docstring)."""
    rows = []
    for u in walk_units(sd, arch):
        wt = u['weight']
        oc, ic, kh, kw = wt.shape
        if u['first']:
            w = wt.numpy().reshape(-1, 27)
            src = w.copy()
            for off in range(27):                       # (c, ky, kx) -> (ky*3+kx, c)
                w[:, (off % 9) * 3 + off // 9] = src[:, off]
            w = np.hstack((w, np.zeros((oc, 5)))).reshape(-1)
            wsize, in_ch = f'{oc}*32*1*1', 32
        elif u['is_dw']:
            w = wt.numpy().reshape(-1, 9).transpose().reshape(-1)
            wsize, in_ch = f'{oc}*{ic}*{kh}*{kw}', oc
        else:
            w = wt.numpy().reshape(-1)
            wsize, in_ch = f'{oc}*{ic}*{kh}*{kw}', ic
        b = u['bias'].numpy().reshape(-1)
        rows.append(dict(name=u['name'], wsize=wsize, w=','.join(_fmt(v) for v in w), bsize=str(oc),
                         b=','.join(_fmt(v) for v in b), in_ch=in_ch, out_ch=oc, is_dw=u['is_dw'],
                         with_bn=u['with_bn']))
    text = ('// Auto generated data file\n// Copyright (c) 2018-2023, Shiqi Yu, all rights reserved.\n'
            '#include "facedetectcnn.h"\n\n')
    for r in rows:
        text += f"float {r['name']}_weight[{r['wsize']}] = {{{r['w']}}};\n"
        text += f"float {r['name']}_bias[{r['bsize']}] = {{{r['b']}}};\n"
    text += '\n//(in_channels, out_channels, is_depthwise, is_pointwise, with_bn, weight_ptr, bias_ptr)\n'
    text += f'ConvInfoStruct param_pConvInfo[{len(rows)}] = {{\n'
    lines = [f"\t{{{r['in_ch']}, {r['out_ch']}, {_cbool(r['is_dw'])}, {_cbool(not r['is_dw'])}, "
             f"{_cbool(r['with_bn'])}, {r['name']}_weight, {r['name']}_bias}}" for r in rows]
    text += ',\n'.join(lines) + '\n};'
    return text


# --------------------------------------------------------------------------- minimal protobuf / ONNX
def _varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wire):
    return _varint((field << 3) | wire)


def _f_varint(field, v):
    return _key(field, 0) + _varint(int(v))


def _f_bytes(field, payload):
    if isinstance(payload, str):
        payload = payload.encode()
    return _key(field, 2) + _varint(len(payload)) + payload


def _f_float(field, v):
    return _key(field, 5) + struct.pack('<f', float(v))


def _tensor(name, arr):
    """TensorProto: dims=1, data_type=2, name=8, raw_data=9 (FLOAT = 1, INT64 = 7)."""
    arr = np.ascontiguousarray(arr)
    dt = {np.dtype('float32'): 1, np.dtype('int64'): 7}[arr.dtype]
    msg = b''.join(_f_varint(1, d) for d in arr.shape)
    return msg + _f_varint(2, dt) + _f_bytes(8, name) + _f_bytes(9, arr.tobytes())


def _attr(name, value):
    """AttributeProto: name=1, f=2, i=3, s=4, floats=7, ints=8, type=20."""
    msg = _f_bytes(1, name)
    if isinstance(value, float):
        return msg + _f_float(2, value) + _f_varint(20, 1)
    if isinstance(value, int):
        return msg + _f_varint(3, value) + _f_varint(20, 2)
    if isinstance(value, str):
        return msg + _f_bytes(4, value) + _f_varint(20, 3)
    if isinstance(value, (list, tuple)) and value and isinstance(value[0], float):
        return msg + b''.join(_f_float(7, v) for v in value) + _f_varint(20, 6)
    return msg + b''.join(_f_varint(8, v) for v in value) + _f_varint(20, 7)


def _node(op, inputs, outputs, name, **attrs):
    """NodeProto: input=1, output=2, name=3, op_type=4, attribute=5."""
    msg = b''.join(_f_bytes(1, i) for i in inputs) + b''.join(_f_bytes(2, o) for o in outputs)
    msg += _f_bytes(3, name) + _f_bytes(4, op)
    return msg + b''.join(_f_bytes(5, _attr(k, v)) for k, v in attrs.items())


def _value_info(name, shape):
    """ValueInfoProto(name=1, type=2{tensor_type=1{elem_type=1, shape=2{dim=1{dim_value=1 | dim_param=2}}}});
    a ``str`` entry of ``shape`` is a symbolic dimension (``dim_param``), an ``int`` a fixed one."""
    dims = b''.join(_f_bytes(1, _f_bytes(2, d) if isinstance(d, str) else _f_varint(1, d)) for d in shape)
    ttype = _f_varint(1, 1) + _f_bytes(2, dims)
    return _f_bytes(1, name) + _f_bytes(2, _f_bytes(1, ttype))


def onnx_model(sd, arch='yunet_n', height=320, width=320, opset=11, dynamic=False):
    """Serialized ONNX ModelProto of the detector's export graph for a fixed ``(1, 3, H, W)`` input;
    outputs ``cls_8, cls_16, cls_32, obj_*, bbox_*, kps_*`` like ``tools/yunet2onnx.py:86-93``.
    ``dynamic=True`` is the tool's ``--dynamic-export`` (``tools/yunet2onnx.py:97-100``, the shipped
    ``onnx/yunet_*_dynamic.onnx``): input axes ``{0: 'batch', 2: 'height', 3: 'width'}``, output axes
    ``{0: 'batch', 1: 'dim'}``; the graph itself is shape-agnostic (convolutions, 2x2 pools, scale-2 nearest
    resize), only the flattening ``Reshape`` takes the batch from its input (target shape ``[0, -1, C]``)
    instead of the constant 1."""
    a = ARCHS[arch] if isinstance(arch, str) else arch
    units = {u['name']: u for u in walk_units(sd, arch)}
    nodes, inits = [], []
    counter = [0]

    def fresh(tag):
        counter[0] += 1
        return f'{tag}_{counter[0]}'

    def conv(x, u, stride=1):
        w, b = u['weight'].numpy().astype(np.float32), u['bias'].numpy().astype(np.float32)
        k = w.shape[2]
        inits.append(_tensor(u['name'] + '_w', w))
        inits.append(_tensor(u['name'] + '_b', b))
        y = fresh(u['name'])
        nodes.append(_node('Conv', [x, u['name'] + '_w', u['name'] + '_b'], [y], y,
                           dilations=[1, 1], group=(w.shape[0] if u['is_dw'] else 1),
                           kernel_shape=[k, k], pads=[k // 2] * 4, strides=[stride, stride]))
        return y

    def relu(x):
        y = fresh('relu')
        nodes.append(_node('Relu', [x], [y], y))
        return y

    def dp(x, name, act=True):
        y = conv(conv(x, units[name + '_pw']), units[name + '_dw'])
        return relu(y) if act else y

    def pool(x):
        y = fresh('pool')
        nodes.append(_node('MaxPool', [x], [y], y, kernel_shape=[2, 2], pads=[0, 0, 0, 0], strides=[2, 2]))
        return y

    x = relu(conv('input', units['backbone__model0_pw'], stride=2))
    x = dp(x, 'backbone__model0_dp')
    nstage = len(a['stage_channels'])
    feats = []
    for s in range(nstage):
        if s > 0:
            x = dp(dp(x, f'backbone__model{s}_dp1'), f'backbone__model{s}_dp2')
        if s in a['out_idx']:
            feats.append(x)
        if s in a['downsample_idx']:
            x = pool(x)
    # TFPN (necks/tfpn.py:33-45): top-down, nearest x2 upsample + add, lateral unit after the add
    inits.append(_tensor('up_roi', np.zeros((0,), np.float32)))
    inits.append(_tensor('up_scales', np.array([1, 1, 2, 2], np.float32)))
    outs = [None, None, None]
    outs[2] = dp(feats[2], 'neck__lateral_convs__2')
    for i in (1, 0):
        up = fresh('up')
        nodes.append(_node('Resize', [outs[i + 1], 'up_roi', 'up_scales'], [up], up,
                           coordinate_transformation_mode='asymmetric', mode='nearest',
                           nearest_mode='floor'))
        s_ = fresh('add')
        nodes.append(_node('Add', [feats[i], up], [s_], s_))
        outs[i] = dp(s_, f'neck__lateral_convs__{i}')
    for i in range(3):
        for j in range(a.get('shared_stacked_convs', 0)):
            outs[i] = dp(outs[i], f'bbox_head__multi_level_share_convs__{i}__{j}')
    graph_outputs = []
    strides = (8, 16, 32)
    for branch, nch, sig in (('cls', 1, True), ('obj', 1, True), ('bbox', 4, False), ('kps', 10, False)):
        for i in range(3):
            y = dp(outs[i], f'bbox_head__multi_level_{branch}__{i}', act=False)
            t = fresh('nhwc')
            nodes.append(_node('Transpose', [y], [t], t, perm=[0, 2, 3, 1]))
            shp = f'shape_{branch}_{i}'
            inits.append(_tensor(shp, np.array([0 if dynamic else 1, -1, nch], np.int64)))
            name = f'{branch}_{strides[i]}'
            r = name if not sig else fresh('flat')
            nodes.append(_node('Reshape', [t, shp], [r], r))
            if sig:
                nodes.append(_node('Sigmoid', [r], [name], name))
            hw = (height // strides[i]) * (width // strides[i])
            graph_outputs.append(_value_info(name, ['batch', 'dim', nch] if dynamic else [1, hw, nch]))
    # GraphProto: node=1, name=2, initializer=5, input=11, output=12
    graph = b''.join(_f_bytes(1, n) for n in nodes) + _f_bytes(2, 'yunet_b200')
    graph += b''.join(_f_bytes(5, t) for t in inits)
    graph += _f_bytes(11, _value_info('input', ['batch', 3, 'height', 'width'] if dynamic else [1, 3, height, width]))
    graph += b''.join(_f_bytes(12, o) for o in graph_outputs)
    # ModelProto: ir_version=1, producer_name=2, graph=7, opset_import=8{version=2}
    return (_f_varint(1, 6) + _f_bytes(2, 'libfacedetection.train_b200') + _f_bytes(7, graph) +
            _f_bytes(8, _f_bytes(1, '') + _f_varint(2, opset)))


def main(argv=None):
    """``python -m libfacedetection.train_b200.export yunet_n weights/yunet_n.pth --cpp out.cpp --onnx out.onnx``
    (the two reference CLIs ``tools/yunet2cpp.py`` / ``tools/yunet2onnx.py`` in one, no mmdet config)."""
    import argparse
    ap = argparse.ArgumentParser(description='Export YuNet weights to libfacedetection cpp data / ONNX')
    ap.add_argument('arch', choices=sorted(ARCHS))
    ap.add_argument('checkpoint', help='reference-format .pth (state_dict or {"state_dict": ...}) or .npz')
    ap.add_argument('--cpp', default=None, help='write facedetectcnn-data.cpp here')
    ap.add_argument('--onnx', default=None, help='write the 12-output ONNX model here')
    ap.add_argument('--shape', type=int, nargs=2, default=[640, 640], help='ONNX input height width')
    ap.add_argument('--dynamic-export', action='store_true',
                    help='ONNX with dynamic batch / height / width axes (tools/yunet2onnx.py --dynamic-export)')
    args = ap.parse_args(argv)
    if args.checkpoint.endswith('.npz'):
        sd = dict(np.load(args.checkpoint))
    else:
        ck = torch.load(args.checkpoint, map_location='cpu', weights_only=False)
        sd = ck.get('state_dict', ck)
    if args.cpp:
        with open(args.cpp, 'w') as f:
            f.write(cpp_data(sd, args.arch))
    if args.onnx:
        with open(args.onnx, 'wb') as f:
            f.write(onnx_model(sd, args.arch, args.shape[0], args.shape[1], dynamic=args.dynamic_export))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
