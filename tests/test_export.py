"""SURVEY §8f N3 — export parity (host code, CPU): ``facedetectcnn-data.cpp`` byte-identical to the
reference's ``tools/yunet2cpp.py`` (golden = SHA-256 recorded from the unmodified tool by
``oracle/gen_golden_export.py``), and the hand-serialised 12-output ONNX graph evaluated by
OpenCV-DNN against the reference forward fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from libfacedetection.train_b200 import export

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _weights(arch):
    return dict(np.load(os.path.join(GOLD, f'weights_{arch}.npz')))


@pytest.mark.parametrize('arch', ['yunet_n', 'yunet_s'])
def test_cpp_data_is_byte_identical_to_reference_tool(arch):
    gold = json.load(open(os.path.join(GOLD, 'export_golden.json')))[arch]
    text = export.cpp_data(_weights(arch), arch)
    assert len(text) == gold['length']
    assert text[:200] == gold['head'] and text[-160:] == gold['tail']
    assert hashlib.sha256(text.encode()).hexdigest() == gold['sha256']


def test_cpp_data_structure():
    text = export.cpp_data(_weights('yunet_n'), 'yunet_n')
    # stem conv + 28 ConvDPUnits with BN or not: 1 + 2 * 29 - 1 ... count the table rows instead
    rows = [l for l in text.splitlines() if l.startswith('\t{')]
    assert len(rows) == 59 and f'param_pConvInfo[{len(rows)}]' in text
    assert rows[0].startswith('\t{32, 16, false, true, true, backbone__model0_pw_weight')
    assert text.count('float ') == 2 * len(rows)


@pytest.mark.parametrize('arch,size', [('yunet_n', 320), ('yunet_s', 320), ('yunet_n', 640)])
def test_onnx_graph_matches_reference_forward(arch, size):
    cv2 = pytest.importorskip('cv2')
    g = np.load(os.path.join(GOLD, f'forward_{arch}_{size}.npz'))
    blob = export.onnx_model(_weights(arch), arch, size, size)
    net = cv2.dnn.readNetFromONNX(np.frombuffer(blob, np.uint8))
    torch.manual_seed(0)
    img = (torch.rand(1, 3, size, size) * 255).numpy()      # the input of oracle/gen_golden.forward_case
    net.setInput(img)
    names = [f'{t}_{s}' for t in ('cls', 'obj', 'bbox', 'kps') for s in (8, 16, 32)]
    outs = dict(zip(names, net.forward(names)))
    preds = g['preds'][0]                                    # (P, 16) = [cls, bbox4, obj, kps10] logits
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    cat = lambda t: np.concatenate([outs[f'{t}_{s}'].reshape(-1, outs[f'{t}_{s}'].shape[-1]) for s in (8, 16, 32)])
    ref = {'cls': sig(preds[:, 0:1]), 'bbox': preds[:, 1:5], 'obj': sig(preds[:, 5:6]), 'kps': preds[:, 6:16]}
    for t in ('cls', 'bbox', 'obj', 'kps'):
        got = cat(t)
        assert got.shape == ref[t].shape
        err = np.abs(got - ref[t]).max() / max(np.abs(ref[t]).max(), 1e-6)
        assert err < 1e-4, (t, err)


def test_onnx_bytes_are_deterministic_and_wellformed():
    sd = _weights('yunet_s')
    a, b = export.onnx_model(sd, 'yunet_s', 320, 320), export.onnx_model(sd, 'yunet_s', 320, 320)
    assert a == b and a[:1] == b'\x08'                       # field 1 (ir_version), varint
    assert b'Conv' in a and b'Resize' in a and b'kps_32' in a
