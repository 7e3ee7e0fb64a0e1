"""TEST INFRASTRUCTURE (development container only: reads /root/reference).

Pins ``libfacedetection.train_b200.export`` against the reference's own export tool: runs the
unmodified ``CppConvertor`` of ``tools/yunet2cpp.py`` on the reference detector with the shipped
weights and records the SHA-256 / length / head of the generated ``facedetectcnn-data.cpp`` in
``tests/golden/export_golden.json``; also asserts here that our exporter reproduces the text
byte for byte, and cross-checks our ONNX writer against the shipped ``onnx/yunet_*_320_320.onnx``
through OpenCV-DNN on a seeded input.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def reference_cpp(arch):
    model, _ = ref_loader.build_reference_model(arch, pretrained=True)
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, 'tools', 'yunet2cpp.py')).read()
    src = src.replace('from mmdet.core.export import build_model_from_cfg', '')
    ns = {'__name__': 'yunet2cpp_reference'}
    exec(compile(src, 'tools/yunet2cpp.py', 'exec'), ns)      # the reference file, run verbatim
    return ns['CppConvertor'](model).data


def main():
    from libfacedetection.train_b200 import export
    out = {}
    for arch in ('yunet_n', 'yunet_s'):
        ref = reference_cpp(arch)
        sd = dict(np.load(os.path.join(GOLD, f'weights_{arch}.npz')))
        ours = export.cpp_data(sd, arch)
        assert ours == ref, f'{arch}: exported cpp data differs from the reference tool'
        out[arch] = {'sha256': hashlib.sha256(ref.encode()).hexdigest(), 'length': len(ref),
                     'head': ref[:200], 'tail': ref[-160:]}
        print(f'[cpp {arch}] identical to tools/yunet2cpp.py: {len(ref)} bytes, sha256 {out[arch]["sha256"][:16]}…')
        try:
            import cv2
        except ImportError:
            continue
        torch.manual_seed(0)
        img = (torch.rand(1, 3, 320, 320) * 255).numpy()
        names = [f'{t}_{s}' for t in ('cls', 'obj', 'bbox', 'kps') for s in (8, 16, 32)]
        ref_net = cv2.dnn.readNetFromONNX(os.path.join(ref_loader.REFERENCE_ROOT, 'onnx', f'{arch}_320_320.onnx'))
        ref_net.setInput(img)
        r = ref_net.forward(names)
        net = cv2.dnn.readNetFromONNX(np.frombuffer(export.onnx_model(sd, arch, 320, 320), np.uint8))
        net.setInput(img)
        o = net.forward(names)
        worst = max(float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-6)) for a, b in zip(r, o))
        print(f'[onnx {arch}] ours vs shipped onnx through OpenCV-DNN: worst rel diff {worst:.2e}')
        assert worst < 1e-4
        out[arch]['onnx_vs_shipped'] = worst
    json.dump(out, open(os.path.join(GOLD, 'export_golden.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
