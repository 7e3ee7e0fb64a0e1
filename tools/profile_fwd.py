"""A few train-mode forwards (bs 256, 320x320) — the command ncu wraps for the forward kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from libfacedetection.train_b200 import YuNetEngine, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eng = YuNetEngine(sys.argv[2] if len(sys.argv) > 2 else 'yunet_n')
eng.init_weights(0)
img = torch.from_numpy(synthetic.make_images(256, 320, 0)).cuda()
for _ in range(n):
    eng.forward(img, train=True)
torch.cuda.synchronize()
