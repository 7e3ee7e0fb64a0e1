"""Run N plain training steps (bs=256, 320x320, yunet_n) and nothing else — the command ncu wraps
for the per-launch list and the `--set full` capture (see /opt/skills/guides/B200_PROFILING.md).
Numbers printed by a run under ncu are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from libfacedetection.train_b200 import YuNetEngine, synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
arch = sys.argv[2] if len(sys.argv) > 2 else 'yunet_n'
B, S = 256, 320
eng = YuNetEngine(arch)
eng.init_weights(0)
img = torch.from_numpy(synthetic.make_images(B, S, 0)).cuda()
gb, gl, gk = synthetic.make_gt(B, S, 0)
gt, offs = synthetic.pack_gt_csr(gb, gk)
gt, offs = torch.from_numpy(gt).cuda(), torch.from_numpy(offs).cuda()
for _ in range(steps):
    losses = eng.train_step(img, gt, offs, lr=1e-5)
torch.cuda.synchronize()
print('losses', losses.cpu().tolist())
