"""Per-phase cycle counts of the tensor-core backward kernel (80x80 plain units), CTA 0 / thread 0.
Needs the library built with `make -C libfacedetection/train_b200/csrc EXTRA=-DYUNET_PHASE_TIMING`
(touch unit_bwd_tc.cu first); counters live in the status words 32..47 of the workspace."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from libfacedetection.train_b200 import YuNetEngine, synthetic, _capi  # noqa: E402

B, S = 256, 320
eng = YuNetEngine('yunet_n')
eng.init_weights(0)
img = torch.from_numpy(synthetic.make_images(B, S, 0)).cuda()
gb, gl, gk = synthetic.make_gt(B, S, 0)
gt, offs = synthetic.pack_gt_csr(gb, gk)
gt, offs = torch.from_numpy(gt).cuda(), torch.from_numpy(offs).cuda()
for _ in range(2):
    eng.train_step(img, gt, offs, lr=1e-5)
torch.cuda.synchronize()
off = _capi.lib.yunet_ws_offset(eng.h, B, S, S, 1, 0, 3)
ws = eng.workspace(B, S, S, True)
st = ws[off:off + 256].view(torch.int32)
st[32:50] = 0
eng.train_step(img, gt, offs, lr=1e-5)
torch.cuda.synchronize()
c = st.cpu().tolist()
names = ['wait du/z_out', 'g pass', 'wait z_in', 'T1 convert a', 'MMA1 issue+wait', 'T3 y->smem',
         'T4 depthwise', 'T5 dy->TMEM (+T)', 'MMA2/3 issue + wait MMA2', 'epilogue', 'collect dW1',
         'end sync']
ntile = max(c[48], 1)
tot = sum(c[33:45])
print(f'tiles timed {c[48]}  cycles/tile {tot / ntile:.0f}  flags {c[:4]}')
for i, n in enumerate(names):
    print(f'  {n:28s} {c[33 + i] / ntile:8.0f} cyc  {100.0 * c[33 + i] / max(tot, 1):5.1f}%')
