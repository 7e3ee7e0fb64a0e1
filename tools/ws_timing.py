"""Per-role cycle counters of the warp-specialised forward kernel (80x80 64->64 plain units), CTA 0.
Needs the library built with `make -C libfacedetection/train_b200/csrc EXTRA=-DYUNET_WS_TIMING`
(touch unit_fwd_ws.cu first); counters live in the status words 32.. of the workspace."""
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
for _ in range(3):
    eng.forward(img, train=True)
torch.cuda.synchronize()
off = _capi.lib.yunet_ws_offset(eng.h, B, S, S, 1, 0, 3)
ws = eng.workspace(B, S, S, True)
st = ws[off:off + 256].view(torch.int32)
c = st.cpu().tolist()
names = {0: 'mma: wait a_full/d_empty (1 warp)', 1: 'mma: issue', 2: 'cv: wait in_full (2 x 4 warps, each every other block)',
         3: 'cv: convert', 4: 'ep: wait mma_done', 5: 'ep: tmem ld + wait y_empty', 6: 'ep: y store',
         7: 'dw: wait y_full (10 warps)', 8: 'dw: stencil + store', 9: 'ep: tcgen05.ld (of the y store time)'}
div = {0: 1, 1: 1, 2: 8, 3: 8, 4: 8, 5: 8, 6: 8, 7: 10, 8: 10, 9: 8}   # warps adding to each counter
nblk = 2 * 98   # two launches (model2.conv1/conv2), ~97 + 1 blocks each for CTA 0
print('flags', c[:4])
for k, n in names.items():
    print(f'  {n:36s} {c[32 + k] / div[k] / nblk:8.0f} cyc/block')
