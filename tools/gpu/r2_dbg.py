import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from libfacedetection.train_b200 import YuNetEngine, evaluation, synthetic, trainer
from oracle import yunet_oracle as orc
GOLDEN = '/root/repo/tests/golden'
d = np.load(os.path.join(GOLDEN, 'weights_yunet_n.npz'))
sd = {k: torch.from_numpy(d[k]) for k in d.files}


def cmp(tag, a, b):
    for name, x, y in (('params', a.params, b.params), ('mom', a.momentum_buf, b.momentum_buf), ('bn', a.bn_running, b.bn_running)):
        dlt = (x - y).abs()
        i = int(dlt.argmax())
        print(tag, name, 'max abs', float(dlt.max()), 'at', i, float(x[i]), float(y[i]), 'rel', float((dlt / (y.abs() + 1e-6)).max()))
    va, vb = a.param_views(a.grads), b.param_views(b.grads)
    worst = sorted(((float((va[n] - vb[n]).abs().max() / (vb[n].abs().max() + 1e-12)), n) for n in va), reverse=True)[:4]
    print(tag, 'grad rel', worst)

def _batches(B, size, n, seed0=0):
    out = []
    for i in range(n):
        img = torch.from_numpy(synthetic.make_images(B, size, seed0 + i)).cuda()
        gb, gl, gk = synthetic.make_gt(B, size, seed0 + i)
        gt, offs = synthetic.pack_gt_csr(gb, gk)
        out.append((img, torch.from_numpy(gt).cuda(), torch.from_numpy(offs).cuda()))
    return out
for mode in ('eager-eager', 'eager-graph', 'graph-graph'):
    a, b = YuNetEngine('yunet_n'), YuNetEngine('yunet_n')
    a.load_state_dict(sd); b.load_state_dict(sd)
    data = _batches(8, 320, 2, seed0=5)
    for it in range(8):
        lr = trainer.lr_at(it, 0)
        la = (a.train_step_graph if mode == 'graph-graph' else a.train_step)(*data[it % 2], lr=lr).clone()
        lb = (b.train_step_graph if mode == 'eager-graph' else b.train_step)(*data[it % 2], lr=lr).clone()
        print(mode, it, (la - lb).abs().max().item())
        cmp(mode + str(it), a, b)

