"""Minimal training loop around the fused engine: the reference's learning-rate
schedule, checkpoint format and iteration order, without the mmcv runner.

Reference behaviour reproduced:
  * SGD lr 0.01 / momentum 0.9 / weight decay 5e-4 on all parameters (``configs/yunet_n.py:1``);
  * ``lr_config``: step policy, x0.1 at epochs 400 and 544 of 640, linear warm-up over the first
    1 500 iterations starting at 0.001 x lr (``configs/yunet_n.py:4-11``; mmcv ``StepLrUpdaterHook``
    with ``warmup='linear'``: lr_warm = lr_regular * (1 - (1 - it/warmup_iters) * (1 - ratio)));
  * checkpoints ``{'meta', 'state_dict', 'optimizer'}`` as written by mmcv's ``CheckpointHook`` and
    read by ``runner.resume`` (``mmdet/apis/train.py:234-243``), optimizer state in
    ``torch.optim.SGD.state_dict()`` layout (parameter order = ``model.parameters()`` order of the
    reference detector), so the shipped ``weights/yunet_{n,s}.pth`` can be resumed from and files
    written here load into the reference.
"""
import time

import numpy as np
import torch


def lr_at(iteration, epoch, base_lr=0.01, steps=(400, 544), gamma=0.1, warmup_iters=1500,
          warmup_ratio=0.001):
    """Learning rate of iteration ``iteration`` (0-based, global) inside epoch ``epoch``."""
    regular = base_lr * (gamma ** sum(1 for s in steps if epoch >= s))
    if iteration < warmup_iters:
        k = (1 - iteration / warmup_iters) * (1 - warmup_ratio)
        return regular * (1 - k)
    return regular


def reference_param_order(engine_or_names):
    """``model.parameters()`` order of the reference detector: backbone (model0.conv1, model0.conv2.*,
    model0.bn1, model1..), neck.lateral_convs.0..2, bbox_head share convs, cls, bbox, obj, kps."""
    names = [n for n, _, _ in engine_or_names.param_table] if hasattr(engine_or_names, 'param_table') \
        else list(engine_or_names)

    def key(n):
        parts = n.split('.')
        if parts[0] == 'backbone':
            stage = int(parts[1][5:])
            sub = {'conv1': 0, 'conv2': 1, 'bn1': 2}[parts[2]] if stage == 0 else \
                {'conv1': 0, 'conv2': 1}[parts[2]]
            return (0, stage, sub, _leaf_rank(parts[3:]))
        if parts[0] == 'neck':
            return (1, int(parts[2]), 0, _leaf_rank(parts[3:]))
        group = {'multi_level_share_convs': 0, 'multi_level_cls': 1, 'multi_level_bbox': 2,
                 'multi_level_obj': 3, 'multi_level_kps': 4}[parts[1]]
        rest = parts[3:] if group else parts[4:]
        return (2, group, int(parts[2]), _leaf_rank(rest))

    return sorted(names, key=key)


def _leaf_rank(parts):
    # inside a ConvDPUnit: conv1.weight, conv1.bias, conv2.weight, conv2.bias, bn.weight, bn.bias;
    # a bare conv / bn: weight, bias
    order = {'conv1': 0, 'conv2': 2, 'bn': 4, 'weight': 0, 'bias': 1}
    return sum(order[p] for p in parts)


def reference_state_dict_order(keys):
    """Key order of the reference detector's ``state_dict()``: parameters in
    ``model.parameters()`` order, every BatchNorm followed by its three buffers."""
    keys = list(keys)
    params = [k for k in keys if not k.endswith(('.running_mean', '.running_var', '.num_batches_tracked'))]
    out = []
    have = set(keys)
    for k in reference_param_order(params):
        out.append(k)
        if k.endswith('.bias'):
            prefix = k[:-len('.bias')]
            for b in ('running_mean', 'running_var', 'num_batches_tracked'):
                if prefix + '.' + b in have:
                    out.append(prefix + '.' + b)
    assert len(out) == len(keys), (len(out), len(keys))
    return out


def save_checkpoint(engine, path, epoch=0, iteration=0, lr=0.01, momentum=0.9, weight_decay=0.0005,
                    meta=None):
    """``state_dict`` is written in the reference's key order, so that optimizer-state index i is the
    i-th parameter key of the file (what ``load_checkpoint`` and the reference's ``runner.resume``
    assume); ``optimizer['param_names']`` additionally records the mapping by name."""
    raw = {k: v.cpu() for k, v in engine.state_dict().items()}
    sd = {k: raw[k] for k in reference_state_dict_order(raw.keys())}
    order = reference_param_order(engine)
    mom = engine.param_views(engine.momentum_buf)
    opt = {'state': {i: {'momentum_buffer': mom[n].detach().cpu().clone()} for i, n in enumerate(order)},
           'param_groups': [{'lr': lr, 'momentum': momentum, 'dampening': 0,
                             'weight_decay': weight_decay, 'nesterov': False,
                             'params': list(range(len(order)))}],
           'param_names': list(order)}
    m = {'epoch': epoch, 'iter': iteration, 'time': time.asctime()}
    m.update(meta or {})
    torch.save({'meta': m, 'state_dict': sd, 'optimizer': opt}, path)


def load_checkpoint(engine, path, resume_optimizer=True):
    ck = torch.load(path, map_location='cpu', weights_only=False)
    engine.load_state_dict(ck['state_dict'], strict=True)
    if resume_optimizer and 'optimizer' in ck:
        mom = engine.param_views(engine.momentum_buf)
        # files written here carry the names; foreign files (the reference's own checkpoints) index
        # the optimizer state by model.parameters() of the code that SAVED the file, which is the
        # parameter-key order of its state_dict (weights/yunet_s.pth lists kps before obj)
        order = ck['optimizer'].get('param_names') or [k for k in ck['state_dict'] if k in mom]
        st = ck['optimizer']['state']
        for i, n in enumerate(order):
            if i in st and 'momentum_buffer' in st[i]:
                buf = st[i]['momentum_buffer']
                if buf.numel() != mom[n].numel():
                    raise ValueError(f'optimizer state {i} ({tuple(buf.shape)}) does not match '
                                     f'parameter {n} {tuple(mom[n].shape)}')
                mom[n].copy_(buf.reshape(mom[n].shape))
    return ck.get('meta', {})


def train(engine, batches, epochs=1, iters_per_epoch=None, base_lr=0.01, momentum=0.9,
          weight_decay=0.0005, start_epoch=0, start_iter=0, log_every=50, log=print):
    """``batches``: iterable of (img (B,3,H,W) CUDA fp32, gt (sumG,19) CUDA, offsets (B+1) CUDA
    int32) — e.g. ``synthetic`` data or a GPU input pipeline.  No host sync inside the loop except
    every ``log_every`` iterations."""
    it = start_iter
    for epoch in range(start_epoch, start_epoch + epochs):
        for k, (img, gt, offs) in enumerate(batches):
            if iters_per_epoch is not None and k >= iters_per_epoch:
                break
            lr = lr_at(it, epoch, base_lr)
            losses = engine.train_step(img, gt, offs, lr=lr, momentum=momentum,
                                       weight_decay=weight_decay)
            if log_every and it % log_every == 0:
                l = losses.cpu().numpy()
                log(f'epoch {epoch} iter {it} lr {lr:.3e} loss_cls {l[0]:.4f} loss_bbox {l[1]:.4f} '
                    f'loss_obj {l[2]:.4f} loss_kps {l[3]:.4f} loss {float(np.sum(l)):.4f}')
            it += 1
    return it
