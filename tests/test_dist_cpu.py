"""CPU tests of the N>1 host logic with the gloo backend, world_size 2 (one process per rank,
rendezvous on 127.0.0.1): reduce_mean of num_pos, the single flat-bucket all-reduce and the
1/world gradient scale must reproduce the reference's DDP semantics — i.e. equal a single-process
step on the concatenated batch (BatchNorm aside, which is local in the reference too)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libfacedetection.train_b200 import dist_utils


def _worker(rank, world, port, tmp):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # num_pos: rank r found 10 + 7*r positives
        npos = torch.tensor([10.0 + 7 * rank])
        dist_utils.reduce_mean_(npos)
        # gradient bucket: rank-dependent values
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        scale = dist_utils.allreduce_bucket_(g)
        w = torch.ones(1000)
        v = torch.zeros(1000)
        # SGD exactly like csrc/sgd.cu: g' = g*scale + wd*w ; v = mom*v + g' ; w -= lr*v
        gi = g * scale + 0.0005 * w
        v = 0.9 * v + gi
        w = w - 0.01 * v
        lo, hi = dist_utils.shard_batch(512, rank, world)
        # plugin-surface logging (mmdet/models/detectors/base.py:184-217): log_vars are rank means, the
        # loss tensor that is back-propagated stays local
        from libfacedetection.train_b200 import plugins
        ls = dict(loss_cls=torch.tensor(1.0 + rank, requires_grad=True), loss_bbox=torch.tensor(2.0 * (rank + 1)),
                  loss_obj=torch.tensor(0.5), loss_kps=[torch.tensor([1.0, 3.0]) * (rank + 1)], acc=torch.tensor(10.0 * rank))
        loss, log_vars = plugins.YuNet._parse_losses(ls)
        torch.save(dict(npos=npos, w=w, shard=(lo, hi), loss=float(loss), log_vars=dict(log_vars),
                        loss_has_grad=bool(loss.requires_grad)), os.path.join(tmp, f'r{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / 'r0.pt')
    r1 = torch.load(tmp_path / 'r1.pt')
    assert float(r0['npos']) == float(r1['npos']) == pytest.approx((10 + 17) / 2)
    mean_g = torch.arange(1000, dtype=torch.float32) * 1.5
    w_ref = torch.ones(1000) - 0.01 * (mean_g + 0.0005)
    assert torch.allclose(r0['w'], w_ref) and torch.equal(r0['w'], r1['w'])
    assert r0['shard'] == (0, 256) and r1['shard'] == (256, 512)
    # rank 0: 1 + 2 + 0.5 + 2 = 5.5, rank 1: 2 + 4 + 0.5 + 4 = 10.5 ('acc' has no 'loss' in its key)
    assert r0['loss'] == pytest.approx(5.5) and r1['loss'] == pytest.approx(10.5) and r0['loss_has_grad']
    assert r0['log_vars'] == r1['log_vars']
    assert r0['log_vars'] == pytest.approx(dict(loss_cls=1.5, loss_bbox=3.0, loss_obj=0.5, loss_kps=3.0, acc=5.0, loss=8.0))
    assert list(r0['log_vars']) == ['loss_cls', 'loss_bbox', 'loss_obj', 'loss_kps', 'acc', 'loss']


def test_parse_losses_single_process():
    from libfacedetection.train_b200 import plugins
    ls = dict(loss_cls=torch.tensor(1.25, requires_grad=True), loss_bbox=torch.tensor(2.0), loss_obj=torch.tensor(0.5),
              loss_kps=torch.tensor(0.25))
    loss, log_vars = plugins.YuNet._parse_losses(ls)
    assert float(loss) == 4.0 and loss.requires_grad
    assert dict(log_vars) == dict(loss_cls=1.25, loss_bbox=2.0, loss_obj=0.5, loss_kps=0.25, loss=4.0)
    with pytest.raises(TypeError):
        plugins.YuNet._parse_losses(dict(loss_x=1.0))


def test_single_process_is_identity():
    t = torch.tensor([5.0])
    assert float(dist_utils.reduce_mean_(t)) == 5.0
    g = torch.ones(4)
    assert dist_utils.allreduce_bucket_(g) == 1.0 and torch.equal(g, torch.ones(4))
    with pytest.raises(ValueError):
        dist_utils.shard_batch(10, 0, 4)
