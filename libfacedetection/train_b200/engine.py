"""Host-side engine: owns the flat parameter / gradient / momentum buckets and the workspace as
torch CUDA tensors and drives ``libyunet_b200.so`` through the C ABI on torch's current stream.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); every FLOP of the hot
path runs in the hand-written sm_100a kernels.  There is no CPU fallback: constructing an engine
without a CUDA device raises.

Reference call sites this replaces (paths in ShiqiYu/libfacedetection.train):
``mmdet/models/detectors/single_stage.py:52-57`` (extract_feat), ``detectors/yunet.py:21-86``,
``dense_heads/yunet_head.py:249-288,290-374,418-534``, ``detectors/base.py:184-252`` (train_step /
_parse_losses), the mmcv ``OptimizerHook`` backward+step and the DDP gradient all-reduce
(``mmdet/apis/train.py:156-161``).
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import lib, check

ARCHS = {
    # configs/yunet_n.py:104-121 / configs/yunet_s.py:104-121
    'yunet_n': dict(stage_channels=[[3, 16, 16], [16, 64], [64, 64], [64, 64], [64, 64], [64, 64]],
                    downsample_idx=[0, 2, 3, 4], out_idx=[3, 4, 5], shared_stacked_convs=1),
    'yunet_s': dict(stage_channels=[[3, 16, 16], [16, 32], [32, 64], [64, 64], [64, 64], [64, 64]],
                    downsample_idx=[0, 2, 3, 4], out_idx=[3, 4, 5], shared_stacked_convs=0),
}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class YuNetEngine:

    def __init__(self, arch='yunet_n', device=None, loss_cfg=None, **arch_kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError('YuNetEngine needs a CUDA device (sm_100a); there is no CPU fallback')
        a = dict(ARCHS[arch]) if isinstance(arch, str) else dict(arch)
        a.update(arch_kwargs)
        self.arch = a
        self.device = torch.device(device if device is not None else
                                   f'cuda:{torch.cuda.current_device()}')
        self.ctx = _capi.Ctx(_capi.make_arch_cfg(a['stage_channels'], a['downsample_idx'],
                                                 a['out_idx'], a['shared_stacked_convs'],
                                                 a.get('feat_channels', 64)))
        self.h = self.ctx.handle
        self.loss_cfg = loss_cfg if loss_cfg is not None else _capi.default_loss_cfg()
        n, nbn = self.ctx.num_params, self.ctx.num_bn_channels
        self.params = torch.zeros(n, device=self.device)
        self.grads = torch.zeros(n, device=self.device)
        self.momentum_buf = torch.zeros(n, device=self.device)
        self.bn_running = torch.cat([torch.zeros(nbn), torch.ones(nbn)]).to(self.device)
        self.num_batches_tracked = 0
        self.param_table = self.ctx.params()
        self.bn_table = self.ctx.bns()
        self._ws = {}
        self._bufs = {}
        self.momentum = 0.1   # BatchNorm momentum (torch default)
        self.launches_per_train_step = None

    def set_option(self, name, value):
        """Library options: ``tc_forward`` (tcgen05/TMEM/TMA unit kernel for the 64-channel
        plain-load units)."""
        check(self.h, lib.yunet_set_option(self.h, name.encode(), int(value)), 'yunet_set_option')

    def status_flags(self, B, H, W, train):
        """Device-side error flags of the tensor-core kernels (all zero = ok)."""
        off = lib.yunet_ws_offset(self.h, B, H, W, 1 if train else 0, 0, 3)
        ws = self.workspace(B, H, W, train)
        return ws[off:off + 256].view(torch.int32).cpu()

    # ------------------------------------------------------------------ export
    def export_cpp(self):
        """``facedetectcnn-data.cpp`` text, byte-identical to the reference's ``tools/yunet2cpp.py``."""
        from . import export
        return export.cpp_data(self.state_dict(), self.arch)

    def export_onnx(self, height=640, width=640, dynamic=False):
        """Serialized 12-output ONNX model (``tools/yunet2onnx.py`` graph, BatchNorm folded);
        ``dynamic=True``: the tool's ``--dynamic-export`` (batch / height / width axes)."""
        from . import export
        return export.onnx_model(self.state_dict(), self.arch, height, width, dynamic=dynamic)

    # ------------------------------------------------------------------ parameters
    def param_views(self, bucket=None):
        bucket = self.params if bucket is None else bucket
        return {name: bucket[off:off + int(np.prod(shape))].view(shape)
                for name, off, shape in self.param_table}

    def load_state_dict(self, sd, strict=True):
        """Reference-format state_dict (keys/shapes of weights/yunet_{n,s}.pth)."""
        views = self.param_views()
        missing = [k for k in views if k not in sd]
        unexpected = [k for k in sd if k not in views and 'running_' not in k and
                      'num_batches_tracked' not in k]
        nbn = self.ctx.num_bn_channels
        for name, off, ch in self.bn_table:
            for suffix, base in (('.running_mean', 0), ('.running_var', nbn)):
                key = name + suffix
                if key in sd:
                    self.bn_running[base + off:base + off + ch].copy_(torch.as_tensor(sd[key]))
                else:
                    missing.append(key)
        if strict and (missing or unexpected):
            raise KeyError(f'state_dict mismatch: missing {missing}, unexpected {unexpected}')
        for k, v in views.items():
            if k in sd:
                v.copy_(torch.as_tensor(sd[k]).reshape(v.shape))
        nbt = [int(v) for k, v in sd.items() if k.endswith('num_batches_tracked')]
        if nbt:
            self.num_batches_tracked = nbt[0]
        return missing, unexpected

    def state_dict(self):
        sd = {k: v.detach().clone() for k, v in self.param_views().items()}
        nbn = self.ctx.num_bn_channels
        for name, off, ch in self.bn_table:
            sd[name + '.running_mean'] = self.bn_running[off:off + ch].clone()
            sd[name + '.running_var'] = self.bn_running[nbn + off:nbn + off + ch].clone()
            sd[name + '.num_batches_tracked'] = torch.tensor(self.num_batches_tracked)
        return sd

    def init_weights(self, seed=0):
        """Xavier-normal conv weights, conv bias 0.02, BN gamma 1 / beta 0
        (yunet_backbone.py:21-31, tfpn.py:21-31, yunet_head.py:158-168)."""
        g = torch.Generator().manual_seed(seed)
        flat = torch.zeros(self.ctx.num_params)
        for name, off, shape in self.param_table:
            n = int(np.prod(shape))
            parent = name.rsplit('.', 1)[0]
            if parent.endswith('.bn') or parent.endswith('.bn1'):
                flat[off:off + n] = 1.0 if name.endswith('.weight') else 0.0
            elif len(shape) == 4:
                fan_in = shape[1] * shape[2] * shape[3]
                fan_out = shape[0] * shape[2] * shape[3]
                std = (2.0 / float(fan_in + fan_out)) ** 0.5
                flat[off:off + n] = torch.empty(n).normal_(0, std, generator=g)
            else:
                flat[off:off + n] = 0.02
        self.params.copy_(flat)
        nbn = self.ctx.num_bn_channels
        self.bn_running[:nbn] = 0
        self.bn_running[nbn:] = 1
        self.momentum_buf.zero_()

    # ------------------------------------------------------------------ buffers
    def workspace(self, B, H, W, train):
        key = (B, H, W, bool(train))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = self.ctx.workspace_bytes(B, H, W, train)
            if nbytes == 0:
                raise ValueError(f'bad input shape {(B, H, W)}: H and W must be multiples of 32')
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    # ------------------------------------------------------------------ entry points
    def forward(self, img, train=False, preds=None):
        """img (B,3,H,W) fp32 CUDA NCHW 0..255 -> preds (B,P,16)."""
        assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and img.shape[1] == 3
        img = img.contiguous()
        B, _, H, W = img.shape
        ws = self.workspace(B, H, W, train)
        P = self.ctx.num_priors(H, W)
        if preds is None:
            preds = torch.empty(B, P, 16, device=self.device)
        check(self.h, lib.yunet_forward(self.h, _ptr(img), _ptr(self.params), _ptr(self.bn_running),
                                        B, H, W, 1 if train else 0, self.momentum, _ptr(preds),
                                        _ptr(ws), ws.numel(), _stream()), 'yunet_forward')
        if train:
            self.num_batches_tracked += 1
        return preds

    def read_activation(self, unit_index, B, H, W, train=False):
        """NCHW post-BN/ReLU output of a unit (what the reference module returns)."""
        d = self.ctx.units(include_stem=True)[unit_index + 1]
        ws = self.workspace(B, H, W, train)
        out = torch.empty(B, d.cout, H // d.div, W // d.div, device=self.device)
        check(self.h, lib.yunet_read_activation(self.h, unit_index, _ptr(self.params),
                                                _ptr(self.bn_running), B, H, W,
                                                1 if train else 0, _ptr(ws), _ptr(out), _stream()),
              'yunet_read_activation')
        return out

    def grid_priors(self, H, W):
        pri = torch.empty(self.ctx.num_priors(H, W), 4, device=self.device)
        check(self.h, lib.yunet_grid_priors(self.h, H, W, _ptr(pri), _stream()), 'yunet_grid_priors')
        return pri

    def assign(self, preds, gt, gt_offsets, H, W):
        """SimOTA for the whole batch.  gt (sumG,19) fp32 CUDA, gt_offsets (B+1) int32 CUDA."""
        B, P, _ = preds.shape
        assigned = self._buf('assigned', (B, P), torch.int32)
        miou = self._buf('miou', (B, P), torch.float32)
        counters = self._buf('counters', (4,), torch.float32)
        nws = lib.yunet_assign_workspace_bytes(self.h, B, H, W)
        aws = self._buf('assign_ws', (max(nws, 16),), torch.uint8)
        check(self.h, lib.yunet_simota_assign(self.h, C.byref(self.loss_cfg), _ptr(preds), _ptr(gt),
                                              _ptr(gt_offsets), B, H, W, _ptr(assigned), _ptr(miou),
                                              _ptr(counters), _ptr(aws), aws.numel(), _stream()),
              'yunet_simota_assign')
        return assigned, miou, counters

    def assign_ext(self, scores, priors, decoded_boxes, gt_boxes):
        """``SimOTAAssigner.assign`` for one image with explicit inputs (sim_ota_assigner.py:38-93):
        scores (P,) = sigmoid(cls)*sigmoid(obj), priors (P,4) offset ``[cx, cy, stride, stride]``,
        decoded boxes (P,4), gt boxes (G,4); all fp32 CUDA.  Returns (assigned (P,) int32 with the
        1-based gt index or 0, matched IoU (P,), counters)."""
        P_ = int(scores.shape[0])
        G = int(gt_boxes.shape[0])
        gt = torch.zeros(max(G, 1), 19, device=self.device)
        if G:
            gt[:, :4] = gt_boxes
        offs = torch.tensor([0, G], dtype=torch.int32, device=self.device)
        assigned = torch.zeros(P_, dtype=torch.int32, device=self.device)
        miou = torch.zeros(P_, device=self.device)
        counters = torch.zeros(4, device=self.device)
        aws = torch.empty(max(P_ * 32 if P_ > 2112 else 16, 16), dtype=torch.uint8, device=self.device)
        check(self.h, lib.yunet_simota_assign_ext(
            self.h, C.byref(self.loss_cfg), P_, _ptr(scores.contiguous().float()),
            _ptr(priors.contiguous().float()), _ptr(decoded_boxes.contiguous().float()), _ptr(gt),
            _ptr(offs), _ptr(assigned), _ptr(miou), _ptr(counters), _ptr(aws), aws.numel(), _stream()),
            'yunet_simota_assign_ext')
        return assigned, miou, counters

    def loss_grad(self, preds, gt, gt_offsets, assigned, miou, counters, num_total, H, W,
                  loss_scale=None, want_grad=True):
        B, P, _ = preds.shape
        losses = self._buf('losses', (4,), torch.float32)
        d_preds = self._buf('d_preds', (B, P, 16), torch.float32) if want_grad else None
        scale = (C.c_float * 4)(*(loss_scale if loss_scale is not None else (1., 1., 1., 1.)))
        check(self.h, lib.yunet_loss_grad(self.h, C.byref(self.loss_cfg), _ptr(preds), _ptr(gt),
                                          _ptr(gt_offsets), _ptr(assigned), _ptr(miou),
                                          _ptr(counters), _ptr(num_total), scale, B, H, W,
                                          _ptr(losses), _ptr(d_preds), _stream()),
              'yunet_loss_grad')
        return losses, d_preds

    def backward(self, img, d_preds):
        B, _, H, W = img.shape
        ws = self.workspace(B, H, W, True)
        check(self.h, lib.yunet_backward(self.h, _ptr(img), _ptr(self.params), _ptr(d_preds), B, H,
                                         W, _ptr(self.grads), _ptr(ws), ws.numel(), _stream()),
              'yunet_backward')
        return self.grads

    def sgd_step(self, lr=0.01, momentum=0.9, weight_decay=0.0005, grad_scale=1.0):
        check(self.h, lib.yunet_sgd_step(self.h, _ptr(self.params), _ptr(self.grads),
                                         _ptr(self.momentum_buf), self.ctx.num_params, lr, momentum,
                                         weight_decay, grad_scale, _stream()), 'yunet_sgd_step')

    def train_step_graph(self, img, gt, gt_offsets, lr=0.01, momentum=0.9, weight_decay=0.0005):
        """``train_step`` replayed as ONE CUDA graph: the whole iteration — forward,
        SimOTA, [num_pos all-reduce], loss + gradient, backward, [bucket all-reduce], SGD — is
        captured once per set of input buffers and relaunched with a single ``cudaGraphLaunch``;
        the learning rate is a device scalar, so the schedule keeps working.  The first call with
        new buffers runs eagerly (it also sizes every workspace), the second captures, later ones
        replay.  Inputs must stay at the same addresses (a real input pipeline writes into fixed
        device slots, like ``bench.py``'s two double-buffered ones)."""
        from . import dist_utils
        if dist_utils.world_size() > 1:
            # the two NCCL all-reduces of a data-parallel step are not captured (every rank would have to
            # capture and replay in lock step): launch the kernels individually
            return self._train_step_impl(img, gt, gt_offsets, lr, momentum, weight_decay, True)
        key = (img.data_ptr(), gt.data_ptr(), gt_offsets.data_ptr(), tuple(img.shape), tuple(gt.shape),
               float(momentum), float(weight_decay))
        if not hasattr(self, '_graphs'):
            self._graphs, self._lr_dev = {}, torch.zeros(1, device=self.device)
        self._lr_dev.fill_(float(lr))
        st = self._graphs.get(key)
        if st is None:                      # first sight of these buffers: eager step
            self._graphs[key] = 'seen'
            return self._train_step_impl(img, gt, gt_offsets, None, momentum, weight_decay, True, lr_dev=self._lr_dev)
        if st == 'seen':                    # second: capture (nothing executes), then fall through
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                out = self._train_step_impl(img, gt, gt_offsets, None, momentum, weight_decay, True,
                                            lr_dev=self._lr_dev)
            st = self._graphs[key] = (g, out)
        st[0].replay()
        return st[1]

    def train_step(self, img, gt, gt_offsets, lr=0.01, momentum=0.9, weight_decay=0.0005,
                   step=True):
        """One full training iteration on this rank: forward (train-mode BN) -> SimOTA ->
        [all-reduce num_pos] -> losses + d_preds -> backward -> [all-reduce gradient bucket] ->
        fused SGD.  Returns the device tensor of the four losses [cls, bbox, obj, kps]."""
        return self._train_step_impl(img, gt, gt_offsets, lr, momentum, weight_decay, step)

    def _train_step_impl(self, img, gt, gt_offsets, lr, momentum, weight_decay, step, lr_dev=None):
        from . import dist_utils
        B, _, H, W = img.shape
        preds = self.forward(img, train=True, preds=self._buf('preds', (B, self.ctx.num_priors(H, W), 16), torch.float32))
        assigned, miou, counters = self.assign(preds, gt, gt_offsets, H, W)
        num_total = counters
        if dist_utils.world_size() > 1:
            # reduce_mean(num_pos) sits between assignment and loss scaling (yunet_head.py:493-497)
            num_total = self._buf('num_total', (1,), torch.float32)
            num_total.copy_(counters[:1])
            dist_utils.reduce_mean_(num_total)
        losses, d_preds = self.loss_grad(preds, gt, gt_offsets, assigned, miou, counters, num_total,
                                         H, W)
        self.backward(img, d_preds)
        grad_scale = dist_utils.allreduce_bucket_(self.grads)   # ONE all-reduce, 303 KB (yunet_n)
        if step and lr_dev is not None:
            check(self.h, lib.yunet_sgd_step_dev(self.h, _ptr(self.params), _ptr(self.grads),
                                                 _ptr(self.momentum_buf), self.ctx.num_params, _ptr(lr_dev),
                                                 momentum, weight_decay, grad_scale, _stream()),
                  'yunet_sgd_step_dev')
        elif step:
            self.sgd_step(lr, momentum, weight_decay, grad_scale)
        return losses

    def detect(self, img, score_thr=0.02, iou_thr=0.45, scale_factors=None, max_det=None,
               with_kps=False, preds=None):
        """Inference: forward (running-stat BN) + decode + NMS.  Returns (dets (B,max_det,5),
        counts (B,), kps (B,max_det,10) or None) — device tensors, no host sync."""
        B, _, H, W = img.shape
        if preds is None:
            preds = self.forward(img, train=False)
        return self.decode_nms(preds, H, W, score_thr, iou_thr, scale_factors, max_det, with_kps)

    def decode_nms(self, preds, H, W, score_thr=0.02, iou_thr=0.45, scale_factors=None,
                   max_det=None, with_kps=False):
        B, P, _ = preds.shape
        max_det = P if max_det is None else max_det
        dets = self._buf('dets', (B, max_det, 5), torch.float32)
        kps = self._buf('det_kps', (B, max_det, 10), torch.float32) if with_kps else None
        counts = self._buf('det_count', (B,), torch.int32)
        nws = lib.yunet_nms_workspace_bytes(self.h, B, H, W)
        ws = self._buf('nms_ws', (max(nws, 16),), torch.uint8)
        check(self.h, lib.yunet_decode_nms(self.h, _ptr(preds), B, H, W, score_thr, iou_thr,
                                           _ptr(scale_factors), max_det, _ptr(dets), _ptr(kps),
                                           _ptr(counts), _ptr(ws), ws.numel(), _stream()),
              'yunet_decode_nms')
        return dets, counts, kps
