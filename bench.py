#!/usr/bin/env python
"""Benchmark of the north-star path: images/sec of one full YuNet training step
(forward + SimOTA + loss + backward + [grad all-reduce] + SGD) at 320x320, batch 256 per GPU, on
synthetic WIDER-shaped batches (BASELINE.json configs[1]; configs[2] under torchrun).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM),
`e2e` = the same step driven from pinned HOST buffers (H2D of images+GT and D2H of the losses
inside the timed region, double-buffered on a copy stream), `roofline` = achieved algorithmic
GB/s of the dominant kernel (per-launch CUDA events on the launching stream, from the library's
profiling hooks) against the measured HBM peak, `cpu_baseline` = the CPU oracle (reference
restatement, all host cores) on a bounded sample.  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'train_images_per_sec_320'
UNIT = 'images/s'
# first-iteration learning rate of the reference schedule: lr 0.01 x warmup_ratio 0.001
# (configs/yunet_n.py:1-11); the un-normalised 0..255 inputs diverge at random init with the
# post-warm-up rate, exactly as they would in the reference without its warm-up
LR = 0.01 * 0.001


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--arch', default='yunet_n')
    ap.add_argument('--batch', type=int, default=256, help='images per GPU')
    ap.add_argument('--size', type=int, default=320)
    ap.add_argument('--cpu-sample', type=int, default=32, help='images per CPU-baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--kernel-table', default='', help='write the per-kernel profile here (json)')
    ap.add_argument('--no-graph', action='store_true',
                    help='launch every kernel of the step individually instead of replaying the captured CUDA graph')
    ap.add_argument('--no-extra', action='store_true',
                    help='skip the extra_configs (yunet_s train, 640x640 inference) and e2e_plugin legs')
    ap.add_argument('--workload', default='train', choices=['train', 'infer'],
                    help="'train' (default, the headline metric) or 'infer': BASELINE.json "
                         "configs[3], eval forward + decode + NMS at 640x640, bs=512")
    return ap.parse_args()


def load_synthetic():
    """``libfacedetection/train_b200/synthetic.py`` loaded from its file, NOT through the package:
    importing the package dlopens libyunet_b200.so, which the CPU reference arm must not do."""
    import importlib.util
    path = os.path.join(ROOT, 'libfacedetection', 'train_b200', 'synthetic.py')
    spec = importlib.util.spec_from_file_location('_yunet_synthetic_standalone', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------- CPU oracle leg
def cpu_reference_steps(arch, sample, size, steps, warmup, seed=0):
    """The reference's own training step (CPU restatement of the unmodified Python path, all host
    threads): zero_grad -> forward_train -> _parse_losses -> backward -> SGD.step."""
    from oracle import yunet_oracle as orc
    synthetic = load_synthetic()
    cores = os.cpu_count() or 1
    P, Bf = orc.init_params(arch, seed=0)
    img = torch.from_numpy(synthetic.make_images(sample, size, seed))
    gb, gl, gk = synthetic.make_gt(sample, size, seed)
    gb = [torch.from_numpy(x) for x in gb]
    gl = [torch.from_numpy(x) for x in gl]
    gk = [torch.from_numpy(x) for x in gk]
    mom = {}

    split = dict(forward=0.0, assign_loss=0.0, backward=0.0, sgd=0.0)

    def one_step():
        # the body of oracle.train_forward_backward + sgd_step with a clock between the phases
        t0 = time.perf_counter()
        Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
        outs = orc.model_forward(img, Pg, Bf, arch, training=True)
        t1 = time.perf_counter()
        losses, _ = orc.head_loss(*outs, gb, gl, gk, strides=orc.ARCH[arch]['strides'], return_assign=True)
        total = sum(losses.values())
        t2 = time.perf_counter()
        total.backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
        t3 = time.perf_counter()
        orc.sgd_step(P, grads, mom, lr=LR)
        t4 = time.perf_counter()
        for k, dt in zip(split, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            split[k] += dt
        return t4 - t0

    # the per-image SimOTA loop is thousands of tiny ops: more threads than ~32 only add
    # synchronisation cost, so probe a few thread counts (all host cores first) and keep the best
    best_t, best_n = None, cores
    for n in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
        torch.set_num_threads(n)
        one_step() if best_t is None and warmup > 0 else None
        dt = one_step()
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
    torch.set_num_threads(best_n)
    for k in split:
        split[k] = 0.0
    times = [one_step() for _ in range(steps)]
    ms = 1e3 * float(np.mean(times))
    return dict(value=sample / (ms / 1e3), ms_per_step=ms, cores=cores, threads=best_n,
                sample=sample, split_ms={k: round(1e3 * v / max(steps, 1), 2) for k, v in split.items()})


def cpu_reference_infer(arch, sample, size, steps, warmup):
    """the reference's ``simple_test`` (eval forward + get_bboxes: decode,
    score filter, NMS) on the host cores, ``sample`` images of ``size`` x ``size`` per step."""
    from oracle import yunet_oracle as orc
    d = np.load(os.path.join(ROOT, 'tests', 'golden', f'weights_{arch}.npz'))
    P, Bf = orc.split_state_dict({k: torch.from_numpy(d[k]) for k in d.files})
    g = torch.Generator().manual_seed(0)
    img = torch.rand(sample, 3, size, size, generator=g) * 255
    cores = os.cpu_count() or 1
    torch.set_num_threads(min(cores, 32))

    def one():
        t0 = time.perf_counter()
        with torch.no_grad():
            outs = orc.model_forward(img, P, Bf, arch, training=False)
            orc.get_bboxes(*outs)
        return time.perf_counter() - t0

    for _ in range(warmup):
        one()
    ms = 1e3 * float(np.mean([one() for _ in range(steps)]))
    return dict(value=sample / (ms / 1e3), ms_per_step=ms, cores=cores, threads=min(cores, 32), sample=sample)


def step_table(eng, step_fn, reps=3):
    """Per-launch CUDA-event profile of ``reps`` calls of ``step_fn`` -> (rows, total ms, bytes)."""
    import ctypes as C
    from libfacedetection.train_b200._capi import lib
    lib.yunet_profile_begin(eng.h)
    for _ in range(reps):
        step_fn()
    torch.cuda.synchronize()
    n = lib.yunet_profile_end(eng.h)
    name, ms, by = C.create_string_buffer(160), C.c_float(), C.c_double()
    kern = {}
    for i in range(n):
        lib.yunet_profile_get(eng.h, i, name, 160, C.byref(ms), C.byref(by))
        k = kern.setdefault(name.value.decode(), [0.0, 0.0, 0])
        k[0] += ms.value
        k[1] = by.value
        k[2] += 1
    rows = [{'kernel': k, 'ms': v[0] / v[2], 'algorithmic_bytes': v[1],
             'gbs': (v[1] / 1e9) / (v[0] / v[2] / 1e3) if v[1] > 0 and v[0] > 0 else None}
            for k, v in kern.items()]
    rows.sort(key=lambda r: -r['ms'])
    return rows


def extra_train_config(arch, B, S, dev, peak):
    """BASELINE.json configs[4] (yunet_s train) measured in the same process: device-resident step."""
    from libfacedetection.train_b200 import YuNetEngine, synthetic
    eng = YuNetEngine(arch, device=dev)
    eng.init_weights(0)
    img = torch.from_numpy(synthetic.make_images(B, S, 0)).to(dev)
    gb, gl, gk = synthetic.make_gt(B, S, 0)
    gt, offs = synthetic.pack_gt_csr(gb, gk)
    gt, offs = torch.from_numpy(gt).to(dev), torch.from_numpy(offs).to(dev)
    for _ in range(3):
        eng.train_step(img, gt, offs, lr=LR)
    torch.cuda.synchronize()
    K = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        eng.train_step(img, gt, offs, lr=LR)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    rows = step_table(eng, lambda: eng.train_step(img, gt, offs, lr=LR))
    streaming = [r for r in rows if r['gbs'] is not None]
    fwd = [r for r in streaming if r['kernel'].startswith('fwd')]
    dom = streaming[0]
    gb_all = sum(r['algorithmic_bytes'] for r in streaming) / 1e9
    out = {'config': f'{arch} {S}x{S} train step bs={B} (BASELINE.json configs[4])',
           'metric': METRIC, 'value': B / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'steps': K,
           'roofline': {'kernel': dom['kernel'], 'achieved': dom['gbs'], 'frac': dom['gbs'] / peak,
                        'ms_per_launch': dom['ms'],
                        'forward_total_frac': (sum(r['algorithmic_bytes'] for r in fwd) / 1e9) /
                        (sum(r['ms'] for r in fwd) / 1e3) / peak,
                        'step_total_frac': gb_all / (sum(r['ms'] for r in streaming) / 1e3) / peak}}
    del eng
    torch.cuda.empty_cache()
    return out


def extra_infer_config(arch, B, S, dev, peak, cpu=True):
    """BASELINE.json configs[3]: eval forward + decode + NMS at 640x640, bs=512, with its own CPU
    baseline (oracle ``simple_test`` at bs=8) and an end-to-end number from pinned host images."""
    from libfacedetection.train_b200 import YuNetEngine
    eng = YuNetEngine(arch, device=dev)
    d = np.load(os.path.join(ROOT, 'tests', 'golden', f'weights_{arch}.npz'))
    eng.load_state_dict({k: torch.from_numpy(d[k]) for k in d.files})
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand(B, 3, S, S, device=dev, generator=g) * 255
    for _ in range(3):
        eng.detect(img)
    torch.cuda.synchronize()
    K = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        dets, counts, _ = eng.detect(img)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    rows = step_table(eng, lambda: eng.detect(img), reps=2)
    streaming = [r for r in rows if r['gbs'] is not None]
    nms_ms = sum(r['ms'] for r in rows if 'nms' in r['kernel'])
    # end to end: images from pinned host memory, detections + counts back to the host
    himg = torch.empty(B, 3, S, S, dtype=torch.float32).pin_memory()
    himg.copy_(img)
    P_ = dets.shape[1]
    hdets = torch.empty(B, P_, 5).pin_memory()
    hcnt = torch.empty(B, dtype=torch.int32).pin_memory()

    def e2e_step():
        img.copy_(himg, non_blocking=True)
        dd, cc, _ = eng.detect(img)
        hdets.copy_(dd, non_blocking=True)
        hcnt.copy_(cc, non_blocking=True)

    e2e_step()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(3):
        e2e_step()
    s1.record()
    torch.cuda.synchronize()
    ms_e2e = s0.elapsed_time(s1) / 3
    alg = sum(r['algorithmic_bytes'] for r in streaming)
    out = {'config': f'{arch} {S}x{S} eval forward + decode + NMS, bs={B} (BASELINE.json configs[3])',
           'metric': 'infer_images_per_sec_640', 'value': B / (ms / 1e3), 'unit': UNIT,
           'ms_per_step': ms, 'steps': K, 'nms_ms': nms_ms,
           'detections_per_image': float(counts.float().mean()),
           'roofline': {'bound': 'hbm', 'achieved': alg / 1e9 / (sum(r['ms'] for r in streaming) / 1e3),
                        'frac': alg / 1e9 / (sum(r['ms'] for r in streaming) / 1e3) / peak,
                        'note': 'forward kernels: algorithmic bytes of all launches / their summed time'},
           'e2e': {'value': B / (ms_e2e / 1e3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                   'h2d_bytes_per_step': himg.numel() * 4,
                   'd2h_bytes_per_step': hdets.numel() * 4 + hcnt.numel() * 4}}
    del eng, img, himg, hdets
    torch.cuda.empty_cache()
    if cpu:
        r = cpu_reference_infer(arch, 8, S, steps=3, warmup=1)
        out['cpu_baseline'] = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                               'sample': f'8 images/step x 3 steps of the oracle simple_test '
                                         f'(eval forward + decode + NMS) at {S}x{S}, torch CPU, '
                                         f'{r["threads"]} threads', 'ms_per_step': r['ms_per_step']}
    return out


def plugin_e2e(arch, B, S, dev, steps, warmup, optimizer='registry', breakdown=False):
    """The same training step through the mmdet plugin surface (``plugins.YuNet.train_step`` +
    ``torch.optim.SGD``), host buffers, H2D of images / GT lists and D2H of the losses per step."""
    from libfacedetection.train_b200 import plugins, synthetic
    cfg = dict(
        type='YuNet',
        backbone=dict(type='YuNetBackbone', stage_channels=plugins_arch(arch)['stage_channels'],
                      downsample_idx=[0, 2, 3, 4], out_idx=[3, 4, 5]),
        neck=dict(type='TFPN', in_channels=[64, 64, 64], out_idx=[0, 1, 2]),
        bbox_head=dict(type='YuNet_Head', num_classes=1, in_channels=64,
                       shared_stacked_convs=plugins_arch(arch)['shared_stacked_convs'], stacked_convs=0,
                       feat_channels=64,
                       prior_generator=dict(type='MlvlPointGenerator', offset=0, strides=[8, 16, 32]),
                       loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, reduction='sum', loss_weight=1.0),
                       loss_bbox=dict(type='EIoULoss', loss_weight=5.0, reduction='sum'),
                       use_kps=True, kps_num=5,
                       loss_kps=dict(type='SmoothL1Loss', beta=0.1111111111111111, loss_weight=0.1),
                       loss_obj=dict(type='CrossEntropyLoss', use_sigmoid=True, reduction='sum', loss_weight=1.0)),
        train_cfg=dict(assigner=dict(type='SimOTAAssigner', center_radius=2.5)),
        test_cfg=dict(nms_pre=-1, min_bbox_size=0, score_thr=0.02, nms=dict(type='nms', iou_threshold=0.45),
                      max_per_img=-1))
    torch.manual_seed(0)
    m = plugins.DETECTORS.build(cfg).to(dev).train()
    ocfg = dict(type='SGD', lr=LR, momentum=0.9, weight_decay=0.0005)      # configs/yunet_n.py:1
    if optimizer == 'registry':      # what mmcv's build_optimizer returns after register_into_mmdet()
        opt = plugins.OPTIMIZERS.build(ocfg, default_args=dict(params=m.parameters()))
    else:                            # the stock class on the same parameters
        opt = torch.optim.SGD(m.parameters(), **{k: v for k, v in ocfg.items() if k != 'type'})
    himg = torch.from_numpy(synthetic.make_images(B, S, 0)).pin_memory()
    gb, gl, gk = synthetic.make_gt(B, S, 0)
    hb = [torch.from_numpy(x).pin_memory() for x in gb]
    hl = [torch.from_numpy(x) for x in gl]
    hk = [torch.from_numpy(x).pin_memory() for x in gk]
    bb_all, kp_all = torch.cat(hb).pin_memory(), torch.cat(hk).pin_memory()
    lb_all = torch.cat(hl).pin_memory()
    counts = [int(x.shape[0]) for x in hb]
    # a prefetching loader: the next batch is uploaded on a copy stream while this one trains
    dimgs = [torch.empty_like(himg, device=dev) for _ in range(2)]
    copy_stream, main = torch.cuda.Stream(device=dev), torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    staged = [None, None]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            dimgs[slot].copy_(himg, non_blocking=True)
            flat = [t.to(dev, non_blocking=True) for t in (bb_all, kp_all, lb_all)]
            for t in flat:
                t.record_stream(main)          # allocated on the copy stream, consumed on the main one
            staged[slot] = tuple(t.split(counts) for t in flat)
            ready[slot].record(copy_stream)

    state = {'i': 0}
    for ev in freed:
        ev.record(main)
    upload(0)

    host = dict(prep=0.0, train_step=0.0, backward=0.0, opt_step=0.0, n=0)

    def step():
        ts = time.perf_counter()
        slot = state['i'] % 2
        state['i'] += 1
        upload(slot ^ 1)
        main.wait_event(ready[slot])
        db, dk, dl = staged[slot]
        data = dict(img=dimgs[slot], img_metas=[{}] * B, gt_bboxes=list(db), gt_labels=list(dl),
                    gt_keypointss=list(dk))
        t0 = time.perf_counter()
        opt.zero_grad()
        out = m.train_step(data)              # log_vars: host floats (D2H of the four losses)
        t1 = time.perf_counter()
        out['loss'].backward()
        t2 = time.perf_counter()
        opt.step()
        t3 = time.perf_counter()
        freed[slot].record(main)
        host['prep'] += t0 - ts; host['train_step'] += t1 - t0; host['backward'] += t2 - t1
        host['opt_step'] += t3 - t2; host['n'] += 1
        return out

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in host:
        host[k] = 0
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    h2d = himg.numel() * 4 + bb_all.numel() * 4 + kp_all.numel() * 4 + lb_all.numel() * 8
    del m, opt, dimgs, staged
    torch.cuda.empty_cache()
    # host-side wall time of the four phases of a step (train_step includes the wait for the losses)
    host_ms = {k: round(1e3 * v / max(host['n'], 1), 3) for k, v in host.items() if k != 'n'}
    return {'value': B / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
            'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 16, 'host_ms': host_ms,
            'optimizer': ('plugins.SGD (torch.optim.SGD subclass, what the OPTIMIZERS registry builds from '
                          "dict(type='SGD', ...))" if optimizer == 'registry' else 'torch.optim.SGD'),
            'api': 'plugins.YuNet.train_step + loss.backward() + optimizer.step() (mmdet plugin surface)'}


def plugins_arch(arch):
    from libfacedetection.train_b200.engine import ARCHS
    return ARCHS[arch]


def run_reference(args):
    # nothing of the product is imported here: oracle/ + the stand-alone synthetic generator only
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference_steps(args.arch, args.cpu_sample, args.size, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.arch} {args.size}x{args.size} train step '
                               f'(fwd+SimOTA+loss+bwd+SGD), bs={args.batch}/GPU; CPU sample of '
                               f'{args.cpu_sample} images per step', 'arch': args.arch,
                   'global_batch': args.batch * args.gpus, 'image_size': args.size},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                         'sample': f'{args.cpu_sample} images/step x {args.steps} steps, torch '
                                   f'{torch.__version__} CPU, best of thread counts up to '
                                   f'{r["cores"]} (used {r["threads"]})',
                         'split_ms': r['split_ms']},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40,
                 'sw_power_cap': 0x4, 'hw_power_brake': 0x80}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, 'nvmlDeviceGetCurrentClocksEventReasons') else \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def reset(self):
        self.samples, self.reasons = [], set()

    def snapshot(self, window):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples), 'window': window}

    def stop(self):
        self._stop_evt.set()


# ------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from libfacedetection.train_b200 import YuNetEngine, synthetic
    from libfacedetection.train_b200._capi import lib
    import ctypes as C

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                                timeout=datetime.timedelta(seconds=180))
    dev = torch.device('cuda', local)
    B, S, K, Wm = args.batch, args.size, args.steps, max(args.warmup, 3)

    eng = YuNetEngine(args.arch, device=dev)
    eng.init_weights(0)                      # reference init, identical on every rank
    P = eng.ctx.num_priors(S, S)

    # synthetic WIDER-shaped data, seed = rank; two pinned host slots
    img_np = synthetic.make_images(B, S, seed=rank)
    gb, gl, gk = synthetic.make_gt(B, S, seed=rank)
    gt_np, offs_np = synthetic.pack_gt_csr(gb, gk)
    host = []
    for _ in range(2):
        host.append((torch.from_numpy(img_np).clone().pin_memory(),
                     torch.from_numpy(gt_np).clone().pin_memory(),
                     torch.from_numpy(offs_np).clone().pin_memory()))
    devb = [(torch.empty_like(h[0], device=dev), torch.empty_like(h[1], device=dev),
             torch.empty_like(h[2], device=dev)) for h in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    loss_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h_bytes = 16

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---------------- device-resident arm: inputs already in HBM
    for d, h in zip(devb, host):
        for a, b_ in zip(d, h):
            a.copy_(b_)
    # eager launches first (they also count the kernels of one step), then the CUDA-graph replay of the
    # same step (engine.train_step_graph: one cudaGraphLaunch per iteration, lr as a device scalar)
    for i in range(Wm):
        eng.train_step(*devb[i % 2], lr=LR)
    sync_all()
    l0 = lib.yunet_launch_count(eng.h)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(max(4, K // 4)):
        eng.train_step(*devb[i % 2], lr=LR)
    g1.record()
    sync_all()
    ms_eager = max_over_ranks(g0.elapsed_time(g1)) / max(4, K // 4)
    launches_per_step = int(lib.yunet_launch_count(eng.h) - l0) // max(4, K // 4)
    step_fn, graph_info = eng.train_step, {'used': False, 'eager_ms_per_step': ms_eager,
                                           'kernel_launches_per_step': launches_per_step}
    if not args.no_graph and world > 1:
        # NCCL collectives inside a stream capture need every rank to capture and replay in lock step and
        # hung on this pool's 2-GPU box: data-parallel runs launch the step eagerly (measured 1.3 % slower
        # than the replay at N = 1)
        graph_info['skipped'] = 'world_size > 1: eager launches around the two NCCL all-reduces'
    if not args.no_graph and world == 1:
        try:
            for i in range(6):             # per input slot: eager, capture, replay
                eng.train_step_graph(*devb[i % 2], lr=LR)
            sync_all()
            step_fn = eng.train_step_graph
            graph_info['used'] = True
        except Exception as ex:            # e.g. a collective that cannot be captured: stay eager
            graph_info['error'] = repr(ex)[:200]
    sampler = ClockSampler(local)      # started before the warm-up: the first NVML queries are the slow ones
    sampler.start()
    for i in range(Wm):
        step_fn(*devb[i % 2], lr=LR)
    sync_all()
    sampler.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        losses = step_fn(*devb[i % 2], lr=LR)
    e1.record()
    sync_all()
    launches = launches_per_step * K       # kernels executed in the timed region (graph nodes or launches)
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / K
    graph_info['ms_per_step'] = ms_dev
    clocks = sampler.snapshot('device-resident timed region')
    loss_vals = losses.cpu().numpy().tolist()

    # ---------------- end-to-end arm: host buffers, H2D + D2H inside the timed region
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for a, b_ in zip(devb[slot], host[slot]):
                a.copy_(b_, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_loop(n):
        prefetch(0)
        for i in range(n):
            slot = i % 2
            if i + 1 < n:
                prefetch((i + 1) % 2)
            main.wait_event(copied[slot])
            ls = step_fn(*devb[slot], lr=LR)
            consumed[slot].record(main)
            loss_host[slot].copy_(ls, non_blocking=True)

    for ev in consumed:
        ev.record(main)
    e2e_loop(Wm)
    sync_all()
    for ev in consumed:
        ev.record(main)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(main)
    copy_stream.wait_event(s0)
    e2e_loop(K)
    s1.record(main)
    sync_all()
    ms_e2e = max_over_ranks(s0.elapsed_time(s1)) / K
    if clocks['samples'] < 3:
        # an NVML query under load can take longer than a short timed region: the sampler kept running
        # through the end-to-end timed region (same kernels, same load), so report both windows together
        clocks = sampler.snapshot('device-resident + end-to-end timed regions (incl. the e2e warm-up between them)')
    sampler.stop()

    # ---------------- per-kernel profile (separate pass, per-launch CUDA events on the stream)
    kern = {}
    reps = 3
    if rank == 0:
        lib.yunet_profile_begin(eng.h)
    for i in range(reps):          # every rank steps (the step contains collectives)
        eng.train_step(*devb[i % 2], lr=LR)
    torch.cuda.synchronize()
    if rank == 0:
        n = lib.yunet_profile_end(eng.h)
        name = C.create_string_buffer(160)
        ms = C.c_float()
        by = C.c_double()
        for i in range(n):
            lib.yunet_profile_get(eng.h, i, name, 160, C.byref(ms), C.byref(by))
            k = kern.setdefault(name.value.decode(), [0.0, 0.0, 0])
            k[0] += ms.value
            k[1] = by.value
            k[2] += 1
    if world > 1:
        dist.barrier()

    # ---------------- the two collectives of the step, timed alone (device time, max over ranks)
    comm = None
    if world > 1:
        from libfacedetection.train_b200 import dist_utils
        scal = torch.ones(1, device=dev)
        for _ in range(5):
            dist_utils.reduce_mean_(scal)
            dist_utils.allreduce_bucket_(eng.grads)
        sync_all()
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record()
        for _ in range(20):
            dist_utils.reduce_mean_(scal)
        c1.record()
        for _ in range(20):
            dist_utils.allreduce_bucket_(eng.grads)
        c2.record()
        sync_all()
        comm = {'num_pos_allreduce_ms': max_over_ranks(c0.elapsed_time(c1)) / 20,
                'grad_bucket_allreduce_ms': max_over_ranks(c1.elapsed_time(c2)) / 20,
                'bucket_bytes': int(eng.grads.numel() * 4),
                'note': 'back-to-back launches of each collective alone (latency bound)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    table = []
    for nme, (tot, byts, cnt) in kern.items():
        avg = tot / cnt
        table.append({'kernel': nme, 'ms': avg, 'algorithmic_bytes': byts,
                      'gbs': (byts / 1e9) / (avg / 1e3) if byts > 0 and avg > 0 else None})
    table.sort(key=lambda r: -r['ms'])
    step_ms = sum(r['ms'] for r in table)
    streaming = [r for r in table if r['gbs'] is not None]
    dom = streaming[0] if streaming else None
    fwd = [r for r in streaming if r['kernel'].startswith('fwd')]
    fwd_bytes = sum(r['algorithmic_bytes'] for r in fwd)
    fwd_ms = sum(r['ms'] for r in fwd)
    all_bytes = sum(r['algorithmic_bytes'] for r in streaming)
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if dom and os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom['kernel'])
    roofline = None
    if dom:
        roofline = {'bound': 'hbm', 'kernel': dom['kernel'], 'achieved': dom['gbs'], 'peak': peak,
                    'unit': 'GB/s', 'frac': dom['gbs'] / peak, 'traffic': traffic,
                    'peak_source': peak_src, 'ms_per_launch': dom['ms'],
                    'share_of_step': dom['ms'] / step_ms if step_ms else None,
                    'forward_total': {'algorithmic_gb': fwd_bytes / 1e9, 'ms': fwd_ms,
                                      'gbs': (fwd_bytes / 1e9) / (fwd_ms / 1e3) if fwd_ms else None,
                                      'frac': ((fwd_bytes / 1e9) / (fwd_ms / 1e3)) / peak if fwd_ms else None},
                    'step_total': {'algorithmic_gb': all_bytes / 1e9, 'ms': step_ms,
                                   'frac': ((all_bytes / 1e9) / (step_ms / 1e3)) / peak if step_ms else None}}
    if args.kernel_table:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernel_table)), exist_ok=True)
        json.dump(table, open(args.kernel_table, 'w'), indent=1)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_steps(args.arch, args.cpu_sample, S, steps=5, warmup=1)
        cpu = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
               'sample': f'{args.cpu_sample} images/step, 5 timed steps of the oracle train step '
                         f'(fwd+SimOTA+loss+bwd+SGD), torch CPU, best thread count of those probed '
                         f'up to {r["cores"]} host cores (used {r["threads"]})',
               'ms_per_step': r['ms_per_step'], 'split_ms': r['split_ms']}

    # ---------------- the other BASELINE configs and the plugin-surface step, same process (N = 1)
    extra, e2e_plugin = None, None
    if world == 1 and not args.no_extra:
        free_cache = torch.cuda.empty_cache
        del eng
        free_cache()
        try:
            e2e_plugin = plugin_e2e(args.arch, B, S, dev, steps=max(5, K // 2), warmup=3)
            e2e_plugin['torch_sgd'] = {k: v for k, v in plugin_e2e(
                args.arch, B, S, dev, steps=max(5, K // 2), warmup=3, optimizer='torch').items()
                if k in ('value', 'ms_per_step', 'host_ms')}
        except Exception as ex:      # reported, never fatal for the headline line
            e2e_plugin = {'error': repr(ex)[:300]}
        extra = {}
        try:
            extra['yunet_s_train_320_bs256'] = extra_train_config('yunet_s', 256, 320, dev, peak)
        except Exception as ex:
            extra['yunet_s_train_320_bs256'] = {'error': repr(ex)[:300]}
        try:
            extra['yunet_n_infer_640_bs512'] = extra_infer_config('yunet_n', 512, 640, dev, peak,
                                                                   cpu=not args.no_cpu_baseline)
        except Exception as ex:
            extra['yunet_n_infer_640_bs512'] = {'error': repr(ex)[:300]}

    gimg = B * world
    line = {
        'metric': METRIC, 'value': gimg / (ms_dev / 1e3), 'unit': UNIT, 'n_gpus': world,
        'steps': K, 'warmup': Wm, 'ms_per_step': ms_dev, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.arch} {S}x{S} train step (fwd + SimOTA + loss + bwd + '
                               f'grad all-reduce + SGD), bs={B}/GPU, synthetic WIDER-shaped GT '
                               f'(BASELINE.json configs[{1 if world == 1 else 2}])',
                   'arch': args.arch, 'global_batch': gimg, 'image_size': S, 'priors': P,
                   'parallelism': f'dp{world}',
                   'l2': 'inputs+activations per step (>5 GB) exceed the 126 MB L2; no flush needed'},
        'e2e': {'value': gimg / (ms_e2e / 1e3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': roofline,
        'cpu_baseline': cpu,
        'cuda_graph': graph_info,
        'e2e_plugin': e2e_plugin,
        'comm_ms': comm,
        'extra_configs': extra,
        'losses_last_step': loss_vals,
        'kernels_top': table[:6],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_infer(args):
    """BASELINE.json configs[3]: yunet_n 640x640 inference-only throughput, bs=512, incl. NMS."""
    from libfacedetection.train_b200 import YuNetEngine
    torch.cuda.set_device(0)
    B = 512 if args.batch == 256 else args.batch
    S = 640 if args.size == 320 else args.size
    eng = YuNetEngine(args.arch)
    gold = os.path.join(ROOT, 'tests', 'golden', f'weights_{args.arch}.npz')
    d = np.load(gold)
    eng.load_state_dict({k: torch.from_numpy(d[k]) for k in d.files})
    g = torch.Generator(device='cuda').manual_seed(0)
    img = torch.rand(B, 3, S, S, device='cuda', generator=g) * 255
    K, Wm = args.steps, max(args.warmup, 3)
    for _ in range(Wm):
        eng.detect(img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        dets, counts, _ = eng.detect(img)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    peak, src = measured_peaks()
    alg = 112.15e6 * B * (S / 640.0) ** 2      # forward algorithmic bytes / image @640
    print(json.dumps({
        'metric': 'infer_images_per_sec_640', 'value': B / (ms / 1e3), 'unit': UNIT, 'n_gpus': 1,
        'steps': K, 'warmup': Wm, 'ms_per_step': ms, 'higher_is_better': True, 'dtype': 'f32',
        'data': 'synthetic', 'config': {'workload': f'{args.arch} {S}x{S} eval forward + decode + '
                                        f'NMS, bs={B} (BASELINE.json configs[3])'},
        'detections_per_image': float(counts.float().mean()),
        'roofline': {'bound': 'hbm', 'achieved': alg / 1e9 / (ms / 1e3), 'peak': peak, 'unit': 'GB/s',
                     'frac': alg / 1e9 / (ms / 1e3) / peak, 'peak_source': src,
                     'note': 'whole forward+NMS step against the forward algorithmic bytes'}}),
        flush=True)


def main():
    args = parse()
    if args.workload == 'infer' and args.impl != 'reference':
        return run_infer(args)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
