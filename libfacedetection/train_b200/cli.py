"""Command-line entry points composed from the tested pieces: what
``tools/train.py`` / ``tools/detect_image.py`` / ``tools/test_widerface.py`` do for YuNet, without
mmcv configs.

    python -m libfacedetection.train_b200.cli train  yunet_n --ann labelv2.txt --img-prefix images/ \\
        --size 320 --batch 256 --epochs 640 --work-dir work_dirs/yunet_n [--resume ckpt.pth]
    python -m libfacedetection.train_b200.cli detect yunet_n weights/yunet_n.pth photo.jpg [--mode 0]

``train`` follows ``configs/yunet_n.py``: SGD 0.01 / 0.9 / 5e-4, linear warm-up over 1 500
iterations, x0.1 at epochs 400 and 544, RandomSquareCrop + Resize + RandomFlip on the GPU
(``pipeline.GpuAugmenter``), a reference-format checkpoint every ``--save-every`` epochs.  Under
``torchrun`` every rank reads its own shard of the shuffled sample order (DistributedSampler-like)
and the step all-reduces the gradient bucket.
"""
import argparse
import os

import numpy as np


def sharded_batches(samples, batch_size, epoch, rank=0, world=1, seed=0):
    """Batches of this rank for ``epoch``: one shuffled order shared by all ranks (seeded by
    ``seed + epoch`` like ``DistributedSampler.set_epoch``).  Face-less samples are removed from
    the SHARED order first and the order is cut to a multiple of ``world * batch_size`` before the
    strided split, so every rank runs the same number of iterations (a rank with one batch more
    would pair its all-reduces with the next epoch of the others and hang)."""
    order = np.random.RandomState(seed + epoch).permutation(len(samples))
    usable = np.array([samples.num_faces(int(i)) > 0 for i in order], bool) if hasattr(samples, 'num_faces') \
        else np.array([samples[int(i)][1].shape[0] > 0 for i in order], bool)
    order = order[usable]
    order = order[:len(order) - len(order) % (world * batch_size)][rank::world]
    for k in range(0, len(order), batch_size):
        cur = [[], [], [], []]
        for i in order[k:k + batch_size]:
            for c, v in zip(cur, samples[int(i)]):
                c.append(v)
        yield tuple(cur)


def num_batches(samples, batch_size, world=1):
    """Iterations per epoch of every rank (independent of the rank by construction)."""
    n = sum(1 for i in range(len(samples))
            if (samples.num_faces(i) if hasattr(samples, 'num_faces') else samples[i][1].shape[0]) > 0)
    return n // (world * batch_size)


def run_training(engine, augment, samples, epochs, batch_size, start_epoch=0, start_iter=0,
                 base_lr=0.01, momentum=0.9, weight_decay=0.0005, rank=0, world=1, seed=0,
                 save=None, save_every=1, log=print, log_every=50):
    """The epoch / iteration loop.  ``augment(images, boxes, kps, labels) -> (img, gt, offsets)``;
    ``save(epoch, iteration, lr)`` is called after every ``save_every`` epochs (rank 0 only)."""
    from . import trainer
    it = start_iter
    lr = base_lr
    for epoch in range(start_epoch, epochs):
        for images, boxes, kps, labels in sharded_batches(samples, batch_size, epoch, rank, world, seed):
            lr = trainer.lr_at(it, epoch, base_lr)
            img, gt, offs = augment(images, boxes, kps, labels)
            losses = engine.train_step(img, gt, offs, lr=lr, momentum=momentum, weight_decay=weight_decay)
            if log_every and it % log_every == 0 and rank == 0:
                l = [float(v) for v in losses.tolist()]
                log(f'epoch {epoch + 1} iter {it} lr {lr:.3e} loss_cls {l[0]:.4f} loss_bbox {l[1]:.4f} '
                    f'loss_obj {l[2]:.4f} loss_kps {l[3]:.4f} loss {sum(l):.4f}')
            it += 1
        if save is not None and rank == 0 and (epoch + 1) % save_every == 0:
            save(epoch + 1, it, lr)
    return it


def cmd_train(args):
    import torch
    from . import YuNetEngine, dataset, dist_utils, pipeline, trainer
    rank, world = dist_utils.init_from_env()
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    eng = YuNetEngine(args.arch)
    start_epoch = start_iter = 0
    if args.resume:
        meta = trainer.load_checkpoint(eng, args.resume)
        start_epoch, start_iter = int(meta.get('epoch', 0)), int(meta.get('iter', 0))
    elif args.load_from:
        trainer.load_checkpoint(eng, args.load_from, resume_optimizer=False)
    else:
        eng.init_weights(args.seed)
    np.random.seed(args.seed + rank)                        # augmentation decisions (numpy global RNG)
    samples = dataset.RetinaFaceSamples(args.ann, args.img_prefix, min_size=args.min_size)
    aug = pipeline.GpuAugmenter(eng, size=args.size,
                                crop_choice=args.crop_choice if args.crop_choice is not None else
                                (pipeline.CROP_CHOICE_N if args.arch == 'yunet_n' else pipeline.CROP_CHOICE_S))
    os.makedirs(args.work_dir, exist_ok=True)

    def save(epoch, iteration, lr):
        trainer.save_checkpoint(eng, os.path.join(args.work_dir, f'epoch_{epoch}.pth'), epoch=epoch,
                                iteration=iteration, lr=lr)

    run_training(eng, aug, samples, args.epochs, args.batch, start_epoch, start_iter, args.lr, rank=rank,
                 world=world, seed=args.seed, save=save, save_every=args.save_every)


def cmd_detect(args):
    import cv2
    import torch
    from . import YuNetEngine, evaluation, trainer
    eng = YuNetEngine(args.arch)
    trainer.load_checkpoint(eng, args.checkpoint, resume_optimizer=False)
    img = cv2.imread(args.image, cv2.IMREAD_COLOR)
    chw, factor = evaluation.prepare_test_image(img, args.mode)
    t = torch.from_numpy(chw).to(eng.device)[None]
    sf = torch.from_numpy(factor).to(eng.device)[None]
    dets, counts, kps = eng.detect(t, args.score_thr, args.iou_thr, scale_factors=sf, with_kps=True)
    n = int(counts[0])
    d, k = dets[0, :n].cpu().numpy(), kps[0, :n].cpu().numpy() / np.tile(factor[:2], 5)
    for b, p in zip(d, k):
        print('%.1f %.1f %.1f %.1f %.4f  ' % tuple(b) + ' '.join('%.1f' % v for v in p))
    return d, k


def main(argv=None):
    ap = argparse.ArgumentParser(prog='libfacedetection.train_b200.cli')
    sub = ap.add_subparsers(dest='cmd', required=True)
    t = sub.add_parser('train')
    t.add_argument('arch', choices=['yunet_n', 'yunet_s'])
    t.add_argument('--ann', required=True, help='RetinaFace labelv2.txt')
    t.add_argument('--img-prefix', default='')
    t.add_argument('--size', type=int, default=640)
    t.add_argument('--batch', type=int, default=16, help='images per GPU (configs: samples_per_gpu)')
    t.add_argument('--epochs', type=int, default=640)
    t.add_argument('--lr', type=float, default=0.01)
    t.add_argument('--min-size', type=float, default=None)
    t.add_argument('--crop-choice', type=float, nargs='+', default=None,
                   help='RandomSquareCrop scales (default: the configs/<arch>.py list)')
    t.add_argument('--work-dir', default='work_dirs/yunet')
    t.add_argument('--save-every', type=int, default=10)
    t.add_argument('--resume', default=None)
    t.add_argument('--load-from', default=None)
    t.add_argument('--seed', type=int, default=0)
    t.set_defaults(fn=cmd_train)
    d = sub.add_parser('detect')
    d.add_argument('arch', choices=['yunet_n', 'yunet_s'])
    d.add_argument('checkpoint')
    d.add_argument('image')
    d.add_argument('--mode', type=int, default=0)
    d.add_argument('--score-thr', type=float, default=0.3)
    d.add_argument('--iou-thr', type=float, default=0.45)
    d.set_defaults(fn=cmd_detect)
    args = ap.parse_args(argv)
    return args.fn(args)


if __name__ == '__main__':
    main()
