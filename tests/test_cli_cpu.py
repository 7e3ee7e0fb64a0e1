"""Host logic of the command-line training loop (``cli.run_training`` / ``sharded_batches``) with a
fake engine: sharding across ranks, LR schedule progression, checkpoint cadence."""
import numpy as np

from libfacedetection.train_b200 import cli, trainer


class FakeSamples:
    def __init__(self, n, empty=()):
        self.n, self.empty = n, set(empty)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = 0 if i in self.empty else 1
        return (np.full((4, 4, 3), i, np.uint8), np.zeros((g, 4), np.float32),
                np.zeros((g, 5, 3), np.float32), np.zeros(g, np.int64))


class FakeEngine:
    def __init__(self):
        self.lrs = []

    def train_step(self, img, gt, offs, lr, momentum, weight_decay):
        self.lrs.append(lr)
        return np.array([0.1, 0.2, 0.3, 0.4])


def test_sharded_batches_partition_the_epoch():
    s = FakeSamples(41, empty=(3, 17))
    seen = []
    for rank in range(2):
        for images, boxes, kps, labels in cli.sharded_batches(s, 4, epoch=5, rank=rank, world=2, seed=1):
            assert len(images) == len(boxes) == len(kps) == len(labels) == 4
            seen += [int(im[0, 0, 0]) for im in images]
    assert len(seen) == len(set(seen))                 # no sample twice, ranks are disjoint
    assert 3 not in seen and 17 not in seen            # samples without a face are skipped
    a = [int(b[0][0][0, 0, 0]) for b in cli.sharded_batches(s, 4, epoch=5, rank=0, world=2, seed=1)]
    b = [int(b[0][0][0, 0, 0]) for b in cli.sharded_batches(s, 4, epoch=6, rank=0, world=2, seed=1)]
    assert a != b                                      # the order changes with the epoch


def test_training_loop_schedule_and_checkpoint_cadence():
    eng, saved, logs = FakeEngine(), [], []
    aug = lambda images, boxes, kps, labels: (images, boxes, kps)      # noqa: E731
    it = cli.run_training(eng, aug, FakeSamples(32), epochs=6, batch_size=8, save=lambda e, i, lr: saved.append((e, i)),
                          save_every=2, log=logs.append, log_every=4)
    assert it == 6 * 4 and len(eng.lrs) == 24
    assert eng.lrs == [trainer.lr_at(i, i // 4) for i in range(24)]    # warm-up ramp of the reference config
    assert saved == [(2, 8), (4, 16), (6, 24)]
    assert len(logs) == 6 and logs[0].startswith('epoch 1 iter 0 ')
    # resume: continues the iteration count and the schedule
    eng2 = FakeEngine()
    it2 = cli.run_training(eng2, aug, FakeSamples(32), epochs=6, batch_size=8, start_epoch=4, start_iter=16, log_every=0)
    assert it2 == 24 and eng2.lrs == eng.lrs[16:]


def test_every_rank_runs_the_same_number_of_batches():
    """face-less samples are dropped from the SHARED order before the strided
    split, so no rank gets an extra batch (which would hang the all-reduce)."""
    for seed in range(20):
        rng = np.random.RandomState(seed)
        s = FakeSamples(200, empty=rng.choice(200, 9, replace=False))
        counts = [sum(1 for _ in cli.sharded_batches(s, 4, epoch=seed, rank=r, world=2, seed=seed)) for r in range(2)]
        assert counts[0] == counts[1] == cli.num_batches(s, 4, world=2) == (200 - 9) // 8
