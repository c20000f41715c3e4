"""Training-control callbacks with the reference's names and behaviour (spnet/callbacks.py):
the per-batch 1-cycle LR schedule, the periodic checkpoint, a text-only progress/diagnostics
callback and augment-on-the-fly. They plug into SPNetModel.fit through the Keras callback
protocol (set_model / on_train_begin / on_epoch_begin / on_batch_begin / on_epoch_end)."""
import os
import random
import time

import numpy as np

from . import config as cf
from . import utils


class Callback:
    def __init__(self):
        self.model = None
        self.params = {}

    def set_model(self, model):
        self.model = model

    def set_params(self, params):
        self.params = params

    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_epoch_begin(self, epoch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass
    def on_batch_begin(self, batch, logs=None): pass
    def on_batch_end(self, batch, logs=None): pass


def get_1cycle_schedule(lr_max=1e-3, n_data_points=8000, epochs=200, batch_size=40, verbose=0):
    """Per-iteration LR look-up table: linear warm-up over the first 30 % of the iterations from
    lr_max/25 to lr_max, cosine anneal to lr_max/25/1e4 (spnet/callbacks.py:346-377)."""
    if verbose > 0:
        print("Setting up 1Cycle LR schedule...")
    pct_start, div_factor = 0.3, 25.0
    lr_start = lr_max / div_factor
    lr_end = lr_start / 1e4
    n_iter = n_data_points * epochs // batch_size
    a1 = int(n_iter * pct_start)
    a2 = n_iter - a1
    up = np.linspace(lr_start, lr_max, a1)
    down = (lr_max - lr_end) * (1 + np.cos(np.linspace(0, np.pi, a2))) / 2 + lr_end
    return np.concatenate((up, down))


class OneCycleScheduler(Callback):
    """Sets model.optimizer.lr from the table on every batch (spnet/callbacks.py:380-406)."""

    def __init__(self, **kwargs):
        super().__init__()
        self.verbose = kwargs.pop("verbose", 0)
        self.lrs = get_1cycle_schedule(**kwargs)
        self.iteration = 0

    def on_batch_begin(self, batch, logs=None):
        self.model.optimizer.lr = float(self.lrs[min(self.iteration, len(self.lrs) - 1)])
        self.iteration += 1

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        lr = self.model.optimizer.lr
        logs["lr"] = lr
        if self.verbose > 0:
            print("\nLearning rate =", lr)


class ParallelCheckpointCallback(Callback):
    """Every save_every epochs: weights of the serial model -> dir/<filepath>, full model ->
    dir/spnet.model (spnet/callbacks.py:20-41). Rank 0 only under data parallelism."""

    def __init__(self, model, filepath="weights.hdf5", save_every=1, dir="."):
        super().__init__()
        self.model_to_save = model
        self.filepath = filepath
        self.save_every = save_every
        self.dir = dir

    def on_epoch_end(self, epoch, logs=None):
        from . import multi_gpu
        if multi_gpu.world()[0] != 0:
            return
        if epoch % self.save_every == 0:
            utils.make_sure_path_exists(self.dir)
            target = multi_gpu.get_serial_part(self.model if self.model is not None else self.model_to_save)
            print("Saving checkpoint to", os.path.join(self.dir, os.path.basename(self.filepath)))
            target.save_weights(os.path.join(self.dir, os.path.basename(self.filepath)))
            target.save(os.path.join(self.dir, "spnet.model"))


class MyProgressCallback(Callback):
    """Text part of the reference's progress callback (spnet/callbacks.py:58-265): per epoch,
    predict on the validation set, append `epoch train val center size angle noobj class` to
    losses.dat and print the FPS line. The plots / sample PNGs are out of scope."""

    def __init__(self, X_val=None, Y_val=None, val_file_list=None, log_dir="./logs", pred_shape=None, **kwargs):
        super().__init__()
        self.X_val, self.Y_val, self.val_file_list = X_val, Y_val, val_file_list
        self.log_dir, self.pred_shape = log_dir, pred_shape
        self.batch_size = kwargs.get("batch_size", 32)

    def on_train_begin(self, logs=None):
        from . import multi_gpu
        self.rank0 = multi_gpu.world()[0] == 0
        if self.rank0:
            utils.make_sure_path_exists(self.log_dir)
            with open(os.path.join(self.log_dir, "losses.dat"), "w") as f:
                f.write("# epoch Train_total Val_total center size angle noobj class\n")

    def on_epoch_end(self, epoch, logs=None):
        from . import models
        logs = logs or {}
        if self.X_val is None or not self.rank0:
            return
        m = self.X_val.shape[0]
        t0 = time.time()
        Y_pred = self.model.predict(self.X_val, batch_size=self.batch_size)
        el = time.time() - t0
        print("  ...elapsed time to predict = ", el, "s.   FPS = ", m * 1.0 / max(el, 1e-9))
        total, parts = models.my_loss(self.Y_val, Y_pred)
        with open(os.path.join(self.log_dir, "losses.dat"), "a") as f:
            f.write("%d %g %g %s\n" % (epoch + 1, logs.get("loss", float("nan")), logs.get("val_loss", total),
                                       " ".join("%g" % p for p in parts)))


class AugmentOnTheFly(Callback):
    """Each aug_every epochs rewrite X in place from a pristine copy: cutout (<= 6 rectangles of
    11..75 px) and salt-and-pepper (50 % chance, 0.4 % of the pixels); the reference's blur is a no-op
    (spnet/callbacks.py:272-341, spnet/augmentation.py:66-71,117-180). Labels are untouched."""

    def __init__(self, X, Y, aug_every=1):
        super().__init__()
        self.X, self.Y, self.aug_every = X, Y, aug_every
        self.X_orig = np.copy(X)

    def on_epoch_begin(self, epoch, logs=None):
        if epoch % self.aug_every != 0:
            return
        X = self.X
        X[...] = self.X_orig
        n, H, W = X.shape[0], X.shape[1], X.shape[2]
        for i in range(n):
            for _ in range(random.randint(0, 6)):
                h, w = random.randint(11, 75), random.randint(11, 75)
                y0, x0 = random.randint(0, max(0, H - h)), random.randint(0, max(0, W - w))
                X[i, y0:y0 + h, x0:x0 + w, :] = 0.0
            if random.random() < 0.5:
                k = int(0.004 * H * W)
                ys, xs = np.random.randint(0, H, k), np.random.randint(0, W, k)
                X[i, ys, xs, :] = np.where(np.random.rand(k, 1) < 0.5, -1.0, 1.0)
