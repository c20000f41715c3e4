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


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def _epoch_seed(seed, epoch):
    """Per-epoch key of the device augmentation. The kernel keys every draw as seed + G*(frame+1) + ...; an ADDITIVE
    per-epoch offset with the same constant made (epoch e, frame f) and (epoch e+1, frame f-1) share their draws, so
    the epoch goes through a proper mix instead: splitmix64(seed ^ splitmix64(epoch))."""
    return _splitmix64((seed & 0xFFFFFFFFFFFFFFFF) ^ _splitmix64(epoch))


class ParallelCheckpointCallback(Callback):
    """Every save_every epochs: weights of the serial model -> dir/<filepath>, full model ->
    dir/spnet.model (spnet/callbacks.py:20-41). Rank 0 only under data parallelism."""

    def __init__(self, model, filepath="weights.hdf5", save_every=1, dir="."):
        super().__init__()
        self.model_to_save = model
        self.save_every = save_every
        self.dir = dir
        self.weights_path = dir + "/" + filepath        # spnet/callbacks.py:31-32
        self.model_path = dir + "/" + "spnet.model"

    def on_epoch_end(self, epoch, logs=None):
        from . import multi_gpu
        if multi_gpu.world()[0] != 0:
            return
        # epoch + 1 agrees with Keras' "Epoch 1/20" display: after the 5th, 10th, ... epoch for save_every = 5
        if (1 == self.save_every) or ((0 == ((epoch + 1) % self.save_every)) and (epoch > 0)):
            utils.make_sure_path_exists(os.path.dirname(self.weights_path) or ".")
            # the model being trained right now (after unfreeze_model it is a new object), else the one given at set-up
            target = multi_gpu.get_serial_part(self.model if self.model is not None else self.model_to_save)
            print("Saving weights checkpoint to", self.weights_path)
            target.save_weights(self.weights_path)
            print("Saving entire model checkpoint to", self.model_path)
            target.save(self.model_path)


class MyProgressCallback(Callback):
    """Text part of the reference's progress callback (spnet/callbacks.py:58-265): per epoch,
    predict on the validation set, append `epoch train val center size angle noobj class` to
    losses.dat, print the FPS line, and compute the ring-count / existence accuracy the reference plots
    (diagnostics.calc_errors on the de-normalised predictions, :156-165; kept in self.acc_hist and printed).
    The plots / sample PNGs are out of scope."""

    def __init__(self, X_val=None, Y_val=None, val_file_list=None, log_dir="./logs", pred_shape=None, **kwargs):
        super().__init__()
        self.X_val, self.Y_val, self.val_file_list = X_val, Y_val, val_file_list
        self.log_dir, self.pred_shape = log_dir, pred_shape
        self.batch_size = kwargs.get("batch_size", 32)
        self.acc_hist = []

    def on_train_begin(self, logs=None):
        from . import multi_gpu
        self.rank0 = multi_gpu.world()[0] == 0
        if self.rank0:
            utils.make_sure_path_exists(self.log_dir)
            path = os.path.join(self.log_dir, "losses.dat")
            if not getattr(self, "_header_written", False):  # a second fit() (after unfreeze_model) keeps the frozen-phase log
                with open(path, "w") as f:
                    f.write("# epoch Train_total Val_total center size angle noobj class\n")
                self._header_written = True

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
        # ring-count / existence accuracy on world values (spnet/callbacks.py:152-165)
        from . import diagnostics
        if len(utils.means) == 0:
            utils.setup_means_and_ranges(self.pred_shape or [6, 6, 2, cf.vars_per_pred])
        Yp = np.array(Y_pred, dtype=np.float32, copy=True)
        if cf.loss_type != "same":  # logits -> probabilities
            Yp[:, cf.ind_noobj::cf.vars_per_pred] = 1.0 / (1.0 + np.exp(-Yp[:, cf.ind_noobj::cf.vars_per_pred]))
        r = diagnostics.calc_errors(utils.denorm_Y(Yp), utils.denorm_Y(np.asarray(self.Y_val, dtype=np.float32)))
        ring_miscounts, total_obj, false_obj_pos, false_obj_neg = r[0], r[2], r[3], r[4]
        mistakes = ring_miscounts + false_obj_pos + false_obj_neg
        class_acc = (total_obj - mistakes) * 1.0 / total_obj * 100 if total_obj else float("nan")
        self.acc_hist.append(class_acc)
        print("  diagnostics: ring_miscounts, total_obj, false_obj_pos, false_obj_neg = %d %d %d %d   class accuracy = %.2f %%"
              % (ring_miscounts, total_obj, false_obj_pos, false_obj_neg, class_acc))


class AugmentOnTheFly(Callback):
    """Each aug_every epochs rewrite the training frames X in place from a pristine copy with the reference's two
    live augmentations (spnet/callbacks.py:272-341): cutout_inplace (spnet/augmentation.py:117-135: 0..6
    rectangles, corner in [0, dim-11), extents 11..74 clipped to dim-1, filled with one grey value drawn between
    the frame's min and max) and salt_n_pepa_inplace (:159-180: with probability 1/2, ceil(0.004*size*0.2) pixels
    set to the frame's max, then ceil(0.004*size*0.8) to its min, coordinates in [0, dim-1)). blur_inplace
    (:66-71) discards its result, i.e. is a no-op, and bp_mixup is commented out at the call site. Labels are
    untouched (they are already grid-assigned and normalised).

    X may be a numpy array (host path: numpy's global RNG consumed in the reference's order, so the same seed gives
    the reference's frames bit for bit) or a CUDA tensor: then the pristine
    copy stays in HBM too and one kernel launch per epoch does the rewrite (csrc/augment.cu,
    spnet_augment_on_the_fly) - the device-resident input path of SPNetModel.fit."""

    def __init__(self, X, Y, orig_img_shape=(384, 512), aug_every=1, seed=None):
        super().__init__()
        self.X, self.Y, self.aug_every, self.orig_img_shape = X, Y, aug_every, orig_img_shape
        self.on_device = not isinstance(X, np.ndarray)
        self.X_orig = X.clone() if self.on_device else np.copy(X)
        self.seed = random.getrandbits(62) if seed is None else int(seed)

    @staticmethod
    def cutout(img, max_regions=6, minsize=11, maxsize=75):
        n = np.random.randint(0, max_regions + 1)
        if n == 0:
            return
        lo, hi = np.min(img), np.max(img)
        for _ in range(n):
            y0, x0 = np.random.randint(0, img.shape[0] - minsize), np.random.randint(0, img.shape[1] - minsize)
            eh, ew = np.random.randint(minsize, maxsize), np.random.randint(minsize, maxsize)
            y1, x1 = min(y0 + eh, img.shape[0] - 1), min(x0 + ew, img.shape[1] - 1)
            img[y0:y1, x0:x1, :] = np.random.uniform(lo, hi)

    @staticmethod
    def salt_n_pepa(img, salt_vs_pepper=0.2, amount=0.004):
        if np.random.choice(["good", "not good"]) != "good":  # the reference's coin, same draw from numpy's stream
            return
        salt, pepper = np.max(img), np.min(img)
        for count, value in ((int(np.ceil(amount * img.size * salt_vs_pepper)), salt),
                             (int(np.ceil(amount * img.size * (1.0 - salt_vs_pepper))), pepper)):
            ys = np.random.randint(0, img.shape[0] - 1, count)
            xs = np.random.randint(0, img.shape[1] - 1, count)
            img[ys, xs, :] = value

    def on_epoch_begin(self, epoch, logs=None):
        if epoch % self.aug_every != 0:
            return
        if self.on_device:
            from . import ops
            ops.augment_on_the_fly(self.X_orig, self.X, _epoch_seed(self.seed, epoch))
            return
        X = self.X
        X[...] = self.X_orig
        for i in range(X.shape[0]):
            self.cutout(X[i])
            self.salt_n_pepa(X[i])
            # blur() / blur_inplace() (callbacks.py:303-305, augmentation.py:66-71) change nothing - GaussianBlur's
            # result is dropped - but they draw from the generators; the same draws keep this loop on the
            # reference's random stream (same seed -> same augmented frames, tests/golden/ref_augment.npz)
            if np.random.rand() < 0.4 and np.random.random() <= 0.3:
                random.choice([3, 7])
