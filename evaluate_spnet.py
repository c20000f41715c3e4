#! /usr/bin/env python3
"""Scoring CLI — same flags and flow as the reference's evaluate_spnet.py (evaluate_network :38-94, flags
:98-112) on the B200 engine: load the test set, predict, de-normalise, print mAP (diagnostics.calc_map), the
ring-count / existence error summary (diagnostics.calc_errors) and write hawley_spnet.csv. The reference loads
`full_model.h5` through Keras; here `-w` names the weights / whole-model file the way predict_spnet.py does.
Extra flags (not in the reference): --no-png, --dtype, --model_type."""
import time

import numpy as np

from spnet import diagnostics, models, utils
from spnet.utils import *    # noqa: F401,F403  (reference: `from spnet.utils import *`)
import spnet.config as cf


def evaluate_network(model=None, weights_file="", datapath="Test/", fraction=1.0, log_dir="", batch_size=32,
                     pred_grid=[6, 6, 2], set_means_ranges=True, draw_images=True):
    np.random.seed(1)
    print("Getting data..., fraction = ", fraction)
    X_test, Y_test, test_file_list, pred_shape = utils.build_dataset(
        path=datapath, load_frac=fraction, set_means_ranges=set_means_ranges, batch_size=batch_size, shuffle=False,
        pred_grid=pred_grid)
    if model is None:
        if ".hdf5" in weights_file:
            print("Setting up model and then loading weights from", weights_file)
            model, _ = models.setup_model(X_test, Y_test[0].size, try_checkpoint=True, no_cp_fatal=True,
                                          weights_file=weights_file, parallel=False, quick_setup=True, freeze_fac=0)
        else:
            print("Loading full model from", weights_file)
            model = models.load_model(weights_file)
    m = X_test.shape[0]
    print("    Predicting... (m = ", m, " frames in this (Test?) dataset)", sep="")
    start_time = time.time()
    Y_pred = model.predict(X_test, batch_size=batch_size)
    elapsed = time.time() - start_time
    print("    ...elapsed time to predict = ", elapsed, "s.   FPS = ", m * 1.0 / elapsed)
    if cf.loss_type != "same":  # convert from logits if needed
        Y_pred[:, cf.ind_noobj::cf.vars_per_pred] = 1.0 / (1.0 + np.exp(-Y_pred[:, cf.ind_noobj::cf.vars_per_pred]))
    Yt, Yp = utils.denorm_Y(Y_test), utils.denorm_Y(Y_pred)  # normalised -> 'world' values
    try:
        mean_ap = diagnostics.calc_map(Yp, Yt)
    except ZeroDivisionError:
        # no (prediction, label) pair with both ellipses present: tp = fp = fn = 0 and precision() divides by zero -
        # the reference's precision() does the same and the script dies there; here the summary goes on without a mAP
        mean_ap = float("nan")
        print("    no detections at all: mean average precision is undefined (the reference raises ZeroDivisionError here)")
    print("mAP = ", mean_ap)
    (ring_miscounts, ring_truecounts, total_obj, false_obj_pos, false_obj_neg, true_obj_pos, true_obj_neg, pix_err,
     ipem) = diagnostics.calc_errors(Yp, Yt)
    mistakes = ring_miscounts + false_obj_pos + false_obj_neg
    class_acc = (total_obj - mistakes) * 1.0 / total_obj * 100
    print("Mean pixel error =", np.mean(pix_err))
    print("    Ring correct counts = ", ring_truecounts, " / ", total_obj, ".   = ", 100 * ring_miscounts / total_obj,
          " % ring-class accuracy", sep="")
    print("         Ring miscounts = ", ring_miscounts, " / ", total_obj, ".   = ", 100 * ring_miscounts / total_obj,
          " % ring-miscount rate", sep="")
    print("        False positives = ", false_obj_pos, " / ", total_obj, ".   = ", 100 * false_obj_pos / total_obj,
          " % FP rate", sep="")
    print("        False negatives = ", false_obj_neg, " / ", total_obj, ".   = ", 100 * false_obj_neg / total_obj,
          " % FN rate", sep="")
    print("         True positives = ", true_obj_pos, " / ", total_obj, ".   = ", 100 * true_obj_pos / total_obj,
          " % TP rate", sep="")
    print("         True negatives = ", true_obj_neg, sep="")
    print("    Total Mistakes = ", mistakes, " / ", total_obj, ".   => ", class_acc,
          " % class. accuracy rate (lack of mistakes)", sep="")
    utils.make_sure_path_exists(log_dir)
    print("    Drawing sample ellipse images...")
    utils.show_pred_ellipses(Yt, Yp, test_file_list, num_draw=m, log_dir=log_dir, out_csv=log_dir + "hawley_spnet.csv",
                             draw_images=draw_images)
    evaluate_network.last = dict(mAP=mean_ap, class_acc=class_acc, total_obj=total_obj, mistakes=mistakes)
    return model


if __name__ == "__main__":
    import argparse
    parser = argparse.ArgumentParser(description="tests network on test dataset",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("-w", "--weights", help="weights file in hdf5 format", default="weights.hdf5")
    parser.add_argument("-d", "--datapath", help="Test dataset directory", default="Test/")
    parser.add_argument("-f", "--fraction", type=float, help="Fraction of dataset to use", default=1.0)
    parser.add_argument("-l", "--logdir", help="Directory to write log files into", default="logs/Testing/")
    parser.add_argument("-b", "--batch_size", type=int, help="Batch size to use", default=16)
    parser.add_argument("--no-png", action="store_true", help="write only hawley_spnet.csv, skip the per-image PNGs")
    parser.add_argument("--dtype", choices=["bf16", "fp32"], default=cf.compute_dtype)
    parser.add_argument("--model_type", default=cf.model_type, help="'big' keeps 384x512 input, default resizes to 331x331")
    args = parser.parse_args()
    cf.compute_dtype = args.dtype
    cf.model_type = args.model_type
    evaluate_network(weights_file=args.weights, datapath=args.datapath, fraction=args.fraction, log_dir=args.logdir,
                     batch_size=args.batch_size, draw_images=not args.no_png)
