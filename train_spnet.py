#! /usr/bin/env python3
"""Training CLI — same flags and flow as the reference's train_spnet.py (train_network :32-85,
__main__ :89-152) on the B200 engine. Under `torchrun --nproc-per-node N` every rank runs this
script and the batch is sharded across the N GPUs (--parallel), replacing make_parallel."""
import os
import random
import sys
import time

import numpy as np

from spnet import callbacks, models, multi_gpu, utils
import spnet.config as cf
from predict_spnet import predict_network, default_image_dir
from evaluate_spnet import evaluate_network


def train_network(weights_file="weights.hdf5", datapath=".", fraction=1.0, batch_size=32, epochs=30, pred_grid=[6, 6, 2],
                  noaugment=False, log_dir=".", lr_max=4e-5, freeze_fac=0.7, frozen_epochs=4, random_seed=1,
                  parallel=False, device_data=False):
    np.random.seed(random_seed)
    print("pred_grid = ", pred_grid)
    trainpath = datapath + "/Train/"
    valpath = datapath + "/Val/"
    X_train, Y_train, train_file_list, pred_shape = utils.build_dataset(
        path=trainpath, load_frac=fraction, set_means_ranges=True, batch_size=batch_size, pred_grid=pred_grid)
    X_val, Y_val, val_file_list, pred_shape = utils.build_dataset(
        path=valpath, load_frac=1.0, set_means_ranges=False, batch_size=batch_size, pred_grid=pred_grid)
    print("Seting up NN model.  model_type = ", cf.model_type)
    model, serial_model = models.setup_model(X_train, Y_train[0].size, no_cp_fatal=False, weights_file=weights_file,
                                             parallel=parallel, freeze_fac=freeze_fac)
    myprogress = callbacks.MyProgressCallback(X_val=X_val, Y_val=Y_val, val_file_list=val_file_list, log_dir=log_dir,
                                              pred_shape=pred_shape, batch_size=batch_size)
    checkpointer = callbacks.ParallelCheckpointCallback(model, filepath=weights_file, save_every=5, dir=log_dir)
    lr_sched = callbacks.OneCycleScheduler(lr_max=lr_max, n_data_points=X_train.shape[0], epochs=epochs,
                                           batch_size=batch_size, verbose=1)
    callback_list = [myprogress, checkpointer, lr_sched]
    if device_data:
        # B200 build only: the training set (and AugmentOnTheFly's pristine copy) stay in HBM for the whole run;
        # batches are gathered on the device and the per-epoch augmentation is one kernel launch
        import torch
        X_train, Y_train = torch.from_numpy(X_train).cuda(), torch.from_numpy(Y_train).cuda()
    if not noaugment:
        print("Adding callback for augment on the fly")
        callback_list.append(callbacks.AugmentOnTheFly(X_train, Y_train, aug_every=1))
    later_epochs = epochs - frozen_epochs
    if (frozen_epochs > 0) and (freeze_fac > 0.0):
        model.fit(X_train, Y_train, batch_size=batch_size, epochs=frozen_epochs, shuffle=True, verbose=1,
                  validation_data=(X_val, Y_val), callbacks=callback_list)
    if freeze_fac > 0.0:
        model = models.unfreeze_model(model, X_train, Y_train, parallel=parallel)
    model.fit(X_train, Y_train, batch_size=batch_size, epochs=later_epochs, shuffle=True, verbose=1,
              validation_data=(X_val, Y_val), callbacks=callback_list)
    return model


if __name__ == "__main__":
    seed = 1
    np.random.seed(seed)
    random.seed(seed)
    import argparse
    parser = argparse.ArgumentParser(description="trains network on training dataset",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("-b", "--batch_size", type=int, help="Batch size to use", default=16)
    parser.add_argument("-d", "--datapath", help="Directory with images in Train/ and Val/ subdirs", default="./")
    parser.add_argument("-e", "--epochs", type=int, help="Number of epochs to run", default=100)
    parser.add_argument("-f", "--fraction", type=float, help="Fraction of dataset to use (for quick testing: -f 0.05)", default=1.0)
    parser.add_argument("--freeze_fac", type=float, help="Fraction of base model (e.g. Xception) to freeze", default=0.0)
    parser.add_argument("--frozen_epochs", type=int, help="Number of starting epochs to run while base model is frozen", default=0)
    parser.add_argument("-g", "--grid", help="Shape of predictor grid", default="6x6x2")
    parser.add_argument("-w", "--weights", help="Weights file in hdf5 format", default="weights.hdf5")
    parser.add_argument("-l", "--lrmax", type=float, help="Maximum learning rate value", default=4e-5)
    parser.add_argument("-n", "--noaugment", action="store_true", help="don't augment on the fly")
    parser.add_argument("--name", help="Descriptive name of the run, prepended to the log directory name", default="")
    parser.add_argument("-r", "--random_seed", type=int, help="Random seed value", default=1)
    parser.add_argument("--dtype", choices=["bf16", "fp32"], default=cf.compute_dtype, help="compute precision (B200 build only)")
    parser.add_argument("--model_type", default=cf.model_type, help="'big' keeps 384x512 input, default resizes to 331x331")
    parser.add_argument("--parallel", action="store_true", help="shard each batch over the ranks of a torchrun launch")
    parser.add_argument("--device_data", action="store_true", help="keep the training set in GPU memory and augment it there (B200 build only)")
    parser.add_argument("--predict_path", default=default_image_dir,
                        help="directory of frames for the post-training predict_network pass (reference: hard-coded Zooniverse path)")
    args = parser.parse_args()
    print("Command line ~= \n", " ".join(s for s in sys.argv))
    print("args = ", args)
    cf.compute_dtype = args.dtype
    cf.model_type = args.model_type
    if args.parallel or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
        args.parallel = True
    pred_grid = [int(i) for i in args.grid.split("x")]
    now = time.strftime("%c").replace("  ", "_").replace(" ", "_")
    log_dir = "./logs/" + args.name + "_" + now if args.name else "./logs/" + now
    print("Logging will go to ", log_dir)
    print("\n----------------------------\nStarting training...")
    model = train_network(weights_file=args.weights, datapath=args.datapath, fraction=args.fraction,
                          batch_size=args.batch_size, epochs=args.epochs, pred_grid=pred_grid, noaugment=args.noaugment, device_data=args.device_data,
                          log_dir=log_dir, lr_max=args.lrmax, freeze_fac=args.freeze_fac,
                          frozen_epochs=args.frozen_epochs, random_seed=args.random_seed, parallel=args.parallel)
    if multi_gpu.world()[0] == 0:
        # run model evaluation (train_spnet.py:130-138)
        print("\n----------------------------\nStarting model evaluation...")
        testpath = args.datapath + "/Test/"
        if not os.path.isdir(testpath):
            testpath = args.datapath + "/Val/"
        evaluate_network(model=model, weights_file="", datapath=testpath, fraction=1.0, log_dir="logs/Evaluation/",
                         batch_size=args.batch_size, pred_grid=pred_grid, set_means_ranges=False)
        # make predictions on the Zooniverse dataset (:140-143); the reference's path is the author's home directory,
        # so the pass is skipped with a notice where that directory holds no frames (the reference would raise there)
        print("\n----------------------------\nStarting Zooniverse predictions...")
        import glob
        if glob.glob(args.predict_path + "/*.png") or glob.glob(args.predict_path + "/*.bmp"):
            predict_network(weights_file="", datapath=args.predict_path, fraction=args.fraction, log_dir="logs/Predicting/",
                            batch_size=args.batch_size, model=model, X_pred="", ranks=(0, 1))
        else:
            print("    no frames under", args.predict_path, "- skipping (give --predict_path)")
        weights2name = "final_" + args.weights
        print("Just to be sure: Saving model to", weights2name)
        model.save_weights(weights2name)
        print("And saving full model too")
        model.save("full_model.h5")
    print("SPNet execution completed.")
