"""predict_spnet end to end on the GPU: PNGs on disk -> build_X -> engine forward -> device decode ->
hawley_spnet.csv, compared with the host numpy decode and with the oracle's CSV text for the same outputs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_predict_cli_path_writes_reference_format_csv(tmp_path):
    from PIL import Image
    import spnet.config as cf
    from spnet import models, utils
    import predict_spnet
    from spnet_b200 import fake_espi
    from oracle import numpy_side as ns

    cf.model_type = "big"          # native 384x512 input (predict_spnet.py:50-52)
    cf.compute_dtype = "bf16"
    n = 6
    for i in range(n):
        img, _ = fake_espi.make_frame(100 + i)
        Image.fromarray(img).save(tmp_path / ("steelpan_%07d.png" % i))
    X = np.zeros((2, 384, 512, 1), np.float32)
    model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0, quick_setup=True)
    wpath = str(tmp_path / "weights.hdf5")
    model.save_weights(wpath)
    log_dir = str(tmp_path / "out") + "/"
    m2 = predict_spnet.predict_network(weights_file=wpath, datapath=str(tmp_path), fraction=1.0, log_dir=log_dir,
                                       batch_size=2, draw_images=True)
    text = open(log_dir + "hawley_spnet.csv").read()
    files = sorted(str(p) for p in tmp_path.glob("*.png"))
    assert len(text.strip().splitlines()) >= n
    for f in files:
        assert os.path.basename(f) in text
    assert os.path.exists(log_dir + "steelpan_pred_00000.png")
    # one set of network outputs through (a) device decode + product CSV writer, (b) the oracle's row
    # formatter, (c) the host numpy decode
    Xp, _ = utils.build_X(n, files, force_dim=None, grayscale=True)
    Y = m2.predict(Xp, batch_size=2)
    utils.setup_means_and_ranges([6, 6, 2, 8])
    Yp_dev, (ints, exists) = predict_spnet.decode_on_device(Y)
    out2 = str(tmp_path / "again.csv")
    utils.show_pred_ellipses(Yp_dev, Yp_dev, files, num_draw=n, log_dir=log_dir, out_csv=out2, show_true=False,
                             draw_images=False, decoded=(ints, exists))
    means, ranges = ns.setup_means_and_ranges([6, 6, 2, 8])[7:9]
    assert ns.pred_csv_text(ns.denorm_Y(Y, means, ranges), files) == open(out2).read()
    np.testing.assert_array_equal(Yp_dev, utils.denorm_Y(Y))
    hi, he = utils.decode_host(Yp_dev)
    np.testing.assert_array_equal(ints, hi)
    np.testing.assert_array_equal(exists, he)
    cf.model_type = "monolithic"


def test_fit_loop_reduces_loss_and_checkpoints(tmp_path):
    import spnet.config as cf
    from spnet import callbacks, models
    from spnet_b200 import fake_espi
    cf.model_type = "big"
    X, Y, _ = fake_espi.make_dataset(16, base_seed=7)
    model, serial = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0)
    sched = callbacks.OneCycleScheduler(lr_max=4e-5, n_data_points=16, epochs=12, batch_size=8)
    ck = callbacks.ParallelCheckpointCallback(model, filepath="weights.hdf5", save_every=5, dir=str(tmp_path))
    prog = callbacks.MyProgressCallback(X_val=X[:8], Y_val=Y[:8], log_dir=str(tmp_path), batch_size=8)
    hist = model.fit(X, Y, batch_size=8, epochs=12, shuffle=True, verbose=0, validation_data=(X[:8], Y[:8]),
                     callbacks=[prog, ck, sched])
    loss = hist.history["loss"]
    assert len(loss) == 12 and np.isfinite(loss).all() and min(loss[6:]) < loss[0], loss
    assert os.path.exists(tmp_path / "weights.hdf5") and os.path.exists(tmp_path / "losses.dat")
    assert len(prog.acc_hist) == 12 and all(a <= 100.0 for a in prog.acc_hist)  # calc_errors ran every epoch
    again, _ = models.setup_model(X, 576, try_checkpoint=True, no_cp_fatal=True, weights_file=str(tmp_path / "weights.hdf5"),
                                  freeze_fac=0.0)
    assert len(again.get_weights()) == len(model.get_weights())
    cf.model_type = "monolithic"


def test_mobilenet_backbone_through_the_model_surface(tmp_path):
    """cf.basemodel = 'MobileNet' (BASELINE configs[2]; reference plug-in point spnet/models.py:43-44,349-355):
    setup_model -> fit -> save_weights -> load -> predict through the same Keras-like surface."""
    import spnet.config as cf
    from spnet import models
    from spnet_b200 import fake_espi
    cf.model_type = "big"
    cf.basemodel = "MobileNet"
    try:
        X, Y, _ = fake_espi.make_dataset(16, base_seed=9)
        model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0)
        assert model.backbone == "MobileNet" and model.count_params() == 31541217
        hist = model.fit(X, Y, batch_size=8, epochs=8, shuffle=True, verbose=0)
        loss = hist.history["loss"]
        assert np.isfinite(loss).all() and min(loss[4:]) < loss[0], loss
        wpath = str(tmp_path / "mobilenet.hdf5")
        model.save_weights(wpath)
        again, _ = models.setup_model(X, 576, try_checkpoint=True, no_cp_fatal=True, weights_file=wpath, freeze_fac=0.0)
        y1 = model.predict(X[:8], batch_size=8)
        y2 = again.predict(X[:8], batch_size=8)
        assert y1.shape == (8, 576) and np.isfinite(y1).all()
        np.testing.assert_array_equal(y1, y2)  # same weights -> bit-identical predictions (fixed-order split-K head)
        cf.basemodel = "NASNetLarge"
        with pytest.raises((NotImplementedError, AttributeError)):
            models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0)
    finally:
        cf.basemodel = "Xception"
        cf.model_type = "monolithic"


def test_irv2_backbone_through_the_model_surface(tmp_path):
    """cf.basemodel = 'InceptionResNetV2' (BASELINE configs[3]; the reference's generic backbone branch,
    spnet/models.py:357-359) through setup_model -> fit -> predict at the native 384x512 input."""
    import spnet.config as cf
    from spnet import models
    from spnet_b200 import fake_espi
    cf.model_type = "big"
    cf.basemodel = "InceptionResNetV2"
    try:
        X, Y, _ = fake_espi.make_dataset(8, base_seed=11)
        model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0)
        assert model.backbone == "InceptionResNetV2" and model.count_params() == 75571201
        hist = model.fit(X, Y, batch_size=4, epochs=6, shuffle=True, verbose=0)
        loss = hist.history["loss"]
        assert np.isfinite(loss).all() and min(loss[3:]) < loss[0], loss
        y = model.predict(X[:4], batch_size=4)
        assert y.shape == (4, 576) and np.isfinite(y).all()
    finally:
        cf.basemodel = "Xception"
        cf.model_type = "monolithic"


def test_evaluate_cli_path(tmp_path, capsys):
    """evaluate_spnet.evaluate_network on a small Test/ directory (PNG + metadata CSV pairs): predict ->
    de-normalise -> mAP + error summary on the device -> hawley_spnet.csv."""
    from PIL import Image
    import spnet.config as cf
    from spnet import models
    import evaluate_spnet
    from spnet_b200 import fake_espi

    cf.model_type = "big"
    cf.compute_dtype = "bf16"
    test_dir = tmp_path / "Test"
    test_dir.mkdir()
    from spnet import utils
    n, i, seed = 4, 0, 300
    while i < n:
        img, rows = fake_espi.make_frame(seed)
        seed += 1
        try:
            utils.build_Y_from_rows([rows], pred_grid=[6, 6, 2])  # labels that overflow a grid cell are redrawn
        except AssertionError:
            continue
        Image.fromarray(img.reshape(img.shape[0], img.shape[1])).save(test_dir / ("steelpan_%07d.png" % i))
        with open(test_dir / ("steelpan_%07d.csv" % i), "w") as f:
            f.write("\n".join(",".join(str(v) for v in r) for r in rows))
        i += 1
    X = np.zeros((2, 384, 512, 1), np.float32)
    engine_model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0, quick_setup=True)
    _, Y_lab, _, _ = utils.build_dataset(path=str(test_dir) + "/", load_frac=1.0, set_means_ranges=True, batch_size=2,
                                         shuffle=False, pred_grid=[6, 6, 2])

    class NearTruth:
        """The engine's forward pass stays in the path, but a random-initialised network finds no ellipse at all
        (the reference's precision() then divides by zero, and so does ours): predictions = labels + a small
        bounded function of the network output."""

        def predict(self, Xb, batch_size=None):
            return (Y_lab + 0.05 * np.tanh(engine_model.predict(Xb, batch_size=batch_size))).astype(np.float32)

    log_dir = str(tmp_path / "out") + "/"
    evaluate_spnet.evaluate_network(model=NearTruth(), datapath=str(test_dir) + "/", fraction=1.0, log_dir=log_dir,
                                    batch_size=2, draw_images=False)
    out = capsys.readouterr().out
    assert "mAP = " in out and "Total Mistakes = " in out and "Mean pixel error =" in out
    last = evaluate_spnet.evaluate_network.last
    assert 0.5 <= last["mAP"] <= 1.0 and last["total_obj"] >= n  # every frame holds at least one antinode
    assert os.path.exists(log_dir + "hawley_spnet.csv")
    cf.model_type = "monolithic"



def test_predict_streaming_matches_whole_set(tmp_path):
    """predict_network(stream_chunk=...) = the same CSV TEXT as loading every frame first."""
    from PIL import Image
    import spnet.config as cf
    from spnet import models
    import predict_spnet
    from spnet_b200 import fake_espi

    cf.model_type = "big"
    cf.compute_dtype = "bf16"
    n = 6
    for i in range(n):
        img, _ = fake_espi.make_frame(500 + i)
        Image.fromarray(img).save(tmp_path / ("steelpan_%07d.png" % i))
    X = np.zeros((2, 384, 512, 1), np.float32)
    model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0, quick_setup=True)
    a_dir, b_dir = str(tmp_path / "whole") + "/", str(tmp_path / "stream") + "/"
    predict_spnet.predict_network(model=model, datapath=str(tmp_path), log_dir=a_dir, batch_size=2, draw_images=False)
    predict_spnet.predict_network(model=model, datapath=str(tmp_path), log_dir=b_dir, batch_size=2, draw_images=False,
                                  stream_chunk=4)
    # raw_u8 (default: frames cross PCIe as bytes, normalised on the device) against the reference's host-side
    # normalisation, and a second whole-set run: the forward pass has a fixed summation order everywhere, so the
    # three CSV files are IDENTICAL TEXT
    c_dir = str(tmp_path / "hostnorm") + "/"
    predict_spnet.predict_network(model=model, datapath=str(tmp_path), log_dir=c_dir, batch_size=2, draw_images=False,
                                  raw_u8=False)
    d_dir = str(tmp_path / "again") + "/"
    predict_spnet.predict_network(model=model, datapath=str(tmp_path), log_dir=d_dir, batch_size=3, draw_images=False)
    a = open(a_dir + "hawley_spnet.csv").read()
    cf.model_type = "monolithic"
    assert len(a.strip().splitlines()) >= n
    assert a == open(b_dir + "hawley_spnet.csv").read()
    assert a == open(c_dir + "hawley_spnet.csv").read()
    # a different batch size (6 = 3 x 2 = 2 x 3 frames; the reference loads a multiple of the batch size) changes nothing
    # either: inference-mode rows are independent of their batch mates
    assert a == open(d_dir + "hawley_spnet.csv").read()


def test_predict_graph_u8_pinned_and_device_inputs_agree():
    """SPNetModel.predict: numpy fp32, numpy uint8, pinned torch uint8 and CUDA-resident inputs give bit-identical
    outputs; ragged last batch; graph replay == eager forward."""
    import torch
    import spnet.config as cf
    from spnet import models
    from spnet_b200 import fake_espi
    cf.compute_dtype = "bf16"
    n = 7
    Xu, _, _ = fake_espi.make_frames_u8(n, base_seed=900)
    Xf = Xu.astype(np.float32)
    Xf = Xf / 255.0
    Xf -= 0.5
    Xf *= 2.0
    model, _ = models.setup_model(Xf[:2], 576, try_checkpoint=False, freeze_fac=0.0, quick_setup=True)
    y_f = model.predict(Xf, batch_size=3)
    y_u = model.predict(Xu, batch_size=3)
    y_p = model.predict(torch.from_numpy(Xu).pin_memory(), batch_size=3)
    y_d = model.predict(torch.from_numpy(Xu).cuda(), batch_size=3)
    np.testing.assert_array_equal(y_f, y_u)
    np.testing.assert_array_equal(y_f, y_p)
    np.testing.assert_array_equal(y_f, y_d)
    eng = model._engine(3, False)
    assert getattr(eng, "fwd_graph", None) is not None
    eng.load_batch(Xf[:3])
    eager = eng.forward(training=False).cpu().numpy()
    np.testing.assert_array_equal(eager, y_f[:3])


def test_predict_cli_sharded_over_ranks_writes_the_same_csv(tmp_path):
    """`torchrun --nproc-per-node 2 predict_spnet.py`: the sorted file list is cut into contiguous shards, every rank
    predicts its shard (no collective on the data path), rank 0 concatenates the per-rank CSV parts in rank order ->
    the same hawley_spnet.csv, byte for byte, as the single-process run (7 frames: uneven shards)."""
    import subprocess
    import sys
    from PIL import Image
    import spnet.config as cf
    from spnet import models
    from spnet_b200 import fake_espi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    data = tmp_path / "frames"
    data.mkdir()
    n = 7
    for i in range(n):
        img, _ = fake_espi.make_frame(700 + i)
        Image.fromarray(img).save(data / ("steelpan_%07d.png" % i))
    cf.model_type = "big"
    cf.compute_dtype = "bf16"
    X = np.zeros((2, 384, 512, 1), np.float32)
    model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0, quick_setup=True)
    wpath = str(tmp_path / "spnet.model")
    model.save(wpath)
    cf.model_type = "monolithic"
    common = ["-w", wpath, "-d", str(data), "-b", "1", "--no-png", "--model_type", "big", "--dtype", "bf16"]
    env = dict(os.environ, PYTHONPATH=root)
    one = subprocess.run([sys.executable, os.path.join(root, "predict_spnet.py"), "-l", str(tmp_path / "one") + "/"] + common,
                         capture_output=True, text=True, env=env, timeout=900)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29571", os.path.join(root, "predict_spnet.py"), "-l",
                          str(tmp_path / "two") + "/"] + common, capture_output=True, text=True, env=env, timeout=900)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    a = open(str(tmp_path / "one") + "/hawley_spnet.csv").read()
    b = open(str(tmp_path / "two") + "/hawley_spnet.csv").read()
    assert a == b and len(a.strip().splitlines()) >= n
    assert not [f for f in os.listdir(str(tmp_path / "two")) if "part" in f]


def test_predict_network_ranks_override_ignores_the_torchrun_environment(tmp_path, monkeypatch):
    """train_spnet.py's post-training prediction pass runs on rank 0 alone while WORLD_SIZE is still 2: with
    ranks=(0, 1) predict_network must predict every frame and not wait for the other rank's CSV part (it used to
    block for merge_csv_parts' ten-minute timeout)."""
    from PIL import Image
    import spnet.config as cf
    from spnet import models
    import predict_spnet
    from spnet_b200 import fake_espi
    cf.model_type, cf.compute_dtype = "big", "bf16"
    n = 5
    for i in range(n):
        img, _ = fake_espi.make_frame(900 + i)
        Image.fromarray(img).save(tmp_path / ("steelpan_%07d.png" % i))
    X = np.zeros((2, 384, 512, 1), np.float32)
    model, _ = models.setup_model(X, 576, try_checkpoint=False, freeze_fac=0.0, quick_setup=True)
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    log_dir = str(tmp_path / "out") + "/"
    predict_spnet.predict_network(weights_file="", datapath=str(tmp_path), fraction=1.0, log_dir=log_dir, batch_size=1,
                                  model=model, X_pred="", draw_images=False, ranks=(0, 1))
    text = open(log_dir + "hawley_spnet.csv").read()
    for i in range(n):
        assert ("steelpan_%07d.png" % i) in text
    assert not [f for f in os.listdir(log_dir) if "part" in f]
    cf.model_type = "monolithic"
