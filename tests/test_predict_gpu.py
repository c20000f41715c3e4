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
    # formatter, (c) the host numpy decode. (The Dense head accumulates split-K partials with fp32 atomics,
    # so two predict() calls may differ in the last bits: every path below consumes the SAME Y.)
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
        np.testing.assert_allclose(y1, y2, rtol=2e-2, atol=2e-2)  # split-K atomics reorder the Dense sums
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
    """predict_network(stream_chunk=...) = the same CSV as loading every frame first (up to the last bits of the
    Dense head's split-K accumulation, which can move a rounded integer by one)."""
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
    a = open(a_dir + "hawley_spnet.csv").read().strip().splitlines()
    b = open(b_dir + "hawley_spnet.csv").read().strip().splitlines()
    cf.model_type = "monolithic"
    assert len(a) == len(b) and a[0] == b[0]  # header + one row per detection (or per empty frame)
    for x, y in zip(a[1:], b[1:]):
        fx, fy = x.split(","), y.split(",")
        assert fx[2] == fy[2]  # file name
        for k in (0, 1, 4, 5):  # cx, cy, a, b: rounded integers
            assert abs(int(fx[k]) - int(fy[k])) <= 1
        for k in (3, 6):  # rings, angle: floats printed at full precision
            assert abs(float(fx[k]) - float(fy[k])) <= 1e-3 * max(1.0, abs(float(fx[k])))
