"""Pins the oracle (oracle/numpy_side.py, oracle/xception_torch.py) to the reference:
fixtures produced by the reference's OWN numpy code (oracle/make_goldens.py) and the constants
held by the reference's tests and run logs."""
import json
import os

import numpy as np
import pytest

from oracle import numpy_side as ns
from oracle import xception_torch as xt

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    a = np.load(os.path.join(G, "ref_numpy_side.npz"))
    t = json.load(open(os.path.join(G, "ref_text.json")))
    return a, t


def test_means_ranges_and_grid_constants(gold):
    a, _ = gold
    r = ns.setup_means_and_ranges([6, 6, 2, 8])
    assert tuple(r[:6]) == tuple(a["smr_scalars"]) == (40, 40, 470, 350, 71, 51)
    np.testing.assert_array_equal(r[6], a["smr_grid_defaults"])
    np.testing.assert_array_equal(r[7], a["means"])
    np.testing.assert_array_equal(r[8], a["ranges"])
    np.testing.assert_array_equal(r[7][:8], [75.5, 65.5, 35.5, 25.5, 0, 0, 0, 5])
    np.testing.assert_array_equal(r[8][:8], [71, 51, 71, 51, 2, 2, 1, 10])


def test_parse_meta_and_grid_assignment_bit_exact(gold):
    a, t = gold
    means, ranges = a["means"], a["ranges"]
    for i, rows in enumerate(t["meta_rows"]):
        parsed = ns.parse_meta_rows(rows)
        ref = t["parsed_meta"][i]
        assert len(parsed) == len(ref)
        if parsed:
            np.testing.assert_array_equal(np.array(parsed, float), np.array(ref, float))
        grid = ns.true_to_pred_grid(np.array(parsed), [6, 6, 2, 8]).flatten()
        np.testing.assert_array_equal(grid, a["grid_Y_raw"][i])
    Y = a["grid_Y_raw"]
    np.testing.assert_array_equal(ns.norm_Y(Y, means, ranges).astype(np.float32), a["grid_Y_norm"])
    np.testing.assert_array_equal(ns.denorm_Y(a["grid_Y_norm"], means, ranges), a["grid_Y_denorm"])
    Yb, shape = ns.build_Y_from_rows(t["meta_rows"])
    np.testing.assert_array_equal(Yb, a["grid_Y_norm"])
    assert t["overflow_asserts"] is True
    with pytest.raises(AssertionError):
        three = np.array([[100, 140, 30, 20, 1, 0, 0, 3], [101, 141, 30, 20, 1, 0, 0, 3], [102, 141, 30, 20, 1, 0, 0, 3]], float)
        ns.true_to_pred_grid(three, [6, 6, 2, 8])


def test_survey_probe_cases():
    # two antinodes at (100,140),(101,141) -> cell (0,1), slots 0 and 1
    g = ns.true_to_pred_grid(np.array([[100, 140, 30, 20, 1, 0, 0, 3], [101, 141, 30, 20, 1, 0, 0, 4]], float), [6, 6, 2, 8])
    assert g[0, 1, 0, 7] == 3 and g[0, 1, 1, 7] == 4
    rows = [[300, 200, 40, 60, 30, 5], [100, 140, 60, 30, 45, 3], [100, 140, 60, 30, 45, 3], [50, 50, 20, 10, 10, 0]]
    p = ns.parse_meta_rows(rows)
    assert len(p) == 2 and p[0][:4] == [100, 140, 60, 30] and p[1][:4] == [300, 200, 60, 40]
    np.testing.assert_allclose(p[1][4:6], [-0.5, -0.8660254], atol=1e-6)
    means, ranges = ns.setup_means_and_ranges([6, 6, 2, 8])[7:9]
    Y, _ = ns.build_Y_from_rows([rows])
    txt = ns.pred_csv_text(ns.denorm_Y(Y, means, ranges), ["a.png"])
    assert txt == "100,140,a.png,3.0,60,30,45.0\n300,200,a.png,5.0,60,40,120.0\n"


@pytest.mark.parametrize("lt", ["same", "hybrid"])
def test_my_loss_matches_reference(gold, lt):
    a, _ = gold
    total, parts = ns.my_loss(a["loss_y_true"], a["loss_y_pred"], lt)
    np.testing.assert_allclose(total, a["loss_total_" + lt], rtol=1e-6)
    np.testing.assert_allclose(parts, a["loss_parts_" + lt], rtol=1e-6)
    # gradient formula vs central differences of the reference-pinned loss
    rng = np.random.default_rng(0)
    yt, yp = a["loss_y_true"].astype(np.float64), a["loss_y_pred"].astype(np.float64)
    g = ns.my_loss_grad(yt, yp, lt)
    for _ in range(20):
        i, j = rng.integers(0, yt.shape[0]), rng.integers(0, 576)
        e = np.zeros_like(yp)
        e[i, j] = 1e-5
        fd = (ns.my_loss(yt, yp + e, lt)[0] - ns.my_loss(yt, yp - e, lt)[0]) / 2e-5
        np.testing.assert_allclose(g[i, j], fd, rtol=1e-5, atol=1e-10)


def test_decode_and_csv_match_reference(gold):
    a, t = gold
    Yp = a["csv_Yp_denorm"]
    assert ns.pred_csv_text(Yp, t["csv_files"]) == t["csv_text"]
    cl = [ns.cleanup_antinode_vars(Yp[j, an * 8:(an + 1) * 8]) for j in range(Yp.shape[0]) for an in range(72)]
    np.testing.assert_array_equal(np.array([[c[0], c[1], c[2], c[3], c[5]] for c in cl]), a["cleanup_ints"])
    np.testing.assert_array_equal(np.array([c[4] for c in cl], np.float32), a["cleanup_angle"])
    assert "0,0,steelpan_0000000.png,0,0,0,0\n" in t["csv_text"]


def test_lr_schedule(gold):
    a, _ = gold
    lrs = ns.get_1cycle_schedule(lr_max=4e-5, n_data_points=40000, epochs=100, batch_size=16)
    probe = np.array([len(lrs), lrs[0], lrs[2499], lrs[4999], lrs.max(), lrs[-1]])
    np.testing.assert_array_equal(probe, a["lrs_probe"])
    # values printed in the reference's run log (paper/run_logs/log_DatasetA_*.txt:207,229)
    assert "%.6e" % lrs[2499] == "2.879505e-06"
    assert "%.6e" % np.float32(lrs[4999]) == "4.159522e-06"
    assert len(lrs) == 250000 and abs(lrs[0] - 1.6e-6) < 1e-18 and abs(lrs[-1] - 1.6e-10) < 1e-20
    np.testing.assert_array_equal(ns.get_1cycle_schedule(1e-3, 64, 5, 8), a["lrs_small"])


def test_reference_test_constants(gold):
    a, _ = gold
    assert ns.nearest_multiple(720, 31) == 713 == int(a["nearest_multiple"])      # tests/test_utils.py:7
    s = ns.add_to_stack(None, 5)
    assert s == [5] and ns.add_to_stack(s, 5) == [5, 5]                           # tests/test_utils.py:10-14
    assert float(a["iou"]) == 0.44227983107795693                                 # tests/test_diagnostics.py:15
    x = np.random.default_rng(0).random((2, 16)) - 0.5                            # tests/test_selectivesigmoid.py
    y = ns.selective_sigmoid(x)
    assert sorted(set(np.argwhere(np.abs(y - x) > 1e-6)[:, 1].tolist())) == [6, 14]
    np.testing.assert_allclose(y[0, 6], 1 / (1 + np.exp(-x[0, 6])))


def test_network_structure_pins():
    # paper/run_logs/log_DatasetA_*.txt:94-101
    spec = xt.xception_spnet_spec(331, 331)
    assert xt.count_params(spec) == (50353481, 50298935, 54546)
    assert xt.keras_layer_count() == 144
    assert xt.feature_hw(331, 331) == (5, 5)
    l2 = sorted(l for l, w, s, t, r in spec if r)
    assert l2 == sorted(["conv2d_%d" % i for i in range(1, 8)] + ["block1_conv1", "block1_conv2", "FinalOutput"])  # log:98
    assert xt.count_params(xt.xception_spnet_spec(384, 512))[0] == 77485385
    w = xt.init_weights(xt.xception_spnet_spec(67, 67))
    m = xt.OracleSPNet(w, 67, 67)
    import torch
    with torch.no_grad():
        y = m.forward(np.zeros((1, 67, 67, 1), np.float32), taps=True)
    assert tuple(y.shape) == (1, 576)
    spec331 = dict(stem=(165, 165), block14=(5, 5))
    assert (331 // 2, 331 // 2) == spec331["stem"]
    # first-epoch loss scale (log:208, 0.2582 incl. L2): the Dense kernel alone contributes ~0.11 at glorot init
    wd = xt.init_weights(spec)["FinalOutput/kernel"]
    assert 0.09 < 1e-4 * float((wd.astype(np.float64) ** 2).sum()) < 0.13


def test_mobilenet_structure_pins():
    """keras.applications.mobilenet.MobileNet(alpha=1, include_top=False) @ Keras 2.1.3: 3,228,864 parameters
    (SURVEY.md section 2.2); the engine's spec and the oracle's spec must describe the same tensors."""
    from spnet_b200 import arch
    spec = xt.mobilenet_spnet_spec(384, 512)
    bb = [int(np.prod(s)) for l, w, s, t, r in spec if l.startswith(("conv1", "conv_dw", "conv_pw"))]
    assert sum(bb) == 3228864
    assert xt.mobilenet_feature_hw(384, 512) == (6, 8)
    eng = arch.mobilenet_param_spec(384, 512)
    assert [(k, tuple(s)) for k, s, _, _ in eng] == [(l + "/" + w, tuple(s)) for l, w, s, _, _ in spec]
    assert [k for k, _, _, r in eng if r] == [l + "/" + w for l, w, _, _, r in spec if r]
    assert arch.count_params(eng) == xt.count_params(spec)
    l2 = sorted(l for l, w, s, t, r in spec if r)
    assert l2 == sorted(["conv2d_1", "conv2d_2", "conv2d_3", "conv1", "FinalOutput"] + ["conv_pw_%d" % i for i in range(1, 14)])
    import torch
    m = xt.OracleMobileNetSPNet(xt.init_weights(xt.mobilenet_spnet_spec(64, 64)), 64, 64)
    with torch.no_grad():
        y = m.forward(np.zeros((1, 64, 64, 1), np.float32))
    assert tuple(y.shape) == (1, 576)
    assert arch.mobilenet_shape_walk(331, 331)["out13"] == xt.mobilenet_feature_hw(331, 331) == (6, 6)


def test_irv2_structure_pins():
    """keras.applications.InceptionResNetV2(include_top=False) @ Keras 2.1.3: 54,336,736 parameters (SURVEY.md
    section 2.2), backbone output (4,6,1536) at 384x512; the oracle's walk of the Keras construction code and
    the engine's layer program must produce the same tensors, names and order."""
    from spnet_b200 import irv2
    spec = xt.irv2_spnet_spec(384, 512)
    stem = ("conv2d_1", "conv2d_2", "conv2d_3", "batch_normalization_1", "batch_normalization_2", "batch_normalization_3", "FinalOutput")
    assert sum(int(np.prod(s)) for l, w, s, t, r in spec if l not in stem) == 54336736
    eng = irv2.param_spec(384, 512)
    assert [(k, tuple(s), t, r) for k, s, t, r in eng] == [(l + "/" + w, tuple(s), t, r) for l, w, s, t, r in spec]
    prog = irv2.build_program(192, 256)
    assert (prog.output.h, prog.output.w, prog.output.c) == (4, 6, 1536)
    assert sum(1 for o in prog.ops if o["kind"] == "conv_bn") == 204 and sum(1 for o in prog.ops if o["kind"] == "conv_bias") == 40
    assert xt.count_params(spec)[0] == 75571201


def test_custom_loss_equals_my_loss(gold):
    import torch
    a, _ = gold
    for lt in ("same", "hybrid"):
        v = xt.OracleSPNet.custom_loss(torch.tensor(a["loss_y_true"]), torch.tensor(a["loss_y_pred"]), lt)
        np.testing.assert_allclose(float(v), float(a["loss_total_" + lt]), rtol=2e-6)
