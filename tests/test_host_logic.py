"""CPU tests: the product's host-side codec (spnet_b200/utils.py, callbacks.py) against the
reference-generated goldens, the C-ABI library (loads, exports every declared symbol), the
Keras-like surface that needs no GPU, and the data-parallel host logic on gloo (world size 2)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "ref_numpy_side.npz")), json.load(open(os.path.join(G, "ref_text.json")))


def test_library_builds_loads_and_exports_every_symbol():
    import __graft_entry__ as ge
    lib_path = ge.build()
    from spnet_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 39
    dll = ctypes.CDLL(lib_path)
    for name in protos:
        assert hasattr(dll, name), name
    assert dll.spnet_version() >= 100
    # error convention: bad arguments -> negative code + message, never a crash (no GPU needed: checked before launch)
    dll.spnet_last_error.restype = ctypes.c_char_p
    rc = dll.spnet_yolo_ellipse_loss(None, None, 0, 0, 0, 0, None, None, None)
    assert rc == -1 and b"null pointer" in dll.spnet_last_error()


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spnet_b200 import models
    with pytest.raises(Exception):
        models.custom_loss(np.zeros((1, 576), np.float32), np.zeros((1, 576), np.float32))


def test_utils_match_reference_goldens(gold):
    a, t = gold
    from spnet import utils
    r = utils.setup_means_and_ranges([6, 6, 2, 8])
    assert tuple(r[:6]) == (40, 40, 470, 350, 71, 51)
    np.testing.assert_array_equal(r[6], a["smr_grid_defaults"])
    np.testing.assert_array_equal(utils.means, a["means"])
    np.testing.assert_array_equal(utils.ranges, a["ranges"])
    for i, rows in enumerate(t["meta_rows"]):
        parsed = utils.parse_meta_rows(rows)
        assert len(parsed) == len(t["parsed_meta"][i])
        if parsed:
            np.testing.assert_array_equal(np.array(parsed, float), np.array(t["parsed_meta"][i], float))
        np.testing.assert_array_equal(utils.true_to_pred_grid(np.array(parsed), [6, 6, 2, 8]).flatten(), a["grid_Y_raw"][i])
    Y, shape = utils.build_Y_from_rows(t["meta_rows"])
    np.testing.assert_array_equal(Y, a["grid_Y_norm"])
    np.testing.assert_array_equal(utils.denorm_Y(a["grid_Y_norm"]), a["grid_Y_denorm"])
    with pytest.raises(AssertionError):
        utils.true_to_pred_grid(np.array([[100, 140, 30, 20, 1, 0, 0, 3]] * 3, float), [6, 6, 2, 8])
    assert utils.nearest_multiple(720, 31) == 713
    assert utils.add_to_stack(utils.add_to_stack(None, 5), 5) == [5, 5]


def test_csv_writer_matches_reference_text(gold, tmp_path):
    a, t = gold
    from spnet import utils
    utils.setup_means_and_ranges([6, 6, 2, 8])
    Yp = a["csv_Yp_denorm"]
    out = tmp_path / "hawley_spnet.csv"
    utils.show_pred_ellipses(Yp, Yp, t["csv_files"], num_draw=Yp.shape[0], log_dir=str(tmp_path), out_csv=str(out),
                             show_true=False, draw_images=False)
    assert out.read_text() == t["csv_text"]
    ints, exists = utils.decode_host(Yp)
    np.testing.assert_array_equal(ints.reshape(-1, 5), a["cleanup_ints"])
    np.testing.assert_array_equal(utils._angles(Yp).reshape(-1), a["cleanup_angle"])
    for j in range(3):
        for an in range(72):
            c = utils.cleanup_antinode_vars(Yp[j, an * 8:(an + 1) * 8])
            assert (c[0], c[1], c[2], c[3], c[5]) == tuple(ints[j, an])


def test_meta_file_and_build_Y(tmp_path, gold):
    a, t = gold
    from spnet import utils
    metas, imgs = [], []
    for i, rows in enumerate(t["meta_rows"][:6]):
        p = tmp_path / ("steelpan_%07d.csv" % i)
        p.write_text("\n".join(",".join(str(v) for v in r) for r in rows))
        metas.append(str(p))
        imgs.append(str(p).replace(".csv", ".png"))
    Y, shape = utils.build_Y(6, metas, imgs)
    np.testing.assert_array_equal(Y, a["grid_Y_norm"][:6])
    assert list(shape) == [6, 6, 2, 8]


def test_build_X_normalisation(tmp_path):
    from PIL import Image
    from spnet import utils
    rng = np.random.default_rng(0)
    files = []
    for i in range(3):
        arr = rng.integers(0, 256, (384, 512), dtype=np.uint8)
        p = tmp_path / ("img%d.png" % i)
        Image.fromarray(arr).save(p)
        files.append((str(p), arr))
    X, dims = utils.build_X(3, [f for f, _ in files], force_dim=None, grayscale=True)
    assert X.shape == (3, 384, 512, 1) and X.dtype == np.float32
    for i, (_, arr) in enumerate(files):
        ref = arr.astype(np.float32) / 255.0
        ref -= 0.5
        ref *= 2.0
        np.testing.assert_array_equal(X[i, :, :, 0], ref)
    X331, _ = utils.build_X(2, [f for f, _ in files], force_dim=331, grayscale=True)
    assert X331.shape == (2, 331, 331, 1) and -1.0 <= X331.min() and X331.max() <= 1.0


def test_one_cycle_schedule_and_callback(gold):
    a, _ = gold
    from spnet import callbacks
    lrs = callbacks.get_1cycle_schedule(lr_max=4e-5, n_data_points=40000, epochs=100, batch_size=16)
    np.testing.assert_array_equal(np.array([len(lrs), lrs[0], lrs[2499], lrs[4999], lrs.max(), lrs[-1]]), a["lrs_probe"])
    np.testing.assert_array_equal(callbacks.get_1cycle_schedule(1e-3, 64, 5, 8), a["lrs_small"])

    class Opt:
        lr = 0.0

    class M:
        optimizer = Opt()
    cb = callbacks.OneCycleScheduler(lr_max=1e-3, n_data_points=64, epochs=5, batch_size=8)
    cb.set_model(M())
    seen = []
    for _ in range(40):
        cb.on_batch_begin(0)
        seen.append(cb.model.optimizer.lr)
    np.testing.assert_array_equal(np.array(seen), a["lrs_small"])
    logs = {}
    cb.on_epoch_end(0, logs)
    assert logs["lr"] == seen[-1]


def test_model_surface_without_gpu(tmp_path):
    from spnet import models
    import spnet.config as cf
    from spnet_b200 import arch
    X = np.zeros((2, 331, 331, 1), np.float32)
    model, serial = models.setup_model(X, 576, try_checkpoint=True, no_cp_fatal=False, weights_file=str(tmp_path / "none.hdf5"),
                                       freeze_fac=0.0)
    assert model is serial and model.count_params() == 50353481          # paper/run_logs/log_DatasetA_*.txt:99
    assert arch.count_params(model.spec) == (50353481, 50298935, 54546)
    assert sorted(k.split("/")[0] for k in model.losses) == sorted(
        ["conv2d_%d" % i for i in range(1, 8)] + ["block1_conv1", "block1_conv2", "FinalOutput"])   # log:98
    assert model.optimizer.lr == 0.00001 and model.loss is models.custom_loss
    assert arch.shape_walk(331, 331)["stem"] == (165, 165) and arch.shape_walk(331, 331)["out13"] == (5, 5)
    assert arch.count_params(arch.param_spec(384, 512))[0] == 77485385
    with pytest.raises(Exception, match="No weights file detected"):
        models.setup_model(X, 576, no_cp_fatal=True, weights_file=str(tmp_path / "none.hdf5"))
    with pytest.raises(ValueError):
        models.SPNetModel((331, 331, 1), Y0size=577)
    # weights round trip through the file format and get/set_weights
    small = models.SPNetModel((67, 67, 1), quick_setup=True)
    path = str(tmp_path / "w.hdf5")
    small.save_weights(path)
    other = models.SPNetModel((67, 67, 1), quick_setup=True, seed=99)
    other.load_weights(path)
    for x, y in zip(small.get_weights(), other.get_weights()):
        np.testing.assert_array_equal(x, y)
    other.set_weights(small.get_weights())
    with pytest.raises(ValueError):
        other.set_weights(small.get_weights()[:-1])
    small.save(str(tmp_path / "full.h5"))
    again = models.load_model(str(tmp_path / "full.h5"))
    np.testing.assert_array_equal(again.get_weights()[0], small.get_weights()[0])
    names = [l.name for l in small.layers]
    assert names[0] == "conv2d_1" and names[-1] == "FinalOutput" and "block8_sepconv2_bn" in names
    assert cf.vars_per_pred == 8 and cf.ind_noobj == 6


def test_fake_espi_generator_statistics():
    from spnet_b200 import fake_espi
    ref = np.load(os.path.join(G, "fake_espi_frames.npz"))["frames"]
    img, rows = fake_espi.make_frame(0)
    assert img.shape == (384, 512) and img.dtype == np.uint8 and 1 <= len(rows) <= 7
    # same recipe as the reference's drawing code: half the pixels dropped, comparable brightness
    assert abs((img == 0).mean() - (ref == 0).mean()) < 0.1
    assert abs(float(img.mean()) - float(ref.mean())) < 15
    X, Y, _ = fake_espi.make_dataset(3, base_seed=5)
    assert X.shape == (3, 384, 512, 1) and Y.shape == (3, 576) and X.min() >= -1 and X.max() <= 1
    X2, Y2, _ = fake_espi.make_dataset(3, base_seed=5)
    np.testing.assert_array_equal(X, X2)
    np.testing.assert_array_equal(Y, Y2)


DP_SCRIPT = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from collections import OrderedDict
from spnet_b200 import multi_gpu
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
assert multi_gpu.world() == (rank, world)
lo, hi = multi_gpu.batch_slice(10, rank, world)      # get_slice: size = shape[0] // parts, remainder dropped
assert (lo, hi) == (rank * 5, rank * 5 + 5)
class FakeEngine:
    device = "cpu"
    def __init__(self, tail_key):
        self.offsets = OrderedDict([("FinalOutput/kernel", (0, 20, (4, 5))), ("entry", (24, 8, (8,))), ("middle", (32, 16, (16,)))])
        self.tail_param_key = tail_key
        self.grads = torch.arange(48, dtype=torch.float32) * (rank + 1)
        self.scale = None
    def optimizer_step(self, grad_scale=1.0):
        self.scale = grad_scale
for tail_key in ("middle", None):        # with and without a tail bucket (Xception / MobileNet engines)
    eng = FakeEngine(tail_key)
    hook = multi_gpu.attach_data_parallel(eng)
    assert hook.buckets["head"].numel() == 24 and hook.buckets["tail"].numel() == (16 if tail_key else 0)
    assert hook.rest.numel() == (8 if tail_key else 24)
    hook.bucket_ready(eng, "head")       # on CPU the buckets are reduced in the final call
    hook.bucket_ready(eng, "tail")
    hook(eng)
    expect = torch.arange(48, dtype=torch.float32) * sum(r + 1 for r in range(world))
    assert torch.equal(eng.grads, expect), (eng.grads, expect)
    assert eng.scale == 1.0 / world
dist.barrier()
if rank == 0:
    print("DP_OK")
"""


def test_data_parallel_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(DP_SCRIPT % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert "DP_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_stream_X_yields_build_X_in_chunks(tmp_path):
    """utils.stream_X: same values as build_X (spnet/utils.py:325-421), chunked, next chunk decoded in the
    background; ragged last chunk; both the native-size grayscale path and the LANCZOS-resize path."""
    from PIL import Image
    from spnet_b200 import utils
    rng = np.random.RandomState(4)
    files = []
    for i in range(7):
        img = rng.randint(0, 256, (48, 64), dtype=np.uint8)
        f = str(tmp_path / ("f_%02d.png" % i))
        Image.fromarray(img).save(f)
        files.append(f)
    for force_dim, gray in ((None, True), (32, True), (32, False)):
        X, _ = utils.build_X(len(files), files, force_dim=force_dim, grayscale=gray)
        got, starts = [], []
        for lo, Xc in utils.stream_X(files, 3, force_dim=force_dim, grayscale=gray):
            starts.append(lo)
            got.append(Xc)
        assert starts == [0, 3, 6] and [g.shape[0] for g in got] == [3, 3, 1]
        np.testing.assert_array_equal(np.concatenate(got, axis=0), X)
    assert list(utils.stream_X([], 3)) == []


def test_every_exported_entry_point_is_documented():
    """INTEGRATION.md is the map from reference operations to C-ABI entry points: nothing the header exports may
    be missing from it (names may be written with {a,b} alternations)."""
    import re
    from spnet_b200._lib import parse_header
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    expanded = set(re.findall(r"spnet_\w+", text))
    for stem, alts, tail in re.findall(r"(spnet_\w*?)\{([^}]*)\}(\w*)", text):
        for a in alts.split(","):
            expanded.add(stem + a.strip() + tail)
    for stem in re.findall(r"(spnet_\w+_)\*", text):
        expanded.update(n for n in parse_header() if n.startswith(stem))
    missing = sorted(n for n in parse_header() if n not in expanded)
    assert not missing, missing
