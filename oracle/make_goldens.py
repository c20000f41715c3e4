"""ORACLE tooling — run ONCE in the authoring container (needs /root/reference):
    python oracle/make_goldens.py
Imports the reference's own numpy-side code under oracle/ref_shim.py and records its outputs on
seeded inputs as small fixtures under tests/golden/. The fixtures travel to the GPU box; the
reference does not."""
import json
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
sys.argv = ["make_goldens"]
import spnet.config as cf  # noqa: E402
from spnet import callbacks, diagnostics, models, utils  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
PRED_SHAPE = [6, 6, 2, 8]


def random_meta_rows(rng, n):
    rows = []
    for _ in range(n):
        a, b = rng.integers(15, 146), rng.integers(15, 109)
        rows.append([int(rng.integers(0, 512)), int(rng.integers(0, 384)), int(a), int(b), int(rng.integers(1, 180)),
                     int(rng.integers(0, 12))])
    return rows


def main():
    arrays, texts = {}, {}
    r = utils.setup_means_and_ranges(PRED_SHAPE)
    arrays["smr_scalars"] = np.array(r[:6])
    arrays["smr_grid_defaults"] = r[6]
    arrays["means"] = np.array(utils.means)
    arrays["ranges"] = np.array(utils.ranges)

    # parse_meta_file + true_to_pred_grid + norm_Y on seeded metadata files (ragged, dups, swaps, rings==0)
    rng = np.random.default_rng(2024)
    metas, grids, parsed = [], [], []
    tmp = tempfile.mkdtemp()
    tries = 0
    while len(metas) < 24:
        tries += 1
        n = int(rng.integers(0, 8)) if len(metas) else 0
        rows = random_meta_rows(rng, n)
        if n >= 2 and len(metas) % 3 == 0:
            rows.append(list(rows[0]))  # duplicate row
        path = os.path.join(tmp, "m%d.csv" % len(metas))
        with open(path, "w") as f:
            f.write("\n".join(",".join(str(v) for v in row) for row in rows))
        if n == 0:
            arrs = []
        else:
            arrs = utils.parse_meta_file(path)
        try:
            grid = utils.true_to_pred_grid(np.array(arrs), PRED_SHAPE) if len(arrs) else utils.setup_means_and_ranges(PRED_SHAPE)[6]
        except AssertionError:
            continue  # >2 per cell: the reference asserts (spnet/utils.py:240); such files are regenerated
        metas.append(rows)
        parsed.append([[float(v) for v in a] for a in arrs])
        grids.append(grid.flatten())
    Y = np.array(grids, dtype=np.float32)
    arrays["grid_Y_raw"] = Y
    arrays["grid_Y_norm"] = utils.norm_Y(Y).astype(np.float32)
    arrays["grid_Y_denorm"] = utils.denorm_Y(utils.norm_Y(Y)).astype(np.float32)
    texts["meta_rows"] = metas
    texts["parsed_meta"] = parsed
    # cell-overflow case: three antinodes in one cell must assert
    three = np.array([[100, 140, 30, 20, 1, 0, 0, 3], [101, 141, 30, 20, 1, 0, 0, 3], [102, 141, 30, 20, 1, 0, 0, 3]], float)
    try:
        utils.true_to_pred_grid(three, PRED_SHAPE)
        texts["overflow_asserts"] = False
    except AssertionError:
        texts["overflow_asserts"] = True

    # my_loss, both loss types
    B = 16
    yt = arrays["grid_Y_norm"][:B].copy()
    yp = (yt + 0.15 * rng.standard_normal(yt.shape)).astype(np.float32)
    arrays["loss_y_true"], arrays["loss_y_pred"] = yt, yp
    for lt in ("same", "hybrid"):
        cf.loss_type = lt
        total, parts = models.my_loss(yt, yp)
        arrays["loss_total_" + lt] = np.array(total)
        arrays["loss_parts_" + lt] = np.array(parts)
    cf.loss_type = "same"

    # decode + CSV through the reference's show_pred_ellipses
    import PIL.Image
    utils.load_img = lambda f, **k: PIL.Image.open(f).convert("RGB")
    utils.img_to_array = lambda im: np.asarray(im, dtype=np.float32)
    files = []
    for i in range(B):
        p = os.path.join(tmp, "steelpan_%07d.png" % i)
        PIL.Image.fromarray(np.full((384, 512), 128, np.uint8)).save(p)
        files.append(p)
    Yp = utils.denorm_Y(yp)
    Yp[0, 6::8] = 1.0  # an image with no detections
    arrays["csv_Yp_denorm"] = Yp.astype(np.float32)
    csv_path = os.path.join(tmp, "out.csv")
    utils.show_pred_ellipses(Yp, Yp, files, num_draw=B, log_dir=tmp, out_csv=csv_path, show_true=False)
    texts["csv_text"] = open(csv_path).read()
    texts["csv_files"] = [os.path.basename(f) for f in files]
    cl = [utils.cleanup_antinode_vars(Yp[j, an * 8:(an + 1) * 8]) for j in range(B) for an in range(72)]
    arrays["cleanup_ints"] = np.array([[c[0], c[1], c[2], c[3], c[5]] for c in cl], dtype=np.int64)
    arrays["cleanup_angle"] = np.array([c[4] for c in cl], dtype=np.float32)

    # LR schedule (paper/run_logs/log_DatasetA_*.txt:207,229)
    lrs = callbacks.get_1cycle_schedule(lr_max=4e-5, n_data_points=40000, epochs=100, batch_size=16)
    arrays["lrs_probe"] = np.array([len(lrs), lrs[0], lrs[2499], lrs[4999], lrs.max(), lrs[-1]])
    arrays["lrs_small"] = callbacks.get_1cycle_schedule(lr_max=1e-3, n_data_points=64, epochs=5, batch_size=8)

    # IoU constant of the reference's own test (tests/test_diagnostics.py:11-15), 8-tuple form
    t8 = (100, 140, 120, 60, np.cos(2 * np.deg2rad(90)), np.sin(2 * np.deg2rad(90)), 0, 10.3)
    p8 = (120, 123, 120, 60, np.cos(2 * np.deg2rad(149.97)), np.sin(2 * np.deg2rad(149.97)), 0, 7.8)
    arrays["iou"] = np.array(diagnostics.compute_iou(p8, t8))
    arrays["nearest_multiple"] = np.array(utils.nearest_multiple(720, 31))

    # two synthetic frames from the reference's own drawing code (bandpass_mixup left out: it
    # needs the author's private images), seeds as SURVEY.md §8(d)
    import cv2
    import gen_fake_espi as gfe
    frames, caps = [], []
    for i in range(2):
        random.seed(i)
        np.random.seed(i)
        cv2.setRNGSeed(i)
        img = 128 * np.ones((gfe.imHeight, gfe.imWidth, 1), np.uint8)
        gfe.draw_waves(img)
        img, caption = gfe.draw_antinodes(img, num_antinodes=random.randint(1, 7))
        gfe.blur_inplace(img)
        noise = cv2.randn(np.zeros((gfe.imHeight, gfe.imWidth, 1), np.uint8), 40, 40)
        img = cv2.add(img, noise)
        mask = np.random.choice([0, 1], size=img.shape).astype(np.float32)
        img = img * mask
        frames.append(img.astype(np.uint8).reshape(gfe.imHeight, gfe.imWidth))
        caps.append(caption)
    np.savez_compressed(os.path.join(OUT, "fake_espi_frames.npz"), frames=np.array(frames))
    texts["fake_espi_captions"] = caps

    np.savez_compressed(os.path.join(OUT, "ref_numpy_side.npz"), **arrays)
    with open(os.path.join(OUT, "ref_text.json"), "w") as f:
        json.dump(texts, f, indent=0)
    print("wrote", OUT, "tries", tries)


if __name__ == "__main__":
    main()
