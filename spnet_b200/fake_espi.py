"""gen_fake_espi-style synthetic ESPI frames (the bench / test input recipe).

Restates the reference generator (gen_fake_espi.py:60-114 draw_waves/draw_rings, :147-207
draw_antinodes, :245-268 frame assembly) without bandpass_mixup (needs the author's private
images) and with one explicit seed per frame for all three RNGs (random, numpy, OpenCV).
Frames are (384, 512) uint8-valued; labels are rows (cx, cy, a, b, angle, rings)."""
import random

import numpy as np

IM_W, IM_H = 512, 384
MIN_LINE_WIDTH = 4


def _cv2():
    import cv2
    return cv2


def _ellipse(img, center, axes, angle, color, thickness):
    cv2 = _cv2()
    shift = 10
    c = (int(round(center[0] * 2 ** shift)), int(round(center[1] * 2 ** shift)))
    ax = (int(round(axes[0] * 2 ** shift)), int(round(axes[1] * 2 ** shift)))
    cv2.ellipse(img, c, ax, -angle, 0, 360, color, thickness, cv2.LINE_AA, shift)


def _waves(img):
    cv2 = _cv2()
    xs = np.arange(0, IM_W)
    amp = random.randint(10, 200)
    wavelength = random.randint(100, IM_W // 2)
    thickness = random.randint(15, 40)
    slope = 3 * (np.random.rand() - 0.5)
    spacing = random.randint(thickness + thickness * int(abs(1.5 * slope)), IM_H // 3)
    for j in range(60 + IM_H // spacing):
        y0 = j * spacing - IM_W * abs(slope)
        ys = (y0 + slope * xs + amp * np.cos(xs / wavelength)).astype(np.int32)
        pts = np.stack([xs.astype(np.int32), ys], 1)
        cv2.polylines(img, [pts], False, 0, thickness=thickness)


def _bbox(center, axes, angle):
    rad = np.radians(angle)
    dx = np.sqrt(axes[0] ** 2 * np.cos(rad) ** 2 + axes[1] ** 2 * np.sin(rad) ** 2)
    dy = np.sqrt(axes[0] ** 2 * np.sin(rad) ** 2 + axes[1] ** 2 * np.cos(rad) ** 2)
    return [center[0] - dx, center[1] - dy, center[0] + dx, center[1] + dy]


def _overlaps(a, b):
    return not (a[2] < b[0] or a[0] > b[2] or a[3] < b[1] or a[1] > b[3])


def _antinodes(img, n):
    boxes, rows = [], []
    for _ in range(n):
        axes = sorted((random.randint(15, int(IM_W / 3.5)), random.randint(15, int(IM_H / 3.5))), reverse=True)
        rings = random.randint(1, max(1, min(axes[1] // 8, 11)))
        if axes[1] / rings < MIN_LINE_WIDTH:
            rings = axes[1] // MIN_LINE_WIDTH
        center = (random.randint(axes[0], IM_W - axes[0]), random.randint(axes[1], IM_H - axes[1]))
        angle = random.randint(1, 179)
        box = _bbox(center, axes, angle)
        tries = 0
        while (any(_overlaps(box, b) for b in boxes) or box[0] < 0 or box[2] > IM_W or box[1] < 0 or box[3] > IM_H) and tries < 2000:
            tries += 1
            axes = sorted((random.randint(25, IM_W // 3), random.randint(25, IM_H // 3)), reverse=True)
            if axes[1] / rings < MIN_LINE_WIDTH:
                rings = axes[1] // MIN_LINE_WIDTH
            center = (random.randint(axes[0], IM_W - axes[0]), random.randint(axes[1], IM_H - axes[1]))
            angle = random.randint(1, 180)
            box = _bbox(center, axes, angle)
        if tries >= 2000:
            continue
        nwb = max(1, 2 * rings)
        thick = int(round(min(axes) / nwb))
        start = np.random.choice([0, 1])
        for j in range(nwb):
            color = 0 if (start + j) % 2 == 0 else 138
            _ellipse(img, center, [a * (j + 1) / (nwb + 1) for a in axes], angle, color, thick)
        rows.append([center[0], center[1], axes[0], axes[1], angle, rings])
        boxes.append(box)
    return rows


def make_frame(seed):
    """One frame: (img uint8 (384,512), rows [[cx,cy,a,b,angle,rings], ...])."""
    cv2 = _cv2()
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    cv2.setRNGSeed(int(seed % (2 ** 31)))
    img = np.full((IM_H, IM_W, 1), 128, np.uint8)
    _waves(img)
    rows = _antinodes(img, random.randint(1, 7))
    if np.random.random() <= 0.3:      # blur_inplace: consumes RNG draws, never changes the image
        random.choice([3, 7])          # (spnet/augmentation.py:66-71 drops GaussianBlur's result)
    noise = cv2.randn(np.zeros((IM_H, IM_W, 1), np.uint8), 40, 40)
    img = cv2.add(img, noise)
    mask = np.random.choice([0, 1], size=img.shape).astype(np.float32)
    return (img.reshape(IM_H, IM_W) * mask.reshape(IM_H, IM_W)).astype(np.uint8), rows


def _frame_job(args):
    seed, pred_grid = args
    from . import utils
    img, rows = make_frame(seed)
    try:
        y, _ = utils.build_Y_from_rows([rows], pred_grid=list(pred_grid))
    except AssertionError:   # labels overflow a grid cell (the reference asserts, spnet/utils.py:240): frame is redrawn
        return None
    return img, y[0], rows


def make_frames_u8(n, base_seed=0, pred_grid=(6, 6, 2), workers=1):
    """n frames as uint8 (n,384,512,1) pixel values, Y float32 (n,576) normalised targets, rows. Frame i is drawn from
    seed base_seed + i + (number of earlier frames that were redrawn), exactly as the sequential loop would; with
    workers > 1 the seeds are drawn in parallel processes (call BEFORE CUDA is initialised) and consumed in seed order."""
    X = np.zeros((n, IM_H, IM_W, 1), np.uint8)
    Ys, all_rows = [], []
    seed, i = base_seed, 0
    pool = None
    if workers > 1:
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(workers)
    try:
        while i < n:
            chunk = list(range(seed, seed + max(n - i, 1)))
            jobs = [(sd, tuple(pred_grid)) for sd in chunk]
            results = pool.map(_frame_job, jobs, chunksize=max(1, len(jobs) // (4 * workers))) if pool else map(_frame_job, jobs)
            for r in results:
                seed += 1
                if r is None:
                    continue
                if i < n:
                    X[i, :, :, 0] = r[0]
                    Ys.append(r[1])
                    all_rows.append(r[2])
                    i += 1
                else:
                    seed -= 1  # not consumed
                    break
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    return X, np.stack(Ys).astype(np.float32), all_rows


def make_dataset(n, base_seed=0, pred_grid=(6, 6, 2), workers=1):
    """n frames -> X float32 (n,384,512,1) in [-1,1], Y float32 (n,576) normalised targets, rows.
    Frames whose labels overflow a grid cell (the reference asserts, spnet/utils.py:240) are redrawn."""
    Xu, Y, rows = make_frames_u8(n, base_seed, pred_grid, workers)
    X = Xu.astype(np.float32)
    X = X / 255.0      # spnet/utils.py:340-342
    X -= 0.5
    X *= 2.0
    return X, Y, rows
