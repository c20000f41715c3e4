"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement of the device AugmentOnTheFly (spnet_b200/csrc/augment.cu). Two things are pinned here:

* the DISTRIBUTIONS are the reference's: spnet/augmentation.py:117-135 (cutout_inplace: np.random.randint(0,
  max_regions+1) regions, corner randint(0, dim-minsize), extents randint(minsize, maxsize), far corner
  min(corner+extent, dim-1), fill np.random.uniform(min, max), rectangles applied in order) and :159-180
  (salt_n_pepa_inplace: a fair coin, ceil(amount*size*svp) salt points at max(img), then
  ceil(amount*size*(1-svp)) pepper points at min(img), coordinates randint(0, dim-1)), called in that order by
  spnet/callbacks.py:322-336;
* the random DRAWS are not numpy's global stream (which depends on the order frames are visited) but a counter
  based generator keyed by (seed, frame, draw index): two rounds of the splitmix64 finaliser. The CUDA kernel must
  reproduce this file bit for bit (fill values: to one float32 rounding).

Parity unpinned against the reference itself: its draws come from np.random's global state, so only the
distributions can agree (tests/test_augment.py checks their moments).
"""
import numpy as np

M64 = (1 << 64) - 1
K_MAX_REGIONS = 16


def rnd64(seed, frame, draw):
    z = (seed + 0x9e3779b97f4a7c15 * (frame + 1) + 0xbf58476d1ce4e5b9 * (draw + 1)) & M64
    for _ in range(2):
        z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & M64
        z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & M64
        z ^= z >> 31
    return z


def rnd_int(r, lo, hi):
    span = hi - lo if hi > lo else 1
    return lo + (((r >> 32) * span) >> 32)


def rnd_unit(r):
    return np.float32(r >> 40) * np.float32(1.0 / 16777216.0)


def frame_plan(seed, frame, H, W, C, lo, hi, max_regions=6, minsize=11, maxsize=75):
    """Rectangles [(y0, y1, x0, x1, value)] of one frame (lo / hi: extrema of the pristine frame)."""
    n = rnd_int(rnd64(seed, frame, 0), 0, max_regions + 1)
    rects = []
    for r in range(n):
        y0 = rnd_int(rnd64(seed, frame, 1 + 5 * r), 0, H - minsize)
        x0 = rnd_int(rnd64(seed, frame, 2 + 5 * r), 0, W - minsize)
        eh = rnd_int(rnd64(seed, frame, 3 + 5 * r), minsize, maxsize)
        ew = rnd_int(rnd64(seed, frame, 4 + 5 * r), minsize, maxsize)
        u = rnd_unit(rnd64(seed, frame, 5 + 5 * r))
        val = np.float32(np.float64(lo) + np.float64(np.float32(hi) - np.float32(lo)) * np.float64(u))
        rects.append((y0, min(y0 + eh, H - 1), x0, min(x0 + ew, W - 1), val))
    return rects


def augment(x_orig, seed, max_regions=6, minsize=11, maxsize=75, sp_prob=0.5, sp_amount=0.004, salt_vs_pepper=0.2):
    """x_orig [n,H,W,C] float32 -> (augmented copy, per-frame info dicts)."""
    x = np.array(x_orig, dtype=np.float32, copy=True)
    n, H, W, C = x.shape
    info = []
    for f in range(n):
        img = x[f]
        rects = frame_plan(seed, f, H, W, C, img.min(), img.max(), max_regions, minsize, maxsize)
        for y0, y1, x0, x1, val in rects:
            img[y0:y1, x0:x1, :] = val
        d0 = 1 + 5 * K_MAX_REGIONS
        did_sp = bool(rnd_unit(rnd64(seed, f, d0)) < np.float32(sp_prob))
        if did_sp:
            salt, pepper = img.max(), img.min()
            n_salt = int(np.ceil(np.float32(sp_amount) * np.float32(img.size) * np.float32(salt_vs_pepper)))
            n_pepper = int(np.ceil(np.float32(sp_amount) * np.float32(img.size) * np.float32(1.0 - np.float32(salt_vs_pepper))))
            for i in range(n_salt):
                py = rnd_int(rnd64(seed, f, d0 + 1 + 2 * i), 0, H - 1)
                px = rnd_int(rnd64(seed, f, d0 + 2 + 2 * i), 0, W - 1)
                img[py, px, :] = salt
            d1 = d0 + 1 + 2 * n_salt
            for i in range(n_pepper):
                py = rnd_int(rnd64(seed, f, d1 + 2 * i), 0, H - 1)
                px = rnd_int(rnd64(seed, f, d1 + 1 + 2 * i), 0, W - 1)
                img[py, px, :] = pepper
        info.append(dict(regions=len(rects), salt_pepper=did_sp))
    return x, info
