"""ORACLE — test infrastructure only (never imported by the product path).

Integer restatement of what cv2.ellipse(img, center, axes, angle, 0, 360, 255, thickness=-1, lineType=LINE_AA,
shift=10) leaves NON-ZERO on a single-channel uint8 canvas — the mask the reference's IoU counts
(spnet/diagnostics.py:61-80 create_ellipse_image -> utils.draw_ellipse, spnet/utils.py:35-53; compute_iou :85-120
counts non-zero pixels of the AND / OR of two such canvases).

The arithmetic lives in the third-party dependency OpenCV (imgproc/src/drawing.cpp; the container has 4.13.0), not in
the reference tree; it is restated from its published algorithm:
  * ellipse():  angle, arc bounds -> cvRound; centre / axes are 1/1024-pixel integers shifted to 16.16 fixed point;
  * EllipseEx(): delta = 5 degrees for max axis >= 15 px (18 / 30 / 90 below), ellipse2Poly() with the float sine table
    of integer degrees, vertices rounded to 16.16, consecutive duplicates dropped;
  * FillConvexPoly(LINE_AA): every edge is drawn with LineAA(), then scanlines ymin..ymax are filled between the two
    edge walkers with the anti-aliased span rule  [ (xl + 0xFFFF) >> 16 , xr >> 16 ];
  * LineAA(): per step along the major axis THREE pixels across it get a blend with alpha = ep_corr * FilterTable >> 8;
    a blend with alpha >= 1 makes a zero pixel non-zero (255 * a + 127 >= 256), so the mask only needs alpha > 0.
This module is pinned against cv2 itself (tests/test_diagnostics.py: identical masks on random and edge-case
ellipses, including ones that leave the canvas) and is the checker for the device kernel csrc/diagnostics.cu.
"""
import math

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT

# sin(i degrees), i = 0..450: drawing.cpp's SinTable is a list of float LITERALS with seven decimals (0.0174524f, ...),
# i.e. float32(round(sin, 7)) - not float32(sin): the difference moves a vertex by ~1e-5 pixel, enough to flip a 16.16
# rounding and with it the odd boundary pixel
SIN_TABLE = np.array([round(math.sin(math.radians(i)), 7) for i in range(451)], np.float64).astype(np.float32)

SLOPE_CORR = [181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
              203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254]
# FilterTable as OpenCV 4.13 uses it; the second half (the two outer pixels of every step) was recovered from cv2.line
# itself: 7,000+ interior samples per entry, each entry the only value consistent with >= 99 % of them
FILTER = [168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
          254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
          158, 149, 140, 131, 122, 114, 105, 97, 89, 82, 75, 68, 62, 56, 50, 45,
          40, 36, 32, 28, 25, 22, 19, 16, 14, 12, 11, 9, 8, 7, 5, 5]


def cv_round(x):
    """cvRound: round half to even (lrint)."""
    return int(np.rint(x))


def _tdiv(a, b):
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def ellipse_poly(cx, cy, a, b, angle_deg, shift=10):
    """Vertices (16.16 fixed point) of the polygon cv2.ellipse fills; utils.draw_ellipse's arguments."""
    centre = (int(round(cx * 2 ** shift)), int(round(cy * 2 ** shift)))
    axes = (int(round(a * 2 ** shift)), int(round(b * 2 ** shift)))
    angle = cv_round(-angle_deg)     # draw_ellipse passes -angle
    ccx, ccy = centre[0] << (XY_SHIFT - shift), centre[1] << (XY_SHIFT - shift)
    aw, ah = abs(axes[0] << (XY_SHIFT - shift)), abs(axes[1] << (XY_SHIFT - shift))
    delta = (max(aw, ah) + (XY_ONE >> 1)) >> XY_SHIFT
    delta = 90 if delta < 3 else 30 if delta < 10 else 18 if delta < 15 else 5
    while angle < 0:
        angle += 360
    while angle > 360:
        angle -= 360
    alpha, beta = SIN_TABLE[450 - angle], SIN_TABLE[angle]      # cos, sin (float32)
    pts = []
    i = 0
    while i < 360 + delta:
        ang = min(i, 360)
        x = float(aw) * float(SIN_TABLE[450 - ang])
        y = float(ah) * float(SIN_TABLE[ang])
        pts.append((float(ccx) + x * float(alpha) - y * float(beta), float(ccy) + x * float(beta) + y * float(alpha)))
        i += delta
    out, prev = [], None
    for px, py in pts:
        qx = cv_round(px / XY_ONE) << XY_SHIFT
        qy = cv_round(py / XY_ONE) << XY_SHIFT
        qx += cv_round(px - qx)
        qy += cv_round(py - qy)
        if (qx, qy) != prev:
            out.append((qx, qy))
            prev = (qx, qy)
    if len(out) == 1:
        out = [(ccx, ccy), (ccx, ccy)]
    return out


def _clip_line(w, h, x1, y1, x2, y2):
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def line_aa(mask, p1, p2):
    """Pixels LineAA() makes non-zero between two 16.16 points (single-channel uint8 canvas, colour 255)."""
    H, W = mask.shape
    ok, x1, y1, x2, y2 = _clip_line(W << XY_SHIFT, H << XY_SHIFT, p1[0], p1[1], p2[0], p2[1])
    if not ok:
        return
    dx, dy = x2 - x1, y2 - y1
    j = -1 if dx < 0 else 0
    ax = (dx ^ j) - j
    i = -1 if dy < 0 else 0
    ay = (dy ^ i) - i
    steep = not (ax > ay)
    if not steep:
        dy = (dy ^ j) - j
        if j:
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = XY_ONE, _tdiv(dy << XY_SHIFT, ax | 1)
        x2 += XY_ONE
        ecount = (x2 >> XY_SHIFT) - (x1 >> XY_SHIFT)
        j = -(x1 & (XY_ONE - 1))
        y1 += ((y_step * j) >> XY_SHIFT) + (XY_ONE >> 1)
        slope = (y_step >> (XY_SHIFT - 5)) & 0x3f
        slope ^= 0x3f if y_step < 0 else 0
        i = (x1 >> (XY_SHIFT - 7)) & 0x78
        j = (x2 >> (XY_SHIFT - 7)) & 0x78
    else:
        dx = (dx ^ i) - i
        if i:
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = _tdiv(dx << XY_SHIFT, ay | 1), XY_ONE
        y2 += XY_ONE
        ecount = (y2 >> XY_SHIFT) - (y1 >> XY_SHIFT)
        j = -(y1 & (XY_ONE - 1))
        x1 += ((x_step * j) >> XY_SHIFT) + (XY_ONE >> 1)
        slope = (x_step >> (XY_SHIFT - 5)) & 0x3f
        slope ^= 0x3f if x_step < 0 else 0
        i = (y1 >> (XY_SHIFT - 7)) & 0x78
        j = (y2 >> (XY_SHIFT - 7)) & 0x78
    slope = 0x100 if (slope & 0x20) else SLOPE_CORR[slope]
    t0 = slope << 7
    t1 = ((0x78 - i) | 4) * slope
    t2 = (j | 4) * slope
    ep = [0] * 9
    ep[8] = slope
    ep[1] = ep[3] = ((((j - i) & 0x78) | 4) * slope >> 8) & 0x1ff
    ep[2] = (t1 >> 8) & 0x1ff
    ep[4] = ((((j - i) + 0x80) | 4) * slope >> 8) & 0x1ff
    ep[5] = ((t1 + t0) >> 8) & 0x1ff
    ep[6] = (t2 >> 8) & 0x1ff
    ep[7] = ((t2 + t0) >> 8) & 0x1ff
    scount = 0
    while ecount >= 0:
        ep_corr = ep[(((scount >= 2) + 1) & (scount | 2)) * 3 + (((ecount >= 2) + 1) & (ecount | 2))]
        if not steep:
            px, base = x1 >> XY_SHIFT, (y1 >> XY_SHIFT) - 1
            dist = (y1 >> (XY_SHIFT - 5)) & 31
        else:
            py, base = y1 >> XY_SHIFT, (x1 >> XY_SHIFT) - 1
            dist = (x1 >> (XY_SHIFT - 5)) & 31
        for k, f in enumerate((FILTER[dist + 32], FILTER[dist], FILTER[63 - dist])):
            a = ((ep_corr * f) >> 8) & 0xff
            if a:
                if not steep:
                    yy, xx = base + k, px
                else:
                    yy, xx = py, base + k
                if 0 <= yy < H and 0 <= xx < W:
                    if mask.dtype == np.uint8:   # ICV_PUT_POINT: two blend steps towards the colour (255)
                        c = int(mask[yy, xx])
                        c += ((255 - c) * a + 127) >> 8
                        c += ((255 - c) * a + 127) >> 8
                        mask[yy, xx] = c
                    else:
                        mask[yy, xx] = True
        if not steep:
            y1 += y_step
            x1 += XY_ONE
        else:
            x1 += x_step
            y1 += XY_ONE
        scount += 1
        ecount -= 1


def fill_convex_poly_aa(mask, v):
    """FillConvexPoly(..., LINE_AA, shift = XY_SHIFT) restricted to what becomes non-zero."""
    H, W = mask.shape
    n = len(v)
    delta = XY_ONE >> 1
    delta1, delta2 = XY_ONE - 1, 0
    p0 = v[n - 1]
    xmin = xmax = v[0][0]
    ymin = ymax = v[0][1]
    imin = 0
    for idx, p in enumerate(v):
        if p[1] < ymin:
            ymin, imin = p[1], idx
        ymax = max(ymax, p[1])
        xmax = max(xmax, p[0])
        xmin = min(xmin, p[0])
        line_aa(mask, p0, p)
        p0 = p
    xmin, xmax = (xmin + delta) >> XY_SHIFT, (xmax + delta) >> XY_SHIFT
    ymin, ymax = (ymin + delta) >> XY_SHIFT, (ymax + delta) >> XY_SHIFT
    if n < 3 or xmax < 0 or ymax < 0 or xmin >= W or ymin >= H:
        return
    ymax = min(ymax, H - 1)
    edge = [dict(idx=imin, di=1, x=-XY_ONE, dx=0, ye=ymin), dict(idx=imin, di=n - 1, x=-XY_ONE, dx=0, ye=ymin)]
    edges = n
    y = ymin
    while True:
        if y < ymax or y == ymin:
            for e in edge:
                if y >= e["ye"]:
                    idx0, di = e["idx"], e["di"]
                    idx = idx0 + di
                    if idx >= n:
                        idx -= n
                    while True:
                        edges -= 1
                        if edges < 0:
                            break
                        ty = (v[idx][1] + delta) >> XY_SHIFT
                        if ty > y:
                            xs, xe = v[idx0][0], v[idx][0]
                            e["ye"] = ty
                            e["dx"] = _tdiv((xe - xs) * 2 + (ty - y), 2 * (ty - y))
                            e["x"] = xs
                            e["idx"] = idx
                            break
                        idx0 = idx
                        idx += di
                        if idx >= n:
                            idx -= n
        if edges < 0:
            break
        if y >= 0:
            l, r = (edge[1], edge[0]) if edge[0]["x"] > edge[1]["x"] else (edge[0], edge[1])
            xx1 = (l["x"] + delta1) >> XY_SHIFT
            xx2 = (r["x"] + delta2) >> XY_SHIFT
            if xx2 >= 0 and xx1 < W:
                xx1 = max(xx1, 0)
                xx2 = min(xx2, W - 1)
                if xx2 >= xx1:
                    mask[y, xx1:xx2 + 1] = 255 if mask.dtype == np.uint8 else True
        edge[0]["x"] += edge[0]["dx"]
        edge[1]["x"] += edge[1]["dx"]
        y += 1
        if y > ymax:
            break


def ellipse_mask(cx, cy, a, b, angle_deg, nx=512, ny=384):
    """Boolean (ny, nx) mask == (canvas > 0) after utils.draw_ellipse(canvas, [cx, cy], [a, b], angle, thickness=-1)."""
    mask = np.zeros((ny, nx), bool)
    fill_convex_poly_aa(mask, ellipse_poly(cx, cy, a, b, angle_deg))
    return mask


def ellipse_image(cx, cy, a, b, angle_deg, nx=512, ny=384):
    """uint8 (ny, nx) canvas, pixel VALUES identical to utils.draw_ellipse(zeros, [cx, cy], [a, b], angle, thickness=-1,
    color=255): the reference's compute_iou ANDs / ORs the VALUES of two such canvases (cv2.bitwise_and of two partially
    covered edge pixels can be zero although both are non-zero), so the values matter, not only the mask."""
    img = np.zeros((ny, nx), np.uint8)
    fill_convex_poly_aa(img, ellipse_poly(cx, cy, a, b, angle_deg))
    return img
