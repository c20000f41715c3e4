"""ORACLE tooling — run ONCE in the authoring container (needs /root/reference and cv2):
    python oracle/make_goldens_diagnostics.py
Records what the reference's own spnet/diagnostics.py (calc_errors :13-60, compute_iou :85-120 on cv2-rasterised
anti-aliased ellipses, precision :125-150, calc_map :153-162) returns on seeded inputs, as
tests/golden/ref_diagnostics.npz. The inputs are the denormalised grid targets of the existing golden file and a
perturbed copy of them standing in for predictions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
sys.argv = ["make_goldens_diagnostics"]
import spnet.config as cf  # noqa: E402
from spnet import diagnostics  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_numpy_side.npz"))
    Yt = g["grid_Y_denorm"].astype(np.float32)[:16]
    rng = np.random.default_rng(77)
    Yp = Yt.copy()
    n, ncols = Yp.shape
    v = cf.vars_per_pred
    for s in range(ncols // v):
        c = s * v
        Yp[:, c + 0] += rng.normal(0, 4.0, n)           # cx
        Yp[:, c + 1] += rng.normal(0, 4.0, n)           # cy
        Yp[:, c + 2] *= rng.uniform(0.85, 1.15, n)      # a
        Yp[:, c + 3] *= rng.uniform(0.85, 1.15, n)      # b
        ang = np.arctan2(Yt[:, c + 5], Yt[:, c + 4]) + rng.normal(0, 0.25, n)
        Yp[:, c + 4], Yp[:, c + 5] = np.cos(ang), np.sin(ang)
        flip = rng.random(n) < 0.08
        Yp[:, c + 6] = np.where(flip, 1.0 - Yt[:, c + 6], Yt[:, c + 6]) + rng.normal(0, 0.1, n)
        Yp[:, c + 7] += rng.normal(0, 0.4, n)
    Yp = Yp.astype(np.float32)
    out = dict(Yp=Yp, Yt=Yt)
    r = diagnostics.calc_errors(Yp, Yt)
    out["calc_errors_counts"] = np.array(r[:7], dtype=np.int64)
    out["calc_errors_pix_err"] = np.asarray(r[7], dtype=np.float64)
    out["calc_errors_ipem"] = np.array(r[8])
    iou = np.zeros((n, ncols // v))
    for i in range(n):
        for s in range(ncols // v):
            iou[i, s] = diagnostics.compute_iou(Yp[i, s * v:(s + 1) * v], Yt[i, s * v:(s + 1) * v])
    out["iou"] = iou
    prec = [diagnostics.precision(Yp, Yt, thresh=t) for t in (0.5, 0.75, 0.9)]
    out["precision_050_075_090"] = np.array(prec, dtype=np.float64)  # rows: (prec, tp, fp, fn)
    out["map"] = np.array(diagnostics.calc_map(Yp, Yt))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_diagnostics.npz"), **out)
    print("objects:", int((Yt[:, 6::8] < 0.5).sum()), "counts:", out["calc_errors_counts"], "map:", out["map"],
          "iou>=0:", int((iou >= 0).sum()))


if __name__ == "__main__":
    main()
