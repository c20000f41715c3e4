"""__graft_entry__.smoke(): one tiny Xception-SPNet training step on cuda:0 through the CUDA
kernels, checked against the CPU oracle (the only place outside tests/ and bench.py's
cpu_baseline that touches oracle/)."""
import os
import sys

import numpy as np


def make_case(H, W, B, seed=0, n_out=576, backbone="Xception"):
    """Seeded inputs shared by the oracle and the engine: weights with non-trivial BN state,
    gen_fake_espi-like images in [-1,1], YOLO-grid targets."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import xception_torch as xt
    rng = np.random.default_rng(seed)
    spec = {"MobileNet": xt.mobilenet_spnet_spec, "InceptionResNetV2": xt.irv2_spnet_spec}.get(backbone, xt.xception_spnet_spec)(H, W, n_out)
    w = xt.init_weights(spec, seed=seed + 1)
    for k in w:
        leaf = k.rsplit("/", 1)[1]
        if leaf == "gamma":
            w[k] = (1.0 + 0.2 * rng.standard_normal(w[k].shape)).astype(np.float32)
        elif leaf in ("beta", "moving_mean", "bias"):
            w[k] = (0.1 * rng.standard_normal(w[k].shape)).astype(np.float32)
        elif leaf == "moving_variance":
            w[k] = (0.5 + rng.random(w[k].shape)).astype(np.float32)
    x = (rng.random((B, H, W, 1)) * 2 - 1).astype(np.float32)
    yt = (0.3 * rng.standard_normal((B, n_out))).astype(np.float32)
    yt[:, 6::8] = (rng.random((B, n_out // 8)) > 0.8).astype(np.float32)
    return w, x, yt


def smoke():
    """bf16 first: the driver's launch list is capped, and the tcgen05 / TMA kernels are what it should show."""
    import torch
    from oracle import xception_torch as xt
    from .engine import XceptionSPNetEngine

    H, W, B = 96, 128, 4
    w, x, yt = make_case(H, W, B)
    plain = xt.OracleSPNet(w, H, W)
    total, data, y_ref, _ = plain.loss_and_grads(x, yt)
    stored = xt.OracleSPNetStored(w, H, W, storage="bf16")   # same network, bf16 where the engine stores bf16
    total_s, data_s, y_s, _ = stored.loss_and_grads(x, yt)
    for dtype, tol, ref_total in (("bf16", 5e-3, total_s), ("fp32", 1e-4, total)):
        eng = XceptionSPNetEngine(H, W, B, dtype=dtype, weights=w, dropout_rate=0.0)
        eng.load_batch(x, yt)
        loss6 = eng.train_step(lr=1e-5)
        torch.cuda.synchronize()
        got = float(loss6[0]) + float(eng.l2_out[0])
        rel = abs(got - ref_total) / abs(ref_total)
        print("smoke %s: loss %.6f (oracle%s %.6f) rel.err %.2e" % (dtype, got, " with bf16 storage" if dtype == "bf16" else "",
                                                                    ref_total, rel))
        assert rel < tol, "smoke(%s): loss mismatch %g vs oracle %g" % (dtype, got, ref_total)
        del eng
    print("smoke OK")
