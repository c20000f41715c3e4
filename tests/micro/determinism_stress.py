"""Repeat one training step many times on fresh engines and report the largest run-to-run deviation of
the outputs and gradients: fp32 atomics reorder sums (1e-6-ish); anything larger is a race."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from spnet_b200.selfcheck import make_case
from spnet_b200 import engine as E

cases = [("MobileNet", E.MobileNetSPNetEngine, 131, 163, 3, "fp32", "hybrid", 15), ("MobileNet", E.MobileNetSPNetEngine, 150, 100, 3, "fp32", "same", 15),
         ("Xception", E.XceptionSPNetEngine, 131, 163, 3, "fp32", "hybrid", 5), ("Xception", E.XceptionSPNetEngine, 131, 163, 3, "bf16", "same", 5),
         ("MobileNet", E.MobileNetSPNetEngine, 131, 163, 3, "bf16", "same", 15)]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for bb, cls, H, W, B, dt, lt, seed in cases:
    w, x, yt = make_case(H, W, B, seed=seed, backbone=bb)
    ref_y = ref_g = None
    worst_y = worst_g = 0.0
    worst_key = None
    for it in range(N):
        # churn the allocator so buffers land on different addresses / stale contents
        junk = [torch.randn(int(np.random.randint(1, 50)) * 100000, device="cuda") for _ in range(int(np.random.randint(0, 4)))]
        eng = cls(H, W, B, dtype=dt, weights=w, dropout_rate=0.0, loss_type=lt)
        eng.load_batch(x, yt)
        eng.grad_hook = lambda e: None
        eng.train_step(lr=1e-3)
        torch.cuda.synchronize()
        y = eng.y_pred.double().cpu().numpy()
        g = {k: v.double().cpu().numpy().copy() for k, v in eng.g.items()}
        if ref_y is None:
            ref_y, ref_g = y, g
            continue
        worst_y = max(worst_y, float(np.abs(y - ref_y).max() / (np.abs(ref_y).max() + 1e-30)))
        for k in g:
            d = float(np.linalg.norm(g[k] - ref_g[k]) / (np.linalg.norm(ref_g[k]) + 1e-12))
            if d > worst_g:
                worst_g, worst_key = d, k
        del eng, junk
    print("%-10s %s %dx%dx%d %-6s runs=%d  worst dy=%.3e  worst dgrad=%.3e (%s)" % (bb, dt, H, W, B, lt, N, worst_y, worst_g, worst_key), flush=True)
