"""Implicit-GEMM convolution timings (CUDA-graph replay, CUDA events). Xception block1_conv2 full size
and InceptionResNetV2 branch shapes at batch 64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops

dev = torch.device("cuda:0")
CASES = [(64, 95, 127, 32, 64, 3, 3, "valid"), (64, 21, 29, 64, 96, 3, 3, "same"), (64, 10, 14, 128, 160, 1, 7, "same"),
         (64, 4, 6, 192, 224, 1, 3, "same")]


def graph_us(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps


for B, H, W, Cin, Cout, KH, KW, pad in CASES:
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    wt = torch.randn(KH, KW, Cin, Cout, device=dev).bfloat16()
    OH, OW, pt, pl = (H, W, (KH - 1) // 2, (KW - 1) // 2) if pad == "same" else (H - KH + 1, W - KW + 1, 0, 0)
    y = torch.empty(B, OH, OW, Cout, device=dev, dtype=torch.bfloat16)
    gx = torch.empty_like(x)
    gw = torch.zeros(KH, KW, Cin, Cout, device=dev)
    st = torch.zeros(2 * Cout, device=dev, dtype=torch.float64)
    flop = 2.0 * B * OH * OW * KH * KW * Cin * Cout
    byt = 2.0 * (x.numel() + y.numel())
    for name, fn in (("fwd+stats", lambda: ops.conv_tc_fwd(x, wt, y, pt, pl, colstats=st)),
                     ("dgrad", lambda: ops.conv_tc_dgrad(y, wt, gx, pt, pl)),
                     ("wgrad", lambda: ops.conv_tc_wgrad(x, y, gw, pt, pl))):
        us = graph_us(fn)
        print("%-28s %-10s %8.1f us  %7.1f TFLOP/s  %6.0f GB/s" % (str((B, H, W, Cin, Cout, KH, KW)), name, us, flop / us * 1e-6, byt / us * 1e-3), flush=True)
