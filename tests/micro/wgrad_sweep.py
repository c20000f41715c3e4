import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
for name, M, N, K in [("mid wgrad", 728, 728, 12288), ("b4 wgrad", 728, 728, 49152), ("b14 wgrad", 1536, 2048, 3072), ("b3 wgrad", 256, 256, 189504)]:
    A = torch.randn(K, M, device=dev).to(torch.bfloat16)
    B = torch.randn(K, N, device=dev).to(torch.bfloat16)
    D = torch.zeros(M, N, device=dev, dtype=torch.float32)
    for splits in (0, 2, 4, 6, 8, 12, 16, 24):
        def run():
            ops.gemm(A, True, B, True, D, M, N, K, out_mode=ops.OUT_ATOMIC, splits=splits)
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print("%s splits=%2d  %6.1f us  %6.1f TFLOP/s" % (name, splits, us, 2.0 * M * N * K / us / 1e6), flush=True)
