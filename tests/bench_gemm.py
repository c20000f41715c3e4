"""GEMM micro-benchmark (not a test): CUDA-graph timed launches of the shapes the Xception-SPNet step uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
SHAPES = [  # (name, M, N, K, a_mn, b_mn, out_mode, splits, stats)
    ("mid fwd  ", 12288, 728, 728, False, True, ops.OUT_T, 1, True),
    ("mid dgrad", 12288, 728, 728, False, False, ops.OUT_T, 1, False),
    ("mid wgrad", 728, 728, 12288, True, True, ops.OUT_ATOMIC, 9, False),
    ("b4s2 fwd ", 49152, 728, 728, False, True, ops.OUT_T, 1, True),
    ("b2s2 fwd ", 744000, 128, 128, False, True, ops.OUT_T, 1, True),
    ("b14 fwd  ", 3072, 2048, 1536, False, True, ops.OUT_T, 1, True),
    ("sq 8192  ", 8192, 8192, 8192, False, False, ops.OUT_T, 1, False),
]
for name, M, N, K, a_mn, b_mn, mode, splits, stats in SHAPES:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if mode == ops.OUT_T else torch.float32)
    cs = torch.zeros(2 * N, device=dev, dtype=torch.float64) if stats else None
    def run():
        ops.gemm(A, a_mn, B, b_mn, D, M, N, K, out_mode=mode, splits=splits, colstats=cs)
    run(); torch.cuda.synchronize()
    reps = 20
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (2 * reps)
    print("%s M=%6d N=%5d K=%6d  %7.1f us  %7.1f TFLOP/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)
