"""Fixed vs per-k-block cost of the tcgen05 GEMM on the middle-flow shape: time as a function of K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")


def graph_us(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)


M, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 12288), (int(sys.argv[2]) if len(sys.argv) > 2 else 728)
for name, a_mn, b_mn, mode, stats in (("fwd+stats", False, True, ops.OUT_T, True), ("fwd", False, True, ops.OUT_T, False),
                                      ("dgrad", False, False, ops.OUT_T, False)):
    for K in ((64, 728) if len(sys.argv) > 1 else (64, 128, 256, 728, 1456, 2912)):
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
        D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        cs = torch.zeros(2 * N, device=dev, dtype=torch.float64) if stats else None
        us = graph_us(lambda: ops.gemm(A, a_mn, B, b_mn, D, M, N, K, out_mode=mode, colstats=cs))
        print("%-10s K=%5d  %7.1f us  %7.1f TFLOP/s" % (name, K, us, 2.0 * M * N * K / us / 1e6), flush=True)
# an empty kernel launch in the same graph harness, for the launch floor
x = torch.zeros(1024, device=dev)
print("fill 4 KB            %7.1f us" % graph_us(lambda: x.zero_()))
