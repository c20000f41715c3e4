"""The HBM-side GEMM shape of the entry flow (block2 pointwise, 744000 x 128 x 128, bf16 out + BN sums)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 744000, 128, 128
stats = len(sys.argv) < 2 or sys.argv[1] != "nostats"
A = torch.randn(M, K, device=dev).bfloat16()
B = torch.randn(K, N, device=dev).bfloat16()
D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(2 * N, device=dev, dtype=torch.float64) if stats else None
for _ in range(5):
    ops.gemm(A, False, B, True, D, M, N, K, colstats=cs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gemm(A, False, B, True, D, M, N, K, colstats=cs)
e1.record(); torch.cuda.synchronize()
print("744000x128x128 stats=%s: %.1f us" % (stats, e0.elapsed_time(e1) * 50))
