"""The middle-flow pointwise forward GEMM (12288 x 728 x 728, bf16 out + BN sums) a few times: ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 12288, 728, 728
A = torch.randn(M, K, device=dev).bfloat16()
B = torch.randn(K, N, device=dev).bfloat16()
D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(2 * N, device=dev, dtype=torch.float64)
for _ in range(6):
    ops.gemm(A, False, B, True, D, M, N, K, colstats=cs)
torch.cuda.synchronize()
print("ok")
