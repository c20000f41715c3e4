import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
B, H, W, C = 64, 12, 16, 728
nb = 8
xs = [torch.randn(B, H, W, C, device=dev).to(torch.bfloat16) for _ in range(nb)]
gs = [torch.randn(B, H, W, C, device=dev).to(torch.bfloat16) for _ in range(nb)]
os_ = [torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
k = torch.randn(3, 3, C, device=dev) * 0.3
dk = torch.zeros(3, 3, C, device=dev)
a = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.2
mean = torch.randn(C, device=dev) * 0.1; rstd = torch.rand(C, device=dev) + 0.5
stats = torch.zeros(2 * C, device=dev, dtype=torch.float64)
for i in range(nb):
    ops.dwconv3x3_bwd_fused(gs[i], xs[i], k, dk, in_a=a, in_b=b, relu=True, bn_mean=mean, bn_rstd=rstd, stats=stats, out=os_[i])
    ops.dwconv3x3_fwd(xs[i], k, a, b, True, out=os_[i])
torch.cuda.synchronize()
print("ok")
