"""Profiling driver (not a test): the kernels written in the second half of round 2 at the bench shapes, two launches each.
   ncu --set full --clock-control none --import-source on \
       -k regex:'gemm_tc_kernel|maxpool_add_fwd_bf16|maxpool_bwd|conv3c3_strip|conv3c3_wgrad_strip' \
       -o gpurun_out/r2b_kernels python tests/micro/r2b_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
bf = torch.bfloat16
B = 64
# (1) block1_conv2 as implicit GEMM with the filter-row taps folded into the channel axis: forward (+BN sums), weight gradient
a11 = torch.randn(B, 95, 127, 32, device=dev).to(bf)
w = (torch.randn(3, 3, 32, 64, device=dev) * 0.05).to(bf)
z12 = torch.empty(B, 93, 125, 64, device=dev, dtype=bf)
g12 = torch.randn(B, 93, 125, 64, device=dev).to(bf)
st = ops.stats_alloc(128, dev)
gw = torch.zeros(3, 3, 32, 64, device=dev)
for _ in range(2):
    ops.conv_tc_fwd_kwfold(a11, w, z12, colstats=st)
    ops.conv_tc_wgrad_kwfold(a11, g12, gw)
# (2) 64-wide GEMM tiles: data gradient into the 64-channel block 1 output
M = B * 93 * 125
gz = torch.randn(M, 128, device=dev).to(bf)
W2 = (torch.randn(64, 128, device=dev) * 0.05).to(bf)
dt = torch.empty(M, 64, device=dev, dtype=bf)
for _ in range(2):
    ops.gemm(gz, False, W2, False, dt, M, 64, 128)
# (3) the middle-flow GEMM (forward with BN sums, data gradient) in its end-of-round state
A = torch.randn(12288, 728, device=dev).to(bf)
Wm = (torch.randn(728, 728, device=dev) * 0.03).to(bf)
D = torch.empty(12288, 728, device=dev, dtype=bf)
st2 = ops.stats_alloc(2 * 728, dev)
for _ in range(2):
    ops.gemm(A, False, Wm, True, D, 12288, 728, 728, colstats=st2)
    ops.gemm(A, False, Wm, False, D, 12288, 728, 728)
# (4) max-pool + BN + residual add forward (bf16: maximum on raw pairs), backward on 2 x 2 input blocks
z = torch.randn(B, 93, 125, 128, device=dev).to(bf)
res = torch.randn(B, 47, 63, 128, device=dev).to(bf)
out = torch.empty_like(res)
am = torch.empty(B, 47, 63, 128, device=dev, dtype=torch.uint8)
a = torch.randn(128, device=dev); b = torch.randn(128, device=dev)
gin = torch.empty_like(z)
for _ in range(2):
    ops.maxpool3s2_add_fwd(z, a, b, res, a, b, out=out, argmax=am)
    ops.maxpool3s2_bwd(res, am, 93, 125, out=gin)
# (5) 3 -> 3 stem convolutions as strip kernels: forward (+BN sums), data gradient (+activation mask), weight gradient
x = torch.randn(B, 192, 256, 3, device=dev).to(bf)
y = torch.empty_like(x)
g = torch.randn(B, 192, 256, 3, device=dev).to(bf)
w33 = torch.randn(3, 3, 3, 3, device=dev) * 0.2
a3 = torch.rand(3, device=dev) + 0.5; b3 = torch.randn(3, device=dev) * 0.1
st3 = ops.stats_alloc(6, dev)
dw = torch.zeros(3, 3, 3, 3, device=dev)
for _ in range(2):
    ops.conv_small_fwd(1, x, w33, y, in_a=a3, in_b=b3, act=2, stats=st3)
    ops.conv_small_dgrad(1, g, w33, y, mask_z=x, mask_a=a3, mask_b=b3, act=2)
    ops.conv_small_wgrad(1, x, g, dw, in_a=a3, in_b=b3, act=2)
torch.cuda.synchronize()
print("done")
