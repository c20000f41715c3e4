"""Times the pooling kernels at the four Xception shapes of the batch-64 384x512 step (entry blocks 2-4, exit block 13).
Rotating buffers larger than L2; CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops

dev = "cuda"
SHAPES = [(64, 93, 125, 128), (64, 47, 63, 256), (64, 24, 32, 728), (64, 12, 16, 1024)]
for (B, H, W, C) in SHAPES:
    OH, OW = (H + 1) // 2, (W + 1) // 2
    nrot = max(2, int(300e6 // (B * H * W * C * 2)) + 1)
    zs = [torch.randn(B, H, W, C, device=dev).bfloat16() for _ in range(nrot)]
    rs = [torch.randn(B, OH, OW, C, device=dev).bfloat16() for _ in range(nrot)]
    outs = [torch.empty(B, OH, OW, C, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    ams = [torch.empty(B, OH, OW, C, device=dev, dtype=torch.uint8) for _ in range(nrot)]
    gins = [torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    a = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    ra = torch.randn(C, device=dev); rb = torch.randn(C, device=dev)
    def fwd(i):
        ops.maxpool3s2_add_fwd(zs[i], a, b, rs[i], ra, rb, out=outs[i], argmax=ams[i])
    def bwd(i):
        ops.maxpool3s2_bwd(rs[i], ams[i], H, W, out=gins[i])
    for name, fn, nbytes in (("fwd", fwd, (B * H * W * C + 2 * B * OH * OW * C) * 2 + B * OH * OW * C),
                             ("bwd", bwd, (B * H * W * C + B * OH * OW * C) * 2 + B * OH * OW * C)):
        for i in range(nrot):
            fn(i)
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            for i in range(nrot):
                fn(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nrot)
        print("%s %s: %.1f us  %.0f GB/s (algorithmic %.0f MB)" % (name, (B, H, W, C), us, nbytes / us / 1e3, nbytes / 1e6), flush=True)
