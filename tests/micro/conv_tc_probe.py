"""Where does the implicit-GEMM convolution's time go? Same-FLOP comparisons on the block1_conv2 shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from spnet_b200 import ops


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


dev = torch.device("cuda:0")
B, H, W = 64, 95, 127
OH, OW = 93, 125
M = B * OH * OW
which = sys.argv[1] if len(sys.argv) > 1 else "all"

def run(name, fn):
    print("%-44s %8.1f us" % (name, graph_us(fn)), flush=True)

for Cin, Cout in ((32, 64), (64, 64), (64, 128)):
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    wt = torch.randn(3, 3, Cin, Cout, device=dev).bfloat16()
    y = torch.empty(B, OH, OW, Cout, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * Cout, device=dev, dtype=torch.float64)
    if which in ("all", "conv"):
        run("conv fwd %d->%d stats" % (Cin, Cout), lambda: ops.conv_tc_fwd(x, wt, y, 0, 0, colstats=st))
        run("conv fwd %d->%d" % (Cin, Cout), lambda: ops.conv_tc_fwd(x, wt, y, 0, 0))
        w1 = torch.randn(1, 1, Cin, Cout, device=dev).bfloat16()
        y1 = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
        run("conv 1x1 %d->%d" % (Cin, Cout), lambda: ops.conv_tc_fwd(x, w1, y1, 0, 0))
    if which in ("all", "gemm"):
        for K in (9 * Cin, 9 * 64):
            col = torch.randn(M, K, device=dev).bfloat16()
            wl = torch.randn(K, Cout, device=dev).bfloat16()
            run("gemm %d x %d x %d stats" % (M, Cout, K), lambda: ops.gemm(col, False, wl, True, y.view(M, Cout), M, Cout, K, colstats=st))
            del col
