import os, sys
sys.path.insert(0, '/root/repo')
import torch
from spnet_b200 import ops
dev = torch.device("cuda:0")
for dbg, use_stats in ((0, True), (0, False)):
    os.environ["SPNET_GEMM_DBG"] = str(dbg)
    for name, M, N, K, a_mn, b_mn, mode, splits, stats in [("b2s2 fwd", 744000, 128, 128, False, True, ops.OUT_T, 1, use_stats), ("b2s2 K64", 744000, 128, 64, False, True, ops.OUT_T, 1, use_stats), ("b3s2 fwd", 189504, 256, 256, False, True, ops.OUT_T, 1, use_stats)]:
        nb = 3
        As = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nb)]
        B = torch.randn(K, N, device=dev).to(torch.bfloat16)
        Ds = [torch.zeros(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
        cs = torch.zeros(2 * N, device=dev, dtype=torch.float64) if stats else None
        def run(i):
            ops.gemm(As[i % nb], a_mn, B, b_mn, Ds[i % nb], M, N, K, out_mode=mode, splits=splits, colstats=cs)
        run(0); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for r in range(12): run(r)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print("dbg=%d stats=%d %s %.1f us" % (dbg, int(use_stats), name, e0.elapsed_time(e1) * 1e3 / 12), flush=True)
