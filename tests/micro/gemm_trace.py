"""Phase timeline of the tcgen05 GEMM on the middle-flow shape. Needs the diagnostic library:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DSPNET_GEMM_TRACE \
       spnet_b200/csrc/*.cu -o tests/micro/libspnet_trace.so
Cycle stamps (clock64, one SM's counter per CTA) relative to the CTA's entry; globaltimer for the spread of CTA starts."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from spnet_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tests", "micro", "libspnet_trace.so")
from spnet_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 12288, 728, 728
for name, a_mn, b_mn, stats in (("fwd+stats", False, True, True), ("dgrad", False, False, False)):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    cs = ops.stats_alloc(2 * N, dev) if stats else None
    big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for it in range(3):
        big.zero_()          # cold L2 like inside the step
        ops.gemm(A, a_mn, B, b_mn, D, M, N, K, out_mode=ops.OUT_T, colstats=cs)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (256 * 16))()
    assert _lib.lib()._dll.spnet_gemm_trace_read(buf) == 0
    t = np.array(buf, dtype=np.uint64).reshape(256, 16)[:148].astype(np.int64)
    rel = (t - t[:, :1]) / 1.965e3   # us at 1965 MHz
    g0 = (t[:, 8] - t[:, 8].min()) / 1e3
    g1 = (t[:, 12] - t[:, 8].min()) / 1e3
    names = {1: "setup done", 2: "first stage full (MMA warp)", 3: "unit0 MMAs issued", 4: "last unit MMAs issued",
             5: "unit0 accumulator ready (epi)", 6: "last unit accumulator ready", 7: "unit0 epilogue done",
             9: "last unit epilogue done", 10: "stores drained", 11: "final sync done"}
    print("==", name)
    for k in (1, 2, 3, 5, 7, 4, 6, 9, 10, 11):
        v = rel[:, k]
        print("  %-34s median %6.2f us   min %6.2f   max %6.2f" % (names[k], np.median(v), v.min(), v.max()))
    print("  CTA start spread (globaltimer): median %.2f us, max %.2f us; CTA end: median %.2f, max %.2f us after the first start" % (
        np.median(g0), g0.max(), np.median(g1), g1.max()))
