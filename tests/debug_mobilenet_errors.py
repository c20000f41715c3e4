"""Debug helper (not a test): per-block error of the MobileNet engine vs the fp64 oracle in training mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import xception_torch as xt
from spnet_b200.selfcheck import make_case
from spnet_b200.engine import MobileNetSPNetEngine
H, W, B = 192, 256, 8
w, x, yt = make_case(H, W, B, seed=17, backbone="MobileNet")
ref = xt.OracleMobileNetSPNet(w, H, W, dtype=torch.float64)
with torch.no_grad():
    y_ref = ref.forward(x, training=True, taps=True)
taps = {k: v.permute(0, 2, 3, 1).numpy() for k, v in ref.taps.items()}
for dtype in ("fp32", "bf16"):
    eng = MobileNetSPNetEngine(H, W, B, dtype=dtype, weights=w, dropout_rate=0.0)
    eng.load_batch(x, yt)
    eng.forward(training=True)
    torch.cuda.synchronize()
    def rms(a, b):
        return float(np.sqrt(((a - b) ** 2).mean()) / np.sqrt((b ** 2).mean()))
    print(dtype, "stem", rms(eng.d.float().cpu().numpy(), taps["stem"]))
    for b in eng.blocks:
        y = torch.clamp(b["zp"].float() * b["bn_pw"].a + b["bn_pw"].b, 0, 6).cpu().numpy()
        print(dtype, "block", b["i"], b["hw"], b["cin"], b["cout"], "%.4f" % rms(y, taps["block%d" % b["i"]]))
    print(dtype, "y", rms(eng.y_pred.cpu().numpy(), y_ref.numpy()))
