"""Debug helper (not a test): per-block RMS relative error of the engine against the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import xception_torch as xt
from spnet_b200.selfcheck import make_case
from spnet_b200.engine import XceptionSPNetEngine

H, W, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (192, 256, 8)
training = True
w, x, yt = make_case(H, W, B, seed=7)
ref = xt.OracleSPNet(w, H, W, dtype=torch.float64)
with torch.no_grad():
    y_ref = ref.forward(x, training=training, taps=True).numpy()
taps = {k: v.permute(0, 2, 3, 1).numpy() for k, v in ref.taps.items()}
ref32 = xt.OracleSPNet(w, H, W, dtype=torch.float32)
with torch.no_grad():
    y32 = ref32.forward(x, training=training, taps=True).numpy()
taps32 = {k: v.permute(0, 2, 3, 1).numpy() for k, v in ref32.taps.items()}
def rms(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)
for dtype in ("fp32", "bf16"):
    eng = XceptionSPNetEngine(H, W, B, dtype=dtype, weights=w, dropout_rate=0.0)
    eng.load_batch(x, yt)
    eng.forward(training=training)
    torch.cuda.synchronize()
    got = {"stem": eng.d, "block1": eng.x2}
    for e in eng.entry: got["block%d" % e["blk"]] = e["out"]
    for i, o in enumerate(eng.mid_out): got["block%d" % (5 + i)] = o
    got["block13"] = eng.exit13["out"]
    got["block14"] = eng.feat.view(B, *eng.shapes["out13"], 2048)
    print(dtype, " ".join("%s:%.1e/%.1e" % (k, rms(got[k].float().cpu().numpy(), taps[k]), rms(taps32[k], taps[k])) for k in taps))
    print(dtype, "y rms %.2e  max-norm %.2e   (fp32 oracle vs fp64: %.1e)" % (rms(eng.y_pred.cpu().numpy(), y_ref), np.abs(eng.y_pred.cpu().numpy() - y_ref).max() / np.abs(y_ref).max(), rms(y32, y_ref)))
