"""Debug aid (not a test): activation gradients around the first Inception-ResNet blocks, engine (fp32) vs fp64 oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xception_torch as xt
import test_parity_configs_gpu as T
from spnet_b200.engine import InceptionResNetV2SPNetEngine
H, W, B = 260, 330, 3
X, Y = T.frames(B, 6000)
X = np.ascontiguousarray(X[:, 62:322, 91:421, :])
w = T.perturbed_weights(xt.irv2_spnet_spec, H, W, 103)
ref = xt.OracleIRv2SPNet(w, H, W, dtype=torch.float64)
for k in ref.trainable:
    ref.p[k].grad = None
y = ref.forward(X, training=True, taps=True)
loss = ref.custom_loss(torch.as_tensor(Y, dtype=torch.float64), y) + ref.l2_term()
loss.backward()
eng = InceptionResNetV2SPNetEngine(H, W, B, dtype="fp32", weights=w, dropout_rate=0.0, deterministic=True)
eng.load_batch(X, Y)
eng.grad_hook = lambda e: None
eng.train_step(lr=1e-3)
torch.cuda.synchronize()
res_ops = [op for op in eng.prog.ops if op["kind"] == "residual"]
def rl2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())
for i, (op, (t, up, yy)) in enumerate(zip(res_ops[:4], ref.taps["residual"][:4])):
    nhwc = lambda v: v.permute(0, 2, 3, 1)
    gx_e, gu_e, gy_e = eng.grad[op["inputs"][0].idx], eng.grad[op["inputs"][1].idx], eng.grad[op["out"].idx]
    print("block", i + 1, "fwd x %.2e up %.2e y %.2e" % (rl2(eng.data[op["inputs"][0].idx], nhwc(t.detach())),
          rl2(eng.data[op["inputs"][1].idx] + eng.w[op["bias_name"]], nhwc(up.detach())), rl2(eng.data[op["out"].idx], nhwc(yy.detach()))),
          "| bwd gy %.2e gu %.2e gx %.2e" % (rl2(gy_e, nhwc(yy.grad)), rl2(gu_e, nhwc(up.grad)), rl2(gx_e, nhwc(t.grad))))
print("---- mask disagreements (engine y > 0 vs oracle y > 0)")
for i, (op, (t, up, yy)) in enumerate(zip(res_ops[:3], ref.taps["residual"][:3])):
    ye = eng.data[op["out"].idx].double().cpu()
    yo = yy.detach().permute(0, 2, 3, 1)
    pre_o = (t.detach() + op["scale"] * up.detach()).permute(0, 2, 3, 1)
    dis = (ye > 0) != (yo > 0)
    print("block", i + 1, "disagree", int(dis.sum()), "of", dis.numel(), "x==0 fraction %.3f" % float((t.detach() == 0).double().mean()),
          "oracle pre at disagreements:", pre_o[dis][:8].tolist(), "engine y there:", ye[dis][:8].tolist())
    xe = eng.data[op["inputs"][0].idx].double().cpu()
    xo = t.detach().permute(0, 2, 3, 1)
    print("      x zero-pattern disagreements:", int(((xe == 0) != (xo == 0)).sum()), " max|xe-xo| %.3e" % float((xe - xo).abs().max()))
    gy_e, gy_o = eng.grad[op["out"].idx].double().cpu(), yy.grad.permute(0, 2, 3, 1)
    gm_o = gy_o * (yo > 0)
    gu_e = eng.grad[op["inputs"][1].idx].double().cpu()
    d = (gu_e - op["scale"] * gm_o)
    idx = torch.nonzero(d.abs() > 10 * d.abs().mean())
    print("      gu outliers:", idx.shape[0], "top |d|", d.abs().flatten().topk(5).values.tolist(), "typ |gu| %.3e" % float(gu_e.abs().mean()))
    if idx.shape[0]:
        j = tuple(idx[0].tolist())
        print("      at", j, "gu_e", float(gu_e[j]), "scale*gy_o", float(op["scale"] * gy_o[j]), "yo", float(yo[j]), "ye", float(ye[j]), "gy_e", float(gy_e[j]))
