"""Profiling driver (not a test): N eager training steps of Xception-SPNet at the bench shape.
Usage: python tests/prof_step.py [steps] [batch] [dtype]   — wrap in `ncu --profile-from-start off` for the launch list of one step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spnet_b200.engine import XceptionSPNetEngine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
eng = XceptionSPNetEngine(384, 512, B, dtype=dtype, seed=1)
rng = np.random.default_rng(0)
x = (rng.random((B, 384, 512, 1)) * 2 - 1).astype(np.float32)
y = (0.3 * rng.standard_normal((B, 576))).astype(np.float32)
y[:, 6::8] = (rng.random((B, 72)) > 0.8)
eng.load_batch(x, y)
# under `ncu --profile-from-start off` only the LAST step is captured (engine construction and the warm-up
# steps launch a few hundred fill / cast kernels of their own)
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    eng.train_step(4e-5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done, loss", float(eng.loss6[0]))
