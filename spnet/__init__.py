"""`spnet` — the reference's package name, served by the B200 implementation so that
`from spnet import models, utils, multi_gpu, callbacks, diagnostics` and `import spnet.config as cf`
(train_spnet.py:24-25, predict_spnet.py:33-34) keep working unchanged."""
import importlib
import sys

for _name in ("config", "utils", "models", "multi_gpu", "callbacks", "diagnostics"):
    _mod = importlib.import_module("spnet_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
