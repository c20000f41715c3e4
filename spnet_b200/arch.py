"""Static description of the Xception-SPNet graph: parameter names/shapes in Keras layouts and
the activation shape walk. Mirrors what the reference builds in create_model_functional
(spnet/models.py:302-424) around keras.applications.Xception (Keras 2.1.3).

Layer names follow Keras auto-naming so weights can be exchanged by name: the three stem
convolutions are conv2d_1..3, the four residual 1x1 convolutions conv2d_4..7 (confirmed by the
L2 list in paper/run_logs/log_DatasetA_*.txt:98), the Dense head is 'FinalOutput'.
"""
import math
from collections import OrderedDict

import numpy as np

BN_EPS = 1e-3          # keras.applications.Xception uses BatchNormalization defaults
BN_MOMENTUM = 0.99
L2_COEF = 1e-4         # regularizers.l2(0.0001), spnet/models.py:47
DROPOUT_RATE = 0.1     # spnet/models.py:338
ENTRY_BLOCKS = ((2, 64, 128), (3, 128, 256), (4, 256, 728))
MIDDLE_BLOCKS = tuple(range(5, 13))


def same_out(n):
    return (n + 1) // 2


def shape_walk(H, W):
    """Spatial sizes: dict stage -> (h, w)."""
    s = OrderedDict()
    s["input"] = (H, W)
    s["stem"] = (H // 2, W // 2)
    h, w = s["stem"]
    s["b1c1"] = ((h - 3) // 2 + 1, (w - 3) // 2 + 1)
    h, w = s["b1c1"]
    s["b1c2"] = (h - 2, w - 2)
    h, w = s["b1c2"]
    for blk in (2, 3, 4):
        s["in%d" % blk] = (h, w)
        h, w = same_out(h), same_out(w)
        s["out%d" % blk] = (h, w)
    s["middle"] = (h, w)
    s["in13"] = (h, w)
    h, w = same_out(h), same_out(w)
    s["out13"] = (h, w)
    return s


def param_spec(H, W, n_out=576):
    """[(key, shape, trainable, l2_regularised)] — keys are '<keras layer>/<weight>'."""
    spec = []

    def conv(name, kh, kw, cin, cout):
        spec.append((name + "/kernel", (kh, kw, cin, cout), True, True))

    def bn(name, c):
        spec.append((name + "/gamma", (c,), True, False))
        spec.append((name + "/beta", (c,), True, False))
        spec.append((name + "/moving_mean", (c,), False, False))
        spec.append((name + "/moving_variance", (c,), False, False))

    def sep(name, cin, cout):
        spec.append((name + "/depthwise_kernel", (3, 3, cin, 1), True, False))
        spec.append((name + "/pointwise_kernel", (1, 1, cin, cout), True, False))
        bn(name + "_bn", cout)

    for i, cin in ((1, 1), (2, 3), (3, 3)):
        conv("conv2d_%d" % i, 3, 3, cin, 3)
        bn("batch_normalization_%d" % i, 3)
    conv("block1_conv1", 3, 3, 3, 32)
    bn("block1_conv1_bn", 32)
    conv("block1_conv2", 3, 3, 32, 64)
    bn("block1_conv2_bn", 64)
    for n, (blk, cin, c) in enumerate(ENTRY_BLOCKS):
        conv("conv2d_%d" % (4 + n), 1, 1, cin, c)
        bn("batch_normalization_%d" % (4 + n), c)
        sep("block%d_sepconv1" % blk, cin, c)
        sep("block%d_sepconv2" % blk, c, c)
    for blk in MIDDLE_BLOCKS:
        for j in (1, 2, 3):
            sep("block%d_sepconv%d" % (blk, j), 728, 728)
    conv("conv2d_7", 1, 1, 728, 1024)
    bn("batch_normalization_7", 1024)
    sep("block13_sepconv1", 728, 728)
    sep("block13_sepconv2", 728, 1024)
    sep("block14_sepconv1", 1024, 1536)
    sep("block14_sepconv2", 1536, 2048)
    fh, fw = shape_walk(H, W)["out13"]
    spec.append(("FinalOutput/kernel", (fh * fw * 2048, n_out), True, True))
    spec.append(("FinalOutput/bias", (n_out,), True, False))
    return spec


# keras.applications.mobilenet.MobileNet (alpha=1, depth_multiplier=1) @ Keras 2.1.3: (cin, cout, stride)
MOBILENET_BLOCKS = ((32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2),
                    (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 1024, 2),
                    (1024, 1024, 1))


def mobilenet_shape_walk(H, W):
    """Spatial sizes; every stride-2 stage is 'same': out = ceil(n / 2)."""
    s = OrderedDict()
    s["input"] = (H, W)
    s["stem"] = (H // 2, W // 2)
    h, w = s["stem"]
    h, w = same_out(h), same_out(w)
    s["conv1"] = (h, w)
    for i, (_, _, stride) in enumerate(MOBILENET_BLOCKS, start=1):
        s["in%d" % i] = (h, w)
        if stride == 2:
            h, w = same_out(h), same_out(w)
        s["out%d" % i] = (h, w)
    return s


def mobilenet_param_spec(H, W, n_out=576):
    """MobileNet-SPNet: [(key, shape, trainable, l2_regularised)]. DepthwiseConv2D drops
    kernel_regularizer from its config, so (as for SeparableConv2D) only the dense convolutions
    and the head are L2-regularised after add_regularization's JSON round trip (spnet/models.py:47-71)."""
    spec = []

    def bn(name, c):
        spec.append((name + "/gamma", (c,), True, False))
        spec.append((name + "/beta", (c,), True, False))
        spec.append((name + "/moving_mean", (c,), False, False))
        spec.append((name + "/moving_variance", (c,), False, False))

    for i, cin in ((1, 1), (2, 3), (3, 3)):
        spec.append(("conv2d_%d/kernel" % i, (3, 3, cin, 3), True, True))
        bn("batch_normalization_%d" % i, 3)
    spec.append(("conv1/kernel", (3, 3, 3, 32), True, True))
    bn("conv1_bn", 32)
    for i, (cin, cout, _) in enumerate(MOBILENET_BLOCKS, start=1):
        spec.append(("conv_dw_%d/depthwise_kernel" % i, (3, 3, cin, 1), True, False))
        bn("conv_dw_%d_bn" % i, cin)
        spec.append(("conv_pw_%d/kernel" % i, (1, 1, cin, cout), True, True))
        bn("conv_pw_%d_bn" % i, cout)
    fh, fw = mobilenet_shape_walk(H, W)["out13"]
    spec.append(("FinalOutput/kernel", (fh * fw * 1024, n_out), True, True))
    spec.append(("FinalOutput/bias", (n_out,), True, False))
    return spec


def count_params(spec):
    tr = sum(int(np.prod(s)) for _, s, t, _ in spec if t)
    nt = sum(int(np.prod(s)) for _, s, t, _ in spec if not t)
    return tr + nt, tr, nt


def glorot_init(spec, seed=1):
    """Keras default initialisers (glorot_uniform kernels, zeros bias, BN 1/0/0/1)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for key, shape, _, _ in spec:
        leaf = key.rsplit("/", 1)[1]
        if leaf.endswith("kernel"):
            if len(shape) == 2:
                fan_in, fan_out = shape
            else:
                rf = shape[0] * shape[1]
                fan_in, fan_out = shape[2] * rf, shape[3] * rf
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            out[key] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif leaf in ("gamma", "moving_variance"):
            out[key] = np.ones(shape, np.float32)
        else:
            out[key] = np.zeros(shape, np.float32)
    return out


# ---- Keras `base_model.layers` order (what `base_model.layers[:N].trainable = False` freezes, spnet/models.py:361-372)
# Keras 2.1.3 sorts Model.layers by depth (distance to the output, largest first) and breaks ties by the order of a
# depth-first walk from the output that follows each layer's inputs in call order (topology.Container.__init__).
# For `add([main, residual])` the main branch is walked first, so a residual branch's layers sit next to the main
# branch's layers of the same depth, main first. Weight-less layers are listed too: they count towards N.
def _stem_layers():
    return ["input_1", "conv2d_1", "average_pooling2d_1", "batch_normalization_1", "leaky_re_lu_1", "conv2d_2",
            "batch_normalization_2", "leaky_re_lu_2", "conv2d_3", "batch_normalization_3", "average_pooling2d_2", "add_1",
            "dropout_1"]


def xception_keras_layers():
    """The 144 layers of Xception-SPNet's base_model (paper/run_logs/log_DatasetA_*.txt:95) in Keras order."""
    L = _stem_layers()
    L += ["block1_conv1", "block1_conv1_bn", "block1_conv1_act", "block1_conv2", "block1_conv2_bn", "block1_conv2_act"]
    nadd = 2
    for n, blk in enumerate((2, 3, 4, 13)):
        res = "conv2d_%d" % (4 + n)
        res_bn = "batch_normalization_%d" % (4 + n)
        if blk == 13:
            for b in MIDDLE_BLOCKS:
                for j in (1, 2, 3):
                    L += ["block%d_sepconv%d_act" % (b, j), "block%d_sepconv%d" % (b, j), "block%d_sepconv%d_bn" % (b, j)]
                L.append("add_%d" % nadd)
                nadd += 1
        if blk != 2:
            L.append("block%d_sepconv1_act" % blk)
        L += ["block%d_sepconv1" % blk, "block%d_sepconv1_bn" % blk, "block%d_sepconv2_act" % blk, "block%d_sepconv2" % blk,
              "block%d_sepconv2_bn" % blk, res, "block%d_pool" % blk, res_bn, "add_%d" % nadd]
        nadd += 1
    L += ["block14_sepconv1", "block14_sepconv1_bn", "block14_sepconv1_act", "block14_sepconv2", "block14_sepconv2_bn",
          "block14_sepconv2_act"]
    assert len(L) == 144, len(L)
    return L


def mobilenet_keras_layers():
    """1 Input + 12 stem layers + 81 MobileNet layers (conv1, conv1_bn, conv1_relu, 13 x 6)."""
    L = _stem_layers() + ["conv1", "conv1_bn", "conv1_relu"]
    for i in range(1, 14):
        L += ["conv_dw_%d" % i, "conv_dw_%d_bn" % i, "conv_dw_%d_relu" % i, "conv_pw_%d" % i, "conv_pw_%d_bn" % i,
              "conv_pw_%d_relu" % i]
    assert len(L) == 94, len(L)
    return L


def frozen_layer_names(backbone, freeze_fac, spec):
    """Names of the layers WITH WEIGHTS that `for i in range(int(num_layers * freeze_fac)): base_model.layers[i].trainable
    = False` freezes (spnet/models.py:361-372). Returns (names, n_frozen_layers, n_layers)."""
    if backbone == "InceptionResNetV2":
        # 792 layers with 3-4-way concatenations: the order of the weight-carrying layers is the construction order
        # (each branch is a chain), the cut is scaled from Keras layers to layers with weights
        total = 792
        nfreeze = int(total * freeze_fac)
        names = []
        for k, _, _, _ in spec:
            n = k.split("/")[0]
            if n != "FinalOutput" and n not in names:
                names.append(n)
        return names[:int(round(len(names) * nfreeze / float(total)))], nfreeze, total
    layers = mobilenet_keras_layers() if backbone == "MobileNet" else xception_keras_layers()
    nfreeze = int(len(layers) * freeze_fac)
    with_weights = set(k.split("/")[0] for k, _, _, _ in spec)
    return [n for n in layers[:nfreeze] if n in with_weights], nfreeze, len(layers)
