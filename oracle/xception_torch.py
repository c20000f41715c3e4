"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement (PyTorch fp32 / fp64, autograd for the backward pass) of the network the
reference builds in spnet/models.py:302-424 (create_model_functional) with cf.basemodel =
'Xception' (spnet/config.py:52), its loss (custom_loss, :564-589), the L2 regulariser of
add_regularization (:47-71) and the Keras-2.1.3 Adam step compiled at :494-502.

The backbone is third-party: keras.applications.xception.Xception @ Keras 2.1.3
(requirements.txt:3) with TF 1.14 semantics (requirements.txt:7); that source is NOT in the
reference tree, so it is restated here from its published architecture:
TF 'SAME' asymmetric padding, BatchNormalization(eps=1e-3, momentum=0.99), separable conv =
depthwise 3x3 then pointwise 1x1 with nothing in between, pre-activation ReLU.

PARITY UNPINNED for the network numerics: TensorFlow/Keras cannot run in this image and the
reference has no golden activations. What IS pinned (tests/test_oracle_goldens.py): the
parameter totals 50,353,481 / 50,298,935 / 54,546, the 144-layer count, the stem output
(165,165,3) and backbone output (5,5,2048) printed in the reference's run logs
(paper/run_logs/log_DatasetA_*.txt:94-101) and the list of the 10 L2-regularised kernels (:98).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
L2 = 1e-4

# (name, kind, shape-fn) in Keras creation order -----------------------------------------------


def xception_spnet_spec(H, W, n_out=576):
    """Ordered list of (keras_layer_name, weight_name, shape, trainable, l2_regularised)."""
    spec = []

    def conv(name, kh, kw, cin, cout, reg=True):
        spec.append((name, "kernel", (kh, kw, cin, cout), True, reg))

    def bn(name, c):
        spec.append((name, "gamma", (c,), True, False))
        spec.append((name, "beta", (c,), True, False))
        spec.append((name, "moving_mean", (c,), False, False))
        spec.append((name, "moving_variance", (c,), False, False))

    def sep(name, cin, cout):
        spec.append((name, "depthwise_kernel", (3, 3, cin, 1), True, False))
        spec.append((name, "pointwise_kernel", (1, 1, cin, cout), True, False))

    # stem (spnet/models.py:321-336)
    conv("conv2d_1", 3, 3, 1, 3)
    bn("batch_normalization_1", 3)
    conv("conv2d_2", 3, 3, 3, 3)
    bn("batch_normalization_2", 3)
    conv("conv2d_3", 3, 3, 3, 3)
    bn("batch_normalization_3", 3)
    # Xception
    conv("block1_conv1", 3, 3, 3, 32)
    bn("block1_conv1_bn", 32)
    conv("block1_conv2", 3, 3, 32, 64)
    bn("block1_conv2_bn", 64)
    nres = 4
    cin = 64
    for blk, c in ((2, 128), (3, 256), (4, 728)):
        conv("conv2d_%d" % nres, 1, 1, cin, c)
        bn("batch_normalization_%d" % nres, c)
        nres += 1
        sep("block%d_sepconv1" % blk, cin, c)
        bn("block%d_sepconv1_bn" % blk, c)
        sep("block%d_sepconv2" % blk, c, c)
        bn("block%d_sepconv2_bn" % blk, c)
        cin = c
    for blk in range(5, 13):
        for j in (1, 2, 3):
            sep("block%d_sepconv%d" % (blk, j), 728, 728)
            bn("block%d_sepconv%d_bn" % (blk, j), 728)
    conv("conv2d_7", 1, 1, 728, 1024)
    bn("batch_normalization_7", 1024)
    sep("block13_sepconv1", 728, 728)
    bn("block13_sepconv1_bn", 728)
    sep("block13_sepconv2", 728, 1024)
    bn("block13_sepconv2_bn", 1024)
    sep("block14_sepconv1", 1024, 1536)
    bn("block14_sepconv1_bn", 1536)
    sep("block14_sepconv2", 1536, 2048)
    bn("block14_sepconv2_bn", 2048)
    fh, fw = feature_hw(H, W)
    spec.append(("FinalOutput", "kernel", (fh * fw * 2048, n_out), True, True))
    spec.append(("FinalOutput", "bias", (n_out,), True, False))
    return spec


def feature_hw(H, W):
    """Spatial size after the stem (avgpool 2) and Xception (valid s2, valid, 4x same s2)."""
    def walk(n):
        n = n // 2              # AveragePooling2D(2), 'valid'
        n = (n - 3) // 2 + 1    # block1_conv1 3x3 s2 valid
        n = n - 2               # block1_conv2 3x3 valid
        for _ in range(4):      # blocks 2,3,4,13: 'same' stride 2
            n = (n + 1) // 2
        return n
    return walk(H), walk(W)


def count_params(spec):
    tr = sum(int(np.prod(s)) for _, _, s, t, _ in spec if t)
    nt = sum(int(np.prod(s)) for _, _, s, t, _ in spec if not t)
    return tr + nt, tr, nt


def keras_layer_count(spec_unused=None):
    """Layers of `base_model` as Keras counts them (log line 95: 144 = 1 Input + 12 stem
    layers + 131 Xception layers after its input)."""
    stem = 12   # conv, avgpool, bn, leaky, conv, bn, leaky, conv, bn, avgpool(inputs), add, dropout
    xc = 0
    xc += 6                      # block1: conv,bn,act,conv,bn,act
    xc += 3 * (2 + 8)            # blocks 2-4: residual conv+bn (2) + act sep bn act sep bn pool add
    xc -= 1                      # block2 has no leading activation
    xc += 8 * 10                 # blocks 5-12: 3x(act,sep,bn) + add
    xc += 2 + 8                  # block13: residual conv+bn + act sep bn act sep bn pool add
    xc += 6                      # block14: sep bn act sep bn act
    return 1 + stem + xc


def init_weights(spec, seed=1):
    """Keras default initialisers: glorot_uniform kernels (fan computed on the Keras kernel
    shape, receptive field included; depthwise: fan_in = kh*kw*cin, fan_out = kh*kw*1 ... Keras
    uses shape[-2], shape[-1] times the receptive field), zero bias, BN gamma 1 / beta 0 /
    mean 0 / var 1."""
    rng = np.random.default_rng(seed)
    w = OrderedDict()
    for layer, wname, shape, _, _ in spec:
        key = layer + "/" + wname
        if wname in ("kernel", "depthwise_kernel", "pointwise_kernel"):
            if len(shape) == 2:
                fan_in, fan_out = shape
            else:
                rf = shape[0] * shape[1]
                fan_in, fan_out = shape[2] * rf, shape[3] * rf
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            w[key] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif wname in ("gamma", "moving_variance"):
            w[key] = np.ones(shape, np.float32)
        else:
            w[key] = np.zeros(shape, np.float32)
    return w


# ---- functional pieces --------------------------------------------------------------------------

def same_pads(n, k, s):
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2


def conv2d_tf(x, kernel, stride=1, padding="valid", groups=1):
    """x NCHW; kernel in Keras layout (kh,kw,cin,cout) (depthwise: (kh,kw,c,1))."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    if groups == 1:
        w = kernel.permute(3, 2, 0, 1)
    else:
        w = kernel.permute(2, 3, 0, 1)  # (c,1,kh,kw)
    if padding == "same":
        pt, pb = same_pads(x.shape[2], kh, stride)
        pl, pr = same_pads(x.shape[3], kw, stride)
        x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w, stride=stride, groups=groups)


def maxpool3s2_same(x):
    pt, pb = same_pads(x.shape[2], 3, 2)
    pl, pr = same_pads(x.shape[3], 3, 2)
    return F.max_pool2d(F.pad(x, (pl, pr, pt, pb), value=float("-inf")), 3, 2)


class OracleSPNet:
    """Holds the weights as torch leaves; forward() follows the Keras graph layer by layer."""

    def make_spec(self, H, W, n_out):
        return xception_spnet_spec(H, W, n_out)

    def __init__(self, weights, H, W, n_out=576, dtype=torch.float32, unbiased_moving_var=True):
        self.H, self.W, self.n_out, self.dtype = H, W, n_out, dtype
        self.spec = self.make_spec(H, W, n_out)
        self.p = OrderedDict()
        for layer, wname, shape, trainable, _ in self.spec:
            key = layer + "/" + wname
            t = torch.tensor(np.asarray(weights[key]), dtype=dtype)
            assert tuple(t.shape) == tuple(shape), (key, t.shape, shape)
            t.requires_grad_(trainable)
            self.p[key] = t
        self.unbiased = unbiased_moving_var
        self.trainable = [l + "/" + w for l, w, _, t, _ in self.spec if t]
        self.l2_keys = [l + "/" + w for l, w, _, t, r in self.spec if r]
        self.m = {k: torch.zeros_like(self.p[k]) for k in self.trainable}
        self.v = {k: torch.zeros_like(self.p[k]) for k in self.trainable}
        self.t = 0
        self.taps = {}

    # -- layers
    def bn(self, x, name, training):
        g, b = self.p[name + "/gamma"], self.p[name + "/beta"]
        mm, mv = self.p[name + "/moving_mean"], self.p[name + "/moving_variance"]
        if training:
            mean = x.mean((0, 2, 3))
            var = x.var((0, 2, 3), unbiased=False)
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                uv = var * n / max(n - 1, 1) if self.unbiased else var
                mm.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * mean)
                mv.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * uv)
        else:
            mean, var = mm, mv
        s = g / torch.sqrt(var + BN_EPS)
        return x * s[None, :, None, None] + (b - mean * s)[None, :, None, None]

    def sep(self, x, name):
        x = conv2d_tf(x, self.p[name + "/depthwise_kernel"], 1, "same", groups=x.shape[1])
        return conv2d_tf(x, self.p[name + "/pointwise_kernel"], 1, "valid")

    def forward(self, x_nhwc, training=False, dropout_mask=None, taps=False):
        """x_nhwc: (B,H,W,1) array. dropout_mask: optional (B,H/2,W/2,3) array of 0/1 keep flags
        (training only; None = no dropout, i.e. rate 0). Returns y (B, n_out)."""
        p = self.p
        x0 = torch.as_tensor(np.asarray(x_nhwc), dtype=self.dtype).permute(0, 3, 1, 2)
        tp = self.taps = {}
        x = conv2d_tf(x0, p["conv2d_1/kernel"], 1, "same")
        x = F.avg_pool2d(x, 2)
        x = F.leaky_relu(self.bn(x, "batch_normalization_1", training), 0.1)
        x = conv2d_tf(x, p["conv2d_2/kernel"], 1, "same")
        x = F.leaky_relu(self.bn(x, "batch_normalization_2", training), 0.1)
        x = conv2d_tf(x, p["conv2d_3/kernel"], 1, "same")
        x = self.bn(x, "batch_normalization_3", training) + F.avg_pool2d(x0, 2)
        if training and dropout_mask is not None:
            m = torch.as_tensor(np.asarray(dropout_mask), dtype=self.dtype).permute(0, 3, 1, 2)
            x = x * m / 0.9
        if taps:
            tp["stem"] = x
        x = self.backbone(x, training, taps)
        flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # Keras Flatten of NHWC
        return flat @ p["FinalOutput/kernel"] + p["FinalOutput/bias"]

    def backbone(self, x, training, taps=False):
        """keras.applications.Xception(include_top=False) on the stem output (NCHW)."""
        p = self.p
        # block 1
        x = torch.relu(self.bn(conv2d_tf(x, p["block1_conv1/kernel"], 2, "valid"), "block1_conv1_bn", training))
        x = torch.relu(self.bn(conv2d_tf(x, p["block1_conv2/kernel"], 1, "valid"), "block1_conv2_bn", training))
        if taps:
            self.taps["block1"] = x
        nres = 4
        for blk in (2, 3, 4):
            res = self.bn(conv2d_tf(x, p["conv2d_%d/kernel" % nres], 2, "same"), "batch_normalization_%d" % nres, training)
            nres += 1
            if blk != 2:
                x = torch.relu(x)
            x = self.bn(self.sep(x, "block%d_sepconv1" % blk), "block%d_sepconv1_bn" % blk, training)
            x = torch.relu(x)
            x = self.bn(self.sep(x, "block%d_sepconv2" % blk), "block%d_sepconv2_bn" % blk, training)
            x = maxpool3s2_same(x) + res
            if taps:
                self.taps["block%d" % blk] = x
        for blk in range(5, 13):
            r = x
            for j in (1, 2, 3):
                x = torch.relu(x)
                x = self.bn(self.sep(x, "block%d_sepconv%d" % (blk, j)), "block%d_sepconv%d_bn" % (blk, j), training)
            x = x + r
            if taps:
                self.taps["block%d" % blk] = x
        res = self.bn(conv2d_tf(x, p["conv2d_7/kernel"], 2, "same"), "batch_normalization_7", training)
        x = torch.relu(x)
        x = self.bn(self.sep(x, "block13_sepconv1"), "block13_sepconv1_bn", training)
        x = torch.relu(x)
        x = self.bn(self.sep(x, "block13_sepconv2"), "block13_sepconv2_bn", training)
        x = maxpool3s2_same(x) + res
        if taps:
            self.taps["block13"] = x
        x = torch.relu(self.bn(self.sep(x, "block14_sepconv1"), "block14_sepconv1_bn", training))
        x = torch.relu(self.bn(self.sep(x, "block14_sepconv2"), "block14_sepconv2_bn", training))
        if taps:
            self.taps["block14"] = x
        return x

    # -- loss (custom_loss, spnet/models.py:564-589) + L2 (:47-71)
    @staticmethod
    def custom_loss(y_true, y_pred, loss_type="same"):
        v = 8
        sq = (y_true - y_pred) ** 2
        pobj = 1 - y_true[:, 6::v]
        if loss_type == "same":
            loss = 0.3 * sq[:, 6::v].sum(-1)
        else:
            t, z = y_true[:, 6::v], y_pred[:, 6::v]
            loss = 0.3 * (torch.clamp(z, min=0) - z * t + torch.log1p(torch.exp(-z.abs()))).sum(-1)
        loss = loss + 2.0 * ((pobj * sq[:, 0::v]).sum(-1) + (pobj * sq[:, 1::v]).sum(-1))
        loss = loss + 1.0 * ((pobj * sq[:, 2::v]).sum(-1) + (pobj * sq[:, 3::v]).sum(-1))
        d2 = (y_true[:, 2::v] - y_true[:, 3::v]) ** 2
        loss = loss + 3.0 * ((pobj * sq[:, 4::v] * d2).sum(-1) + (pobj * sq[:, 5::v] * d2).sum(-1))
        loss = loss + 5.0 * (pobj * sq[:, 7::v]).sum(-1)
        return (loss / y_pred.shape[-1]).mean()

    def l2_term(self):
        return sum(L2 * (self.p[k] ** 2).sum() for k in self.l2_keys)

    def loss_and_grads(self, x, y_true, loss_type="same", dropout_mask=None, with_l2=True):
        for k in self.trainable:
            self.p[k].grad = None
        y = self.forward(x, training=True, dropout_mask=dropout_mask)
        yt = torch.as_tensor(np.asarray(y_true), dtype=self.dtype)
        data = self.custom_loss(yt, y, loss_type)
        total = data + (self.l2_term() if with_l2 else 0.0)
        total.backward()
        grads = {k: self.p[k].grad.detach().clone() for k in self.trainable}
        return float(total.detach()), float(data.detach()), y.detach(), grads

    def adam_step(self, grads, lr, beta1=0.9, beta2=0.999, eps=1e-7):
        """keras.optimizers.Adam.get_updates @ 2.1.3: eps is added to sqrt(v) un-corrected."""
        self.t += 1
        lr_t = lr * math.sqrt(1 - beta2 ** self.t) / (1 - beta1 ** self.t)
        with torch.no_grad():
            for k in self.trainable:
                g = grads[k]
                self.m[k] = beta1 * self.m[k] + (1 - beta1) * g
                self.v[k] = beta2 * self.v[k] + (1 - beta2) * g * g
                self.p[k] -= lr_t * self.m[k] / (self.v[k].sqrt() + eps)

    def weights_numpy(self):
        return OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self.p.items())


# =================================================================================================
# Storage-faithful mode: the same networks with every tensor the B200 engine keeps in HBM rounded to bf16
# at the point where the engine stores it (spnet_b200/engine.py):
#   * every convolution output z (pre-BatchNorm) is stored as bf16, and its BatchNorm batch statistics are
#     taken over the STORED values;
#   * BatchNorm (+ReLU) is applied by the consumer in fp32 as y = a*z + b and is NOT rounded when the consumer
#     is a depthwise / small convolution (on-load transform); it IS rounded where the engine materialises the
#     tensor: GEMM A operands (im2col of block1, MobileNet's relu6(BN(dw)), every conv+BN+ReLU of
#     InceptionResNetV2), block inputs / outputs (x + BN(z), maxpool(BN(z)) + BN(res)), the flattened features;
#   * weights of the tensor-core GEMMs (pointwise 1x1, residual 1x1, block1_conv2, Dense, every
#     InceptionResNetV2 convolution but the first) are the bf16 working copy; depthwise and stem / first-layer
#     kernels stay fp32.
# With storage="fp32" every rounding is the identity and these classes compute exactly what the plain oracles
# above compute (tests/test_oracle_goldens.py checks that), so the bf16 comparison separates the error of the
# KERNELS (accumulation order, a few flipped roundings) from the amplification of bf16 storage itself, which
# any implementation that stores bf16 activations shares.
# Backward: straight-through (the rounding has gradient 1), i.e. exact gradients of the rounded forward pass.
# =================================================================================================
def bf16_round(x):
    r = x.detach().to(torch.bfloat16).to(x.dtype)
    return x + (r - x.detach())


class _StoredMixin:
    storage = "bf16"

    def q(self, x):
        return bf16_round(x) if self.storage == "bf16" else x

    def wq(self, key):
        return self.q(self.p[key])

    def bn_ab(self, z, name, training):
        """(a, b) of y = a*z + b per channel, fp32, from fp64 statistics of the stored z (NCHW)."""
        g = self.p[name + "/gamma"] if (name + "/gamma") in self.p else None
        b = self.p[name + "/beta"]
        mm, mv = self.p[name + "/moving_mean"], self.p[name + "/moving_variance"]
        if training:
            zd = z.double()
            mean = zd.mean((0, 2, 3))
            var = zd.var((0, 2, 3), unbiased=False)
            n = z.numel() // z.shape[1]
            with torch.no_grad():
                uv = var * n / max(n - 1, 1) if self.unbiased else var
                mm.mul_(BN_MOMENTUM).add_(((1 - BN_MOMENTUM) * mean).to(mm.dtype))
                mv.mul_(BN_MOMENTUM).add_(((1 - BN_MOMENTUM) * uv).to(mv.dtype))
            rstd = (1.0 / torch.sqrt(var + BN_EPS)).to(z.dtype)
            mean = mean.to(z.dtype)
        else:
            rstd = 1.0 / torch.sqrt(mv + BN_EPS)
            mean = mm
        a = rstd if g is None else g * rstd
        return a, b - mean * a

    def bnz(self, z, name, training):
        a, b = self.bn_ab(z, name, training)
        return z * a[None, :, None, None] + b[None, :, None, None]

    def stem(self, x_nhwc, training, dropout_mask):
        p = self.p
        x0 = torch.as_tensor(np.asarray(x_nhwc), dtype=self.dtype).permute(0, 3, 1, 2)
        p1 = self.q(F.avg_pool2d(conv2d_tf(x0, p["conv2d_1/kernel"], 1, "same"), 2))
        s0 = self.q(F.avg_pool2d(x0, 2))
        x = F.leaky_relu(self.bnz(p1, "batch_normalization_1", training), 0.1)
        c2 = self.q(conv2d_tf(x, p["conv2d_2/kernel"], 1, "same"))
        x = F.leaky_relu(self.bnz(c2, "batch_normalization_2", training), 0.1)
        c3 = self.q(conv2d_tf(x, p["conv2d_3/kernel"], 1, "same"))
        x = self.bnz(c3, "batch_normalization_3", training) + s0
        if training and dropout_mask is not None:
            m = torch.as_tensor(np.asarray(dropout_mask), dtype=self.dtype).permute(0, 3, 1, 2)
            x = x * m / 0.9
        return self.q(x)

    def forward(self, x_nhwc, training=False, dropout_mask=None, taps=False):
        self.taps = {}
        d = self.stem(x_nhwc, training, dropout_mask)
        if taps:
            self.taps["stem"] = d
        feat = self.backbone(d, training, taps)  # already materialised (rounded) by the backbone
        flat = feat.permute(0, 2, 3, 1).reshape(feat.shape[0], -1)
        return flat @ self.wq("FinalOutput/kernel") + self.p["FinalOutput/bias"]


class OracleSPNetStored(_StoredMixin, OracleSPNet):
    """Xception-SPNet with the engine's storage points (XceptionSPNetEngine._backbone_fwd)."""

    def __init__(self, weights, H, W, n_out=576, dtype=torch.float32, unbiased_moving_var=True, storage="bf16"):
        OracleSPNet.__init__(self, weights, H, W, n_out, dtype, unbiased_moving_var)
        self.storage = storage

    def sepz(self, x, name):
        """x: the depthwise input AFTER the on-load transform (fp32, not rounded) -> stored z of the pointwise."""
        t = self.q(conv2d_tf(x, self.p[name + "/depthwise_kernel"], 1, "same", groups=x.shape[1]))
        return self.q(conv2d_tf(t, self.wq(name + "/pointwise_kernel"), 1, "valid"))

    def entry(self, x, blk, nres, relu_in, training):
        xs = x[:, :, ::2, ::2]  # 1x1 stride 2 'same' reads the even pixels
        zr = self.q(conv2d_tf(xs, self.wq("conv2d_%d/kernel" % nres), 1, "valid"))
        z1 = self.sepz(torch.relu(x) if relu_in else x, "block%d_sepconv1" % blk)
        z2 = self.sepz(torch.relu(self.bnz(z1, "block%d_sepconv1_bn" % blk, training)), "block%d_sepconv2" % blk)
        y = maxpool3s2_same(self.bnz(z2, "block%d_sepconv2_bn" % blk, training))
        return self.q(y + self.bnz(zr, "batch_normalization_%d" % nres, training))

    def backbone(self, x, training, taps=False):
        p = self.p
        z11 = self.q(conv2d_tf(x, p["block1_conv1/kernel"], 2, "valid"))
        col = self.q(torch.relu(self.bnz(z11, "block1_conv1_bn", training)))          # im2col operand
        z12 = self.q(conv2d_tf(col, self.wq("block1_conv2/kernel"), 1, "valid"))
        x = self.q(torch.relu(self.bnz(z12, "block1_conv2_bn", training)))            # x2
        if taps:
            self.taps["block1"] = x
        for blk, nres in ((2, 4), (3, 5), (4, 6)):
            x = self.entry(x, blk, nres, blk != 2, training)
            if taps:
                self.taps["block%d" % blk] = x
        for blk in range(5, 13):
            z = self.sepz(torch.relu(x), "block%d_sepconv1" % blk)
            z = self.sepz(torch.relu(self.bnz(z, "block%d_sepconv1_bn" % blk, training)), "block%d_sepconv2" % blk)
            z = self.sepz(torch.relu(self.bnz(z, "block%d_sepconv2_bn" % blk, training)), "block%d_sepconv3" % blk)
            x = self.q(self.bnz(z, "block%d_sepconv3_bn" % blk, training) + x)
            if taps:
                self.taps["block%d" % blk] = x
        x = self.entry(x, 13, 7, True, training)
        if taps:
            self.taps["block13"] = x
        z = self.sepz(x, "block14_sepconv1")
        z = self.sepz(torch.relu(self.bnz(z, "block14_sepconv1_bn", training)), "block14_sepconv2")
        x = self.q(torch.relu(self.bnz(z, "block14_sepconv2_bn", training)))
        if taps:
            self.taps["block14"] = x
        return x


# =================================================================================================
# MobileNet backbone (BASELINE configs[2]): keras.applications.mobilenet.MobileNet @ Keras 2.1.3,
# alpha = 1, depth_multiplier = 1, include_top=False (reference call site spnet/models.py:349-355).
# Not in the reference tree either; restated from its published architecture:
# conv1 3x3 s2 'same' -> BN -> relu6, then 13 x [DepthwiseConv2D 3x3 'same' stride s -> BN -> relu6
# -> Conv 1x1 -> BN -> relu6]. Pinned structurally by the no-top parameter total 3,228,864.
# =================================================================================================
MOBILENET_BLOCKS = ((32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2),
                    (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 1024, 2),
                    (1024, 1024, 1))


def mobilenet_feature_hw(H, W):
    def walk(n):
        n = n // 2            # AveragePooling2D(2)
        for _ in range(5):    # conv1 and four stride-2 depthwise stages, all 'same'
            n = -(-n // 2)
        return n
    return walk(H), walk(W)


def mobilenet_spnet_spec(H, W, n_out=576):
    spec = []

    def bn(name, c):
        spec.append((name, "gamma", (c,), True, False))
        spec.append((name, "beta", (c,), True, False))
        spec.append((name, "moving_mean", (c,), False, False))
        spec.append((name, "moving_variance", (c,), False, False))

    for i, cin in ((1, 1), (2, 3), (3, 3)):
        spec.append(("conv2d_%d" % i, "kernel", (3, 3, cin, 3), True, True))
        bn("batch_normalization_%d" % i, 3)
    spec.append(("conv1", "kernel", (3, 3, 3, 32), True, True))
    bn("conv1_bn", 32)
    for i, (cin, cout, _) in enumerate(MOBILENET_BLOCKS, start=1):
        spec.append(("conv_dw_%d" % i, "depthwise_kernel", (3, 3, cin, 1), True, False))
        bn("conv_dw_%d_bn" % i, cin)
        spec.append(("conv_pw_%d" % i, "kernel", (1, 1, cin, cout), True, True))
        bn("conv_pw_%d_bn" % i, cout)
    fh, fw = mobilenet_feature_hw(H, W)
    spec.append(("FinalOutput", "kernel", (fh * fw * 1024, n_out), True, True))
    spec.append(("FinalOutput", "bias", (n_out,), True, False))
    return spec


class OracleMobileNetSPNet(OracleSPNet):
    def make_spec(self, H, W, n_out):
        return mobilenet_spnet_spec(H, W, n_out)

    def backbone(self, x, training, taps=False):
        p = self.p
        relu6 = lambda v: torch.clamp(v, 0.0, 6.0)  # noqa: E731
        x = relu6(self.bn(conv2d_tf(x, p["conv1/kernel"], 2, "same"), "conv1_bn", training))
        for i, (cin, cout, stride) in enumerate(MOBILENET_BLOCKS, start=1):
            x = conv2d_tf(x, p["conv_dw_%d/depthwise_kernel" % i], stride, "same", groups=cin)
            x = relu6(self.bn(x, "conv_dw_%d_bn" % i, training))
            x = conv2d_tf(x, p["conv_pw_%d/kernel" % i], 1, "valid")
            x = relu6(self.bn(x, "conv_pw_%d_bn" % i, training))
            if taps:
                self.taps["block%d" % i] = x
        return x


# =================================================================================================
# InceptionResNetV2 backbone (BASELINE configs[3]): keras.applications.inception_resnet_v2 @ Keras 2.1.3,
# include_top=False (reference call site spnet/models.py:18,357-359). Not in the reference tree; restated
# from its published architecture. conv2d_bn = Conv2D(use_bias=False) ->
# BatchNormalization(scale=False) -> ReLU; Inception-ResNet blocks end in a 1x1 convolution WITH bias and
# no BN/activation, `x + scale * up`, then ReLU. Pinned structurally by the no-top total 54,336,736.
# =================================================================================================
class _IRv2Walker:
    """Walks the Keras construction code once; `emit` decides what a layer does (spec or tensor math)."""

    def __init__(self, conv_bn, conv_bias, maxpool, avgpool, concat, residual):
        self.conv_bn, self.conv_bias, self.maxpool, self.avgpool = conv_bn, conv_bias, maxpool, avgpool
        self.concat, self.residual = concat, residual

    def run(self, x):
        cb = self.conv_bn
        x = cb(x, 32, 3, 2, "valid")
        x = cb(x, 32, 3, 1, "valid")
        x = cb(x, 64, 3, 1, "same")
        x = self.maxpool(x)
        x = cb(x, 80, 1, 1, "valid")
        x = cb(x, 192, 3, 1, "valid")
        x = self.maxpool(x)
        b0 = cb(x, 96, 1, 1, "same")
        b1 = cb(cb(x, 48, 1, 1, "same"), 64, 5, 1, "same")
        b2 = cb(cb(cb(x, 64, 1, 1, "same"), 96, 3, 1, "same"), 96, 3, 1, "same")
        bp = cb(self.avgpool(x), 64, 1, 1, "same")
        x = self.concat([b0, b1, b2, bp])
        for i in range(1, 11):
            x = self.block(x, 0.17, "block35", i, True)
        b0 = cb(x, 384, 3, 2, "valid")
        b1 = cb(cb(cb(x, 256, 1, 1, "same"), 256, 3, 1, "same"), 384, 3, 2, "valid")
        x = self.concat([b0, b1, self.maxpool(x)])
        for i in range(1, 21):
            x = self.block(x, 0.1, "block17", i, True)
        b0 = cb(cb(x, 256, 1, 1, "same"), 384, 3, 2, "valid")
        b1 = cb(cb(x, 256, 1, 1, "same"), 288, 3, 2, "valid")
        b2 = cb(cb(cb(x, 256, 1, 1, "same"), 288, 3, 1, "same"), 320, 3, 2, "valid")
        x = self.concat([b0, b1, b2, self.maxpool(x)])
        for i in range(1, 10):
            x = self.block(x, 0.2, "block8", i, True)
        x = self.block(x, 1.0, "block8", 10, False)
        return cb(x, 1536, 1, 1, "same", name="conv_7b")

    def block(self, x, scale, kind, idx, relu):
        cb = self.conv_bn
        if kind == "block35":
            br = [cb(x, 32, 1, 1, "same"), cb(cb(x, 32, 1, 1, "same"), 32, 3, 1, "same"),
                  cb(cb(cb(x, 32, 1, 1, "same"), 48, 3, 1, "same"), 64, 3, 1, "same")]
        elif kind == "block17":
            br = [cb(x, 192, 1, 1, "same"), cb(cb(cb(x, 128, 1, 1, "same"), 160, (1, 7), 1, "same"), 192, (7, 1), 1, "same")]
        else:
            br = [cb(x, 192, 1, 1, "same"), cb(cb(cb(x, 192, 1, 1, "same"), 224, (1, 3), 1, "same"), 256, (3, 1), 1, "same")]
        up = self.conv_bias(self.concat(br), "%s_%d_conv" % (kind, idx))
        return self.residual(x, up, scale, relu)


def irv2_spnet_spec(H, W, n_out=576):
    """Ordered (layer, weight, shape, trainable, l2) list; shapes are propagated symbolically as (h, w, c)."""
    spec = []
    cnt = [4, 4]  # next auto index for Conv2D / BatchNormalization (1..3 belong to the SPNet stem)

    def bn(name, c, scale=True):
        if scale:
            spec.append((name, "gamma", (c,), True, False))
        spec.append((name, "beta", (c,), True, False))
        spec.append((name, "moving_mean", (c,), False, False))
        spec.append((name, "moving_variance", (c,), False, False))

    def osz(n, k, s, pad):
        return (n - k) // s + 1 if pad == "valid" else -(-n // s)

    def conv_bn(x, cout, k, s, pad, name=None):
        kh, kw = (k, k) if isinstance(k, int) else k
        if name is None:
            name, bname = "conv2d_%d" % cnt[0], "batch_normalization_%d" % cnt[1]
            cnt[0] += 1
            cnt[1] += 1
        else:
            bname = name + "_bn"
        spec.append((name, "kernel", (kh, kw, x[2], cout), True, True))
        bn(bname, cout, scale=False)
        return (osz(x[0], kh, s, pad), osz(x[1], kw, s, pad), cout)

    def conv_bias(x, name):
        return ("pending", x, name)

    def residual(x, up, scale, relu):
        _, m, name = up
        spec.append((name, "kernel", (1, 1, m[2], x[2]), True, True))
        spec.append((name, "bias", (x[2],), True, False))
        return x

    for i, cin in ((1, 1), (2, 3), (3, 3)):
        spec.append(("conv2d_%d" % i, "kernel", (3, 3, cin, 3), True, True))
        bn("batch_normalization_%d" % i, 3)
    w = _IRv2Walker(conv_bn, conv_bias, lambda x: (osz(x[0], 3, 2, "valid"), osz(x[1], 3, 2, "valid"), x[2]), lambda x: x,
                    lambda xs: (xs[0][0], xs[0][1], sum(t[2] for t in xs)), residual)
    o = w.run((H // 2, W // 2, 3))
    spec.append(("FinalOutput", "kernel", (o[0] * o[1] * o[2], n_out), True, True))
    spec.append(("FinalOutput", "bias", (n_out,), True, False))
    return spec


class OracleIRv2SPNet(OracleSPNet):
    def make_spec(self, H, W, n_out):
        return irv2_spnet_spec(H, W, n_out)

    def bn(self, x, name, training):
        if (name + "/gamma") in self.p:
            return super().bn(x, name, training)
        b = self.p[name + "/beta"]                      # BatchNormalization(scale=False): gamma == 1
        mm, mv = self.p[name + "/moving_mean"], self.p[name + "/moving_variance"]
        if training:
            mean = x.mean((0, 2, 3))
            var = x.var((0, 2, 3), unbiased=False)
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                uv = var * n / max(n - 1, 1) if self.unbiased else var
                mm.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * mean)
                mv.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * uv)
        else:
            mean, var = mm, mv
        s = 1.0 / torch.sqrt(var + BN_EPS)
        return x * s[None, :, None, None] + (b - mean * s)[None, :, None, None]

    def backbone(self, x, training, taps=False):
        p = self.p
        cnt = [4, 4]

        def conv_bn(t, cout, k, s, pad, name=None):
            if name is None:
                name, bname = "conv2d_%d" % cnt[0], "batch_normalization_%d" % cnt[1]
                cnt[0] += 1
                cnt[1] += 1
            else:
                bname = name + "_bn"
            return torch.relu(self.bn(conv2d_tf(t, p[name + "/kernel"], s, pad), bname, training))

        def conv_bias(t, name):
            return conv2d_tf(t, p[name + "/kernel"], 1, "valid") + p[name + "/bias"][None, :, None, None]

        def avgpool(t):
            # AveragePooling2D(3, 1, 'same'): TF averages over the in-image cells only
            return F.avg_pool2d(t, 3, 1, padding=1, count_include_pad=False)

        def residual(t, up, scale, relu):
            y = t + scale * up
            y = torch.relu(y) if relu else y
            if taps:   # debugging aid: (input, up-branch, output) of every Inception-ResNet block, gradients retained
                for v in (t, up, y):
                    if v.requires_grad:
                        v.retain_grad()
                self.taps.setdefault("residual", []).append((t, up, y))
            return y

        w = _IRv2Walker(conv_bn, conv_bias, lambda t: F.max_pool2d(t, 3, 2), avgpool, lambda ts: torch.cat(ts, 1), residual)
        return w.run(x)


class OracleMobileNetSPNetStored(_StoredMixin, OracleMobileNetSPNet):
    """MobileNet-SPNet with the engine's storage points (MobileNetSPNetEngine._backbone_fwd)."""

    def __init__(self, weights, H, W, n_out=576, dtype=torch.float32, unbiased_moving_var=True, storage="bf16"):
        OracleMobileNetSPNet.__init__(self, weights, H, W, n_out, dtype, unbiased_moving_var)
        self.storage = storage

    def backbone(self, x, training, taps=False):
        p = self.p
        relu6 = lambda v: torch.clamp(v, 0.0, 6.0)  # noqa: E731
        z, zbn = self.q(conv2d_tf(x, p["conv1/kernel"], 2, "same")), "conv1_bn"
        for i, (cin, cout, stride) in enumerate(MOBILENET_BLOCKS, start=1):
            y = relu6(self.bnz(z, zbn, training))                                      # on load, not rounded
            zd = self.q(conv2d_tf(y, p["conv_dw_%d/depthwise_kernel" % i], stride, "same", groups=cin))
            t = self.q(relu6(self.bnz(zd, "conv_dw_%d_bn" % i, training)))             # GEMM operand: materialised
            z, zbn = self.q(conv2d_tf(t, self.wq("conv_pw_%d/kernel" % i), 1, "valid")), "conv_pw_%d_bn" % i
            if taps:
                self.taps["block%d" % i] = relu6(self.bnz(z, zbn, False)) if not training else z
        return self.q(relu6(self.bnz(z, zbn, training)))


class OracleIRv2SPNetStored(_StoredMixin, OracleIRv2SPNet):
    """InceptionResNetV2-SPNet with the engine's storage points (InceptionResNetV2SPNetEngine._backbone_fwd): every
    conv output z and every conv+BN+ReLU output, pool, residual sum is a stored tensor."""

    def __init__(self, weights, H, W, n_out=576, dtype=torch.float32, unbiased_moving_var=True, storage="bf16"):
        OracleIRv2SPNet.__init__(self, weights, H, W, n_out, dtype, unbiased_moving_var)
        self.storage = storage

    def backbone(self, x, training, taps=False):
        p = self.p
        cnt = [4, 4]
        first = [True]

        def conv_bn(t, cout, k, s, pad, name=None):
            if name is None:
                name, bname = "conv2d_%d" % cnt[0], "batch_normalization_%d" % cnt[1]
                cnt[0] += 1
                cnt[1] += 1
            else:
                bname = name + "_bn"
            w = p[name + "/kernel"] if first[0] else self.wq(name + "/kernel")  # the 3-channel first layer keeps fp32 weights
            first[0] = False
            z = self.q(conv2d_tf(t, w, s, pad))
            return self.q(torch.relu(self.bnz(z, bname, training)))

        def conv_bias(t, name):
            return self.q(conv2d_tf(t, self.wq(name + "/kernel"), 1, "valid")), p[name + "/bias"]

        def avgpool(t):
            return self.q(F.avg_pool2d(t, 3, 1, padding=1, count_include_pad=False))

        def residual(t, up, scale, relu):
            u, bias = up
            y = t + scale * (u + bias[None, :, None, None])
            return self.q(torch.relu(y) if relu else y)

        w = _IRv2Walker(conv_bn, conv_bias, lambda t: F.max_pool2d(t, 3, 2), avgpool, lambda ts: torch.cat(ts, 1), residual)
        return w.run(x)
