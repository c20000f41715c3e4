"""keras.applications.inception_resnet_v2.InceptionResNetV2(include_top=False) @ Keras 2.1.3 as a static
program of layer records (BASELINE configs[3]; reference call site spnet/models.py:18,357-359 — the
generic backbone branch). The source is not in the reference tree; this restates its published
architecture: conv2d_bn = Conv2D(no bias) -> BatchNormalization(scale=False) -> ReLU,
the stem, mixed_5b, 10 x block35 (scale 0.17), mixed_6a, 20 x block17 (scale 0.1), mixed_7a,
9 x block8 (scale 0.2) + 1 x block8 (scale 1, no activation), conv_7b. Pinned structurally by the
no-top parameter total 54,336,736.

Layer names follow Keras auto-naming inside the SPNet graph: the three stem convolutions are
conv2d_1..3 / batch_normalization_1..3, so the backbone's unnamed layers continue at 4.
"""
from collections import OrderedDict


def conv_out(n, k, s, pad):
    """TF output size and leading pad of one dimension."""
    if pad == "valid":
        return (n - k) // s + 1, 0
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return out, tot // 2


class Sym:
    """Symbolic NHWC activation: spatial size, channels, producer index, consumer count."""
    __slots__ = ("h", "w", "c", "idx", "consumers")

    def __init__(self, h, w, c, idx):
        self.h, self.w, self.c, self.idx, self.consumers = h, w, c, idx, 0


class Program:
    def __init__(self, h, w, c=3, first_auto=4):
        self.ops = []          # dicts: kind, inputs (Sym list), out (Sym), + parameters
        self.syms = []
        self.n_conv = first_auto   # next auto-name index for Conv2D
        self.n_bn = first_auto     # ... for BatchNormalization
        self.input = self._sym(h, w, c)

    def _sym(self, h, w, c):
        s = Sym(h, w, c, len(self.syms))
        self.syms.append(s)
        return s

    def _add(self, kind, inputs, out, **kw):
        for x in inputs:
            x.consumers += 1
        op = dict(kind=kind, inputs=list(inputs), out=out, **kw)
        self.ops.append(op)
        return out

    def conv_bn(self, x, cout, k, stride=1, pad="same", name=None, act=True):
        kh, kw = (k, k) if isinstance(k, int) else k
        oh, pt = conv_out(x.h, kh, stride, pad)
        ow, pl = conv_out(x.w, kw, stride, pad)
        if name is None:
            cname, bname = "conv2d_%d" % self.n_conv, "batch_normalization_%d" % self.n_bn
        else:
            cname, bname = name, name + "_bn"
        self.n_conv += 1 if name is None else 0
        self.n_bn += 1 if name is None else 0
        return self._add("conv_bn", [x], self._sym(oh, ow, cout), name=cname, bn=bname, kh=kh, kw=kw, stride=stride,
                         pt=pt, pl=pl, cin=x.c, cout=cout, act=act)

    def conv_bias(self, x, cout, name):
        return self._add("conv_bias", [x], self._sym(x.h, x.w, cout), name=name, cin=x.c, cout=cout)

    def maxpool(self, x):
        oh, _ = conv_out(x.h, 3, 2, "valid")
        ow, _ = conv_out(x.w, 3, 2, "valid")
        return self._add("maxpool", [x], self._sym(oh, ow, x.c))

    def avgpool(self, x):
        return self._add("avgpool", [x], self._sym(x.h, x.w, x.c))

    def concat(self, xs):
        return self._add("concat", xs, self._sym(xs[0].h, xs[0].w, sum(x.c for x in xs)))

    def residual(self, x, up_op_out, scale, relu):
        # y = act(x + scale * (up + bias)); the bias belongs to the conv_bias that produced `up`
        return self._add("residual", [x, up_op_out], self._sym(x.h, x.w, x.c), scale=scale, relu=relu)


def build_program(h, w):
    """InceptionResNetV2 on a (h, w, 3) input."""
    p = Program(h, w)
    x = p.conv_bn(p.input, 32, 3, stride=2, pad="valid")
    x = p.conv_bn(x, 32, 3, pad="valid")
    x = p.conv_bn(x, 64, 3)
    x = p.maxpool(x)
    x = p.conv_bn(x, 80, 1, pad="valid")
    x = p.conv_bn(x, 192, 3, pad="valid")
    x = p.maxpool(x)
    # mixed_5b
    b0 = p.conv_bn(x, 96, 1)
    b1 = p.conv_bn(p.conv_bn(x, 48, 1), 64, 5)
    b2 = p.conv_bn(p.conv_bn(p.conv_bn(x, 64, 1), 96, 3), 96, 3)
    bp = p.conv_bn(p.avgpool(x), 64, 1)
    x = p.concat([b0, b1, b2, bp])

    def block(x, scale, kind, idx, relu=True):
        if kind == "block35":
            b0 = p.conv_bn(x, 32, 1)
            b1 = p.conv_bn(p.conv_bn(x, 32, 1), 32, 3)
            b2 = p.conv_bn(p.conv_bn(p.conv_bn(x, 32, 1), 48, 3), 64, 3)
            branches = [b0, b1, b2]
        elif kind == "block17":
            b0 = p.conv_bn(x, 192, 1)
            b1 = p.conv_bn(p.conv_bn(p.conv_bn(x, 128, 1), 160, (1, 7)), 192, (7, 1))
            branches = [b0, b1]
        else:
            b0 = p.conv_bn(x, 192, 1)
            b1 = p.conv_bn(p.conv_bn(p.conv_bn(x, 192, 1), 224, (1, 3)), 256, (3, 1))
            branches = [b0, b1]
        mixed = p.concat(branches)
        up = p.conv_bias(mixed, x.c, "%s_%d_conv" % (kind, idx))
        return p.residual(x, up, scale, relu)

    for i in range(1, 11):
        x = block(x, 0.17, "block35", i)
    # mixed_6a
    b0 = p.conv_bn(x, 384, 3, stride=2, pad="valid")
    b1 = p.conv_bn(p.conv_bn(p.conv_bn(x, 256, 1), 256, 3), 384, 3, stride=2, pad="valid")
    bp = p.maxpool(x)
    x = p.concat([b0, b1, bp])
    for i in range(1, 21):
        x = block(x, 0.1, "block17", i)
    # mixed_7a
    b0 = p.conv_bn(p.conv_bn(x, 256, 1), 384, 3, stride=2, pad="valid")
    b1 = p.conv_bn(p.conv_bn(x, 256, 1), 288, 3, stride=2, pad="valid")
    b2 = p.conv_bn(p.conv_bn(p.conv_bn(x, 256, 1), 288, 3), 320, 3, stride=2, pad="valid")
    bp = p.maxpool(x)
    x = p.concat([b0, b1, b2, bp])
    for i in range(1, 10):
        x = block(x, 0.2, "block8", i)
    x = block(x, 1.0, "block8", 10, relu=False)
    x = p.conv_bn(x, 1536, 1, name="conv_7b")
    p.output = x
    return p


def param_spec(H, W, n_out=576):
    """InceptionResNetV2-SPNet: [(key, shape, trainable, l2_regularised)] in Keras creation order. Every
    Conv2D keeps its kernel_regularizer through add_regularization's JSON round trip (spnet/models.py:47-71)."""
    spec = []

    def bn(name, c, scale=True):
        if scale:
            spec.append((name + "/gamma", (c,), True, False))
        spec.append((name + "/beta", (c,), True, False))
        spec.append((name + "/moving_mean", (c,), False, False))
        spec.append((name + "/moving_variance", (c,), False, False))

    for i, cin in ((1, 1), (2, 3), (3, 3)):
        spec.append(("conv2d_%d/kernel" % i, (3, 3, cin, 3), True, True))
        bn("batch_normalization_%d" % i, 3)
    prog = build_program(H // 2, W // 2)
    for op in prog.ops:
        if op["kind"] == "conv_bn":
            spec.append((op["name"] + "/kernel", (op["kh"], op["kw"], op["cin"], op["cout"]), True, True))
            bn(op["bn"], op["cout"], scale=False)
        elif op["kind"] == "conv_bias":
            spec.append((op["name"] + "/kernel", (1, 1, op["cin"], op["cout"]), True, True))
            spec.append((op["name"] + "/bias", (op["cout"],), True, False))
    o = prog.output
    spec.append(("FinalOutput/kernel", (o.h * o.w * o.c, n_out), True, True))
    spec.append(("FinalOutput/bias", (n_out,), True, False))
    return spec


def shape_walk(H, W):
    prog = build_program(H // 2, W // 2)
    s = OrderedDict()
    s["input"] = (H, W)
    s["stem"] = (H // 2, W // 2)
    s["features"] = (prog.output.h, prog.output.w)
    return s
