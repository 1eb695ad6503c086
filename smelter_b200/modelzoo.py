"""Synthetic ONNX models of the shapes BASELINE.json names, written with the in-repo protobuf codec.

The reference ships no model files; these generators emit opset-9 graphs restricted to the op
set Smelter registers (Sources/Smelter/ONNXGraph.swift:110-155) plus `Clip` (MobileNetV2's ReLU6, a documented
extension).  Node/attribute conventions are the ones torch's opset-9 exporter produces for the same
architectures (SURVEY.md §8d "Model files"), which tests/test_modelzoo.py cross-checks against real torch
exports: Conv{dilations,group,kernel_shape,pads,strides}, Gemm{alpha,beta,transB}, Flatten{axis},
MaxPool{kernel_shape,pads,strides}, Pad{mode,pads}, Upsample{mode} + Constant scales, Clip{min,max}.

Weights are seeded random (no network for checkpoints): He-normal convolutions, randomised norm statistics
(gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5)) so BN folding is a meaningful test.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import onnx_proto as op


class GraphBuilder:
    def __init__(self, seed: int = 0, fold_bn: bool = False, name: str = "g"):
        self.rng = np.random.default_rng(seed)
        self.fold_bn = fold_bn
        self.graph = op.Graph(name=name)
        self._n = 0
        self.channels = {}  # value name -> channel count

    # ---- plumbing
    def _name(self, base: str) -> str:
        self._n += 1
        return f"{base}_{self._n}"

    def _init(self, base: str, a: np.ndarray) -> str:
        name = self._name(base)
        self.graph.initializer.append(op.Tensor.from_numpy(name, a))
        return name

    def _node(self, op_type: str, inputs: Sequence[str], attrs: dict, c_out: int, n_out: int = 1) -> str:
        out = self._name(op_type.lower())
        self.graph.node.append(op.Node(op_type=op_type, input=list(inputs), output=[out], name=out,
                                       attribute=[op.attr(k, v) for k, v in attrs.items()]))
        self.channels[out] = c_out
        return out

    def input(self, name: str, shape: Sequence[int]) -> str:
        self.graph.input.append(op.ValueInfo(name=name, elem_type=op.FLOAT, dims=list(shape)))
        self.channels[name] = shape[1] if len(shape) == 4 else shape[0]
        return name

    def output(self, name: str, shape: Sequence[int]) -> None:
        self.graph.output.append(op.ValueInfo(name=name, elem_type=op.FLOAT, dims=list(shape)))

    def model(self, producer: str = "smelter_b200.modelzoo") -> op.Model:
        return op.Model(ir_version=4, producer_name=producer, producer_version="1", graph=self.graph, opset_import=[("", 9)])

    # ---- ops
    def conv(self, x: str, c_out: int, k: int, stride: int = 1, pad: int = 0, groups: int = 1, bias: bool = True,
             dilation: int = 1, gain: float = 2.0) -> str:
        c_in = self.channels[x]
        fan_in = (c_in // groups) * k * k
        w = self.rng.standard_normal((c_out, c_in // groups, k, k)).astype(np.float32) * np.float32(np.sqrt(gain / fan_in))
        ins = [x, self._init("w", w)]
        if bias:
            ins.append(self._init("b", (self.rng.standard_normal(c_out) * 0.1).astype(np.float32)))
        return self._node("Conv", ins, {"dilations": [dilation, dilation], "group": groups, "kernel_shape": [k, k],
                                       "pads": [pad, pad, pad, pad], "strides": [stride, stride]}, c_out)

    def conv_transpose(self, x: str, c_out: int, k: int, stride: int = 1, pad: int = 0, output_padding: int = 0, bias: bool = True) -> str:
        c_in = self.channels[x]
        w = self.rng.standard_normal((c_in, c_out, k, k)).astype(np.float32) * np.float32(np.sqrt(2.0 * stride * stride / (c_in * k * k)))
        ins = [x, self._init("wt", w)]
        if bias:
            ins.append(self._init("b", (self.rng.standard_normal(c_out) * 0.1).astype(np.float32)))
        return self._node("ConvTranspose", ins, {"dilations": [1, 1], "group": 1, "kernel_shape": [k, k], "pads": [pad, pad, pad, pad],
                                                "strides": [stride, stride], "output_padding": [output_padding, output_padding]}, c_out)

    def group_norm(self, x: str, groups: int) -> str:
        """`custom_group_norm` as the reference's exporter emits it: X, groups, gamma, beta (Converters.swift:1273-1300)."""
        c = self.channels[x]
        gamma = self.rng.uniform(0.5, 1.5, c).astype(np.float32)
        beta = (self.rng.standard_normal(c) * 0.1).astype(np.float32)
        return self._node("custom_group_norm", [x, self._init("groups", np.asarray([groups], dtype=np.int64)), self._init("gamma", gamma),
                                                self._init("beta", beta)], {}, c)

    def pow(self, x: str, exponent: float) -> str:
        return self._node("Pow", [x, self._init("exponent", np.asarray([exponent], dtype=np.float32))], {}, self.channels[x])

    def _bn_params(self, c: int, gamma_scale: float = 1.0):
        gamma = (self.rng.uniform(0.5, 1.5, c) * gamma_scale).astype(np.float32)
        beta = (self.rng.standard_normal(c) * 0.1).astype(np.float32)
        mean = (self.rng.standard_normal(c) * 0.1).astype(np.float32)
        var = self.rng.uniform(0.5, 1.5, c).astype(np.float32)
        return gamma, beta, mean, var

    def bn(self, x: str, gamma_scale: float = 1.0, eps: float = 1e-5) -> str:
        c = self.channels[x]
        gamma, beta, mean, var = self._bn_params(c, gamma_scale)
        return self._node("BatchNormalization", [x, self._init("gamma", gamma), self._init("beta", beta), self._init("mean", mean),
                                                 self._init("var", var)], {"epsilon": float(eps), "momentum": 0.9}, c)

    def conv_bn(self, x: str, c_out: int, k: int, stride: int = 1, pad: int = 0, groups: int = 1, gamma_scale: float = 1.0,
                gain: float = 2.0) -> str:
        """Conv (no bias) + BatchNormalization; with fold_bn the pair is emitted as one biased Conv, the way torch's
        exporter constant-folds eval-mode BN."""
        if not self.fold_bn:
            return self.bn(self.conv(x, c_out, k, stride, pad, groups, bias=False, gain=gain), gamma_scale)
        c_in = self.channels[x]
        fan_in = (c_in // groups) * k * k
        w = self.rng.standard_normal((c_out, c_in // groups, k, k)).astype(np.float32) * np.float32(np.sqrt(gain / fan_in))
        gamma, beta, mean, var = self._bn_params(c_out, gamma_scale)
        s = gamma / np.sqrt(var + np.float32(1e-5))
        wf = (w * s[:, None, None, None]).astype(np.float32)
        bf = (beta - mean * s).astype(np.float32)
        return self._node("Conv", [x, self._init("w", wf), self._init("b", bf)],
                          {"dilations": [1, 1], "group": groups, "kernel_shape": [k, k], "pads": [pad, pad, pad, pad],
                           "strides": [stride, stride]}, c_out)

    def relu(self, x: str) -> str:
        return self._node("Relu", [x], {}, self.channels[x])

    def sigmoid(self, x: str) -> str:
        return self._node("Sigmoid", [x], {}, self.channels[x])

    def clip(self, x: str, lo: float = 0.0, hi: float = 6.0) -> str:
        return self._node("Clip", [x], {"max": float(hi), "min": float(lo)}, self.channels[x])

    def add(self, a: str, b: str) -> str:
        return self._node("Add", [a, b], {}, self.channels[a])

    def maxpool(self, x: str, k: int, stride: int, pad: int) -> str:
        return self._node("MaxPool", [x], {"kernel_shape": [k, k], "pads": [pad, pad, pad, pad], "strides": [stride, stride]}, self.channels[x])

    def avgpool(self, x: str, k: int, stride: int, pad: int) -> str:
        return self._node("AveragePool", [x], {"kernel_shape": [k, k], "pads": [pad, pad, pad, pad], "strides": [stride, stride]}, self.channels[x])

    def gap(self, x: str) -> str:
        return self._node("GlobalAveragePool", [x], {}, self.channels[x])

    def flatten(self, x: str, c_out: Optional[int] = None) -> str:
        return self._node("Flatten", [x], {"axis": 1}, c_out or self.channels[x])

    def gemm(self, x: str, c_out: int, gain: float = 0.1) -> str:
        c_in = self.channels[x]
        w = self.rng.standard_normal((c_out, c_in)).astype(np.float32) * np.float32(np.sqrt(gain / c_in))
        b = (self.rng.standard_normal(c_out) * 0.1).astype(np.float32)
        return self._node("Gemm", [x, self._init("fc_w", w), self._init("fc_b", b)], {"alpha": 1.0, "beta": 1.0, "transB": 1}, c_out)

    def softmax(self, x: str) -> str:
        return self._node("Softmax", [x], {"axis": 1}, self.channels[x])

    def pad(self, x: str, p: int, mode: str = "reflect") -> str:
        return self._node("Pad", [x], {"mode": mode, "pads": [0, 0, p, p, 0, 0, p, p]}, self.channels[x])

    def instancenorm(self, x: str, eps: float = 1e-5) -> str:
        c = self.channels[x]
        gamma = self.rng.uniform(0.5, 1.5, c).astype(np.float32)
        beta = (self.rng.standard_normal(c) * 0.1).astype(np.float32)
        return self._node("InstanceNormalization", [x, self._init("in_g", gamma), self._init("in_b", beta)], {"epsilon": float(eps)}, c)

    def upsample(self, x: str, scale: int, mode: str = "nearest") -> str:
        # opset 9: scales arrive through a Constant node (what torch emits; Converters.swift:505-514 reads it via tensor(name:))
        cname = self._name("scales")
        t = op.Tensor.from_numpy("", np.asarray([1, 1, scale, scale], dtype=np.float32))
        self.graph.node.append(op.Node(op_type="Constant", input=[], output=[cname], name=cname, attribute=[op.attr("value", t)]))
        return self._node("Upsample", [x, cname], {"mode": mode}, self.channels[x])

    def concat(self, xs: Sequence[str]) -> str:
        return self._node("Concat", list(xs), {"axis": 1}, sum(self.channels[x] for x in xs))

    def reshape(self, x: str, shape: Sequence[int], c_out: int) -> str:
        sname = self._init("shape", np.asarray(shape, dtype=np.int64))
        return self._node("Reshape", [x, sname], {}, c_out)


# ------------------------------------------------------------------------------------------------ models
def conv_bn_relu(seed: int = 0, c_in: int = 3, c_out: int = 8, hw: int = 16) -> op.Model:
    """BASELINE.json configs[0]: Conv(3->8, 3x3, pad 1, bias) -> BatchNormalization(8) -> Relu on 1x3x16x16."""
    g = GraphBuilder(seed, fold_bn=False, name="conv_bn_relu")
    x = g.input("input", [1, c_in, hw, hw])
    y = g.relu(g.bn(g.conv(x, c_out, 3, 1, 1, bias=True)))
    g.output(y, [1, c_out, hw, hw])
    return g.model()


def resnet50(seed: int = 0, fold_bn: bool = True, batch: int = 1, num_classes: int = 1000, hw: int = 224,
             widths: Sequence[int] = (64, 128, 256, 512), depths: Sequence[int] = (3, 4, 6, 3)) -> op.Model:
    """torchvision-style ResNet-50 v1.5 (stride on the 3x3): 53 Conv + 49 Relu + 16 Add + MaxPool + GAP + Flatten + Gemm."""
    g = GraphBuilder(seed, fold_bn=fold_bn, name="resnet50")
    x = g.input("input", [batch, 3, hw, hw])
    y = g.relu(g.conv_bn(x, widths[0], 7, 2, 3))
    y = g.maxpool(y, 3, 2, 1)
    for stage, (w, d) in enumerate(zip(widths, depths)):
        for blk in range(d):
            stride = 2 if (blk == 0 and stage > 0) else 1
            identity = y
            z = g.relu(g.conv_bn(y, w, 1))
            z = g.relu(g.conv_bn(z, w, 3, stride, 1))
            z = g.conv_bn(z, w * 4, 1, gamma_scale=0.5, gain=1.0)  # damped last BN keeps the residual sum O(1)
            if blk == 0:
                identity = g.conv_bn(y, w * 4, 1, stride, gain=1.0)
            y = g.relu(g.add(z, identity))
    y = g.flatten(g.gap(y))
    y = g.gemm(y, num_classes)
    g.output(y, [batch, num_classes])
    return g.model()


def mobilenet_v2(seed: int = 0, fold_bn: bool = True, batch: int = 1, num_classes: int = 1000, hw: int = 224, width_div: int = 1) -> op.Model:
    """torchvision MobileNetV2: 52 Conv (17 depthwise) + 35 Clip + 10 Add + GAP + Flatten + Gemm.
    width_div > 1 shrinks every channel count (test fixtures only)."""
    g = GraphBuilder(seed, fold_bn=fold_bn, name="mobilenet_v2")
    x = g.input("input", [batch, 3, hw, hw])
    d = width_div
    y = g.clip(g.conv_bn(x, 32 // d, 3, 2, 1))
    cfg = [(1, 16 // d, 1, 1), (6, 24 // d, 2, 2), (6, 32 // d, 3, 2), (6, 64 // d, 4, 2), (6, 96 // d, 3, 1), (6, 160 // d, 3, 2), (6, 320 // d, 1, 1)]
    c_in = 32 // d
    for t, c, n, s in cfg:
        for i in range(n):
            stride = s if i == 0 else 1
            z = y
            hidden = c_in * t
            if t != 1:
                z = g.clip(g.conv_bn(z, hidden, 1))
            z = g.clip(g.conv_bn(z, hidden, 3, stride, 1, groups=hidden))
            z = g.conv_bn(z, c, 1, gain=1.0, gamma_scale=0.7)
            y = g.add(y, z) if (stride == 1 and c_in == c) else z
            c_in = c
    y = g.clip(g.conv_bn(y, 1280 // d, 1))
    y = g.flatten(g.gap(y))
    y = g.gemm(y, num_classes)
    g.output(y, [batch, num_classes])
    return g.model()


def transformer_net(seed: int = 0, batch: int = 1, hw: int = 512, width_div: int = 1) -> op.Model:
    """pytorch/examples fast_neural_style TransformerNet: 16 Pad + 16 Conv + 15 InstanceNorm + 10 Relu + 5 Add + 2 Upsample.
    width_div > 1 shrinks every channel count (test fixtures only)."""
    g = GraphBuilder(seed, name="transformer_net")
    x = g.input("input", [batch, 3, hw, hw])
    c32, c64, c128 = 32 // width_div, 64 // width_div, 128 // width_div

    def conv_layer(v, c, k, s, gain=2.0):
        return g.conv(g.pad(v, k // 2), c, k, s, 0, gain=gain)

    y = g.relu(g.instancenorm(conv_layer(x, c32, 9, 1)))
    y = g.relu(g.instancenorm(conv_layer(y, c64, 3, 2)))
    y = g.relu(g.instancenorm(conv_layer(y, c128, 3, 2)))
    for _ in range(5):
        z = g.relu(g.instancenorm(conv_layer(y, c128, 3, 1)))
        z = g.instancenorm(conv_layer(z, c128, 3, 1, gain=1.0))
        y = g.add(z, y)
    y = g.relu(g.instancenorm(conv_layer(g.upsample(y, 2), c64, 3, 1)))
    y = g.relu(g.instancenorm(conv_layer(g.upsample(y, 2), c32, 3, 1)))
    y = conv_layer(y, 3, 9, 1, gain=1.0)
    g.output(y, [batch, 3, hw, hw])
    return g.model()


def synthetic_ops(seed: int = 0, hw: int = 12, c: int = 16) -> op.Model:
    """Sigmoid / Concat / Reshape / Softmax / AveragePool path named by north_star but absent from the three CNNs."""
    g = GraphBuilder(seed, name="synthetic_ops")
    x = g.input("input", [1, c, hw, hw])
    a = g.sigmoid(g.conv(x, 24, 3, 1, 1))
    b = g.relu(g.conv(x, 8, 1))
    y = g.concat([a, b])                       # 32 channels
    y = g.avgpool(y, 2, 2, 0)
    y = g.conv(y, 10, 1)
    y = g.flatten(g.gap(y))
    y = g.reshape(y, [-1, 10, 1, 1], 10)
    y = g.softmax(y)
    g.output(y, [1, 10, 1, 1])
    return g.model()


def decoder_ops(seed: int = 0, hw: int = 10, c: int = 32) -> op.Model:
    """ConvTranspose / custom_group_norm / Pow: the registry entries (ONNXGraph.swift:116,143,154) no BASELINE model uses."""
    g = GraphBuilder(seed, name="decoder_ops")
    x = g.input("input", [1, c, hw, hw])
    y = g.relu(g.group_norm(g.conv(x, 64, 3, 1, 1), 8))
    y = g.relu(g.conv_transpose(y, 48, 3, stride=2, pad=1, output_padding=1))   # hw -> 2 hw
    y = g.pow(g.sigmoid(g.conv_transpose(y, 24, 4, stride=2, pad=1)), 2.0)      # 2 hw -> 4 hw
    y = g.conv_transpose(y, 8, 1, stride=1, pad=0)                              # 1x1: the tiled path
    g.output(y, [1, 8, 4 * hw, 4 * hw])
    return g.model()


def count_ops(model: op.Model) -> dict:
    out = {}
    for n in model.graph.node:
        out[n.op_type] = out.get(n.op_type, 0) + 1
    return out


def macs(model: op.Model, input_shape: Sequence[int]) -> int:
    """Algorithmic multiply-accumulates per batch for Conv + Gemm (the metric's FLOP numerator is 2x this)."""
    from .shape_infer import infer_shapes  # local import: shape_infer uses this module's codec only

    shapes = infer_shapes(model, {model.graph.input[0].name: tuple(input_shape)})
    inits = model.initializers()
    total = 0
    for n in model.graph.node:
        if n.op_type == "Conv":
            w = inits[n.input[1]]
            out = shapes[n.output[0]]
            total += int(np.prod(out)) * int(np.prod(w.dims[1:]))
        elif n.op_type == "Gemm":
            w = inits[n.input[1]]
            total += shapes[n.output[0]][0] * int(np.prod(w.dims))
    return total
