"""ORACLE — TEST INFRASTRUCTURE ONLY.  fp32 torch-CPU interpreter of the ONNX graphs on Smelter's inference path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
nothing under smelter_b200/ does.  It is the checker, never the thing shipped.

What it restates.  In the reference every FLOP happens inside Apple's closed MetalPerformanceShaders (`MPSCNN*Node`
call sites in Sources/Smelter/Converters.swift, graph assembly in ONNXGraph.swift:169-193, execution by
`MPSNNGraph.encode(to:sourceImages:)`, README.md:43-44) — the arithmetic is not in the reference tree, so the
oracle is the ONNX operator specification evaluated in fp32, standing in for the
"onnxruntime-CPU" BASELINE.json names (onnxruntime is not installed in this image and cannot be: no network).
Per-op citations give the reference converter whose MPS node the op replaces:

  Conv / Gemm            ConvolutionConverter        Converters.swift:187-338
  BatchNormalization     BatchNormalizationConverter :797-827  (epsilon honoured — ONNX spec; reference ignores it, Q10)
  InstanceNormalization  InstanceNormConverter       :992-1017
  Relu/Sigmoid/...       :342-359, :466-476, :386-428, :1056-1175
  Clip                   (extension; MobileNetV2 ReLU6)
  Add/Sub/Mul/Div        :430-464, :1177-1211
  MaxPool/AveragePool    :607-695 + Padding/PyTorchPoolPadding.swift:94-103 (floor mode, symmetric pads[0..1])
  GlobalAveragePool      :578-605
  Upsample               :478-552 (integer scales; bilinear honours Configuration.alignCorners, ONNXGraph.swift:20)
  Concat                 :554-574 (ONNX semantics: channel sum; reference hard-codes 2x, Q11)
  Reshape/Flatten        :830-915 (ONNX semantics on NCHW)
  Softmax/LogSoftmax     :697-714, :1213-1231 (axis 1)
  Pad                    :942-989 (constant/reflect/edge; `value` honoured)
  ConvTranspose          ConvolutionConverter :266-287 (+ Array+Extensions.swift:70-76 flip; ONNX2MPS.py:54-79 swizzle undone)
  custom_group_norm      GroupNormConverter :1273-1300;  Pow  PowConverter :1160-1175 (ONNX meaning: scalar exponent)
  Constant / Dropout / Identity   :716-727, :918-939

The `.mpsFlavor` format (producer_name == "ONNX2MPS", ONNXGraph.swift:98-103) is honoured: conv weights are OHWI
(ONNX2MPS.py:75) and every initializer may be fp16, including shape tensors (Onnx_TensorProto+Extensions.swift:29-30).

PARITY STATUS: **parity unpinned** against the reference itself (it has no tests, golden vectors or runnable
backend here).  The interpreter is pinned instead against independent torch eager `nn.Module` forwards of the
same architectures (tests/test_oracle.py) and against the committed fixtures in tests/golden/.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from smelter_b200 import onnx_proto as op


def _ints(t: np.ndarray) -> List[int]:
    # `.integers` (Onnx_TensorProto+Extensions.swift:2-34): Int(Float) truncates toward zero
    return [int(v) for v in np.trunc(t.astype(np.float64)).reshape(-1)]


class Interpreter:
    def __init__(self, model_bytes: bytes, align_corners: bool = True, dtype: torch.dtype = torch.float32):
        self.model = op.Model.parse(model_bytes)
        self.mps = self.model.producer_name == "ONNX2MPS"
        self.align_corners = align_corners
        self.dtype = dtype
        self.inits: Dict[str, np.ndarray] = {t.name: t.numpy() for t in self.model.graph.initializer}
        self._weights: Dict[str, torch.Tensor] = {}
        self.input_names = [v.name for v in self.model.graph.input if v.name not in self.inits]

    def _w(self, name: str) -> torch.Tensor:
        if name not in self._weights:
            self._weights[name] = torch.from_numpy(self.inits[name].astype(np.float32)).to(self.dtype)
        return self._weights[name]

    def _conv_weight(self, name: str) -> torch.Tensor:
        key = name + "#oihw"
        if key not in self._weights:
            w = self._w(name)
            if self.mps:
                w = w.permute(0, 3, 1, 2).contiguous()  # OHWI -> OIHW (inverse of ONNX2MPS.py:75)
            self._weights[key] = w
        return self._weights[key]

    @torch.no_grad()
    def run(self, *inputs: torch.Tensor, keep: Optional[Sequence[str]] = None) -> torch.Tensor:
        """Evaluate the graph.  Returns the single graph output (ONNXGraph.swift:178-180)."""
        if len(inputs) != len(self.input_names):
            raise ValueError(f"graph expects {len(self.input_names)} input(s)")
        env: Dict[str, torch.Tensor] = {n: x.to(self.dtype) for n, x in zip(self.input_names, inputs)}
        consts: Dict[str, np.ndarray] = {}
        kept: Dict[str, torch.Tensor] = {}

        def host(name: str) -> np.ndarray:
            return consts[name] if name in consts else self.inits[name]

        for n in self.model.graph.node:
            a = {x.name: x for x in n.attribute}
            t = n.op_type
            if t == "Constant":
                consts[n.output[0]] = a["value"].t.numpy()
                continue
            if t == "Identity" and n.input[0] not in env and (n.input[0] in self.inits or n.input[0] in consts):
                # torch's exporter de-duplicates equal initializers (e.g. all-zero biases) through Identity nodes: an alias of a weight
                self.inits[n.output[0]] = host(n.input[0])
                continue
            x = env[n.input[0]]
            if t == "Conv":
                w = self._conv_weight(n.input[1])
                b = self._w(n.input[2]) if len(n.input) > 2 and n.input[2] else None
                st = a["strides"].ints if "strides" in a else [1, 1]
                dl = a["dilations"].ints if "dilations" in a else [1, 1]
                pd = a["pads"].ints if "pads" in a else [0, 0, 0, 0]
                grp = a["group"].i if "group" in a else 1
                if pd[0] != pd[2] or pd[1] != pd[3]:
                    x = F.pad(x, (pd[1], pd[3], pd[0], pd[2]))
                    pd = [0, 0, 0, 0]
                y = F.conv2d(x, w, b, stride=tuple(st), padding=(pd[0], pd[1]), dilation=tuple(dl), groups=grp)
            elif t == "ConvTranspose":  # Converters.swift:266-287; ONNX weight layout [Cin, Cout, kH, kW]
                w = self._w(n.input[1])
                if self.mps:  # ONNX2MPS.py:54-79: [Cout, kH, kW, Cin], spatially flipped -> undo both
                    w = w.flip(1, 2).permute(3, 0, 1, 2).contiguous()
                b = self._w(n.input[2]) if len(n.input) > 2 and n.input[2] else None
                st = a["strides"].ints if "strides" in a else [1, 1]
                dl = a["dilations"].ints if "dilations" in a else [1, 1]
                pd = a["pads"].ints if "pads" in a else [0, 0, 0, 0]
                op_ = a["output_padding"].ints if "output_padding" in a else [0, 0]
                if pd[0] != pd[2] or pd[1] != pd[3]:
                    raise NotImplementedError("asymmetric ConvTranspose pads")
                y = F.conv_transpose2d(x, w, b, stride=tuple(st), padding=(pd[0], pd[1]), output_padding=tuple(op_), dilation=tuple(dl))
            elif t == "custom_group_norm":  # Converters.swift:1273-1300: X, groups, gamma, beta
                groups = _ints(host(n.input[1]))[0]
                eps = a["epsilon"].f if "epsilon" in a else 1e-5
                y = F.group_norm(x, groups, weight=self._w(n.input[2]), bias=self._w(n.input[3]), eps=eps)
            elif t == "Pow":
                e = float(host(n.input[1]).reshape(-1)[0]) if len(n.input) > 1 and n.input[1] else 1.0
                y = torch.pow(x, e)
            elif t == "Gemm":
                w = self._w(n.input[1])
                alpha = a["alpha"].f if "alpha" in a else 1.0
                beta = a["beta"].f if "beta" in a else 1.0
                if "transA" in a and a["transA"].i:
                    raise NotImplementedError("Gemm transA")
                x2 = x.reshape(x.shape[0], -1)
                wt = w if ("transB" in a and a["transB"].i) else w.t()
                y = alpha * (x2 @ wt.t())
                if len(n.input) > 2 and n.input[2]:
                    y = y + beta * self._w(n.input[2])
            elif t == "BatchNormalization":
                g, b, m, v = (self._w(i) for i in n.input[1:5])
                eps = a["epsilon"].f if "epsilon" in a else 1e-5
                y = F.batch_norm(x, m, v, g, b, training=False, eps=eps)
            elif t == "InstanceNormalization":
                eps = a["epsilon"].f if "epsilon" in a else 1e-5
                y = F.instance_norm(x, weight=self._w(n.input[1]), bias=self._w(n.input[2]), eps=eps)
            elif t == "Relu":
                y = F.relu(x)
            elif t == "Sigmoid":
                y = torch.sigmoid(x)
            elif t == "Tanh":
                y = torch.tanh(x)
            elif t == "Abs":
                y = torch.abs(x)
            elif t == "Exp":
                y = torch.exp(x)
            elif t == "Log":
                y = torch.log(x)
            elif t == "Softplus":
                y = F.softplus(x)
            elif t == "Softsign":
                y = F.softsign(x)
            elif t == "Elu":
                y = F.elu(x, alpha=a["alpha"].f if "alpha" in a else 1.0)
            elif t == "HardSigmoid":
                al = a["alpha"].f if "alpha" in a else 0.2
                be = a["beta"].f if "beta" in a else 0.5
                y = torch.clamp(al * x + be, 0.0, 1.0)
            elif t == "PRelu":
                slope = float(host(n.input[1]).reshape(-1)[0])  # single scalar slope, Converters.swift:372-374
                y = torch.where(x > 0, x, x * slope)
            elif t == "Clip":
                lo = a["min"].f if "min" in a else -3.402823466e38
                hi = a["max"].f if "max" in a else 3.402823466e38
                if len(n.input) > 1 and n.input[1]:
                    lo = float(host(n.input[1]).reshape(-1)[0])
                if len(n.input) > 2 and n.input[2]:
                    hi = float(host(n.input[2]).reshape(-1)[0])
                y = torch.clamp(x, lo, hi)
            elif t in ("Add", "Sub", "Mul", "Div"):
                z = env[n.input[1]]
                y = {"Add": torch.add, "Sub": torch.sub, "Mul": torch.mul, "Div": torch.div}[t](x, z)
            elif t in ("MaxPool", "AveragePool"):
                k = a["kernel_shape"].ints
                st = a["strides"].ints if "strides" in a else [1, 1]          # ONNX defaults
                pd = a["pads"].ints if "pads" in a else [0, 0, 0, 0]
                if t == "MaxPool":
                    y = F.max_pool2d(x, tuple(k), tuple(st), (pd[0], pd[1]), ceil_mode=False)
                else:
                    y = F.avg_pool2d(x, tuple(k), tuple(st), (pd[0], pd[1]), ceil_mode=False, count_include_pad=True)
            elif t == "GlobalAveragePool":
                y = x.mean(dim=(2, 3), keepdim=True)
            elif t == "ReduceMean":  # extension (not in the reference registry): spatial mean only
                axes = sorted(v % 4 for v in (a["axes"].ints if "axes" in a else _ints(host(n.input[1]))))
                assert axes == [2, 3], axes
                y = x.mean(dim=(2, 3), keepdim=True)
            elif t == "Flatten":
                y = x.reshape(x.shape[0], -1)
            elif t == "Reshape":
                tgt = _ints(host(n.input[1]))
                tgt = [x.shape[i] if v == 0 else v for i, v in enumerate(tgt)]
                y = x.reshape(tgt)
            elif t == "Softmax":
                y = torch.softmax(x, dim=1)
            elif t == "LogSoftmax":
                y = torch.log_softmax(x, dim=1)
            elif t == "Upsample":
                if "scales" in a:
                    sc = [int(v) for v in a["scales"].floats]       # Converters.swift:498-499 (truncation)
                else:
                    sc = _ints(host(n.input[1]))                    # :505-514
                mode = a["mode"].s.decode()
                if mode == "nearest":
                    y = x.repeat_interleave(sc[2], dim=2).repeat_interleave(sc[3], dim=3)
                elif self.align_corners:
                    y = F.interpolate(x, scale_factor=(sc[2], sc[3]), mode="bilinear", align_corners=True)
                else:
                    y = _bilinear_asymmetric(x, sc[2], sc[3])
            elif t == "Concat":
                y = torch.cat([env[i] for i in n.input], dim=1)
            elif t == "Pad":
                p = a["pads"].ints
                mode = a["mode"].s.decode() if "mode" in a else "constant"
                value = a["value"].f if "value" in a else 0.0
                tp = (p[3], p[7], p[2], p[6])
                if mode == "constant":
                    y = F.pad(x, tp, mode="constant", value=value)
                else:
                    y = F.pad(x, tp, mode={"reflect": "reflect", "edge": "replicate"}[mode])
            elif t in ("Dropout", "Identity"):
                y = x
            else:
                raise NotImplementedError(f"unknownNodeOpType({t})")  # ONNXGraph.swift:173-174
            env[n.output[0]] = y
            if keep and n.output[0] in keep:
                kept[n.output[0]] = y
        outs = self.model.graph.output
        if len(outs) != 1:
            raise ValueError("unsupportedOutput")
        self.kept = kept
        return env[outs[0].name]


def _bilinear_asymmetric(x: torch.Tensor, sh: int, sw: int) -> torch.Tensor:
    """ONNX opset-9 Upsample 'linear' (asymmetric coordinates: src = dst / scale, clamped at the far edge)."""
    n, c, h, w = x.shape
    ys = torch.arange(h * sh, dtype=torch.float32) / sh
    xs = torch.arange(w * sw, dtype=torch.float32) / sw
    y0 = ys.floor().long().clamp(max=h - 1)
    x0 = xs.floor().long().clamp(max=w - 1)
    y1 = (y0 + 1).clamp(max=h - 1)
    x1 = (x0 + 1).clamp(max=w - 1)
    wy = (ys - y0).view(1, 1, -1, 1).to(x.dtype)
    wx = (xs - x0).view(1, 1, 1, -1).to(x.dtype)
    top = x[:, :, y0][:, :, :, x0] * (1 - wx) + x[:, :, y0][:, :, :, x1] * wx
    bot = x[:, :, y1][:, :, :, x0] * (1 - wx) + x[:, :, y1][:, :, :, x1] * wx
    return top * (1 - wy) + bot * wy


def run_model(model_bytes: bytes, x: torch.Tensor, align_corners: bool = True) -> torch.Tensor:
    return Interpreter(model_bytes, align_corners=align_corners).run(x)
