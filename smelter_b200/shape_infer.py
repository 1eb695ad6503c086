"""NCHW shape walk over the op set the engine supports (host tooling: FLOP accounting, ONNX2MPS value-infos).

Size formulas restate Sources/Smelter/Padding/ONNXConvolutionPadding.swift:91-113 (with the dilation term the
reference omits) and Padding/PyTorchPoolPadding.swift:94-103 (floor mode).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

from . import onnx_proto as op


def conv_out(i: int, k: int, s: int, d: int, p0: int, p1: int) -> int:
    return (i + p0 + p1 - (d * (k - 1) + 1)) // s + 1


def pool_out(i: int, k: int, s: int, p: int) -> int:
    return int(float(i + 2 * p - k) / float(s) + 1.0)


def infer_shapes(model: op.Model, inputs: Dict[str, Tuple[int, ...]]) -> Dict[str, Tuple[int, ...]]:
    inits = model.initializers()
    mps = model.producer_name == "ONNX2MPS"
    shapes: Dict[str, Tuple[int, ...]] = dict(inputs)
    consts: Dict[str, np.ndarray] = {}
    for n in model.graph.node:
        a = {x.name: x for x in n.attribute}
        t = n.op_type
        if t == "Constant":
            consts[n.output[0]] = a["value"].t.numpy()
            continue
        x = shapes[n.input[0]]
        if t == "Conv":
            w = inits[n.input[1]].dims
            co, kh, kw = (w[0], w[1], w[2]) if mps else (w[0], w[2], w[3])
            st = a["strides"].ints if "strides" in a else [1, 1]
            dl = a["dilations"].ints if "dilations" in a else [1, 1]
            pd = a["pads"].ints if "pads" in a else [0, 0, 0, 0]
            shapes[n.output[0]] = (x[0], co, conv_out(x[2], kh, st[0], dl[0], pd[0], pd[2]), conv_out(x[3], kw, st[1], dl[1], pd[1], pd[3]))
        elif t == "ConvTranspose":
            w = inits[n.input[1]].dims
            co, kh, kw = (w[0], w[1], w[2]) if mps else (w[1], w[2], w[3])
            st = a["strides"].ints if "strides" in a else [1, 1]
            dl = a["dilations"].ints if "dilations" in a else [1, 1]
            pd = a["pads"].ints if "pads" in a else [0, 0, 0, 0]
            opd = a["output_padding"].ints if "output_padding" in a else [0, 0]
            shapes[n.output[0]] = (x[0], co, (x[2] - 1) * st[0] - pd[0] - pd[2] + dl[0] * (kh - 1) + 1 + opd[0],
                                   (x[3] - 1) * st[1] - pd[1] - pd[3] + dl[1] * (kw - 1) + 1 + opd[1])
        elif t == "Gemm":
            w = inits[n.input[1]].dims
            trans_b = a["transB"].i if "transB" in a else 0
            shapes[n.output[0]] = (x[0], w[0] if trans_b else w[1])
        elif t in ("MaxPool", "AveragePool"):
            k, st, pd = a["kernel_shape"].ints, a["strides"].ints, a["pads"].ints
            shapes[n.output[0]] = (x[0], x[1], pool_out(x[2], k[0], st[0], pd[0]), pool_out(x[3], k[1], st[1], pd[1]))
        elif t == "GlobalAveragePool":
            shapes[n.output[0]] = (x[0], x[1], 1, 1)
        elif t == "Flatten":
            shapes[n.output[0]] = (x[0], int(np.prod(x[1:])))
        elif t == "Reshape":
            tgt = inits[n.input[1]].numpy() if n.input[1] in inits else consts[n.input[1]]
            tgt = [int(round(float(v))) for v in tgt.reshape(-1)]
            tgt = [x[i] if v == 0 else v for i, v in enumerate(tgt)]
            if -1 in tgt:
                known = int(np.prod([v for v in tgt if v != -1]))
                tgt[tgt.index(-1)] = int(np.prod(x)) // known
            shapes[n.output[0]] = tuple(tgt)
        elif t == "Pad":
            p = a["pads"].ints
            shapes[n.output[0]] = tuple(x[i] + p[i] + p[i + len(x)] for i in range(len(x)))
        elif t == "Upsample":
            if "scales" in a:
                sc = a["scales"].floats
            else:
                src = inits[n.input[1]].numpy() if n.input[1] in inits else consts[n.input[1]]
                sc = [float(v) for v in src.reshape(-1)]
            shapes[n.output[0]] = (x[0], x[1], x[2] * int(sc[2]), x[3] * int(sc[3]))
        elif t == "Concat":
            shapes[n.output[0]] = (x[0], sum(shapes[i][1] for i in n.input)) + tuple(x[2:])
        else:  # elementwise / normalisation / softmax / identity
            shapes[n.output[0]] = x
    return shapes
