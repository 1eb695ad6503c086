"""ONNX2MPS restated without the `onnx` package (which is not installable here).

Follows /ONNX2MPS.py of the reference step by step:
  optimize_model (:104-109)   check -> strip doc strings -> `fuse_bn_into_conv` -> modelToMPS
  modelToMPS     (:70-101)    swizzle plan per Conv / ConvTranspose weight (:73-79), cast *every* initializer and the
                              graph input/output infos to the float type (:14-31, :84-96), transpose Conv weights
                              [0,2,3,1] and ConvTranspose weights [1,2,3,0] + 180 degree spatial flip (:54-67),
                              stamp producer_name='ONNX2MPS', producer_version='1.0.0' (:97-99)
  main           (:112-134)   --half / --input / --output

`fuse_bn_into_conv` lived in `onnx.optimizer` (removed from onnx >= 1.9, never pinned by the reference); its
published algorithm is restated in fuse_bn_into_conv() below:  s = gamma / sqrt(var + eps);
W' = W * s[:,None,None,None];  b' = (b - mean) * s + beta  (b = 0 when the Conv has no bias).

Deviations (documented): value_info is dropped exactly as the reference does (make_graph is called without it);
the opset_import of the source model is preserved (the reference re-stamps whatever opset its installed `onnx`
defaults to; the engine never reads it).
"""
from __future__ import annotations

import argparse
import sys
from typing import Dict, List

import numpy as np

from . import onnx_proto as op

DATA_TYPES = {np.float32: op.FLOAT, np.float16: op.FLOAT16}  # ONNX2MPS.py:8-11


def fuse_bn_into_conv(model: op.Model) -> int:
    """In-place Conv+BatchNormalization fold.  Returns the number of fused pairs.

    Preconditions of the onnx.optimizer pass: the Conv output feeds only the BN (and is not a graph output), the
    BN's four parameters and the Conv's weight (and bias) are initializers, the BN has a single used output."""
    g = model.graph
    inits = model.initializers()
    uses: Dict[str, int] = {}
    for n in g.node:
        for i in n.input:
            uses[i] = uses.get(i, 0) + 1
    for o in g.output:
        uses[o.name] = uses.get(o.name, 0) + 1
    producer = {o: n for n in g.node for o in n.output}
    fused = 0
    dead: List[op.Node] = []
    for bn in g.node:
        if bn.op_type != "BatchNormalization" or len(bn.input) < 5:
            continue
        conv = producer.get(bn.input[0])
        if conv is None or conv.op_type != "Conv" or uses.get(conv.output[0], 0) != 1:
            continue
        if any(name not in inits for name in bn.input[1:5]) or conv.input[1] not in inits:
            continue
        if len(conv.input) > 2 and conv.input[2] not in inits:
            continue
        if any(uses.get(o, 0) for o in bn.output[1:]):
            continue
        gamma, beta, mean, var = (inits[nm].numpy().astype(np.float32) for nm in bn.input[1:5])
        eps_attr = bn.attr("epsilon")
        eps = np.float32(eps_attr.f if eps_attr is not None else 1e-5)
        w_t = inits[conv.input[1]]
        w = w_t.numpy().astype(np.float32)
        s = gamma / np.sqrt(var + eps)
        b = inits[conv.input[2]].numpy().astype(np.float32) if len(conv.input) > 2 else np.zeros(w.shape[0], np.float32)
        w_new = (w * s.reshape((-1,) + (1,) * (w.ndim - 1))).astype(np.float32)
        b_new = ((b - mean) * s + beta).astype(np.float32)
        # weights may be shared between nodes: write fresh tensors, leave the originals to dead-initializer cleanup
        w_name, b_name = conv.output[0] + "_fused_w", conv.output[0] + "_fused_b"
        g.initializer.append(op.Tensor.from_numpy(w_name, w_new))
        g.initializer.append(op.Tensor.from_numpy(b_name, b_new))
        conv.input[:] = [conv.input[0], w_name, b_name]
        conv.output[0] = bn.output[0]
        dead.append(bn)
        fused += 1
    if dead:
        dead_ids = {id(n) for n in dead}
        g.node[:] = [n for n in g.node if id(n) not in dead_ids]
        live = {i for n in g.node for i in n.input}
        g.initializer[:] = [t for t in g.initializer if t.name in live]
        g.input[:] = [v for v in g.input if v.name in live or v.name not in inits]
    return fused


def tensors_to_type(tensors: List[op.Tensor], data_type) -> List[op.Tensor]:
    """ONNX2MPS.py:24-31 — every initializer (also int64 shape tensors) is cast with numpy astype."""
    return [op.Tensor.from_numpy(t.name, t.numpy().astype(data_type)) for t in tensors]


def tensor_infos_to_type(infos: List[op.ValueInfo], data_type) -> List[op.ValueInfo]:
    """ONNX2MPS.py:14-21 — dim_param dims become dim_value 0, as `dim.dim_value` reads them."""
    return [op.ValueInfo(name=v.name, elem_type=DATA_TYPES[data_type], dims=[d if isinstance(d, int) else 0 for d in (v.dims or [])])
            for v in infos]


def resolve_initializer_aliases(model: op.Model) -> int:
    """torch's exporter de-duplicates equal initializers (typically all-zero biases) through `Identity` nodes whose input is an
    initializer.  Point every reader at the original tensor and drop those nodes, so that the weight swizzle below and the engine's
    converters see plain initializers (onnx.optimizer's eliminate_identity does the same for the reference's script)."""
    inits = model.initializers()
    alias: Dict[str, str] = {}
    kept = []
    for n in model.graph.node:
        if n.op_type == "Identity" and n.input and (n.input[0] in inits or n.input[0] in alias):
            alias[n.output[0]] = alias.get(n.input[0], n.input[0])
            continue
        n.input = [alias.get(i, i) for i in n.input]
        kept.append(n)
    model.graph.node = kept
    return len(alias)


def model_to_mps(model: op.Model, data_type) -> op.Model:
    is_transpose: Dict[str, bool] = {}
    swizzle_plan: Dict[str, List[int]] = {}
    for node in model.graph.node:  # ONNX2MPS.py:73-79
        if node.op_type == "Conv":
            swizzle_plan[node.input[1]] = [0, 2, 3, 1]
            is_transpose[node.input[1]] = False
        if node.op_type == "ConvTranspose":
            swizzle_plan[node.input[1]] = [1, 2, 3, 0]
            is_transpose[node.input[1]] = True

    def swizzle_infos(infos: List[op.ValueInfo]) -> List[op.ValueInfo]:  # :39-51
        for v in infos:
            if v.name in swizzle_plan and v.dims is not None and len(v.dims) == 4:
                v.dims = [v.dims[i] for i in swizzle_plan[v.name]]
        return infos

    tensors = tensors_to_type(model.graph.initializer, data_type)
    for idx, t in enumerate(tensors):  # :54-67
        if t.name not in swizzle_plan:
            continue
        a = t.numpy().transpose(swizzle_plan[t.name])
        if is_transpose[t.name]:
            a = a[:, ::-1, ::-1, :]
        tensors[idx] = op.Tensor.from_numpy(t.name, np.ascontiguousarray(a))
    graph = op.Graph(node=model.graph.node, name=model.graph.name, initializer=tensors,
                     input=swizzle_infos(tensor_infos_to_type(model.graph.input, data_type)),
                     output=swizzle_infos(tensor_infos_to_type(model.graph.output, data_type)))
    return op.Model(ir_version=model.ir_version, producer_name="ONNX2MPS", producer_version="1.0.0", graph=graph,
                    opset_import=list(model.opset_import))


def check_model(model: op.Model) -> None:
    """The structural subset of onnx.checker.check_model that matters on this path."""
    names = set()
    inits = model.initializers()
    for v in model.graph.input:
        names.add(v.name)
    names.update(inits)
    for n in model.graph.node:
        if not n.op_type:
            raise ValueError("node without op_type")
        for i in n.input:
            if i and i not in names:
                raise ValueError(f"node '{n.name or n.op_type}' reads '{i}' before it is produced (graph must be topologically sorted)")
        names.update(n.output)
    for o in model.graph.output:
        if o.name not in names:
            raise ValueError(f"graph output '{o.name}' is never produced")


def optimize_model(model: op.Model, data_type=np.float32) -> op.Model:
    check_model(model)          # ONNX2MPS.py:105
    resolve_initializer_aliases(model)
    # strip_doc_string (:106): this codec never keeps doc strings, nothing to do
    fuse_bn_into_conv(model)    # :107
    return model_to_mps(model, data_type)  # :108


def convert_bytes(data: bytes, half: bool = False) -> bytes:
    return optimize_model(op.Model.parse(data), np.float16 if half else np.float32).serialize()


def main(argv=None) -> int:
    parser = argparse.ArgumentParser(description="Convert ONNX model to the ONNX2MPS flavour (BN folded, OHWI weights, optional fp16)")
    parser.add_argument("--half", required=False, help="Use FP16 weights", action="store_true")
    parser.add_argument("--input", required=True, help="Path to ONNX model")
    parser.add_argument("--output", required=True, help="Path to MPS model")
    args = parser.parse_args(argv)
    print("Parsed args")
    print("half: {}".format(args.half))
    print("input: {}".format(args.input))
    print("output: {}".format(args.output))
    print()
    print("Started convertion")
    with open(args.input, "rb") as f:
        data = f.read()
    out = convert_bytes(data, args.half)
    with open(args.output, "wb") as f:
        f.write(out)
    print("Success")
    return 0


if __name__ == "__main__":
    sys.exit(main())
