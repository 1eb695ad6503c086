"""The Python wire codec (host tooling) against hand-built bytes, itself, and a real torch ONNX export."""
import io
import struct

import numpy as np
import pytest
import torch

from smelter_b200 import modelzoo, onnx_proto as op


def test_roundtrip_preserves_everything():
    m = modelzoo.synthetic_ops(seed=2)
    again = op.Model.parse(m.serialize())
    assert again.serialize() == m.serialize()
    assert again.producer_name == "smelter_b200.modelzoo" and again.opset_import == [("", 9)]
    assert [n.op_type for n in again.graph.node] == [n.op_type for n in m.graph.node]
    assert again.graph.input[0].dims == [1, 16, 12, 12]


def test_hand_built_bytes():
    # TensorProto{dims:[2], data_type:FLOAT, float_data:[1.5, -2.0] (packed), name:"t"}
    raw = bytes([0x0A, 0x01, 0x02, 0x10, 0x01, 0x22, 0x08]) + struct.pack("<2f", 1.5, -2.0) + bytes([0x42, 0x01]) + b"t"
    t = op.Tensor.parse(memoryview(raw))
    assert t.dims == [2] and t.data_type == op.FLOAT and t.name == "t" and t.float_data == [1.5, -2.0]
    assert t.numpy().tolist() == [1.5, -2.0]
    # unpacked repeated varints + negative int64 (10-byte varint)
    raw = bytes([0x08, 0x03, 0x10, 0x07, 0x38]) + bytes([0xFF] * 9 + [0x01]) + bytes([0x38, 0x05])
    t = op.Tensor.parse(memoryview(raw))
    assert t.dims == [3] and t.int64_data == [-1, 5]
    with pytest.raises(ValueError):
        op.Model.parse(bytes([0x3A, 0x7F, 0x00]))


def test_dim_param_inputs():
    v = op.ValueInfo("x", op.FLOAT, ["batch", 3, 224, 224])
    assert op.ValueInfo.parse(memoryview(v.serialize())).dims == ["batch", 3, 224, 224]


def _torch_export(module, x, **kw):
    """torch's legacy exporter emits the bytes before its onnx-package hook; neutralise the hook."""
    try:
        from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    except Exception:  # pragma: no cover
        pytest.skip("torch exporter internals moved")
    saved = onnx_proto_utils._add_onnxscript_fn
    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    try:
        f = io.BytesIO()
        torch.onnx.export(module, x, f, opset_version=9, dynamo=False, **kw)
        return f.getvalue()
    except Exception as e:  # pragma: no cover
        pytest.skip(f"torch legacy exporter unavailable: {e}")
    finally:
        onnx_proto_utils._add_onnxscript_fn = saved


def test_parses_a_real_torch_export_and_conventions_match_modelzoo():
    import torch.nn as nn

    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.BatchNorm2d(8), nn.ReLU(), nn.MaxPool2d(3, 2, 1), nn.ReflectionPad2d(1),
                        nn.Conv2d(8, 8, 3), nn.InstanceNorm2d(8, affine=True), nn.Upsample(scale_factor=2, mode="nearest"),
                        nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(8, 5)).eval()
    with torch.no_grad():
        net[1].running_mean.normal_(0, 0.1); net[1].running_var.uniform_(0.5, 1.5); net[1].weight.uniform_(0.5, 1.5); net[1].bias.normal_(0, 0.1)
        net[6].weight.uniform_(0.5, 1.5); net[6].bias.normal_(0, 0.1)
    x = torch.rand(1, 3, 32, 32)
    data = _torch_export(net, x, do_constant_folding=True)  # folds eval-mode BN into the Conv and the Upsample scales into a Constant
    m = op.Model.parse(data)
    ops = [n.op_type for n in m.graph.node]
    assert set(ops) <= {"Conv", "BatchNormalization", "Relu", "MaxPool", "Pad", "InstanceNormalization", "Upsample", "Constant",
                        "GlobalAveragePool", "Flatten", "Gemm"}, ops
    by = {n.op_type: n for n in m.graph.node}
    assert {a.name for a in by["Conv"].attribute} == {"dilations", "group", "kernel_shape", "pads", "strides"}
    assert {a.name for a in by["Gemm"].attribute} >= {"alpha", "beta", "transB"}
    assert {a.name for a in by["MaxPool"].attribute} >= {"kernel_shape", "pads", "strides"}
    assert by["Pad"].attr("mode").s == b"reflect" and len(by["Pad"].attr("pads").ints) == 8
    assert by["Upsample"].attr("mode").s == b"nearest" and len(by["Upsample"].input) == 2
    # and the oracle agrees with torch eager on the exported graph
    from oracle.onnx_interp import Interpreter

    with torch.no_grad():
        want = net(x)
    assert torch.allclose(Interpreter(data).run(x), want, atol=1e-5)
