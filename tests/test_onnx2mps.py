"""ONNX2MPS restatement (smelter_b200/onnx2mps.py) — BASELINE.json configs[0]: parse + fuse on the host CPU."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.fusion import fuse_conv_bn_eval

from oracle.onnx_interp import Interpreter
from smelter_b200 import modelzoo, onnx2mps, onnx_proto as op


def _t(model, name):
    return torch.from_numpy(model.initializers()[name].numpy().astype(np.float32))


def test_config1_fold_matches_torch_fuse_conv_bn_eval():
    m = modelzoo.conv_bn_relu(seed=0)
    conv_n, bn_n, _ = m.graph.node
    conv, bn = nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8, eps=bn_n.attr("epsilon").f)
    with torch.no_grad():
        conv.weight.copy_(_t(m, conv_n.input[1])); conv.bias.copy_(_t(m, conv_n.input[2]))
        bn.weight.copy_(_t(m, bn_n.input[1])); bn.bias.copy_(_t(m, bn_n.input[2]))
        bn.running_mean.copy_(_t(m, bn_n.input[3])); bn.running_var.copy_(_t(m, bn_n.input[4]))
    fused = fuse_conv_bn_eval(conv.eval(), bn.eval())
    out = op.Model.parse(onnx2mps.convert_bytes(m.serialize(), half=False))
    assert out.producer_name == "ONNX2MPS" and out.producer_version == "1.0.0"           # ONNX2MPS.py:97-99
    assert [n.op_type for n in out.graph.node] == ["Conv", "Relu"]
    w, b = (out.initializers()[i] for i in out.graph.node[0].input[1:3])
    assert w.dims == [8, 3, 3, 3] and w.data_type == op.FLOAT                             # OHWI (same numbers here: 3x3x3)
    assert np.abs(w.numpy() - fused.weight.detach().numpy().transpose(0, 2, 3, 1)).max() <= 1e-6
    assert np.abs(b.numpy() - fused.bias.detach().numpy()).max() <= 1e-6
    assert out.graph.value_info == []                                                     # dropped like the reference


def test_half_casts_every_initializer_and_io_info():
    m = modelzoo.synthetic_ops(seed=2)                      # has an int64 Reshape shape tensor
    out = op.Model.parse(onnx2mps.convert_bytes(m.serialize(), half=True))
    assert all(t.data_type == op.FLOAT16 for t in out.graph.initializer)
    assert all(v.elem_type == op.FLOAT16 for v in out.graph.input + out.graph.output)
    shape_t = [t for t in out.graph.initializer if t.name.startswith("shape")][0]
    assert shape_t.numpy().tolist() == [-1.0, 10.0, 1.0, 1.0]
    x = torch.rand(1, 16, 12, 12)
    a, b = Interpreter(m.serialize()).run(x), Interpreter(out.serialize()).run(x)
    assert torch.allclose(a.reshape(-1), b.reshape(-1), atol=2e-3)


def test_conv_weights_are_rounded_like_numpy_astype():
    m = modelzoo.resnet50(seed=1, fold_bn=False, num_classes=10, hw=32, widths=(8, 16, 32, 64), depths=(1, 1, 1, 1))
    fp32 = op.Model.parse(onnx2mps.convert_bytes(m.serialize(), half=False))
    fp16 = op.Model.parse(onnx2mps.convert_bytes(m.serialize(), half=True))
    assert modelzoo.count_ops(fp16) == {"Conv": 17, "Relu": 13, "MaxPool": 1, "Add": 4, "GlobalAveragePool": 1, "Flatten": 1, "Gemm": 1}
    for a, b in zip(fp32.graph.initializer, fp16.graph.initializer):
        assert a.name == b.name and a.dims == b.dims
        assert np.array_equal(a.numpy().astype(np.float16).view(np.uint16), b.numpy().view(np.uint16))
    # fold + layout preserved the function
    x = torch.rand(1, 3, 32, 32)
    assert torch.allclose(Interpreter(m.serialize()).run(x), Interpreter(fp32.serialize()).run(x), atol=1e-5)


def test_shared_or_multi_consumer_convs_are_not_folded():
    b = modelzoo.GraphBuilder(seed=0)
    x = b.input("input", [1, 4, 8, 8])
    c = b.conv(x, 4, 3, 1, 1)
    y = b.add(b.bn(c), c)            # conv output has two consumers
    b.output(y, [1, 4, 8, 8])
    out = op.Model.parse(onnx2mps.convert_bytes(b.model().serialize()))
    assert [n.op_type for n in out.graph.node] == ["Conv", "BatchNormalization", "Add"]


def test_convtranspose_swizzle_and_flip(host_oracle):
    g = op.Graph(name="t")
    w = np.random.default_rng(0).standard_normal((6, 4, 3, 2)).astype(np.float32)  # [Cin, Cout, kH, kW]
    g.initializer.append(op.Tensor.from_numpy("w", w))
    g.input.append(op.ValueInfo("x", op.FLOAT, [1, 6, 5, 5]))
    g.node.append(op.Node("ConvTranspose", ["x", "w"], ["y"], attribute=[op.attr("kernel_shape", [3, 2])]))
    g.output.append(op.ValueInfo("y", op.FLOAT, [1, 4, 7, 6]))
    out = op.Model.parse(onnx2mps.convert_bytes(op.Model(graph=g).serialize()))
    got = out.initializers()["w"].numpy()
    assert got.shape == (4, 3, 2, 6)
    assert np.array_equal(got, w.transpose(1, 2, 3, 0)[:, ::-1, ::-1, :])               # ONNX2MPS.py:58-62
    want = np.empty(w.size, np.float32)                                                   # == Array+Extensions.swift:70-76
    host_oracle.oracle_reformat_conv_weight(w.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), 4, 4, 6, 3, 2, 1)
    assert np.array_equal(got.reshape(-1), want)


def test_cli_matches_the_reference_flags(tmp_path):
    src, dst = tmp_path / "in.onnx", tmp_path / "out.onnx"
    src.write_bytes(modelzoo.conv_bn_relu(seed=0).serialize())
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-m", "smelter_b200.onnx2mps", "--half", "--input", str(src), "--output", str(dst)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "Success" in r.stdout
    assert op.load(str(dst)).producer_name == "ONNX2MPS"


def test_check_model_rejects_unsorted_graphs():
    m = modelzoo.conv_bn_relu(seed=0)
    m.graph.node.reverse()
    try:
        onnx2mps.optimize_model(m)
    except ValueError as e:
        assert "topologically" in str(e)
    else:
        raise AssertionError("expected a ValueError")
