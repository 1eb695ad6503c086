"""Pins the ORACLE (test infrastructure) before it is trusted: the ONNX interpreter against independent torch eager
modules built from the same weights, and against the committed golden fixtures.  The reference has no golden vectors
(SURVEY.md §8c): parity against the reference itself stays unpinned."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle.onnx_interp import Interpreter
from smelter_b200 import modelzoo, onnx_proto as op

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_interpreter_reproduces_golden(path):
    z = np.load(path)
    y = Interpreter(z["model"].tobytes()).run(torch.from_numpy(z["x"].astype(np.float32))).numpy()
    assert np.abs(y - z["y"]).max() <= 1e-5


def _t(model, name):
    return torch.from_numpy(model.initializers()[name].numpy().astype(np.float32))


def test_conv_bn_relu_against_torch_modules():
    m = modelzoo.conv_bn_relu(seed=0)
    conv_n, bn_n, _ = m.graph.node
    conv = nn.Conv2d(3, 8, 3, padding=1)
    bn = nn.BatchNorm2d(8, eps=bn_n.attr("epsilon").f)
    with torch.no_grad():
        conv.weight.copy_(_t(m, conv_n.input[1]))
        conv.bias.copy_(_t(m, conv_n.input[2]))
        bn.weight.copy_(_t(m, bn_n.input[1]))
        bn.bias.copy_(_t(m, bn_n.input[2]))
        bn.running_mean.copy_(_t(m, bn_n.input[3]))
        bn.running_var.copy_(_t(m, bn_n.input[4]))
    net = nn.Sequential(conv, bn, nn.ReLU()).eval()
    x = torch.rand(2, 3, 16, 16)
    with torch.no_grad():
        want = net(x)
    assert torch.allclose(Interpreter(m.serialize()).run(x), want, atol=1e-6)


class _Bottleneck(nn.Module):
    def __init__(self, cin, w, stride, down):
        super().__init__()
        self.c1, self.c2, self.c3 = nn.Conv2d(cin, w, 1), nn.Conv2d(w, w, 3, stride, 1), nn.Conv2d(w, 4 * w, 1)
        self.down = nn.Conv2d(cin, 4 * w, 1, stride) if down else None

    def forward(self, x):
        z = self.c3(torch.relu(self.c2(torch.relu(self.c1(x)))))
        return torch.relu(z + (self.down(x) if self.down is not None else x))


def test_resnet_graph_against_independent_torch_module():
    """An eager nn.Module with torchvision's ResNet-50 v1.5 topology, loaded with the generated initializers in file order,
    must agree with the interpreter walking the ONNX graph (independent path: module topology vs graph walk)."""
    widths, depths = (8, 16, 32, 64), (2, 1, 2, 1)
    m = modelzoo.resnet50(seed=7, fold_bn=True, num_classes=10, hw=64, widths=widths, depths=depths)
    blocks, cin = [], widths[0]
    for stage, (w, d) in enumerate(zip(widths, depths)):
        for b in range(d):
            blocks.append(_Bottleneck(cin, w, 2 if (b == 0 and stage > 0) else 1, b == 0))
            cin = 4 * w
    stem = nn.Conv2d(3, widths[0], 7, 2, 3)
    fc = nn.Linear(cin, 10)
    convs = [stem]
    for blk in blocks:  # generation order inside a block: c1, c2, c3, (down)
        convs += [blk.c1, blk.c2, blk.c3] + ([blk.down] if blk.down is not None else [])
    conv_nodes = [n for n in m.graph.node if n.op_type == "Conv"]
    assert len(conv_nodes) == len(convs)
    with torch.no_grad():
        for mod, node in zip(convs, conv_nodes):
            mod.weight.copy_(_t(m, node.input[1]))
            mod.bias.copy_(_t(m, node.input[2]))
        gemm = [n for n in m.graph.node if n.op_type == "Gemm"][0]
        fc.weight.copy_(_t(m, gemm.input[1]))
        fc.bias.copy_(_t(m, gemm.input[2]))
    x = torch.rand(2, 3, 64, 64)
    with torch.no_grad():
        y = torch.max_pool2d(torch.relu(stem(x)), 3, 2, 1)
        for blk in blocks:
            y = blk(y)
        want = fc(y.mean(dim=(2, 3)))
    got = Interpreter(m.serialize()).run(x)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


def test_transformer_net_against_independent_torch_module():
    m = modelzoo.transformer_net(seed=5, hw=32, width_div=8)
    convs = [n for n in m.graph.node if n.op_type == "Conv"]
    norms = [n for n in m.graph.node if n.op_type == "InstanceNormalization"]
    ci, ni = iter(convs), iter(norms)

    def conv_layer(x, stride):
        n = next(ci)
        w, b = _t(m, n.input[1]), _t(m, n.input[2])
        p = w.shape[-1] // 2
        return torch.nn.functional.conv2d(torch.nn.functional.pad(x, (p, p, p, p), mode="reflect"), w, b, stride=stride)

    def inorm(x):
        n = next(ni)
        return torch.nn.functional.instance_norm(x, weight=_t(m, n.input[1]), bias=_t(m, n.input[2]), eps=n.attr("epsilon").f)

    x = torch.rand(1, 3, 32, 32)
    y = torch.relu(inorm(conv_layer(x, 1)))
    y = torch.relu(inorm(conv_layer(y, 2)))
    y = torch.relu(inorm(conv_layer(y, 2)))
    for _ in range(5):
        z = torch.relu(inorm(conv_layer(y, 1)))
        y = inorm(conv_layer(z, 1)) + y
    y = torch.relu(inorm(conv_layer(torch.nn.functional.interpolate(y, scale_factor=2, mode="nearest"), 1)))
    y = torch.relu(inorm(conv_layer(torch.nn.functional.interpolate(y, scale_factor=2, mode="nearest"), 1)))
    want = conv_layer(y, 1)
    assert torch.allclose(Interpreter(m.serialize()).run(x), want, atol=1e-5)


def test_depthwise_clip_path_against_torch():
    b = modelzoo.GraphBuilder(seed=3)
    x = b.input("input", [1, 12, 9, 9])
    y = b.clip(b.conv(x, 12, 3, 2, 1, groups=12))
    b.output(y, [1, 12, 5, 5])
    m = b.model()
    n = m.graph.node[0]
    xin = torch.rand(1, 12, 9, 9) * 10
    want = torch.nn.functional.conv2d(xin, _t(m, n.input[1]), _t(m, n.input[2]), stride=2, padding=1, groups=12).clamp(0, 6)
    assert torch.allclose(Interpreter(m.serialize()).run(xin), want, atol=1e-6)


def test_model_op_inventories_match_the_survey():
    """op counts and algorithmic MACs of the generated models (FLOP numerators of the metric)."""
    r = modelzoo.resnet50(seed=0, fold_bn=True)
    assert modelzoo.count_ops(r) == {"Conv": 53, "Relu": 49, "MaxPool": 1, "Add": 16, "GlobalAveragePool": 1, "Flatten": 1, "Gemm": 1}
    assert modelzoo.macs(r, (1, 3, 224, 224)) == 4_089_184_256
    mb = modelzoo.mobilenet_v2(seed=0, fold_bn=True)
    assert modelzoo.count_ops(mb) == {"Conv": 52, "Clip": 35, "Add": 10, "GlobalAveragePool": 1, "Flatten": 1, "Gemm": 1}
    assert modelzoo.macs(mb, (1, 3, 224, 224)) == 300_774_272
    t = modelzoo.transformer_net(seed=0)
    assert modelzoo.count_ops(t) == {"Pad": 16, "Conv": 16, "InstanceNormalization": 15, "Relu": 10, "Add": 5, "Constant": 2, "Upsample": 2}
    assert modelzoo.macs(t, (1, 3, 512, 512)) == 40_315_650_048


def test_conv_transpose_group_norm_pow_against_torch_modules():
    """Pins the oracle's ConvTranspose / custom_group_norm / Pow (the registry entries no BASELINE model uses) against eager
    nn.ConvTranspose2d / nn.GroupNorm loaded with the graph's initializers, in ONNX flavour and after the ONNX2MPS swizzle."""
    from smelter_b200 import onnx2mps

    m = modelzoo.decoder_ops(seed=3)
    nodes = {n.op_type + str(i): n for i, n in enumerate(m.graph.node)}
    conv_n = m.graph.node[0]
    gn_n = m.graph.node[1]
    ct = [n for n in m.graph.node if n.op_type == "ConvTranspose"]
    assert [n.op_type for n in m.graph.node] == ["Conv", "custom_group_norm", "Relu", "ConvTranspose", "Relu", "ConvTranspose", "Sigmoid", "Pow",
                                                 "ConvTranspose"] and nodes
    conv = nn.Conv2d(32, 64, 3, padding=1)
    gn = nn.GroupNorm(8, 64, eps=1e-5)
    t1 = nn.ConvTranspose2d(64, 48, 3, stride=2, padding=1, output_padding=1)
    t2 = nn.ConvTranspose2d(48, 24, 4, stride=2, padding=1)
    t3 = nn.ConvTranspose2d(24, 8, 1)
    with torch.no_grad():
        conv.weight.copy_(_t(m, conv_n.input[1])); conv.bias.copy_(_t(m, conv_n.input[2]))
        gn.weight.copy_(_t(m, gn_n.input[2])); gn.bias.copy_(_t(m, gn_n.input[3]))
        for mod, n in zip((t1, t2, t3), ct):
            mod.weight.copy_(_t(m, n.input[1])); mod.bias.copy_(_t(m, n.input[2]))
    x = torch.randn(2, 32, 10, 10)
    with torch.no_grad():
        want = t3(torch.sigmoid(t2(torch.relu(t1(torch.relu(gn(conv(x))))))) ** 2.0)
    got = Interpreter(m.serialize()).run(x)
    assert got.shape == want.shape == (2, 8, 40, 40)
    assert torch.allclose(got, want, atol=1e-5)
    got_mps = Interpreter(onnx2mps.convert_bytes(m.serialize(), half=False)).run(x)  # [1,2,3,0] swizzle + flip, fp32 kept
    assert torch.allclose(got_mps, want, atol=1e-5)


@pytest.mark.parametrize("arch", ["squeezenet1_1", "densenet121", "googlenet", "resnext50_32x4d", "regnet_y_400mf", "mnasnet1_0", "mobilenet_v3_small"])
def test_oracle_on_models_exported_by_torch(arch):
    """The oracle against EAGER torchvision modules on files written by torch's own exporter (a writer and model definitions this
    repository did not author): fp32 on the plain export, and within the fp16 tolerance after the ONNX2MPS restatement with --half.
    squeezenet1_1 carries weights aliased through Identity nodes (the exporter de-duplicates equal initializers), densenet121
    BatchNorms that no convolution precedes, Concat, AveragePool and Pad."""
    from real_export import export, torchvision_model
    from smelter_b200 import onnx2mps

    net = torchvision_model(arch)
    x = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(1)).half().float()
    data = export(net, x, do_constant_folding=False)
    with torch.no_grad():
        want = net(x)
    scale = max(1.0, float(want.abs().max()))
    got = Interpreter(data).run(x).reshape(want.shape)
    assert float((got - want).abs().max()) <= 2e-3 * scale
    mps = onnx2mps.convert_bytes(data, half=True)
    got16 = Interpreter(mps).run(x).reshape(want.shape)
    assert float((got16 - want).abs().max()) <= 1e-2 * scale


def test_oracle_on_transformer_net_exported_by_torch():
    """BASELINE.json configs[3] from an eager module restated from pytorch/examples (tests/real_export.py) and torch's own exporter
    (constant folding on: the Upsample scales arrive as a Constant, the only form the reference's UpsampleConverter reads)."""
    from real_export import export, transformer_net
    from smelter_b200 import onnx2mps

    net = transformer_net(16)
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1)).half().float()
    data = export(net, x, do_constant_folding=True)
    assert modelzoo.count_ops(op.Model.parse(data)) == {"Pad": 16, "Conv": 16, "InstanceNormalization": 15, "Relu": 10, "Add": 5, "Constant": 2,
                                                        "Upsample": 2}
    with torch.no_grad():
        want = net(x)
    got = Interpreter(data).run(x).reshape(want.shape)
    assert float((got - want).abs().max()) <= 1e-4
    got16 = Interpreter(onnx2mps.convert_bytes(data, half=True)).run(x).reshape(want.shape)
    assert float((got16 - want).abs().max()) <= 1e-2
