"""The reference's workflow on models this repository did not author (README.md:54, ONNX2MPS.py:112-134): a torchvision model is
exported by torch's own (legacy, opset 9) ONNX exporter, taken through the ONNX2MPS restatement with --half, parsed by the C++ wire
reader, built and run by the engine, and compared with the EAGER torch module in fp32.  This pins the parser (field order, packed /
unpacked repeated fields, value_info clutter, doc strings of a real exporter), the BN fold, every converter on the path and the
kernels against a writer and a model definition outside this repository.  Tolerance: 1e-2 max-abs relative to the logit range."""
import numpy as np
import pytest
import torch

from real_export import export as _export, torchvision_model, transformer_net

pytestmark = pytest.mark.gpu


# densenet121: BatchNorm in front of its convolution (un-fused scale/shift kernel), Concat, AveragePool, Pad; squeezenet1_1: Concat and
# weights aliased through Identity nodes (the exporter de-duplicates equal initializers); googlenet: four-branch Concat, pools with
# ceil_mode emulated by the exporter; resnext50_32x4d / regnet_x: grouped convolutions as block-diagonal dense ones; regnet_y,
# efficientnet_b0, mobilenet_v3: squeeze-and-excitation (ReduceMean or GlobalAveragePool -> 1x1 convolutions -> Sigmoid / HardSigmoid ->
# broadcast Mul), SiLU / Hardswish as Sigmoid / HardSigmoid x Mul, depthwise 5x5; mnasnet: ReduceMean as the global pool; vgg11_bn /
# alexnet: AdaptiveAvgPool2d written as an AveragePool without `pads` (ONNX default), Flatten -> 25088 / 9216-wide Gemm, Dropout
@pytest.mark.parametrize("arch,fold_in_exporter", [("resnet50", False), ("resnet50", True), ("mobilenet_v2", False), ("resnet18", False),
                                                   ("resnet34", False), ("densenet121", False), ("squeezenet1_1", False), ("googlenet", False), ("resnext50_32x4d", False),
                                                   ("wide_resnet50_2", False), ("regnet_x_400mf", False), ("regnet_y_400mf", False), ("mnasnet1_0", False),
                                                   ("efficientnet_b0", False), ("mobilenet_v3_small", False), ("vgg11_bn", False), ("alexnet", False)])
def test_torchvision_model_exported_by_torch_matches_eager(ctx, arch, fold_in_exporter):
    from smelter_b200 import onnx2mps
    from smelter_b200 import onnx_proto as op
    from smelter_b200.api import Format, Image, ONNXGraph

    net = torchvision_model(arch)
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(1)).half().float()
    # do_constant_folding=False keeps the BatchNormalization nodes for the ONNX2MPS fold; True lets the exporter fold them itself
    data = _export(net, x, do_constant_folding=fold_in_exporter)
    ops = {n.op_type for n in op.Model.parse(data).graph.node}
    if arch not in ("squeezenet1_1", "alexnet"):  # no BatchNorm in SqueezeNet / AlexNet
        assert ("BatchNormalization" in ops) == (not fold_in_exporter)
    mps = onnx2mps.convert_bytes(data, half=True)
    g = ONNXGraph(mps, context=ctx)
    assert g.modelFormat == Format.mpsFlavor
    nn = g.metalGraph()
    out = nn.encode(sourceImages=[Image.fromArray(ctx, x.numpy().astype(np.float16))]).toFloatArray().reshape(2, -1)
    with torch.no_grad():
        want = net(x).numpy()
    g.close()
    assert out.shape == want.shape == (2, 1000)
    assert np.isfinite(out).all()
    # tolerance relative to the logit range (not clamped to 1: the randomly initialised SE / MobileNetV3 nets have logits of 0.02-0.4)
    scale = float(np.abs(want).max()) if arch in ("regnet_y_400mf", "mnasnet1_0", "efficientnet_b0", "mobilenet_v3_small", "vgg11_bn", "alexnet") else max(1.0, float(np.abs(want).max()))
    assert np.abs(out - want).max() <= 1e-2 * scale, (float(np.abs(out - want).max()), scale)
    # and the un-converted fp32 export (plain ONNX flavour, OIHW weights re-laid-out by the engine) gives the same answer
    g2 = ONNXGraph(data, context=ctx)
    assert g2.modelFormat == Format.onnx
    out2 = g2.metalGraph().encode(sourceImages=[Image.fromArray(ctx, x.numpy().astype(np.float16))]).toFloatArray().reshape(2, -1)
    g2.close()
    assert np.abs(out2 - want).max() <= 1e-2 * scale


@pytest.mark.parametrize("hw,width", [(128, 16), (256, 32)])
def test_transformer_net_exported_by_torch_matches_eager(ctx, hw, width):
    """BASELINE.json configs[3] (reflection Pad + Conv + InstanceNorm + nearest Upsample + Add) from an eager module restated from
    pytorch/examples and torch's own exporter, through ONNX2MPS --half and as the plain export, against the eager module."""
    from smelter_b200 import onnx2mps
    from smelter_b200.api import Image, ONNXGraph

    net = transformer_net(width)
    x = torch.rand(1, 3, hw, hw, generator=torch.Generator().manual_seed(2)).half().float()
    data = _export(net, x, do_constant_folding=True)
    with torch.no_grad():
        want = net(x).numpy()
    scale = max(1.0, float(np.abs(want).max()))
    for model in (onnx2mps.convert_bytes(data, half=True), data):
        g = ONNXGraph(model, context=ctx)
        out = g.metalGraph().encode(sourceImages=[Image.fromArray(ctx, x.numpy().astype(np.float16))]).toFloatArray()
        g.close()
        assert out.shape == want.shape and np.isfinite(out).all()
        assert np.abs(out - want).max() <= 1e-2 * scale, float(np.abs(out - want).max())
