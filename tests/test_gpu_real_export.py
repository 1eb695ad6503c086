"""The reference's workflow on models this repository did not author (README.md:54, ONNX2MPS.py:112-134): a torchvision model is
exported by torch's own (legacy, opset 9) ONNX exporter, taken through the ONNX2MPS restatement with --half, parsed by the C++ wire
reader, built and run by the engine, and compared with the EAGER torch module in fp32.  This pins the parser (field order, packed /
unpacked repeated fields, value_info clutter, doc strings of a real exporter), the BN fold, every converter on the path and the
kernels against a writer and a model definition outside this repository.  Tolerance: 1e-2 max-abs relative to the logit range."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _export(module, x, **kw):
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


def _randomise_norms(net):
    """non-trivial BatchNorm statistics so that the fold is tested and activations stay O(1)."""
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)


@pytest.mark.parametrize("arch,fold_in_exporter", [("resnet50", False), ("resnet50", True), ("mobilenet_v2", False), ("resnet18", False)])
def test_torchvision_model_exported_by_torch_matches_eager(ctx, arch, fold_in_exporter):
    torchvision = pytest.importorskip("torchvision")
    from smelter_b200 import onnx2mps
    from smelter_b200 import onnx_proto as op
    from smelter_b200.api import Format, Image, ONNXGraph

    torch.manual_seed(0)
    net = getattr(torchvision.models, arch)(weights=None).eval()
    _randomise_norms(net)
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(1)).half().float()
    # do_constant_folding=False keeps the BatchNormalization nodes for the ONNX2MPS fold; True lets the exporter fold them itself
    data = _export(net, x, do_constant_folding=fold_in_exporter)
    ops = {n.op_type for n in op.Model.parse(data).graph.node}
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
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(out - want).max() <= 1e-2 * scale, (float(np.abs(out - want).max()), scale)
    # and the un-converted fp32 export (plain ONNX flavour, OIHW weights re-laid-out by the engine) gives the same answer
    g2 = ONNXGraph(data, context=ctx)
    assert g2.modelFormat == Format.onnx
    out2 = g2.metalGraph().encode(sourceImages=[Image.fromArray(ctx, x.numpy().astype(np.float16))]).toFloatArray().reshape(2, -1)
    g2.close()
    assert np.abs(out2 - want).max() <= 1e-2 * scale
