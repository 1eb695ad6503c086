"""Shared by the CPU and GPU real-export tests: torch's own (legacy, opset 9) ONNX exporter on torchvision models."""
import io

import pytest
import torch


def export(module, x, **kw):
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


def randomise_norms(net):
    """non-trivial BatchNorm statistics so that the fold is tested and activations stay O(1)."""
    g = torch.Generator().manual_seed(123)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)


def torchvision_model(arch: str):
    """Seeded torchvision model in eval mode with non-trivial BatchNorm statistics."""
    torchvision = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    kw = {"aux_logits": False, "init_weights": True} if arch == "googlenet" else {}
    net = getattr(torchvision.models, arch)(weights=None, **kw).eval()
    randomise_norms(net)
    return net
