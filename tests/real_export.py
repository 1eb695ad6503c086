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


class TransformerNet(torch.nn.Module):
    """pytorch/examples fast_neural_style TransformerNet (BASELINE.json configs[3]) as an eager module: reflection-padded convolutions,
    affine InstanceNorm, five residual blocks, two nearest x2 upsample-convolutions."""

    class _Conv(torch.nn.Module):
        def __init__(self, cin, cout, k, stride, upsample=None):
            super().__init__()
            self.upsample = upsample
            self.pad = torch.nn.ReflectionPad2d(k // 2)
            self.conv = torch.nn.Conv2d(cin, cout, k, stride)

        def forward(self, x):
            if self.upsample:
                x = torch.nn.functional.interpolate(x, mode="nearest", scale_factor=self.upsample)
            return self.conv(self.pad(x))

    class _Res(torch.nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv1, self.in1 = TransformerNet._Conv(c, c, 3, 1), torch.nn.InstanceNorm2d(c, affine=True)
            self.conv2, self.in2 = TransformerNet._Conv(c, c, 3, 1), torch.nn.InstanceNorm2d(c, affine=True)

        def forward(self, x):
            return self.in2(self.conv2(torch.relu(self.in1(self.conv1(x))))) + x

    def __init__(self, width=32):
        super().__init__()
        c1, c2, c3 = width, 2 * width, 4 * width
        C, N = TransformerNet._Conv, torch.nn.InstanceNorm2d
        self.conv1, self.in1 = C(3, c1, 9, 1), N(c1, affine=True)
        self.conv2, self.in2 = C(c1, c2, 3, 2), N(c2, affine=True)
        self.conv3, self.in3 = C(c2, c3, 3, 2), N(c3, affine=True)
        self.res = torch.nn.Sequential(*[TransformerNet._Res(c3) for _ in range(5)])
        self.deconv1, self.in4 = C(c3, c2, 3, 1, upsample=2), N(c2, affine=True)
        self.deconv2, self.in5 = C(c2, c1, 3, 1, upsample=2), N(c1, affine=True)
        self.deconv3 = C(c1, 3, 9, 1)

    def forward(self, x):
        y = torch.relu(self.in1(self.conv1(x)))
        y = torch.relu(self.in2(self.conv2(y)))
        y = torch.relu(self.in3(self.conv3(y)))
        y = self.res(y)
        y = torch.relu(self.in4(self.deconv1(y)))
        y = torch.relu(self.in5(self.deconv2(y)))
        return self.deconv3(y)


def transformer_net(width=32):
    torch.manual_seed(0)
    net = TransformerNet(width).eval()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.InstanceNorm2d):
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)
    return net
