"""Parity of every HBM-bound kernel on the path (NHWC fp16, 128-bit vectors) against fp32 torch-CPU, through
smelter_run_elementwise.  Channel counts include non-multiples of 8 (padded lanes) and odd spatial sizes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(2, 16, 9, 11), (1, 3, 17, 5), (3, 20, 8, 8), (1, 64, 28, 28)]


def _img(ctx, a):
    from smelter_b200.api import Image

    return Image.fromArray(ctx, a)


def _rand(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float16)


def _t(a):
    return torch.from_numpy(a.astype(np.float32))


def _close(out, ref, tol=2e-3):
    assert out.shape == tuple(ref.shape)
    assert np.isfinite(out).all()
    assert np.abs(out - ref.numpy()).max() <= tol * max(1.0, float(ref.abs().max()))


UNARY = {0: torch.relu, 1: torch.sigmoid, 2: lambda x: x.clamp(-0.5, 0.75), 3: torch.tanh, 4: torch.abs, 5: torch.exp,
         7: lambda x: F.elu(x, 0.7), 8: lambda x: F.leaky_relu(x, 0.7), 9: lambda x: (0.7 * x + 0.75).clamp(0, 1), 10: F.softplus,
         11: F.softsign, 12: lambda x: x}


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", sorted(UNARY))
def test_unary(ctx, shape, kind):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, kind)
    alpha, beta = (-0.5, 0.75) if kind == 2 else (0.7, 0.75)
    y, _ = run_elementwise(ctx, "unary", _img(ctx, x), out_shape=shape, sub=kind, alpha=alpha, beta=beta)
    _close(y.toFloatArray(), UNARY[kind](_t(x)), 3e-3)


def test_log(ctx):
    from smelter_b200.api import run_elementwise

    x = np.abs(_rand((2, 16, 6, 6), 3)) + np.float16(0.1)
    y, _ = run_elementwise(ctx, "unary", _img(ctx, x), out_shape=x.shape, sub=6)
    _close(y.toFloatArray(), torch.log(_t(x)), 3e-3)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind,act", [(0, 0), (0, 1), (1, 0), (2, 0), (3, 0)])
def test_binary(ctx, shape, kind, act):
    from smelter_b200.api import run_elementwise

    a, b = _rand(shape, 1), _rand(shape, 2)
    if kind == 3:
        b = (np.abs(b) + np.float16(0.5)).astype(np.float16)
    ref = [torch.add, torch.sub, torch.mul, torch.div][kind](_t(a), _t(b))
    if act:
        ref = ref.relu()
    y, _ = run_elementwise(ctx, "binary", _img(ctx, a), x2=_img(ctx, b), out_shape=shape, sub=kind, act=act)
    _close(y.toFloatArray(), ref, 3e-3)


@pytest.mark.parametrize("shape", SHAPES)
def test_scale_shift_batchnorm(ctx, shape):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 4)
    rng = np.random.default_rng(5)
    sc, sh = rng.uniform(0.5, 1.5, shape[1]).astype(np.float32), rng.standard_normal(shape[1]).astype(np.float32)
    ref = (_t(x) * torch.from_numpy(sc).view(1, -1, 1, 1) + torch.from_numpy(sh).view(1, -1, 1, 1)).relu()
    y, _ = run_elementwise(ctx, "scale_shift", _img(ctx, x), p0=sc, p1=sh, out_shape=shape, act=1)
    _close(y.toFloatArray(), ref, 3e-3)


@pytest.mark.parametrize("shape", [(2, 16, 56, 56), (1, 24, 13, 15), (2, 64, 112, 112)])
@pytest.mark.parametrize("k,s,p,is_max", [(3, 2, 1, 1), (2, 2, 0, 1), (3, 1, 1, 0), (2, 2, 0, 0)])
def test_pool(ctx, shape, k, s, p, is_max):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 6)
    ref = F.max_pool2d(_t(x), k, s, p) if is_max else F.avg_pool2d(_t(x), k, s, p, count_include_pad=True)
    y, _ = run_elementwise(ctx, "pool", _img(ctx, x), out_shape=tuple(ref.shape), sub=is_max, k_h=k, k_w=k, stride_h=s, stride_w=s, pad_h=p, pad_w=p)
    _close(y.toFloatArray(), ref)


@pytest.mark.parametrize("shape", [(2, 2048, 7, 7), (1, 1280, 7, 7), (3, 24, 30, 30), (32, 16, 9, 9), (1, 5, 64, 64)])
def test_global_avgpool(ctx, shape):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 7)
    y, _ = run_elementwise(ctx, "global_avgpool", _img(ctx, x), out_shape=(shape[0], shape[1], 1, 1))
    _close(y.toFloatArray(), _t(x).mean(dim=(2, 3), keepdim=True), 2e-3)


@pytest.mark.parametrize("shape", [(4, 1000, 1, 1), (2, 10, 1, 1), (1, 21, 6, 7), (32, 1000, 1, 1)])
@pytest.mark.parametrize("log", [0, 1])
def test_softmax(ctx, shape, log):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 8, 3.0)
    ref = torch.log_softmax(_t(x), 1) if log else torch.softmax(_t(x), 1)
    y, _ = run_elementwise(ctx, "softmax", _img(ctx, x), out_shape=shape, sub=log)
    _close(y.toFloatArray(), ref, 2e-3)


@pytest.mark.parametrize("shape", [(1, 128, 16, 16), (2, 12, 7, 9)])
@pytest.mark.parametrize("mode,align,scale", [(0, 1, 2), (0, 1, 3), (1, 1, 2), (1, 0, 2)])
def test_upsample(ctx, shape, mode, align, scale):
    from oracle.onnx_interp import _bilinear_asymmetric
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 9)
    if mode == 0:
        ref = _t(x).repeat_interleave(scale, 2).repeat_interleave(scale, 3)
    elif align:
        ref = F.interpolate(_t(x), scale_factor=scale, mode="bilinear", align_corners=True)
    else:
        ref = _bilinear_asymmetric(_t(x), scale, scale)
    y, _ = run_elementwise(ctx, "upsample", _img(ctx, x), out_shape=tuple(ref.shape), sub=mode, scale_h=scale, scale_w=scale, align_corners=align)
    _close(y.toFloatArray(), ref, 2e-3)


@pytest.mark.parametrize("shape", [(1, 3, 20, 24), (2, 32, 9, 9)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_pad(ctx, shape, mode):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 10)
    pt, pl, pb, pr = 4, 1, 2, 3
    tm = ["constant", "reflect", "replicate"][mode]
    ref = F.pad(_t(x), (pl, pr, pt, pb), mode=tm, value=1.5) if mode == 0 else F.pad(_t(x), (pl, pr, pt, pb), mode=tm)
    y, _ = run_elementwise(ctx, "pad", _img(ctx, x), out_shape=tuple(ref.shape), sub=mode, pad_h=pt, pad_w=pl, pad_b=pb, pad_r=pr, alpha=1.5)
    _close(y.toFloatArray(), ref, 1e-6)  # pure data movement: exact


@pytest.mark.parametrize("c1,c2", [(16, 8), (24, 8), (5, 3), (12, 20), (64, 64)])
def test_concat(ctx, c1, c2):
    """Includes the C % 4 != 0 cases the reference warns are wrong in MPS (README.md:63-64)."""
    from smelter_b200.api import run_elementwise

    a, b = _rand((2, c1, 6, 7), 11), _rand((2, c2, 6, 7), 12)
    y, _ = run_elementwise(ctx, "concat", _img(ctx, a), x2=_img(ctx, b), out_shape=(2, c1 + c2, 6, 7), c2=c2)
    _close(y.toFloatArray(), torch.cat([_t(a), _t(b)], 1), 1e-6)


@pytest.mark.parametrize("shape", [(1, 32, 64, 64), (2, 128, 16, 16), (1, 3, 40, 40), (1, 64, 128, 128)])
@pytest.mark.parametrize("act", [0, 1])
def test_instance_norm(ctx, shape, act):
    from smelter_b200.api import run_elementwise

    x = (_rand(shape, 13).astype(np.float32) * 2 + 0.5).astype(np.float16)
    rng = np.random.default_rng(14)
    g, b = rng.uniform(0.5, 1.5, shape[1]).astype(np.float32), rng.standard_normal(shape[1]).astype(np.float32)
    ref = F.instance_norm(_t(x), weight=torch.from_numpy(g), bias=torch.from_numpy(b), eps=1e-5)
    if act:
        ref = ref.relu()
    y, _ = run_elementwise(ctx, "instance_norm", _img(ctx, x), p0=g, p1=b, out_shape=shape, alpha=1e-5, act=act)
    _close(y.toFloatArray(), ref, 4e-3)


@pytest.mark.parametrize("shape,act", [((2, 128, 128, 128), 1), ((1, 24, 61, 47), 0), ((3, 40, 9, 200), 1), ((1, 32, 300, 300), 0),
                                       ((2, 64, 256, 256), 1), ((1, 32, 512, 512), 1), ((1, 8, 1024, 640), 0)])
def test_instance_norm_cluster_form_matches_three_launch_form(ctx, shape, act, monkeypatch):
    """Small images take the one-launch cluster kernel (partial sums meet through distributed shared memory), large ones and
    SMELTER_NO_CLUSTER_NORM=1 the partials -> finalize -> apply form: same statistics in a different fixed summation order, so the
    outputs agree to fp16 rounding (and both with torch); each form is deterministic."""
    from smelter_b200.api import run_elementwise

    x = (_rand(shape, 21).astype(np.float32) * 3 - 0.25).astype(np.float16)
    rng = np.random.default_rng(22)
    g, b = rng.uniform(0.5, 1.5, shape[1]).astype(np.float32), rng.standard_normal(shape[1]).astype(np.float32)
    ref = F.instance_norm(_t(x), weight=torch.from_numpy(g), bias=torch.from_numpy(b), eps=1e-5)
    if act:
        ref = ref.relu()

    def run():
        y, _ = run_elementwise(ctx, "instance_norm", _img(ctx, x), p0=g, p1=b, out_shape=shape, alpha=1e-5, act=act)
        return y.toHalfArray()

    one, again = run(), run()
    assert np.array_equal(one.view(np.uint16), again.view(np.uint16))
    _close(one.astype(np.float32), ref, 4e-3)
    monkeypatch.setenv("SMELTER_NO_CLUSTER_NORM", "1")
    three = run()
    _close(three.astype(np.float32), ref, 4e-3)
    assert np.abs(one.astype(np.float32) - three.astype(np.float32)).max() <= 4e-3


@pytest.mark.parametrize("shape,act", [((2, 128, 128, 128), 1), ((1, 24, 61, 47), 0), ((3, 40, 9, 200), 1), ((2, 32, 256, 256), 0), ((1, 320, 16, 16), 1)])
def test_instance_norm_one_pass_form_with_supplied_statistics(ctx, shape, act):
    """The apply pass of the norm behind a convolution that accumulated the statistics (k::instance_norm_from_stats; here the fp64
    sums come from the norm's own statistics kernel): scale / shift derived per block, one read and one write of the tensor."""
    from smelter_b200.api import run_elementwise

    x = (_rand(shape, 31).astype(np.float32) * 3 - 0.25).astype(np.float16)
    rng = np.random.default_rng(32)
    g, b = rng.uniform(0.5, 1.5, shape[1]).astype(np.float32), rng.standard_normal(shape[1]).astype(np.float32)
    ref = F.instance_norm(_t(x), weight=torch.from_numpy(g), bias=torch.from_numpy(b), eps=1e-5)
    if act:
        ref = ref.relu()
    y, _ = run_elementwise(ctx, "instance_norm", _img(ctx, x), p0=g, p1=b, out_shape=shape, alpha=1e-5, act=act, sub=1, iters=2)
    _close(y.toFloatArray(), ref, 4e-3)


@pytest.mark.parametrize("shape", SHAPES + [(32, 3, 224, 224)])
def test_layout_roundtrip_is_exact(ctx, shape):
    from smelter_b200.api import run_elementwise

    x = _rand(shape, 15)
    y, _ = run_elementwise(ctx, "layout_roundtrip", _img(ctx, x), out_shape=shape)
    assert np.array_equal(y.toHalfArray().view(np.uint16), x.view(np.uint16))


@pytest.mark.parametrize("src_channels", [3, 4])
def test_uint8_image_ingestion(ctx, src_channels):
    """smelter_tensor_from_u8: interleaved 8-bit pixels -> fp16 NCHW with per-channel scale/bias (the `texture(from:)` step of
    README.md:33-39); exact against numpy in fp32 rounded once to fp16."""
    from smelter_b200.api import Image

    rng = np.random.default_rng(7)
    px = rng.integers(0, 256, size=(2, 19, 23, src_channels), dtype=np.uint8)
    scale, bias = [1 / 255.0, 2 / 255.0, 0.5], [-0.485, 0.0, 1.25]
    got = Image.fromBytes(ctx, px, channels=3, scale=scale, bias=bias).toHalfArray()
    want = (px[..., :3].astype(np.float32) * np.asarray(scale, np.float32) + np.asarray(bias, np.float32)).astype(np.float16).transpose(0, 3, 1, 2)
    assert got.shape == (2, 3, 19, 23)
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
    default = Image.fromBytes(ctx, px[0], channels=3).toHalfArray()
    assert np.array_equal(default.view(np.uint16), (px[:1, ..., :3].astype(np.float32) * np.float32(1 / 255.0)).astype(np.float16).transpose(0, 3, 1, 2).view(np.uint16))


@pytest.mark.parametrize("c", [1, 2, 3, 6, 8])
def test_read_back_in_mps_slice_order(ctx, c):
    """smelter_tensor_to_float_mps: the element order MPSImage.toFloatArray() returns (MPSImage+Extensions.swift:26-59): per image,
    slices of four channels stored [H][W][4] with zero padding; one [H][W][C] slice when C < 3."""
    from smelter_b200.api import Image

    x = np.random.default_rng(c).standard_normal((2, c, 5, 7)).astype(np.float16)
    got = Image.fromArray(ctx, x).toFloatArrayMPS()
    comps, slices = (c, 1) if c < 3 else (4, (c + 3) // 4)
    want = np.zeros((2, slices, 5, 7, comps), dtype=np.float32)
    for ch in range(c):
        want[:, ch // 4 if c >= 3 else 0, :, :, ch % 4 if c >= 3 else ch] = x[:, ch].astype(np.float32)
    assert got.shape == (want.size,)
    assert np.array_equal(got, want.reshape(-1))
