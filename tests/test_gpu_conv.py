"""Parity of the tcgen05 implicit-GEMM convolution (and the depthwise kernel) against fp32 torch-CPU, through the C ABI
(smelter_run_conv).  Shapes cover every A-operand mode (tiled 1x1, im2col TMA, packed-row stems, depthwise), strides,
dilation, padding, M/N tails, channel counts that are not multiples of 8/64, and the fused epilogues."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tools.gpu_probe import CONV_CASES

pytestmark = pytest.mark.gpu

# tolerance: fp16 inputs, fp32 accumulation, one fp16 rounding of the output (relative 2^-11 ~ 4.9e-4) => 4e-3 of the
# output range is > 8 ulp at the largest magnitude
REL_TOL = 4e-3


def _reference(x, w, b, s, p, d, g, act, res):
    y = F.conv2d(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(w.astype(np.float32)), torch.from_numpy(b) if b is not None else None,
                 stride=s, padding=p, dilation=d, groups=g)
    if res is not None:
        y = y + torch.from_numpy(res.astype(np.float32))
    if act == 1:
        y = y.relu()
    elif act == 2:
        y = y.clamp(0.0, 6.0)
    elif act == 3:
        y = torch.sigmoid(y)
    return y.numpy()


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_matches_fp32_reference(ctx, case):
    from smelter_b200.api import Image, run_conv

    name, shape, co, k, s, p, d, g, act, has_bias, has_res, force = case
    rng = np.random.default_rng(abs(hash(name)) % (1 << 31))
    n, c, h, w = shape
    x = rng.standard_normal(shape).astype(np.float16)
    wt = (rng.standard_normal((co, c // g, k, k)) * np.sqrt(2.0 / (c // g * k * k))).astype(np.float16)
    b = rng.standard_normal(co).astype(np.float32) if has_bias else None
    oh = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
    ow = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
    r = rng.standard_normal((n, co, oh, ow)).astype(np.float16) if has_res else None
    ref = _reference(x, wt, b, s, p, d, g, act, r)
    y, _ = run_conv(ctx, Image.fromArray(ctx, x), wt, b, stride=(s, s), pads=(p, p, p, p), dilation=(d, d), groups=g, act=act,
                    clip=(0.0, 6.0), residual=Image.fromArray(ctx, r) if has_res else None, force_path=force)
    out = y.toFloatArray()
    assert out.shape == ref.shape
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() <= REL_TOL * max(1.0, np.abs(ref).max())


def test_asymmetric_padding_and_sigmoid_epilogue(ctx):
    from smelter_b200.api import Image, run_conv

    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 40, 11, 9)).astype(np.float16)
    wt = (rng.standard_normal((24, 40, 3, 3)) * 0.05).astype(np.float16)
    b = rng.standard_normal(24).astype(np.float32)
    xt = F.pad(torch.from_numpy(x.astype(np.float32)), (0, 2, 1, 0))  # left 0, right 2, top 1, bottom 0
    ref = torch.sigmoid(F.conv2d(xt, torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b))).numpy()
    y, _ = run_conv(ctx, Image.fromArray(ctx, x), wt, b, pads=(1, 0, 0, 2), act=3)
    assert np.abs(y.toFloatArray() - ref).max() <= 2e-3


@pytest.mark.parametrize("res", [False, True], ids=["plain", "residual"])
def test_every_block_n_gives_the_same_answer(ctx, monkeypatch, res):
    """The N-tile width is a scheduling choice; results must not depend on it (same k order and fp32 accumulation per output).
    SMELTER_FORCE_BN (read when a launch is prepared) forces 64 / 128 / 256-column tiles of the two-CTA kernel."""
    from smelter_b200.api import Image, run_conv

    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 96, 14, 14)).astype(np.float16)
    wt = (rng.standard_normal((200, 96, 3, 3)) * 0.04).astype(np.float16)
    b = rng.standard_normal(200).astype(np.float32)
    r = rng.standard_normal((2, 200, 14, 14)).astype(np.float16) if res else None
    xi = Image.fromArray(ctx, x)

    def run():
        return run_conv(ctx, xi, wt, b, pads=(1, 1, 1, 1), act=1, residual=Image.fromArray(ctx, r) if res else None)[0].toHalfArray()

    base = run()
    again = run()
    assert np.array_equal(base.view(np.uint16), again.view(np.uint16))  # deterministic
    for bn in (64, 128, 256):
        monkeypatch.setenv("SMELTER_FORCE_BN", str(bn))
        forced = run()
        assert np.array_equal(base.view(np.uint16), forced.view(np.uint16)), f"BLOCK_N={bn} changes the result"


def test_linearity_at_full_size(ctx):
    """Size-independent property at a BASELINE-sized layer (batch 32, 256->64 1x1 @56x56; too big for a CPU reference in
    seconds): conv(a*x1 + x2) == a*conv(x1) + conv(x2) within fp16 rounding, bias off."""
    from smelter_b200.api import Image, run_conv

    rng = np.random.default_rng(11)
    shape = (32, 256, 56, 56)
    x1 = (rng.standard_normal(shape) * 0.5).astype(np.float16)
    x2 = (rng.standard_normal(shape) * 0.5).astype(np.float16)
    wt = (rng.standard_normal((64, 256, 1, 1)) / 16).astype(np.float16)
    xs = (x1.astype(np.float32) * 2.0 + x2.astype(np.float32)).astype(np.float16)
    y1 = run_conv(ctx, Image.fromArray(ctx, x1), wt, None)[0].toFloatArray()
    y2 = run_conv(ctx, Image.fromArray(ctx, x2), wt, None)[0].toFloatArray()
    ys = run_conv(ctx, Image.fromArray(ctx, xs), wt, None)[0].toFloatArray()
    assert np.abs(ys - (2.0 * y1 + y2)).max() <= 2e-2
    # and a sampled exact check of 64 random outputs against float64 dot products
    idx = rng.integers(0, [32, 64, 56, 56], size=(64, 4))
    for n_, co, hh, ww in idx:
        want = float(np.dot(x1[n_, :, hh, ww].astype(np.float64), wt[co, :, 0, 0].astype(np.float64)))
        assert abs(y1[n_, co, hh, ww] - want) <= 4e-3 * max(1.0, abs(want))


_SINGLE_CTA = [c for c in CONV_CASES if c[0] in ("tiled_64_256_56", "tiled_256_64_56_res", "tiled_1024_2048_bn256", "im2col_3x3_64_56", "im2col_3x3_s2_128_56",
                                                 "im2col_3x3_odd_16", "im2col_3x3_512_7", "rows_7x7_s2_stem")]


@pytest.mark.parametrize("case", _SINGLE_CTA, ids=[c[0] for c in _SINGLE_CTA])
def test_single_cta_kernel_matches_the_two_cta_default(ctx, case, monkeypatch):
    """The default for 64/128/256-column tiles is the two-CTA cluster kernel (conv_pair.cu, tcgen05.mma.cta_group::2); the single-CTA
    kernel (SMELTER_NO_PAIR=1) must give the same results bit for bit (same k order, same fp32 accumulation, same epilogue)."""
    from smelter_b200.api import Image, run_conv

    name, shape, co, k, s, p, d, g, act, has_bias, has_res, force = case
    rng = np.random.default_rng(11)
    n, c, h, w = shape
    x = rng.standard_normal(shape).astype(np.float16)
    wt = (rng.standard_normal((co, c // g, k, k)) * np.sqrt(2.0 / (c // g * k * k))).astype(np.float16)
    b = rng.standard_normal(co).astype(np.float32) if has_bias else None
    oh = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
    ow = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
    r = rng.standard_normal((n, co, oh, ow)).astype(np.float16) if has_res else None

    def run():
        y, _ = run_conv(ctx, Image.fromArray(ctx, x), wt, b, stride=(s, s), pads=(p, p, p, p), dilation=(d, d), groups=g, act=act, clip=(0.0, 6.0),
                        residual=Image.fromArray(ctx, r) if has_res else None, force_path=force)
        return y.toHalfArray()

    pair = run()
    monkeypatch.setenv("SMELTER_NO_PAIR", "1")
    single = run()
    assert np.array_equal(pair.view(np.uint16), single.view(np.uint16))
