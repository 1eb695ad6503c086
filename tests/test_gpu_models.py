"""End-to-end parity of ONNXGraph(data:) -> metalGraph -> encode -> toFloatArray on the B200 against the fp32 oracle:
golden fixtures, the four BASELINE.json model families, both model flavours, fusion on/off, CUDA graph on/off, batch
sharding, and the reference's error behaviour.  Tolerance: 1e-2 max-abs (BASELINE.json north_star), outputs O(1)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
TOL = 1e-2


def _run(ctx, model: bytes, x: np.ndarray, **cfg):
    from smelter_b200.api import Configuration, Image, ONNXGraph

    g = ONNXGraph(model, Configuration(**cfg), context=ctx)
    nn = g.metalGraph()
    out = nn.encode(sourceImages=[Image.fromArray(ctx, x)]).toFloatArray()
    launches = nn.numLaunches(x.shape[0])
    _run.folded = nn.planDump(x.shape[0]).count("+conv1x1(")  # projection shortcuts that run inside their block's last conv
    g.close()
    return out, launches


def _oracle(model: bytes, x: np.ndarray, **kw):
    from oracle.onnx_interp import Interpreter

    return Interpreter(model, **kw).run(torch.from_numpy(x.astype(np.float32))).numpy()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixture(ctx, path):
    z = np.load(path)
    model, x, y = z["model"].tobytes(), z["x"], z["y"]
    out, _ = _run(ctx, model, x)
    assert np.isfinite(out).all()
    assert np.abs(out.reshape(y.shape) - y).max() <= TOL


@pytest.mark.parametrize("fusion", [True, False])
@pytest.mark.parametrize("graph", [True, False])
def test_fusion_and_cuda_graph_do_not_change_results(ctx, fusion, graph):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "resnet_tiny.npz"))
    out, launches = _run(ctx, z["model"].tobytes(), z["x"], enableFusion=fusion, useCudaGraph=graph)
    assert np.abs(out.reshape(z["y"].shape) - z["y"]).max() <= TOL
    # un-fused: 53-ish extra BN/ReLU/Add launches; fused: one launch per conv + pool/gap/fc/boundary
    assert (launches < 30) == fusion


def test_config1_pipeline_conv_bn_relu(ctx):
    """BASELINE.json configs[0] on the device: raw model == ONNX2MPS fp32 == ONNX2MPS --half within tolerance."""
    from smelter_b200 import modelzoo, onnx2mps
    from smelter_b200.api import Format, ONNXGraph

    raw = modelzoo.conv_bn_relu(seed=0).serialize()
    x = np.random.default_rng(0).random((1, 3, 16, 16), dtype=np.float32).astype(np.float16)
    want = _oracle(raw, x)
    for model, fmt in ((raw, Format.onnx), (onnx2mps.convert_bytes(raw, half=False), Format.mpsFlavor), (onnx2mps.convert_bytes(raw, half=True), Format.mpsFlavor)):
        g = ONNXGraph(model, context=ctx)
        assert g.modelFormat == fmt
        g.close()
        out, launches = _run(ctx, model, x)
        assert launches == 3  # nchw->nhwc(+zero pad for the packed-row stem), conv(+bn)+relu, nhwc->nchw
        assert np.abs(out - want).max() <= 5e-3


def test_resnet50_batch4_matches_oracle(ctx):
    from smelter_b200 import modelzoo, onnx2mps

    model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
    x = np.random.default_rng(1).random((4, 3, 224, 224), dtype=np.float32).astype(np.float16)
    out, launches = _run(ctx, model, x)
    want = _oracle(model, x)
    assert out.shape == (4, 1000, 1, 1)
    out = out.reshape(4, 1000)
    assert np.abs(out - want).max() <= TOL
    assert (out.argmax(1) == want.argmax(1)).all()
    # boundary + conv (all ReLU/Add fused) + maxpool + gap + fc (logits are a view), minus the folded 1x1 shortcuts
    assert 0 <= _run.folded <= 4 and launches == 1 + 53 + 1 + 1 + 1 - _run.folded


def test_resnet50_batch32_properties(ctx, monkeypatch):
    """BASELINE size (batch 32): every image's logits equal the logits of the same image run in a smaller batch (images never
    interact).  With one k-reduction order per layer (split-K off) that holds bit for bit whatever the batch; with
    the default plan the 16-image shards may split K where the 32-image batch does not, which re-associates fp32 sums only."""
    from smelter_b200 import modelzoo, onnx2mps
    from smelter_b200.api import Image, ONNXGraph

    model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
    x = np.random.default_rng(2).random((32, 3, 224, 224), dtype=np.float32).astype(np.float16)

    def run_all():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        full = nn.encode(sourceImages=[Image.fromArray(ctx, x)]).toHalfArray().reshape(32, 1000).copy()
        again = nn.encode(sourceImages=[Image.fromArray(ctx, x)]).toHalfArray().reshape(32, 1000).copy()
        assert np.array_equal(full.view(np.uint16), again.view(np.uint16))  # replay determinism
        lo = nn.encode(sourceImages=[Image.fromArray(ctx, x[:16])]).toHalfArray().reshape(16, 1000).copy()
        hi = nn.encode(sourceImages=[Image.fromArray(ctx, x[16:])]).toHalfArray().reshape(16, 1000).copy()
        g.close()
        return full, np.concatenate([lo, hi])

    full, shards = run_all()
    assert np.abs(shards.astype(np.float32) - full.astype(np.float32)).max() <= 4e-3  # default plans: fp32 re-association only
    want = _oracle(model, x[:2])
    assert np.abs(full[:2].astype(np.float32) - want).max() <= TOL
    monkeypatch.setenv("SMELTER_NO_SPLITK", "1")  # read when a plan is made
    full1, shards1 = run_all()
    assert np.array_equal(shards1.view(np.uint16), full1.view(np.uint16))  # shard == whole, bit for bit
    assert np.abs(full1[:2].astype(np.float32) - want).max() <= TOL


def test_projection_shortcut_runs_inside_the_block_output_conv(ctx, monkeypatch):
    """The four 1x1 "downsample" convolutions of ResNet-50 are extra k-blocks of their block's last convolution (one GEMM over
    the concatenated K axis, kernels/conv_pair.cu): four launches fewer, the shortcut tensor never exists, and the sum is
    taken in fp32 before the single fp16 rounding -- so results agree with the two-launch form to fp16 rounding only."""
    from smelter_b200 import modelzoo, onnx2mps
    from smelter_b200.api import Image, ONNXGraph

    model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
    x = np.random.default_rng(7).random((32, 3, 224, 224), dtype=np.float32).astype(np.float16)

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        outs = [nn.encode(sourceImages=[Image.fromArray(ctx, x)]).toHalfArray().reshape(32, 1000).copy() for _ in range(2)]
        dump, n = nn.planDump(32), nn.numLaunches(32)
        g.close()
        assert np.array_equal(outs[0].view(np.uint16), outs[1].view(np.uint16))
        return outs[0], dump, n

    folded, dump, n = run()
    monkeypatch.setenv("SMELTER_NO_SIDE", "1")  # read when a plan is made
    plain, dump0, n0 = run()
    assert dump.count("+conv1x1(") == 4 and dump0.count("+conv1x1(") == 0 and n0 - n == 4
    assert np.abs(folded.astype(np.float32) - plain.astype(np.float32)).max() <= 4e-3
    want = _oracle(model, x[:2])
    assert np.abs(folded[:2].astype(np.float32) - want).max() <= TOL
    assert np.abs(plain[:2].astype(np.float32) - want).max() <= TOL


@pytest.mark.parametrize("batch,hw", [(3, 40), (1, 64), (5, 32)])
def test_folded_projection_shortcut_on_ragged_tiles(ctx, monkeypatch, batch, hw):
    """Small, odd-sized bottleneck stacks: the last m-tile pair is partly (or wholly) out of range, stride-2 shortcuts sample odd
    image sizes, and every shortcut is folded (split-K off so that the two-CTA kernel is the plan for these tiny layers)."""
    from smelter_b200 import modelzoo

    monkeypatch.setenv("SMELTER_NO_SPLITK", "1")  # read when a plan is made
    model = modelzoo.resnet50(seed=batch, fold_bn=True, num_classes=24, hw=hw, depths=(1, 2, 1, 1)).serialize()
    x = np.random.default_rng(batch).random((batch, 3, hw, hw), dtype=np.float32).astype(np.float16)
    out, launches = _run(ctx, model, x)
    assert _run.folded == 4
    want = _oracle(model, x)
    assert np.abs(out.reshape(want.shape) - want).max() <= TOL


def test_sm_share_plans_and_encodes_in_flight(ctx):
    """Configuration(smShare=2) sizes every kernel for half of the SMs so that encodes on different streams co-run.
    The plan changes tile widths and grid sizes only: results stay within the tolerance of the oracle and within fp32
    re-association of the whole-chip plan, and three encodes in flight on three streams (each with its own activation arena and
    CUDA graph) give exactly what the same encodes give one at a time."""
    import torch

    from smelter_b200 import modelzoo, onnx2mps
    from smelter_b200.api import Configuration, Image, ONNXGraph

    model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
    rng = np.random.default_rng(9)
    xs = [rng.random((8, 3, 224, 224), dtype=np.float32).astype(np.float16) for _ in range(3)]
    whole = ONNXGraph(model, Configuration(smShare=1), context=ctx)
    half = ONNXGraph(model, Configuration(smShare=2), context=ctx)
    nn1, nn2 = whole.metalGraph(), half.metalGraph()
    base = [nn1.encode(sourceImages=[Image.fromArray(ctx, x)]).toHalfArray().reshape(8, 1000).copy() for x in xs]
    serial = [nn2.encode(sourceImages=[Image.fromArray(ctx, x)]).toHalfArray().reshape(8, 1000).copy() for x in xs]
    for a, b in zip(base, serial):
        assert np.abs(a.astype(np.float32) - b.astype(np.float32)).max() <= 4e-3
    assert np.abs(serial[0][:2].astype(np.float32) - _oracle(model, xs[0][:2])).max() <= TOL
    streams = [torch.cuda.Stream() for _ in xs]
    images = [Image.fromArray(ctx, x) for x in xs]
    ctx.synchronize()
    for _ in range(3):  # several rounds: plans are created on first use of a stream, later rounds replay their graphs concurrently
        results = [nn2.encode(to=s.cuda_stream, sourceImages=[im]) for s, im in zip(streams, images)]
        torch.cuda.synchronize()
        for r, want in zip(results, serial):
            assert np.array_equal(r.toHalfArray().reshape(8, 1000).view(np.uint16), want.view(np.uint16))
    whole.close()
    half.close()


def test_mobilenet_v2_batch1(ctx):
    from smelter_b200 import modelzoo, onnx2mps

    model = onnx2mps.convert_bytes(modelzoo.mobilenet_v2(seed=0, fold_bn=False).serialize(), half=True)
    x = np.random.default_rng(1).random((1, 3, 224, 224), dtype=np.float32).astype(np.float16)
    out, launches = _run(ctx, model, x)
    want = _oracle(model, x)
    assert np.abs(out.reshape(want.shape) - want).max() <= TOL
    assert launches == 1 + 52 + 1 + 1  # boundary + conv (Clip and the 10 residual Adds fused into epilogues) + gap + fc


def test_transformer_net_256(ctx):
    from smelter_b200 import modelzoo, onnx2mps

    model = onnx2mps.convert_bytes(modelzoo.transformer_net(seed=0, hw=256).serialize(), half=True)
    x = np.random.default_rng(1).random((1, 3, 256, 256), dtype=np.float32).astype(np.float16)
    out, _ = _run(ctx, model, x)
    want = _oracle(model, x)
    assert out.shape == (1, 3, 256, 256)
    # InstanceNorm re-normalises every layer, so fp16 storage error does not grow; output magnitudes are O(1)
    assert np.abs(out - want).max() <= 2e-2
    assert np.abs(out - want).mean() <= 2e-3


def test_transformer_net_512_named_shape(ctx):
    """BASELINE.json configs[3] at its named shape, 1x3x512x512, against the fp32 oracle with the north_star tolerance (1e-2 max-abs)."""
    from smelter_b200 import modelzoo, onnx2mps

    model = onnx2mps.convert_bytes(modelzoo.transformer_net(seed=0, hw=512).serialize(), half=True)
    x = np.random.default_rng(1).random((1, 3, 512, 512), dtype=np.float32).astype(np.float16)
    out, _ = _run(ctx, model, x)
    want = _oracle(model, x)
    assert out.shape == (1, 3, 512, 512)
    err = np.abs(out - want)
    print(f"TransformerNet 512x512: max-abs {err.max():.3e}, mean-abs {err.mean():.3e}, |ref| max {np.abs(want).max():.2f}")
    assert err.max() <= TOL * max(1.0, float(np.abs(want).max()))
    assert err.mean() <= 2e-3


@pytest.mark.parametrize("half", [False, True], ids=["onnx", "onnx2mps-half"])
def test_conv_transpose_group_norm_pow(ctx, half):
    """The registry entries no BASELINE model uses (ONNXGraph.swift:116,143,154): ConvTranspose (3x3/2 with output_padding,
    4x4/2, 1x1) as a stride-1 convolution over the zero-stuffed input with the flipped filter, custom_group_norm, Pow — as plain
    ONNX and through the ONNX2MPS weight swizzle ([1,2,3,0] + 180 degree flip, ONNX2MPS.py:54-79)."""
    from smelter_b200 import modelzoo, onnx2mps

    model = modelzoo.decoder_ops(seed=3).serialize()
    if half:
        model = onnx2mps.convert_bytes(model, half=True)
    x = np.random.default_rng(4).standard_normal((3, 32, 10, 10)).astype(np.float16)
    out, _ = _run(ctx, model, x)
    want = _oracle(model, x)
    assert out.shape == (3, 8, 40, 40)
    assert np.abs(out - want).max() <= TOL


def test_batch_override_through_configuration_dims(ctx):
    """Configuration.dims overrides input dims by axis (ONNXGraph.swift:200-202); spatial override re-plans shapes."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Configuration, Image, ONNXGraph

    model = modelzoo.transformer_net(seed=5, hw=32, width_div=8).serialize()
    x = np.random.default_rng(3).random((1, 3, 48, 40), dtype=np.float32).astype(np.float16)
    g = ONNXGraph(model, Configuration(dims={2: 48, 3: 40}), context=ctx)
    out = g.metalGraph().encode(sourceImages=[Image.fromArray(ctx, x)]).toFloatArray()
    assert out.shape == (1, 3, 48, 40)
    assert np.abs(out - _oracle(model, x)).max() <= 2e-2
    g.close()


def test_bilinear_upsample_honours_align_corners(ctx):
    from smelter_b200 import modelzoo
    from smelter_b200.api import BillinearUpsampling, Configuration, Image, ONNXGraph

    b = modelzoo.GraphBuilder(seed=1)
    xin = b.input("input", [1, 8, 6, 5])
    y = b.upsample(b.conv(xin, 8, 3, 1, 1), 2, mode="linear")
    b.output(y, [1, 8, 12, 10])
    model = b.model().serialize()
    x = np.random.default_rng(4).random((1, 8, 6, 5), dtype=np.float32).astype(np.float16)
    for align in (True, False):
        g = ONNXGraph(model, Configuration(billinearUpsamplingConfiguration=BillinearUpsampling(alignCorners=align)), context=ctx)
        out = g.metalGraph().encode(sourceImages=[Image.fromArray(ctx, x)]).toFloatArray()
        assert np.abs(out - _oracle(model, x, align_corners=align)).max() <= 5e-3
        g.close()


def test_error_behaviour_mirrors_the_reference(ctx):
    from smelter_b200 import modelzoo, onnx_proto as op
    from smelter_b200.api import Errors, Image, ONNXGraph

    m = modelzoo.conv_bn_relu(seed=0)
    m.graph.node[2].op_type = "Gelu"                       # not in the registry -> unknownNodeOpType(opType:) (ONNXGraph.swift:173-174)
    with pytest.raises(Errors) as e:
        ONNXGraph(m.serialize(), context=ctx).metalGraph()
    assert e.value.case == "unknownNodeOpType" and e.value.opType == "Gelu"

    m = modelzoo.conv_bn_relu(seed=0)
    m.graph.output.append(op.ValueInfo(name=m.graph.node[0].output[0], elem_type=op.FLOAT, dims=[1, 8, 16, 16]))
    with pytest.raises(Errors) as e:                       # exactly one output (ONNXGraph.swift:178-180)
        ONNXGraph(m.serialize(), context=ctx).metalGraph()
    assert e.value.case == "unsupportedOutput"

    m = modelzoo.conv_bn_relu(seed=0)
    m.graph.node[0].attribute = [a for a in m.graph.node[0].attribute if a.name != "kernel_shape"]
    with pytest.raises(Errors) as e:                       # Converters.swift:234-237
        ONNXGraph(m.serialize(), context=ctx).metalGraph()
    assert e.value.case == "notEnoughAttributes"

    m = modelzoo.conv_bn_relu(seed=0)
    m.graph.node[1].input[0] = "nowhere"
    with pytest.raises(Errors) as e:
        ONNXGraph(m.serialize(), context=ctx).metalGraph()
    assert e.value.case == "noSuchOutput"

    with pytest.raises(Errors) as e:
        ONNXGraph(b"\x0a\x7f\x01", context=ctx)
    assert e.value.case == "parse"

    g = ONNXGraph(modelzoo.conv_bn_relu(seed=0).serialize(), context=ctx)
    nn = g.metalGraph()
    with pytest.raises(Errors) as e:                       # wrong source shape
        nn.encode(sourceImages=[Image.fromArray(ctx, np.zeros((1, 3, 8, 8), np.float16))])
    assert e.value.case == "unsupportedInput"
    with pytest.raises(Errors) as e:
        nn.encode(sourceImages=[])
    assert e.value.code in (6, 100)
    g.close()


def test_host_language_converter_plugin(ctx):
    """NodeConverter protocol through the C ABI (NodeConverter.swift:3-5): a Python converter for a custom op."""
    import ctypes as C

    from smelter_b200 import _lib as L, modelzoo
    from smelter_b200.api import Image, ONNXGraph

    m = modelzoo.conv_bn_relu(seed=0)
    m.graph.node[2].op_type = "MyRelu6"
    g = ONNXGraph(m.serialize(), context=ctx)
    lib = L.lib()

    def convert(handle, node):
        name_in, name_out = C.c_char_p(), C.c_char_p()
        assert lib.smelter_node_input(handle, node, 0, C.byref(name_in)) == 0
        assert lib.smelter_node_output(handle, node, 0, C.byref(name_out)) == 0
        return lib.smelter_add_unary(handle, name_in.value, 2, 0.0, 6.0, name_out.value)  # Clip(0, 6)

    assert not g.hasConverter("MyRelu6")
    g.register("MyRelu6", convert)
    assert g.hasConverter("MyRelu6")
    x = (np.random.default_rng(0).random((1, 3, 16, 16), dtype=np.float32) * 8).astype(np.float16)
    out = g.metalGraph().encode(sourceImages=[Image.fromArray(ctx, x)]).toFloatArray()
    m.graph.node[2].op_type = "Relu"
    want = np.minimum(_oracle(m.serialize(), x), 6.0)
    assert np.abs(out - want).max() <= 1e-2
    g.close()


def test_weight_arena_checksum_and_deferred_weights(ctx):
    from smelter_b200 import modelzoo
    from smelter_b200.api import Configuration, ONNXGraph

    model = modelzoo.conv_bn_relu(seed=0).serialize()
    a = ONNXGraph(model, context=ctx)
    b = ONNXGraph(model, context=ctx)
    c = ONNXGraph(model, Configuration(deferWeights=True), context=ctx)
    ca, cb, cc = a.metalGraph().weightChecksum(), b.metalGraph().weightChecksum(), c.metalGraph().weightChecksum()
    assert ca == cb and ca[1] > 0
    assert cc[0] == 0 and cc[1] == ca[1]  # zero-filled arena waiting for the broadcast
    ptr, nbytes = a.metalGraph().weightArena()
    assert ptr != 0 and nbytes == ca[1]
    for g in (a, b, c):
        g.close()


@pytest.mark.parametrize("scale", ["bilinear", "lanczos"])
def test_force_input_scale_resizes_sources(ctx, scale):
    """Configuration(inputConstraint: .forceInputScale(...)) (ONNXGraph.swift:219-241): sources of any H x W are resampled to the graph
    input in front of the path.  MPS's scale nodes are closed source; the kernels are defined as half-pixel-centre bilinear (== torch
    align_corners=False) and Lanczos-3 with edge clamping, restated in numpy below."""
    from smelter_b200 import modelzoo, onnx2mps

    model = onnx2mps.convert_bytes(modelzoo.conv_bn_relu(seed=0).serialize(), half=True)   # graph input [N,3,16,16]
    x = np.random.default_rng(9).random((2, 3, 23, 29), dtype=np.float32).astype(np.float16)

    def lanczos(t):
        t = np.abs(t)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = 3.0 * np.sin(np.pi * t) * np.sin(np.pi * t / 3.0) / (np.pi * t) ** 2
        return np.where(t < 1e-6, 1.0, np.where(t >= 3.0, 0.0, v))

    def resample_axis(a, n_out, axis):
        n_in = a.shape[axis]
        pos = (np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5
        base = np.floor(pos).astype(int)
        taps = range(0, 2) if scale == "bilinear" else range(-2, 4)
        out, wsum = 0.0, 0.0
        for d in taps:
            idx = np.clip(base + d, 0, n_in - 1)
            w = (1.0 - np.abs(pos - (base + d))) if scale == "bilinear" else lanczos(pos - (base + d))
            shape = [1] * a.ndim
            shape[axis] = n_out
            out = out + np.take(a, idx, axis=axis) * w.reshape(shape)
            wsum = wsum + w.reshape(shape)
        return out / wsum

    xr = resample_axis(resample_axis(x.astype(np.float64), 16, 2), 16, 3).astype(np.float16)  # separable: same as the 2-D kernel
    if scale == "bilinear":
        want_t = torch.nn.functional.interpolate(torch.from_numpy(x.astype(np.float32)), size=(16, 16), mode="bilinear", align_corners=False)
        assert np.abs(want_t.numpy() - xr.astype(np.float32)).max() <= 2e-3  # the numpy restatement is torch's definition
    out, _ = _run(ctx, model, x, inputConstraint=scale)
    want = _oracle(model, xr)
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= TOL
    with pytest.raises(Exception):  # without the constraint a mis-sized source is still unsupportedInput
        _run(ctx, model, x)


@pytest.mark.parametrize("c_in,c_out,k,hw,pad_mode", [(3, 32, 9, (40, 56), "reflect"), (16, 24, 5, (21, 28), "edge"), (32, 64, 5, (18, 20), "reflect"),
                                                      (20, 16, 3, (11, 34), "reflect"), (8, 40, 9, (26, 48), "reflect")])
def test_width_folded_input_convolution(ctx, monkeypatch, c_in, c_out, k, hw, pad_mode):
    """A narrow-input stride-1 convolution behind an explicit Pad runs on the width-folded view of its NHWC buffers (64 / pitch
    neighbouring pixels as one; engine.cc "width-folded"; needs k_w - 1 and the width divisible by the fold factor): same result as the
    plain packed-row form and the oracle, fold factors 8 / 4 / 2, a second convolution behind it so that the re-interpreted output is read as plain NHWC again."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c_in + k, name="wfold")
    x = b.input("input", [2, c_in, h, w])
    y = b.relu(b.conv(b.pad(x, k // 2, pad_mode), c_out, k, 1, 0))
    y = b.conv(y, 8, 1)
    b.output(y, [2, 8, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(k).standard_normal((2, c_in, h, w)).astype(np.float16)
    want = _oracle(model, xin)

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray()
        dump = nn.planDump(2)
        g.close()
        return out, dump

    out, dump = run()
    assert "width-fold" in dump
    assert np.abs(out - want).max() <= TOL * max(1.0, np.abs(want).max())
    monkeypatch.setenv("SMELTER_NO_WIDTH_FOLD", "1")
    plain, dump0 = run()
    assert "width-fold" not in dump0
    assert np.abs(out - plain).max() <= 4e-3 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("c_in,c_out,k,hw,fold", [(32, 3, 9, (24, 40), 4), (16, 8, 5, (30, 18), 2), (24, 1, 7, (16, 16), 2),
                                                  (32, 2, 13, (20, 28), 4), (16, 4, 9, (32, 32), 4), (32, 3, 9, (26, 40), 2)])
def test_phase_folded_output_convolution(ctx, monkeypatch, c_in, c_out, k, hw, fold):
    """A narrow-output stride-1 convolution that produces the graph output behind a Pad runs on the F x F space-to-depth fold of its
    padded input (F = 4 where the sizes divide, else 2) with the F^2 output phases as GEMM columns (engine.cc "phase-folded"): the
    Pad writes the fold, the final conversion un-folds; same result as the unfolded plan, the 2 x 2 fold and the oracle."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c_in + k, name="pfold")
    x = b.input("input", [2, c_in, h, w])
    y = b.conv(b.pad(b.relu(b.conv(x, c_in, 1)), k // 2, "reflect"), c_out, k, 1, 0)
    b.output(y, [2, c_out, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(k).standard_normal((2, c_in, h, w)).astype(np.float16)
    want = _oracle(model, xin)

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray()
        dump = nn.planDump(2)
        g.close()
        return out, dump

    out, dump = run()
    assert "phase-fold" in dump and "pad+s2d" in dump and "phase_to_nchw" in dump
    assert ("phase-fold4" in dump) == (fold == 4)
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= TOL * max(1.0, np.abs(want).max())
    if fold == 4:
        monkeypatch.setenv("SMELTER_PHASE_FOLD_2", "1")
        by2, dump2 = run()
        assert "phase-fold" in dump2 and "phase-fold4" not in dump2
        assert np.abs(out - by2).max() <= 4e-3 * max(1.0, np.abs(want).max())
    monkeypatch.setenv("SMELTER_NO_PHASE_FOLD", "1")
    plain, dump0 = run()
    assert "phase-fold" not in dump0
    assert np.abs(out - plain).max() <= 4e-3 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("c_in,c_out,hw,batch,cluster", [(64, 32, (12, 20), 2, True), (16, 8, (9, 7), 1, True), (128, 64, (16, 16), 1, False),
                                                        (32, 16, (33, 18), 3, True)])
def test_upsample_folded_convolution(ctx, monkeypatch, c_in, c_out, hw, batch, cluster):
    """nearest Upsample x2 -> reflect Pad 1 -> Conv 3x3 -> InstanceNorm (TransformerNet's decoder stages) runs as one 3x3 convolution
    on the edge-padded low-resolution image with the four output phases as GEMM columns and pre-summed taps; the normalisation
    un-permutes the pixels (engine.cc "upsample-folded").  Same result as the literal plan and the oracle, in both forms of the
    normalisation (cluster / three launches)."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c_in + h, name="upfold")
    x = b.input("input", [batch, c_in, h, w])
    y = b.relu(b.instancenorm(b.conv(b.pad(b.upsample(b.relu(b.conv(x, c_in, 1)), 2), 1, "reflect"), c_out, 3, 1, 0)))
    y = b.conv(y, 8, 1)
    b.output(y, [batch, 8, 2 * h, 2 * w])
    model = b.model().serialize()
    xin = np.random.default_rng(h).standard_normal((batch, c_in, h, w)).astype(np.float16)
    want = _oracle(model, xin)
    if not cluster:
        monkeypatch.setenv("SMELTER_NO_CLUSTER_NORM", "1")
        monkeypatch.setenv("SMELTER_NO_CONV_STATS", "1")  # the norm's own statistics in every plan below

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray()
        dump, n = nn.planDump(batch), nn.numLaunches(batch)
        g.close()
        return out, dump, n

    out, dump, n = run()
    assert "upsample-fold" in dump and "instance_norm+unfold" in dump and "upsample " not in dump
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= TOL * max(1.0, np.abs(want).max())
    monkeypatch.setenv("SMELTER_NO_UPSAMPLE_FOLD", "1")
    plain, dump0, n0 = run()
    assert "upsample-fold" not in dump0 and "upsample " in dump0 and n0 == n + 1
    assert np.abs(out - plain).max() <= 6e-3 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("c,hw,pad,k,batch,cluster,mode,second_reader",
                         [(32, (16, 24), 1, 3, 2, True, "reflect", False), (16, (20, 12), 4, 9, 1, True, "reflect", False),
                          (64, (9, 11), 2, 5, 3, True, "reflect", True), (32, (16, 24), 1, 3, 2, False, "reflect", True),
                          (32, (24, 40), 4, 9, 2, False, "reflect", False), (24, (7, 5), 3, 7, 1, True, "reflect", False),
                          (32, (10, 14), 2, 5, 2, True, "edge", False), (16, (12, 8), 1, 3, 1, False, "edge", True)])
def test_pad_written_by_the_instance_norm(ctx, monkeypatch, c, hw, pad, k, batch, cluster, mode, second_reader):
    """InstanceNorm (+ReLU) -> reflect / edge Pad: the norm stores the padded image itself (interior at an offset, border pixels
    re-read from their mirror sources; space-to-depth layout when the Pad feeds a phase-folded convolution; the plain image too while
    a skip connection reads it) in both of its forms.  Same result as the plan with the Pad kernel, bit for bit (same values, same
    positions), and the oracle."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c + pad, name="normpad")
    x = b.input("input", [batch, c, h, w])
    y = b.relu(b.instancenorm(b.conv(x, c, 3, 1, 1)))
    c_out = 3 if k == 9 else c if second_reader else 16
    t = b.conv(b.pad(y, pad, mode), c_out, k, 1, 0)
    if second_reader:
        t = b.add(t, y)
    b.output(t, [batch, c_out, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(pad).standard_normal((batch, c, h, w)).astype(np.float16)
    want = _oracle(model, xin)
    if not cluster:
        monkeypatch.setenv("SMELTER_NO_CLUSTER_NORM", "1")
        monkeypatch.setenv("SMELTER_NO_CONV_STATS", "1")  # the norm's own statistics in every plan below

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toHalfArray().copy()
        dump, n = nn.planDump(batch), nn.numLaunches(batch)
        g.close()
        return out, dump, n

    out, dump, n = run()
    assert "instance_norm+pad" in dump and "pad Pad" not in dump
    assert ("+pad+s2d" in dump) == (k == 9) and ("+pad+plain" in dump) == second_reader
    assert np.abs(out.astype(np.float32) - want).max() <= TOL * max(1.0, np.abs(want).max())
    monkeypatch.setenv("SMELTER_NO_NORM_PAD", "1")
    plain, dump0, n0 = run()
    assert "instance_norm+pad" not in dump0 and n0 == n + 1
    assert np.array_equal(out.view(np.uint16), plain.view(np.uint16))


@pytest.mark.parametrize("mode,relu,second_reader,hw,batch", [("reflect", False, True, (16, 24), 2), ("reflect", True, False, (9, 7), 1),
                                                             ("constant", False, True, (12, 12), 3), ("edge", True, True, (20, 10), 1)])
def test_residual_add_inside_the_pad_kernel(ctx, monkeypatch, mode, relu, second_reader, hw, batch):
    """Add (-> ReLU) -> Pad: the Pad kernel adds on its way and also stores the plain sum while another filter reads it (the next
    block's skip connection).  Bit-identical to the two-kernel plan; oracle within tolerance."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    c = 32
    b = modelzoo.GraphBuilder(seed=h, name="padadd")
    x = b.input("input", [batch, c, h, w])
    a = b.relu(b.conv(x, c, 3, 1, 1))
    z = b.instancenorm(b.conv(a, c, 1))  # a norm in between: the Add is not a convolution epilogue
    y = b.add(z, a)
    if relu:
        y = b.relu(y)
    t = b.conv(b.pad(y, 2, mode), c, 5, 1, 0)
    if second_reader:
        t = b.add(t, y)
    b.output(t, [batch, c, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(h).standard_normal((batch, c, h, w)).astype(np.float16)
    want = _oracle(model, xin)

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toHalfArray().copy()
        dump, n = nn.planDump(batch), nn.numLaunches(batch)
        g.close()
        return out, dump, n

    out, dump, n = run()
    assert ("add+pad+sum" in dump) == second_reader and "add+pad" in dump
    assert np.abs(out.astype(np.float32) - want).max() <= TOL * max(1.0, np.abs(want).max())
    monkeypatch.setenv("SMELTER_NO_PAD_ADD", "1")
    plain, dump0, n0 = run()
    assert "add+pad" not in dump0 and n0 == n + 1
    assert np.array_equal(out.view(np.uint16), plain.view(np.uint16))


@pytest.mark.parametrize("c_in,c_out,k,hw,mode,batch,cluster", [(3, 32, 9, (24, 40), "reflect", 2, True), (1, 8, 5, (16, 12), "constant", 1, True),
                                                               (4, 16, 13, (32, 28), "edge", 3, True), (3, 32, 9, (32, 32), "reflect", 1, False)])
def test_input_folded_convolution(ctx, monkeypatch, c_in, c_out, k, hw, mode, batch, cluster):
    """graph input -> Pad -> Conv k x k (k = 1 mod 4) -> InstanceNorm (TransformerNet's input layer): the boundary conversion writes
    the padded image folded 4 x 4 with 4 channels per pixel, the convolution computes the 16 output phases of a folded pixel as
    GEMM columns, the norm un-permutes (engine.cc "input-folded").  Same result as the width-folded / plain plans and the oracle."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=k + c_in, name="infold")
    x = b.input("input", [batch, c_in, h, w])
    y = b.relu(b.instancenorm(b.conv(b.pad(x, k // 2, mode), c_out, k, 1, 0)))
    y = b.conv(b.pad(y, 1, "reflect"), 8, 3, 1, 0)
    b.output(y, [batch, 8, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(k).standard_normal((batch, c_in, h, w)).astype(np.float16)
    want = _oracle(model, xin)
    if not cluster:
        monkeypatch.setenv("SMELTER_NO_CLUSTER_NORM", "1")
        monkeypatch.setenv("SMELTER_NO_CONV_STATS", "1")  # the norm's own statistics in every plan below

    def run():
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray()
        dump, n = nn.planDump(batch), nn.numLaunches(batch)
        g.close()
        return out, dump, n

    out, dump, n = run()
    assert "input-fold4" in dump and "nchw_to_s2d4+pad" in dump and "instance_norm+unfold+pad" in dump
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= TOL * max(1.0, np.abs(want).max())
    monkeypatch.setenv("SMELTER_NO_INPUT_FOLD", "1")
    plain, dump0, n0 = run()
    assert "input-fold4" not in dump0 and n0 == n + 1
    assert np.abs(out - plain).max() <= 6e-3 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("kind,c_in,c_out,hw,batch", [("plain", 32, 32, (16, 24), 2), ("plain", 32, 32, (128, 128), 1), ("plain", 64, 128, (32, 32), 1), ("plain", 16, 24, (8, 16), 3),
                                                      ("plain", 128, 256, (16, 16), 5), ("plain1x1", 64, 320, (16, 8), 2), ("stride2", 32, 64, (32, 32), 2),
                                                      ("upfold", 64, 32, (16, 16), 2), ("infold", 3, 32, (64, 32), 2), ("ragged", 32, 32, (10, 12), 2)])
def test_instance_norm_statistics_from_the_convolution_epilogue(ctx, monkeypatch, kind, c_in, c_out, hw, batch):
    """Conv -> InstanceNorm with output rows per image a multiple of 128: the two-CTA convolution kernel adds every tile's column sums
    and sums of squares to per-image fp64 accumulators (ConvKernelParams::stats) and the norm is one pass
    (k::instance_norm_from_stats; phase-column convolutions keep a column per (phase, channel)).  Same result as the plan with the
    norm's own statistics and the oracle; repeated encodes agree (the norm's last block re-zeroes the accumulators); ragged image
    sizes keep the norm's own statistics."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c_in + c_out, name="convstats")
    x = b.input("input", [batch, c_in, h, w])
    oh, ow = h, w
    if kind == "upfold":
        y = b.conv(b.pad(b.upsample(b.relu(b.conv(x, c_in, 1)), 2), 1, "reflect"), c_out, 3, 1, 0)
        oh, ow = 2 * h, 2 * w
    elif kind == "infold":
        y = b.conv(b.pad(x, 4, "reflect"), c_out, 9, 1, 0)
    elif kind == "plain1x1":
        y = b.conv(b.relu(b.conv(x, c_in, 3, 1, 1)), c_out, 1)
    elif kind == "stride2":
        y = b.conv(b.relu(b.conv(x, c_in, 3, 1, 1)), c_out, 3, 2, 1)
        oh, ow = h // 2, w // 2
    else:
        y = b.conv(b.relu(b.conv(x, c_in, 3, 1, 1)), c_out, 3, 1, 1)
    y = b.relu(b.instancenorm(y))
    y = b.conv(b.pad(y, 1, "reflect"), 8, 3, 1, 0)
    b.output(y, [batch, 8, oh, ow])
    model = b.model().serialize()
    xin = (np.random.default_rng(h).standard_normal((batch, c_in, h, w)) * 2 + 0.5).astype(np.float16)
    want = _oracle(model, xin)

    def run(repeats=1):
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        outs = [nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray().copy() for _ in range(repeats)]
        dump, n = nn.planDump(batch), nn.numLaunches(batch)
        g.close()
        return outs, dump, n

    outs, dump, n = run(3)
    src = [l for l in dump.splitlines() if " Conv " in l][-2]  # the convolution in front of the norm
    fused = kind != "ragged" and src.startswith("conv_pair[")   # small problems are planned on the single-CTA kernel (32-column tiles / split-K)
    assert fused or kind == "ragged" or c_out < 64
    assert ("+stats" in src) == fused and ("<-stats" in dump) == fused
    scale = max(1.0, np.abs(want).max())
    assert outs[0].shape == want.shape
    assert np.abs(outs[0] - want).max() <= TOL * scale
    for o in outs[1:]:  # fp64 reductions in a different order at most: far below one fp16 step of the result
        assert np.abs(o - outs[0]).max() <= 1e-3 * scale
    monkeypatch.setenv("SMELTER_NO_CONV_STATS", "1")
    (plain,), dump0, n0 = run()
    assert "stats" not in dump0 and n0 >= n
    assert np.abs(outs[0] - plain).max() <= 4e-3 * scale


@pytest.mark.parametrize("mode,relu,second_reader,c,hw,batch", [("reflect", False, True, 128, (32, 32), 1), ("reflect", True, False, 64, (16, 24), 2),
                                                               ("edge", False, True, 64, (32, 16), 3), ("reflect", False, False, 128, (16, 16), 2),
                                                               ("constant", False, True, 64, (16, 16), 1)])
def test_residual_tail_inside_the_one_pass_norm(ctx, monkeypatch, mode, relu, second_reader, c, hw, batch):
    """Conv -> InstanceNorm -> Add(skip) (-> ReLU) -> Pad, the end of a TransformerNet residual block: with the statistics from the
    convolution's epilogue the norm's single pass also reads the skip tensor and stores the padded sum (and the plain sum while the
    next block reads it) -- the Add+Pad kernel runs inside the norm's launch (engine.cc "norm tail").  Same result as the plan with
    the separate Add+Pad kernel and the oracle; constant pads keep their own kernel."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=h + c, name="normtail")
    x = b.input("input", [batch, c, h, w])
    a = b.relu(b.conv(x, c, 3, 1, 1))
    z = b.instancenorm(b.conv(b.pad(a, 1, "reflect"), c, 3, 1, 0))
    y = b.add(z, a)
    if relu:
        y = b.relu(y)
    t = b.conv(b.pad(y, 1, mode), c, 3, 1, 0)
    if second_reader:
        t = b.add(t, y)
    b.output(t, [batch, c, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(h).standard_normal((batch, c, h, w)).astype(np.float16)
    want = _oracle(model, xin)

    def run(repeats=1):
        g = ONNXGraph(model, context=ctx)
        nn = g.metalGraph()
        outs = [nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray().copy() for _ in range(repeats)]
        dump, n = nn.planDump(batch), nn.numLaunches(batch)
        g.close()
        return outs, dump, n

    outs, dump, n = run(2)
    fused = mode != "constant"
    assert "<-stats" in dump
    assert ("+add+pad" in dump and "add+pad " not in dump) == fused, dump
    assert ("+add+pad+sum<-stats" in dump) == (fused and second_reader)
    scale = max(1.0, np.abs(want).max())
    assert np.abs(outs[0] - want).max() <= TOL * scale
    assert np.abs(outs[1] - outs[0]).max() <= 1e-3 * scale
    monkeypatch.setenv("SMELTER_NO_NORM_TAIL", "1")
    (plain,), dump0, n0 = run()
    assert "+add+pad" not in dump0 and n0 == n + (1 if fused else 0)
    assert np.abs(outs[0] - plain).max() <= 2e-3 * scale


@pytest.mark.parametrize("c_in,c_out,groups,k,stride,hw,batch", [(64, 64, 2, 3, 1, (14, 14), 2), (32, 64, 4, 3, 2, (16, 12), 1), (128, 128, 32, 3, 1, (8, 8), 3),
                                                                 (24, 48, 3, 1, 1, (9, 7), 2), (16, 32, 16, 3, 1, (10, 10), 1)])
def test_grouped_convolution_as_block_diagonal_dense(ctx, c_in, c_out, groups, k, stride, hw, batch):
    """Grouped convolutions other than depthwise (ResNeXt-style cardinality, channel multipliers; Converters.swift:57-75) run as the
    dense convolution with block-diagonal weights on the tensor-core path; depthwise keeps its own kernel.  Engine vs the oracle's
    grouped F.conv2d."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c_in + groups, name="grouped")
    x = b.input("input", [batch, c_in, h, w])
    y = b.relu(b.conv(x, c_in, 1))
    y = b.relu(b.conv(y, c_out, k, stride, k // 2, groups=groups))
    y = b.conv(y, c_out, 3, 1, 1, groups=c_out)  # depthwise: its own kernel
    oh, ow = (h + 2 * (k // 2) - k) // stride + 1, (w + 2 * (k // 2) - k) // stride + 1
    b.output(y, [batch, c_out, oh, ow])
    model = b.model().serialize()
    xin = np.random.default_rng(groups).standard_normal((batch, c_in, h, w)).astype(np.float16)
    want = _oracle(model, xin)
    g = ONNXGraph(model, context=ctx)
    nn = g.metalGraph()
    out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray()
    dump = nn.planDump(batch)
    g.close()
    assert f"/groups{groups}-as-dense" in dump and "depthwise" in dump
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= TOL * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("c,hw,batch,op,gate_first", [(64, (14, 14), 2, "Mul", False), (20, (9, 7), 3, "Mul", True), (40, (28, 28), 1, "Add", False),
                                                     (24, (5, 5), 2, "Sub", False), (136, (7, 7), 4, "Div", False)])
def test_pixel_operand_broadcast_in_binary_ops(ctx, c, hw, batch, op, gate_first):
    """[N, C, 1, 1] against [N, C, H, W] (the gate of a squeeze-and-excitation block: GlobalAveragePool / ReduceMean -> 1x1
    convolutions -> Sigmoid -> Mul): the one-pixel operand is broadcast over the image, also with C % 8 != 0 and as the first operand
    of a commutative op; ReduceMean over {2, 3} is the global pool.  Engine vs oracle."""
    from smelter_b200 import modelzoo
    from smelter_b200.api import Image, ONNXGraph

    h, w = hw
    b = modelzoo.GraphBuilder(seed=c + h, name="se")
    x = b.input("input", [batch, c, h, w])
    a = b.relu(b.conv(x, c, 3, 1, 1))
    pooled = b._node("ReduceMean", [a], {"axes": [2, 3], "keepdims": 1}, c) if c % 16 == 8 else b.gap(a)
    gate = b.sigmoid(b.conv(b.relu(b.conv(pooled, max(8, c // 4), 1)), c, 1))
    y = b._node(op, [gate, a] if gate_first else [a, gate], {}, c)
    y = b.conv(y, 16, 1)
    b.output(y, [batch, 16, h, w])
    model = b.model().serialize()
    xin = np.random.default_rng(c).standard_normal((batch, c, h, w)).astype(np.float16)
    want = _oracle(model, xin)
    g = ONNXGraph(model, context=ctx)
    nn = g.metalGraph()
    out = nn.encode(sourceImages=[Image.fromArray(ctx, xin)]).toFloatArray()
    dump = nn.planDump(batch)
    g.close()
    assert "binary/broadcast" in dump
    assert out.shape == want.shape and np.isfinite(out).all()
    assert np.abs(out - want).max() <= TOL * max(1.0, np.abs(want).max())
