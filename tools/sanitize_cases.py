"""A few small cases of every kernel family for `compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_cases.py`
(pytest itself does not start under the sanitizer in this image).  Checks results against fp32 torch like the tests do."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from oracle.onnx_interp import Interpreter
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph, run_conv
from tools.gpu_probe import CONV_CASES

ctx = Context(0)
want = {"tiled_16_10_5_tail", "im2col_3x3_odd_16", "rows_3x3_c8", "im2col_3x3_512_7", "splitk_gemm_like_b4", "dw_3x3_c20", "tiled_256_64_56_res",
        "rows_7x7_s2_stem", "im2col_3x3_dil2"}
for name, shape, co, k, s, p, d, g, act, has_bias, has_res, force in CONV_CASES:
    if name not in want:
        continue
    rng = np.random.default_rng(1)
    n, c, h, w = shape
    x = rng.standard_normal(shape).astype(np.float16)
    wt = (rng.standard_normal((co, c // g, k, k)) * np.sqrt(2.0 / (c // g * k * k))).astype(np.float16)
    b = rng.standard_normal(co).astype(np.float32) if has_bias else None
    ref = F.conv2d(torch.from_numpy(x.astype(np.float32)), torch.from_numpy(wt.astype(np.float32)), torch.from_numpy(b) if has_bias else None,
                   stride=s, padding=p, dilation=d, groups=g)
    r = rng.standard_normal(tuple(ref.shape)).astype(np.float16) if has_res else None
    if has_res:
        ref = ref + torch.from_numpy(r.astype(np.float32))
    ref = ref.relu() if act == 1 else (ref.clamp(0, 6) if act == 2 else ref)
    y, _ = run_conv(ctx, Image.fromArray(ctx, x), wt, b, stride=(s, s), pads=(p, p, p, p), dilation=(d, d), groups=g, act=act, clip=(0.0, 6.0),
                    residual=Image.fromArray(ctx, r) if has_res else None, force_path=force)
    err = float(np.abs(y.toFloatArray() - ref.numpy()).max())
    print(f"conv {name}: max_abs_err {err:.3e}", flush=True)
    assert err < 4e-3 * max(1.0, float(ref.abs().max()))


def model(bytes_, x, tol, what, **env):
    for k_, v in env.items():
        os.environ[k_] = v
    gph = ONNXGraph(bytes_, Configuration(), context=ctx)
    nn = gph.metalGraph()
    out = nn.encode(sourceImages=[Image.fromArray(ctx, x)]).toFloatArray()
    ref = Interpreter(bytes_).run(torch.from_numpy(x.astype(np.float32))).numpy().reshape(out.shape)
    err = float(np.abs(out - ref).max())
    print(f"model {what}: launches {nn.numLaunches(x.shape[0])} folded shortcuts {nn.planDump(x.shape[0]).count('+conv1x1(')} max_abs_err {err:.3e}", flush=True)
    assert err <= tol
    gph.close()
    for k_ in env:
        os.environ.pop(k_)


rng = np.random.default_rng(0)
small = modelzoo.resnet50(seed=0, fold_bn=True, num_classes=16, hw=64, depths=(1, 1, 1, 1)).serialize()
x = rng.random((2, 3, 64, 64), dtype=np.float32).astype(np.float16)
model(small, x, 1e-2, "resnet bottlenecks (per-layer)")
ragged = modelzoo.resnet50(seed=1, fold_bn=True, num_classes=16, hw=40, depths=(1, 1, 1, 1)).serialize()
model(ragged, rng.random((3, 3, 40, 40), dtype=np.float32).astype(np.float16), 1e-2, "resnet bottlenecks, ragged tiles, shortcuts folded", SMELTER_NO_SPLITK="1")
model(modelzoo.decoder_ops(seed=3).serialize(), rng.standard_normal((1, 32, 10, 10)).astype(np.float16), 1e-2, "conv_transpose / group_norm / pow")
model(modelzoo.synthetic_ops(seed=0).serialize(), rng.random((2, 16, 12, 12), dtype=np.float32).astype(np.float16), 1e-2, "sigmoid / concat / avgpool / softmax")
tn = onnx2mps.convert_bytes(modelzoo.transformer_net(seed=0, hw=64, width_div=4).serialize(), half=True)
model(tn, rng.random((1, 3, 64, 64), dtype=np.float32).astype(np.float16), 3e-2, "transformer_net (pad / instance norm / upsample)")
# epilogue statistics, one-pass norms, residual tails (round 2): a TransformerNet large enough for the two-CTA kernel at every layer
tn2 = onnx2mps.convert_bytes(modelzoo.transformer_net(seed=1, hw=128).serialize(), half=True)
gph = ONNXGraph(tn2, Configuration(), context=ctx)
dump = gph.metalGraph().planDump(2)
print(f"transformer_net 128x128 batch 2: {dump.count('+stats')} convolutions with statistics, {dump.count('<-stats')} one-pass norms, {dump.count('+add+pad')} tails", flush=True)
assert dump.count("+stats") >= 10 and dump.count("+add+pad") >= 4
gph.close()
model(tn2, rng.random((2, 3, 128, 128), dtype=np.float32).astype(np.float16), 3e-2, "transformer_net 128x128 (epilogue statistics / one-pass norm / tails)")
# grouped convolution as block-diagonal dense, broadcast gate, ReduceMean
b = modelzoo.GraphBuilder(seed=5, name="se_grouped")
xi = b.input("input", [2, 24, 9, 7])
a_ = b.relu(b.conv(xi, 48, 3, 1, 1, groups=3))
gate = b.sigmoid(b.conv(b._node("ReduceMean", [a_], {"axes": [2, 3], "keepdims": 1}, 48), 48, 1))
yo = b.conv(b._node("Mul", [a_, gate], {}, 48), 20, 1)
b.output(yo, [2, 20, 9, 7])
model(b.model().serialize(), rng.standard_normal((2, 24, 9, 7)).astype(np.float16), 1e-2, "grouped conv / ReduceMean / broadcast Mul")
px = rng.integers(0, 256, size=(1, 9, 11, 4), dtype=np.uint8)
got = Image.fromBytes(ctx, px, channels=3).toHalfArray()
assert got.shape == (1, 3, 9, 11)
print("ok")
