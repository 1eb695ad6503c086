"""Times individual ResNet-50 batch-32 conv layers through smelter_run_conv (kernel-only CUDA-event time, 20 back-to-back
launches).  Usage: python tools/conv_layers.py [layer-substring]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200.api import Context, Image, run_conv

B = int(os.environ.get("BATCH", "32"))
# name, c_in, hw, c_out, k, stride, pad, residual
LAYERS = [
    ("stem7x7", 3, 224, 64, 7, 2, 3, False),
    ("s1_1x1_64_64", 64, 56, 64, 1, 1, 0, False),
    ("s1_3x3_64", 64, 56, 64, 3, 1, 1, False),
    ("s1_1x1_64_256", 64, 56, 256, 1, 1, 0, False),
    ("s1_1x1_64_256_res", 64, 56, 256, 1, 1, 0, True),
    ("s1_1x1_256_64", 256, 56, 64, 1, 1, 0, False),
    ("s2_3x3_128_s2", 128, 56, 128, 3, 2, 1, False),
    ("s2_1x1_128_512_res", 128, 28, 512, 1, 1, 0, True),
    ("s2_3x3_128", 128, 28, 128, 3, 1, 1, False),
    ("s3_1x1_256_1024_res", 256, 14, 1024, 1, 1, 0, True),
    ("s3_3x3_256", 256, 14, 256, 3, 1, 1, False),
    ("s3_1x1_1024_256", 1024, 14, 256, 1, 1, 0, False),
    ("s4_3x3_512", 512, 7, 512, 3, 1, 1, False),
    ("s4_1x1_512_2048_res", 512, 7, 2048, 1, 1, 0, True),
    ("s4_1x1_2048_512", 2048, 7, 512, 1, 1, 0, False),
]
ctx = Context(0)
rng = np.random.default_rng(0)
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, ci, hw, co, k, s, p, res in LAYERS:
    if only and only not in name:
        continue
    x = Image.fromArray(ctx, rng.standard_normal((B, ci, hw, hw)).astype(np.float16))
    w = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float16)
    b = rng.standard_normal(co).astype(np.float32)
    oh = (hw + 2 * p - k) // s + 1
    r = Image.fromArray(ctx, rng.standard_normal((B, co, oh, oh)).astype(np.float16)) if res else None
    y, ms = run_conv(ctx, x, w, b, stride=(s, s), pads=(p, p, p, p), act=1, residual=r, iters=20)
    flops = 2.0 * B * oh * oh * co * ci * k * k
    byts = 2.0 * B * (hw * hw * max(ci, 8) + oh * oh * co * (2 if res else 1)) + 2.0 * co * ci * k * k
    print(json.dumps({"layer": name, "us": round(ms * 1e3, 2), "tflops": round(flops / ms / 1e9, 1), "gbs": round(byts / ms / 1e6, 0),
                      "hbm_bound_us": round(byts / 6.5e6, 1), "tc_bound_us": round(flops / 1.4e9, 1), "dbg": os.environ.get("SMELTER_CONV_DEBUG", "0")}), flush=True)
    del x, y, r
