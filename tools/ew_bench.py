"""HBM-bound kernels on the path at sizes larger than L2 (126 MB): CUDA-event time of the kernel alone through
smelter_run_elementwise, algorithmic bytes = 2 B x (input elements + output elements), each counted once, fraction of the measured HBM copy peak.
Usage: python tools/ew_bench.py [substring] ; writes gpurun_out/ew_bench.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200.api import Context, Image, run_elementwise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
ctx = Context(0)
only = sys.argv[1] if len(sys.argv) > 1 else ""
rng = np.random.default_rng(0)

def img(shape):
    # random fp16 without a giant fp32 temporary
    a = rng.integers(0, 0x3C00, size=shape, dtype=np.uint16).view(np.float16)
    return Image.fromArray(ctx, a)

# name, op, in shape, out shape, kwargs, bytes(in,out elements incl. second input)
N = 64
CASES = [
    ("relu 64x256x56x56", "unary", (N, 256, 56, 56), (N, 256, 56, 56), dict(sub=0), 1, False),
    ("sigmoid 64x256x56x56", "unary", (N, 256, 56, 56), (N, 256, 56, 56), dict(sub=1), 1, False),
    ("add+relu 64x256x56x56", "binary", (N, 256, 56, 56), (N, 256, 56, 56), dict(sub=0, act=1), 2, False),
    ("batchnorm 64x256x56x56", "scale_shift", (N, 256, 56, 56), (N, 256, 56, 56), dict(act=1), 1, True),
    ("maxpool3x3s2 128x64x112x112", "pool", (128, 64, 112, 112), (128, 64, 56, 56), dict(sub=1, k_h=3, k_w=3, stride_h=2, stride_w=2, pad_h=1, pad_w=1), 1, False),
    ("global_avgpool 1024x2048x7x7", "global_avgpool", (1024, 2048, 7, 7), (1024, 2048, 1, 1), dict(), 1, False),
    ("softmax 65536x1000", "softmax", (65536, 1000, 1, 1), (65536, 1000, 1, 1), dict(sub=0), 1, False),
    ("upsample_nearest_x2 16x128x128x128", "upsample", (16, 128, 128, 128), (16, 128, 256, 256), dict(sub=0, scale_h=2, scale_w=2), 1, False),
    ("reflect_pad4 16x32x512x512", "pad", (16, 32, 512, 512), (16, 32, 520, 520), dict(sub=1, pad_h=4, pad_w=4, pad_b=4, pad_r=4), 1, False),
    ("concat 64x128+128x56x56", "concat", (N, 128, 56, 56), (N, 256, 56, 56), dict(c2=128), 2, False),
    ("instance_norm+relu 16x32x512x512", "instance_norm", (16, 32, 512, 512), (16, 32, 512, 512), dict(alpha=1e-5, act=1), 1, True),
    ("instance_norm 64x128x128x128", "instance_norm", (64, 128, 128, 128), (64, 128, 128, 128), dict(alpha=1e-5), 1, True),
    ("instance_norm+relu<-conv statistics 16x32x512x512", "instance_norm", (16, 32, 512, 512), (16, 32, 512, 512), dict(alpha=1e-5, act=1, sub=1), 1, True),
    ("instance_norm<-conv statistics 64x128x128x128", "instance_norm", (64, 128, 128, 128), (64, 128, 128, 128), dict(alpha=1e-5, sub=1), 1, True),
    ("layout nhwc copy 64x256x56x56", "layout_roundtrip", (N, 256, 56, 56), (N, 256, 56, 56), dict(), 1, False),
]
out = []
TRIALS = int(os.environ.get("EW_TRIALS", "3"))
for name, op, ishape, oshape, kw, n_in, params in CASES:
    if only and only not in name:
        continue
    # Where cudaMalloc places the operands matters on this part (lock-step streams that land on the same DRAM banks: the same
    # kernel measured 48 us and 1.6 ms for the two-input add depending on allocation history), so every case is timed with TRIALS
    # fresh placements (different buffer staggers) and the best and the median are reported.
    times = []
    for trial in range(TRIALS):
        os.environ["SMELTER_EW_STAGGER"] = str(trial)
        x = img(ishape)
        x2 = None
        if n_in == 2:
            s2 = (ishape[0], kw.get("c2", ishape[1]), ishape[2], ishape[3])
            x2 = img(s2)
        p0 = p1 = None
        if params:
            p0 = rng.uniform(0.5, 1.5, ishape[1]).astype(np.float32)
            p1 = rng.standard_normal(ishape[1]).astype(np.float32)
        y, ms = run_elementwise(ctx, op, x, x2=x2, p0=p0, p1=p1, out_shape=oshape, iters=10, **kw)
        times.append(ms)
        in_elems = int(np.prod(ishape)) + (int(np.prod(x2.shape)) if x2 is not None else 0)
        del x, x2, y
    byts = 2.0 * (in_elems + int(np.prod(oshape)))   # algorithmic: every input element read once, every output element written once
    ms = min(times)
    gbs = byts / (ms * 1e-3) / 1e9
    rec = {"kernel": name, "ms": round(ms, 4), "ms_median": round(sorted(times)[len(times) // 2], 4), "ms_all": [round(t, 4) for t in times],
           "algorithmic_MB": round(byts / 1e6, 1), "GB/s": round(gbs, 0), "frac_of_measured_hbm": round(gbs / PEAK, 3), "frac_of_8TBs": round(gbs / 8000.0, 3)}
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"hbm_peak_gbs_measured": PEAK, "trials_per_kernel": TRIALS, "reported": "best of the trials (ms); ms_median / ms_all beside it", "rows": out},
          open(os.path.join(ROOT, "gpurun_out", "ew_bench.json"), "w"), indent=1)
