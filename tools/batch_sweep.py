"""ResNet-50 device-resident throughput vs batch size (fixed per-kernel cost shows up as the batch->0 intercept)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smelter_b200.api import Context, Image, ONNXGraph
from bench import model_bytes
stream = torch.cuda.Stream()
ctx = Context(0, stream=stream.cuda_stream)
g = ONNXGraph(model_bytes(), context=ctx); nn = g.metalGraph()
rng = np.random.default_rng(0)
out = []
for b in [int(v) for v in (sys.argv[1:] or "1 2 4 8 16 32 64 128 256".split())]:
    imgs = [Image.fromArray(ctx, rng.random((b, 3, 224, 224), dtype=np.float32).astype(np.float16)) for _ in range(max(2, min(16, 512 // b)))]
    for i in range(5): nn.encode(sourceImages=[imgs[i % len(imgs)]])
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 30
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(K): nn.encode(sourceImages=[imgs[i % len(imgs)]])
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out.append({"batch": b, "ms": ms, "img_s": b / ms * 1e3})
    print(json.dumps(out[-1]), flush=True)
    del imgs
