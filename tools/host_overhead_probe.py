"""Host cost of one encode() call (Python + ctypes + boundary launch + cudaGraphLaunch): after a sync, time 20 back-to-back calls
(the GPU is slower than the host, so the launch queue does not block yet)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph
model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
rng = np.random.default_rng(0)
ctx = Context(0)
g = ONNXGraph(model, Configuration(), context=ctx)
nn = g.metalGraph()
for B in (1, 32):
    im = Image.fromArray(ctx, rng.random((B, 3, 224, 224), dtype=np.float32).astype(np.float16))
    for _ in range(5): nn.encode(sourceImages=[im])
    torch.cuda.synchronize()
    for rep in range(3):
        ts = []
        t_all = time.perf_counter()
        for _ in range(20):
            t0 = time.perf_counter(); nn.encode(sourceImages=[im]); ts.append(time.perf_counter() - t0)
        t_host = time.perf_counter() - t_all
        torch.cuda.synchronize()
        t_total = time.perf_counter() - t_all
        print(f"batch {B}: host {t_host/20*1e6:7.1f} us per encode call (min {min(ts)*1e6:6.1f}, max {max(ts)*1e6:6.1f}); 20 encodes done after {t_total*1e3:.3f} ms", flush=True)
