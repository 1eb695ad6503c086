"""Does running two half batches on two streams beat one full batch?  (Kernel-boundary bubbles of one stream could be filled by
the other stream's kernels.)  Prints images/s for 1 x 32 and 2 x 16 (and 4 x 8)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
rng = np.random.default_rng(0)


def run(parts, iters=200):
    b = 32 // parts
    streams = [torch.cuda.Stream() for _ in range(parts)]
    ctxs = [Context(0, stream=s.cuda_stream) for s in streams]
    graphs = [ONNXGraph(model, Configuration(), context=c) for c in ctxs]
    nns = [g.metalGraph() for g in graphs]
    imgs = [Image.fromArray(c, rng.random((b, 3, 224, 224), dtype=np.float32).astype(np.float16)) for c in ctxs]
    for _ in range(5):
        for nn, im in zip(nns, imgs):
            nn.encode(sourceImages=[im])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        for nn, im in zip(nns, imgs):
            nn.encode(sourceImages=[im])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{parts} stream(s) x batch {b}: {32 * iters / dt:9.0f} images/s  ({dt / iters * 1e3:.4f} ms per 32 images)", flush=True)
    for g in graphs:
        g.close()


for parts in (1, 2, 4, 1, 2):
    run(parts)
