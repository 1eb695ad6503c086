"""Independent encodes in flight on separate streams: do the kernels of one batch fill the kernel-boundary bubbles of another?
Each stream has its own graph instance (own plan / activation arena) and runs full batches of 32.  Prints images/s for 1, 2, 3 streams.
Usage: python tools/dual_encode_probe.py [batch]   (env SMELTER_DUO=1: half-footprint conv kernels, two CTAs per SM)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
rng = np.random.default_rng(0)


def run(parts, iters=150):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    ctxs = [Context(0, stream=s.cuda_stream) for s in streams]
    graphs = [ONNXGraph(model, Configuration(), context=c) for c in ctxs]
    nns = [g.metalGraph() for g in graphs]
    imgs = [Image.fromArray(c, rng.random((B, 3, 224, 224), dtype=np.float32).astype(np.float16)) for c in ctxs]
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
    print(f"{parts} stream(s) x batch {B} (duo={'on' if os.environ.get('SMELTER_DUO') else 'off'}): {B * parts * iters / dt:9.0f} images/s  "
          f"({dt / iters / parts * 1e3:.4f} ms per batch)", flush=True)
    for g in graphs:
        g.close()


for parts in (1, 2, 3, 1, 2):
    run(parts)
