"""Do two launches of the same convolution on two streams overlap on the SMs?  Two host threads, each with its own context / stream,
run `iters` back-to-back launches of one layer through smelter_run_conv; prints the per-launch time alone and when both run."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smelter_b200.api import Context, Image, run_conv

LAYERS = [("s3_3x3_256", 256, 14, 256, 3, 1, 1, False), ("s3_1x1_1024_256", 1024, 14, 256, 1, 1, 0, False), ("s4_3x3_512", 512, 7, 512, 3, 1, 1, False),
          ("s3_1x1_256_1024_res", 256, 14, 1024, 1, 1, 0, True), ("s1_3x3_64", 64, 56, 64, 3, 1, 1, False)]
B = 32
rng = np.random.default_rng(0)
streams = [torch.cuda.Stream() for _ in range(2)]
ctxs = [Context(0, stream=s.cuda_stream) for s in streams]
for name, ci, hw, co, k, s, p, res in LAYERS:
    data = []
    for c in ctxs:
        x = Image.fromArray(c, rng.standard_normal((B, ci, hw, hw)).astype(np.float16))
        w = (rng.standard_normal((co, ci, k, k)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float16)
        b = rng.standard_normal(co).astype(np.float32)
        oh = (hw + 2 * p - k) // s + 1
        r = Image.fromArray(c, rng.standard_normal((B, co, oh, oh)).astype(np.float16)) if res else None
        data.append((x, w, b, r))
    out = [0.0, 0.0]

    def work(i, iters):
        x, w, b, r = data[i]
        _, ms = run_conv(ctxs[i], x, w, b, stride=(s, s), pads=(p, p, p, p), act=1, residual=r, iters=iters)
        out[i] = ms * 1e3

    work(0, 20)
    work(0, 2000)
    alone = out[0]
    ths = [threading.Thread(target=work, args=(i, 2000)) for i in range(2)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    wall = (time.perf_counter() - t0) * 1e6 / 2000
    print(f"{name:22s} alone {alone:6.2f} us/launch | two streams: {out[0]:6.2f} / {out[1]:6.2f} us/launch each, wall {wall:6.2f} us per pair of launches", flush=True)
