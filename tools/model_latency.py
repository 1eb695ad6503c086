"""Latency / throughput of the other BASELINE.json configurations through the public API (CUDA-graph replay, inputs resident,
CUDA events around `iters` back-to-back encodes after 5 warm-ups).  Writes gpurun_out/model_latency.json.
Usage: python tools/model_latency.py [mobilenet|tnet|resnet ...]   (default: all)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

stream = torch.cuda.Stream()
ctx = Context(0, stream=stream.cuda_stream)
rows = []


def run(name, model, shape, iters=200, gflop_per_image=None):
    g = ONNXGraph(model, Configuration(), context=ctx)
    nn = g.metalGraph()
    x = np.random.default_rng(1).random(shape, dtype=np.float32).astype(np.float16)
    img = Image.fromArray(ctx, x)
    for _ in range(5):
        nn.encode(sourceImages=[img])
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(iters):
            nn.encode(sourceImages=[img])
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    row = {"config": name, "input": list(shape), "launches": nn.numLaunches(shape[0]), "ms_per_encode": round(ms, 4),
           "images_per_s": round(shape[0] / ms * 1e3, 1)}
    if gflop_per_image:
        row["tflops"] = round(gflop_per_image * shape[0] / ms, 1)
    rows.append(row)
    print(json.dumps(row), flush=True)
    g.close()


which = set(sys.argv[1:]) or {"mobilenet", "tnet", "resnet"}
if "mobilenet" in which:
    mb = onnx2mps.convert_bytes(modelzoo.mobilenet_v2(seed=0, fold_bn=False).serialize(), half=True)
    run("MobileNetV2 fp16 batch 1 (BASELINE configs[1])", mb, (1, 3, 224, 224), 500, 0.6015)
    run("MobileNetV2 fp16 batch 32", mb, (32, 3, 224, 224), 200, 0.6015)
if "tnet" in which:
    tn = onnx2mps.convert_bytes(modelzoo.transformer_net(seed=0, hw=512).serialize(), half=True)
    run("TransformerNet fp16 1x3x512x512 (BASELINE configs[3])", tn, (1, 3, 512, 512), 100, 80.63)
    if "tnet8" in which:
        run("TransformerNet fp16 8x3x512x512", tn, (8, 3, 512, 512), 30, 80.63)
if "resnet" in which:
    rn = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
    for b in (1, 8, 32, 128, 256):
        run(f"ResNet-50 fp16 batch {b}", rn, (b, 3, 224, 224), 100 if b <= 32 else 30, 8.178)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/model_latency.json", "w"), indent=1)
