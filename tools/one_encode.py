"""A few ResNet-50 encodes and nothing else: the command the ncu captures of tools/gpu_round.sh profile.
Usage: python tools/one_encode.py [batch=32] [smShare=1] [encodes=5] [useCudaGraph=1] [model=resnet|tnet|mobilenet]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
share = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
use_graph = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
ctx = Context(0)
which = sys.argv[5] if len(sys.argv) > 5 else "resnet"
shape = (B, 3, 512, 512) if which == "tnet" else (B, 3, 224, 224)
net = modelzoo.transformer_net(seed=0, hw=512) if which == "tnet" else modelzoo.mobilenet_v2(seed=0, fold_bn=False) if which == "mobilenet" else modelzoo.resnet50(seed=0, fold_bn=False)
model = onnx2mps.convert_bytes(net.serialize(), half=True)
g = ONNXGraph(model, Configuration(smShare=share, useCudaGraph=use_graph), context=ctx)
nn = g.metalGraph()
img = Image.fromArray(ctx, np.random.default_rng(0).random(shape, dtype=np.float32).astype(np.float16))
for _ in range(n):
    out = nn.encode(sourceImages=[img])
ctx.synchronize()
print("launches per encode:", nn.numLaunches(B), "out[0,0] =", float(out.toFloatArray().reshape(B, -1)[0, 0]))
