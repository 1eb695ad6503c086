"""Per-layer timeline of the persistent conv kernel on ResNet-50 batch 32 (instrumented build, no CUDA graph).
Usage: SMELTER_CONV_INSTRUMENT=1 python -m smelter_b200.build && SMELTER_CONV_INSTRUMENT=1 SMELTER_MEGA_TIMELINE=1 python tools/mega_timeline.py"""
import os, sys
os.environ.setdefault("SMELTER_MEGA", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

ctx = Context(0)
model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
g = ONNXGraph(model, Configuration(useCudaGraph=False), context=ctx)
nn = g.metalGraph()
x = np.random.default_rng(0).random((32, 3, 224, 224), dtype=np.float32).astype(np.float16)
img = Image.fromArray(ctx, x)
for i in range(3):
    print(f"--- encode {i}", file=sys.stderr, flush=True)
    nn.encode(sourceImages=[img]).toFloatArray()
