"""Per-step event-bracketed times of one ResNet-50 batch-32 encode planned for a share of the chip (nothing else in flight).
Usage: python tools/share_profile.py [smShare=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

share = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = Context(0)
model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
g = ONNXGraph(model, Configuration(smShare=share), context=ctx)
nn = g.metalGraph()
img = Image.fromArray(ctx, np.random.default_rng(1).random((32, 3, 224, 224), dtype=np.float32).astype(np.float16))
for _ in range(3):
    nn.encode(sourceImages=[img])
ctx.synchronize()
prof = nn.profile([img], iters=10)
print(f"smShare {share}: sum of event-bracketed steps {sum(p['ms'] for p in prof) * 1e3:.1f} us")
for p in prof:
    fl = p["flops"]
    print(f"{p['ms'] * 1e3:8.2f} us  {fl / max(p['ms'], 1e-9) / 1e9:7.1f} TF/s  {p['desc']}")
