"""Per-step in-situ times (event pairs around every launch, graph off) of the BASELINE configs[1] / configs[3] models.
Usage: python tools/step_profile.py [mobilenet|tnet|resnet] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

which = sys.argv[1] if len(sys.argv) > 1 else "tnet"
ctx = Context(0)
if which == "mobilenet":
    model, shape = modelzoo.mobilenet_v2(seed=0, fold_bn=False), (int(sys.argv[2]) if len(sys.argv) > 2 else 1, 3, 224, 224)
elif which == "resnet":
    model, shape = modelzoo.resnet50(seed=0, fold_bn=False), (int(sys.argv[2]) if len(sys.argv) > 2 else 32, 3, 224, 224)
else:
    model, shape = modelzoo.transformer_net(seed=0, hw=512), (1, 3, 512, 512)
g = ONNXGraph(onnx2mps.convert_bytes(model.serialize(), half=True), Configuration(), context=ctx)
nn = g.metalGraph()
img = Image.fromArray(ctx, np.random.default_rng(1).random(shape, dtype=np.float32).astype(np.float16))
for _ in range(3):
    nn.encode(sourceImages=[img])
ctx.synchronize()
prof = nn.profile([img], iters=10)
tot = sum(p["ms"] for p in prof)
print(f"{which} {shape}: {len(prof)} steps, sum of event-bracketed steps {tot * 1e3:.1f} us (each carries a few us of event overhead)")
for p in prof:
    print(f"{p['ms'] * 1e3:8.2f} us  {p['bytes'] / 1e6:8.2f} MB  {p['flops'] / 1e9:8.3f} GF  {p['desc']}")
