"""Launch-chain timeline of one ResNet-50 encode (instrumented build): per conv launch, min / max over CTAs of entry, griddepcontrol.wait
returned, first operands landed, last accumulator ready, last store issued, exit.  Where the per-launch fixed cost sits.
Usage (GPU box): SMELTER_CONV_INSTRUMENT=1 python -m smelter_b200.build --force && SMELTER_CHAIN_TIMELINE=8 python tools/chain_timeline.py [batch]
(rebuild without the switch afterwards: the instrumented library is not the product)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = Context(0)
model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
g = ONNXGraph(model, Configuration(), context=ctx)
nn = g.metalGraph()
img = Image.fromArray(ctx, np.random.default_rng(0).random((B, 3, 224, 224), dtype=np.float32).astype(np.float16))
for _ in range(10):
    nn.encode(sourceImages=[img])
ctx.synchronize()
print(nn.planDump(B))
