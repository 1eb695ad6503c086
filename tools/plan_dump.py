import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from smelter_b200 import modelzoo, onnx2mps
from smelter_b200.api import Configuration, Context, Image, ONNXGraph
ctx = Context(0)
model = onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)
for share in (1, 2):
    g = ONNXGraph(model, Configuration(smShare=share), context=ctx)
    nn = g.metalGraph()
    print("share", share)
    for l in nn.planDump(32).splitlines():
        if "conv_" in l: print("  ", l[:60])
