import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200.api import Configuration, Context, Image, ONNXGraph
from bench import model_bytes
ctx = Context(0)
graph = bool(int(os.environ.get("USE_GRAPH", "0")))
g = ONNXGraph(model_bytes(), Configuration(useCudaGraph=graph), context=ctx); nn = g.metalGraph()
rng = np.random.default_rng(0)
for b in [int(v) for v in sys.argv[1:]]:
    x = Image.fromArray(ctx, rng.random((b, 3, 224, 224), dtype=np.float32).astype(np.float16))
    for i in range(3):
        out = nn.encode(sourceImages=[x]).toFloatArray()
    print("batch", b, "ok", float(np.abs(out).max()), flush=True)
