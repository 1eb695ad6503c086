"""Max-abs error of every golden fixture on the GPU (quick regression probe)."""
import glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from smelter_b200.api import Configuration, Context, Image, ONNXGraph
ctx = Context(0)
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "*.npz"))):
    z = np.load(path)
    for graph in (False, True):
        g = ONNXGraph(z["model"].tobytes(), Configuration(useCudaGraph=graph), context=ctx)
        out = g.metalGraph().encode(sourceImages=[Image.fromArray(ctx, z["x"])]).toFloatArray().reshape(z["y"].shape)
        print(f"{os.path.basename(path):28s} graph={int(graph)} max_abs_err={np.abs(out - z['y']).max():.3e}", flush=True)
        g.close()
