"""Generates the golden fixtures in this directory (run here, on CPU):  python tests/golden/make_golden.py

Each .npz holds the serialized model (`model`, uint8), a seeded fp16 input (`x`) and the fp32 output of the oracle
interpreter (`y`, oracle/onnx_interp.py).  The reference itself has no golden vectors and cannot run here (SURVEY.md
§8c), so these pin the ORACLE (against accidental change) and give the GPU tests fixed vectors; they are derived, not
reference-produced — "parity unpinned" still applies.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.onnx_interp import Interpreter  # noqa: E402
from smelter_b200 import modelzoo, onnx2mps  # noqa: E402


def fixtures():
    yield "conv_bn_relu", modelzoo.conv_bn_relu(seed=0).serialize(), (1, 3, 16, 16)
    yield "conv_bn_relu_mps_half", onnx2mps.convert_bytes(modelzoo.conv_bn_relu(seed=0).serialize(), half=True), (1, 3, 16, 16)
    yield "synthetic_ops", modelzoo.synthetic_ops(seed=2).serialize(), (1, 16, 12, 12)
    yield "resnet_tiny", modelzoo.resnet50(seed=3, fold_bn=False, num_classes=10, hw=32, widths=(8, 16, 32, 64), depths=(1, 1, 1, 1)).serialize(), (2, 3, 32, 32)
    yield "mobilenet_tiny", modelzoo.mobilenet_v2(seed=4, fold_bn=True, num_classes=10, hw=32, width_div=8).serialize(), (1, 3, 32, 32)
    yield "transformer_tiny", modelzoo.transformer_net(seed=5, hw=32, width_div=8).serialize(), (1, 3, 32, 32)


def main():
    for name, model, shape in fixtures():
        rng = np.random.default_rng(abs(hash(name)) % 1000 if False else len(name))
        x = rng.random(shape, dtype=np.float32).astype(np.float16)
        y = Interpreter(model).run(torch.from_numpy(x.astype(np.float32))).numpy().astype(np.float32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, model=np.frombuffer(model, dtype=np.uint8), x=x, y=y)
        print(f"{name}: model {len(model)} B, x {x.shape}, y {y.shape}, |y|max {np.abs(y).max():.3f} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
