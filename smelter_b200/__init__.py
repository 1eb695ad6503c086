"""smelter_b200 — B200-native drop-in for Smelter's ONNX-graph inference path.

The engine is libsmelter_b200.so (C ABI in include/smelter_b200.h: C++ host logic + hand-written sm_100a
kernels).  This package holds the build script, the ctypes binding that mirrors the reference's public names
(api.py), and the Python host tools on the path (ONNX2MPS restatement, model generators).
"""
__all__ = ["api", "onnx_proto", "onnx2mps", "modelzoo", "shape_infer", "build"]
