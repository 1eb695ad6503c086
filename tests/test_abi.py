"""The C-ABI library loads and exports every symbol include/smelter_b200.h declares (no compute calls)."""
import ctypes as C
import subprocess

import pytest

from smelter_b200 import _lib


def test_every_declared_symbol_is_exported(native_lib):
    declared = _lib.header_functions()
    assert len(declared) >= 60
    for name in declared:
        assert hasattr(native_lib, name), f"{name} declared in include/smelter_b200.h but not exported"


def test_binding_table_matches_header(native_lib):
    assert sorted(_lib.SIGNATURES) == _lib.header_functions()


def test_no_torch_or_python_linkage(native_lib):
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out and "c10" not in out


def test_abi_version_and_defaults(native_lib):
    assert native_lib.smelter_abi_version() == 2
    cfg = _lib.smelter_config()
    native_lib.smelter_config_default(C.byref(cfg))
    # Configuration.init defaults, ONNXGraph.swift:20,27-35
    assert cfg.input_constraint == 0 and cfg.bilinear_align_corners == 1 and cfg.n_dims == 0
    assert cfg.enable_fusion == 1 and cfg.use_cuda_graph == 1


def test_null_arguments_are_errors_not_crashes(native_lib):
    assert native_lib.smelter_context_create(0, None, None) == 100
    assert native_lib.smelter_graph_build(None) == 100
    assert native_lib.smelter_tensor_dims(None, None) == 100
    assert b"invalid argument" in native_lib.smelter_last_error()


def test_no_cpu_fallback_without_a_device(native_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = native_lib.smelter_context_create(0, None, C.byref(h))
    assert rc == 102  # SMELTER_ERR_CUDA
    assert b"no CPU fallback" in native_lib.smelter_last_error()


def test_environment_switches_are_documented():
    """Every SMELTER_* variable the native code reads appears in INTEGRATION.md's table (or, for the MEGA_* / build-time ones, by
    prefix), and the table names no switch the code does not read."""
    import glob
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    read = set()
    for path in glob.glob(os.path.join(root, "smelter_b200", "csrc", "**", "*.c*"), recursive=True):
        read |= set(re.findall(r'getenv\("(SMELTER_[A-Z0-9_]+)"\)', open(path).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    table = doc[doc.index("## Environment switches"):]
    named = set(re.findall(r"`(SMELTER_[A-Z0-9_]+\*?)", table))
    prefixes = [n[:-1] for n in named if n.endswith("*")]
    for var in sorted(read):
        assert var in named or any(var.startswith(p) for p in prefixes), var
    for var in sorted(n for n in named if not n.endswith("*")):
        assert var in read or var == "SMELTER_CONV_INSTRUMENT", var
