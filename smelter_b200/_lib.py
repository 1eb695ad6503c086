"""ctypes view of libsmelter_b200.so (the C ABI declared in include/smelter_b200.h).

There is no fallback: if the shared library has not been built, importing the binding raises.  Build it with
`python -m smelter_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# SMELTER_LIB_PATH: kernel experiments load a differently compiled build of the same library (tools/build_variant.sh)
LIB_PATH = os.environ.get("SMELTER_LIB_PATH") or os.path.join(HERE, "libsmelter_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "smelter_b200.h")

i32, i64, u64, f32 = C.c_int32, C.c_int64, C.c_uint64, C.c_float
vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
P = C.POINTER


class smelter_config(C.Structure):
    _fields_ = [("input_constraint", i32), ("bilinear_align_corners", i32), ("n_dims", i32), ("dims_axis", i32 * 8),
                ("dims_value", i64 * 8), ("enable_fusion", i32), ("use_cuda_graph", i32), ("defer_weights", i32), ("sm_share", i32)]


class smelter_shape(C.Structure):
    _fields_ = [("channels", i32), ("width", i32), ("height", i32), ("depth", i32)]


class smelter_conv_desc(C.Structure):
    _fields_ = [("c_out", i32), ("c_in_per_group", i32), ("k_h", i32), ("k_w", i32), ("stride_h", i32), ("stride_w", i32),
                ("dil_h", i32), ("dil_w", i32), ("groups", i32), ("pads", i32 * 4), ("weight_dtype", i32), ("weight_layout", i32),
                ("bias_dtype", i32), ("is_gemm", i32)]


class smelter_conv_problem(C.Structure):
    _fields_ = [(n, i32) for n in ("n", "h", "w", "c_in", "c_out", "k_h", "k_w", "stride_h", "stride_w", "dil_h", "dil_w", "pad_t",
                                   "pad_l", "pad_b", "pad_r", "groups", "act")] + [("clip_lo", f32), ("clip_hi", f32)] + \
               [(n, i32) for n in ("has_bias", "has_residual", "force_path")]


class smelter_ew_problem(C.Structure):
    _fields_ = [(n, i32) for n in ("op", "n", "c", "h", "w", "c2", "sub", "act", "k_h", "k_w", "stride_h", "stride_w", "pad_h", "pad_w",
                                   "pad_b", "pad_r", "scale_h", "scale_w", "align_corners")] + [("alpha", f32), ("beta", f32)]


CONVERTER_FN = C.CFUNCTYPE(i32, vp, i32, vp)

# name -> (restype, argtypes).  Kept in step with include/smelter_b200.h; tests/test_abi.py checks both directions.
SIGNATURES = {
    "smelter_last_error": (cp, []),
    "smelter_abi_version": (i32, []),
    "smelter_config_default": (None, [P(smelter_config)]),
    "smelter_context_create": (i32, [i32, vp, P(vp)]),
    "smelter_context_destroy": (i32, [vp]),
    "smelter_context_stream": (i32, [vp, P(vp)]),
    "smelter_context_synchronize": (i32, [vp]),
    "smelter_nccl_unique_id": (i32, [P(C.c_uint8)]),
    "smelter_context_init_nccl": (i32, [vp, P(C.c_uint8), i32, i32]),
    "smelter_tensor_create": (i32, [vp, i32, i32, i32, i32, P(vp)]),
    "smelter_tensor_wrap": (i32, [vp, vp, i32, i32, i32, i32, P(vp)]),
    "smelter_tensor_destroy": (i32, [vp]),
    "smelter_tensor_dims": (i32, [vp, P(i32)]),
    "smelter_tensor_device_ptr": (i32, [vp, P(vp)]),
    "smelter_tensor_from_float": (i32, [vp, vp, vp, sz]),
    "smelter_tensor_from_half": (i32, [vp, vp, vp, sz]),
    "smelter_tensor_from_u8": (i32, [vp, vp, vp, i32, vp, vp]),
    "smelter_tensor_to_float": (i32, [vp, vp, vp, sz]),
    "smelter_tensor_to_float_async": (i32, [vp, vp, vp, sz]),
    "smelter_tensor_to_float_mps": (i32, [vp, vp, vp, sz]),
    "smelter_tensor_to_half": (i32, [vp, vp, vp, sz]),
    "smelter_graph_create": (i32, [vp, vp, sz, P(smelter_config), P(vp)]),
    "smelter_graph_build": (i32, [vp]),
    "smelter_graph_destroy": (i32, [vp]),
    "smelter_graph_format": (i32, [vp, P(i32)]),
    "smelter_graph_num_outputs": (i32, [vp, P(i32)]),
    "smelter_graph_output_shape": (i32, [vp, i32, P(smelter_shape)]),
    "smelter_graph_num_nodes": (i32, [vp, P(i32)]),
    "smelter_graph_node_op_type": (i32, [vp, i32, P(cp)]),
    "smelter_graph_has_converter": (i32, [vp, cp, P(i32)]),
    "smelter_graph_num_launches": (i32, [vp, i32, P(i32)]),
    "smelter_graph_plan_dump": (i32, [vp, i32, C.c_char_p, sz]),
    "smelter_graph_profile": (i32, [vp, vp, P(vp), i32, i32, P(f32), P(C.c_double), P(C.c_double), P(i32), i32, P(i32)]),
    "smelter_graph_encode": (i32, [vp, vp, P(vp), i32, P(vp)]),
    "smelter_graph_broadcast_weights": (i32, [vp, i32]),
    "smelter_graph_weight_checksum": (i32, [vp, P(u64), P(u64)]),
    "smelter_graph_weight_arena": (i32, [vp, P(vp), P(u64)]),
    "smelter_graph_has_output": (i32, [vp, cp, P(i32)]),
    "smelter_graph_shape": (i32, [vp, cp, P(smelter_shape)]),
    "smelter_graph_has_tensor": (i32, [vp, cp, P(i32)]),
    "smelter_add_conv": (i32, [vp, cp, P(smelter_conv_desc), vp, vp, cp]),
    "smelter_add_batchnorm": (i32, [vp, cp, i32, vp, vp, vp, vp, f32, cp]),
    "smelter_add_instancenorm": (i32, [vp, cp, i32, vp, vp, f32, cp]),
    "smelter_add_unary": (i32, [vp, cp, i32, f32, f32, cp]),
    "smelter_add_binary": (i32, [vp, cp, cp, i32, cp]),
    "smelter_add_pool": (i32, [vp, cp, i32, i32, i32, i32, i32, i32, i32, cp]),
    "smelter_add_global_avgpool": (i32, [vp, cp, cp]),
    "smelter_add_upsample": (i32, [vp, cp, i32, i32, i32, i32, cp]),
    "smelter_add_concat": (i32, [vp, P(cp), i32, cp]),
    "smelter_add_reshape": (i32, [vp, cp, i32, i32, i32, cp]),
    "smelter_add_softmax": (i32, [vp, cp, i32, cp]),
    "smelter_add_pad": (i32, [vp, cp, i32, P(i32), f32, cp]),
    "smelter_add_alias": (i32, [vp, cp, cp]),
    "smelter_graph_register_converter": (i32, [vp, cp, CONVERTER_FN, vp]),
    "smelter_node_num_inputs": (i32, [vp, i32, P(i32)]),
    "smelter_node_input": (i32, [vp, i32, i32, P(cp)]),
    "smelter_node_num_outputs": (i32, [vp, i32, P(i32)]),
    "smelter_node_output": (i32, [vp, i32, i32, P(cp)]),
    "smelter_node_attr_int": (i32, [vp, i32, cp, P(i64), P(i32)]),
    "smelter_node_attr_float": (i32, [vp, i32, cp, P(f32), P(i32)]),
    "smelter_node_attr_ints": (i32, [vp, i32, cp, P(i64), i32, P(i32)]),
    "smelter_reformat_conv_weight": (i32, [vp, vp, i32, i32, i32, i32, i32, i32]),
    "smelter_float16_to_32": (i32, [vp, vp, sz]),
    "smelter_float32_to_16": (i32, [vp, vp, sz]),
    "smelter_conv_output_size": (i32, [i32, i32, i32, i32, i32, i32, i32, i32, P(i32)]),
    "smelter_pool_output_size": (i32, [i32, i32, i32, i32, P(i32)]),
    "smelter_tensorproto_integers": (i32, [vp, sz, P(i64), sz, P(sz)]),
    "smelter_tensorproto_floats": (i32, [vp, sz, P(f32), sz, P(sz)]),
    "smelter_run_conv": (i32, [vp, P(smelter_conv_problem), vp, vp, vp, vp, vp, i32, P(f32)]),
    "smelter_run_elementwise": (i32, [vp, P(smelter_ew_problem), vp, vp, vp, vp, vp, i32, P(f32)]),
    "smelter_l2_flush": (i32, [vp]),
    "smelter_tma_probe": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, P(f32)]),
}

_lib = None


def header_functions() -> list:
    """Names of every function include/smelter_b200.h declares."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smelter_[a-z0-9_]+)\s*\(", text)) - {"smelter_converter_fn"})


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -m smelter_b200.build`; this package has no CPU or "
                           "Python fallback for the engine")
    handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError = ABI drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib
