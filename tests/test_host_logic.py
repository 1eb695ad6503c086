"""Host-side index maps / coercions of the engine (through the C ABI, no GPU) against the C oracle and numpy.

Mirrors the reference host loops: Array+Extensions.swift:52-93, Float16.swift, ONNXConvolutionPadding.swift:91-113,
PyTorchPoolPadding.swift:94-103, Onnx_TensorProto+Extensions.swift:2-62."""
import ctypes as C

import numpy as np
import pytest

from smelter_b200 import onnx_proto as op

vp = C.c_void_p


def ptr(a):
    return a.ctypes.data_as(vp)


@pytest.mark.parametrize("shape", [(8, 3, 3, 3), (4, 5, 1, 1), (7, 2, 3, 5), (16, 16, 7, 7), (1, 1, 1, 1), (3, 9, 9, 2)])
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_reformat_conv_weight_matches_oracle_and_numpy(native_lib, host_oracle, shape, dtype):
    o, i, kh, kw = shape
    rng = np.random.default_rng(sum(shape))
    w = rng.standard_normal(shape).astype(dtype)
    got = np.empty(w.size, dtype=dtype)
    want = np.empty(w.size, dtype=dtype)
    assert native_lib.smelter_reformat_conv_weight(ptr(w), ptr(got), w.itemsize, o, i, kh, kw, 0) == 0
    host_oracle.oracle_reformat_conv_weight(ptr(w), ptr(want), w.itemsize, o, i, kh, kw, 0)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))                      # bit exact vs the C restatement
    assert np.array_equal(got.reshape(o, kh, kw, i), w.transpose(0, 2, 3, 1))           # == ONNX2MPS.py:75 [0,2,3,1]


@pytest.mark.parametrize("shape", [(3, 8, 3, 3), (5, 4, 2, 4), (16, 2, 1, 1)])
def test_reformat_conv_transpose_weight(native_lib, host_oracle, shape):
    # ConvTranspose weights are [Cin, Cout, kH, kW]; result OHWI with a 180 degree flip (Array+Extensions.swift:70-76,
    # ONNX2MPS.py:58-62)
    i, o, kh, kw = shape
    w = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
    got = np.empty(w.size, np.float32)
    want = np.empty(w.size, np.float32)
    assert native_lib.smelter_reformat_conv_weight(ptr(w), ptr(got), 4, o, i, kh, kw, 1) == 0
    host_oracle.oracle_reformat_conv_weight(ptr(w), ptr(want), 4, o, i, kh, kw, 1)
    assert np.array_equal(got, want)
    assert np.array_equal(got.reshape(o, kh, kw, i), w.transpose(1, 2, 3, 0)[:, ::-1, ::-1, :])
    # the oracle's ONNX2MPS swizzle agrees with numpy too
    sw = np.empty(w.size, np.float32)
    dims = (C.c_int * 4)(*shape)
    perm = (C.c_int * 4)(1, 2, 3, 0)
    host_oracle.oracle_onnx2mps_swizzle(ptr(w), ptr(sw), 4, dims, perm, 1)
    assert np.array_equal(sw, got)


def test_reformat_rejects_bad_arguments(native_lib):
    a = np.zeros(4, np.float32)
    assert native_lib.smelter_reformat_conv_weight(ptr(a), ptr(a), 4, 1, 1, 2, 2, 0) == 100  # in-place
    b = np.zeros(4, np.float32)
    assert native_lib.smelter_reformat_conv_weight(ptr(a), ptr(b), 3, 1, 1, 2, 2, 0) == 100  # element size


def test_float16_to_32_all_bit_patterns(native_lib, host_oracle):
    h = np.arange(65536, dtype=np.uint16)
    got = np.empty(65536, np.float32)
    want = np.empty(65536, np.float32)
    assert native_lib.smelter_float16_to_32(ptr(h), ptr(got), h.size) == 0
    host_oracle.oracle_float16_to_32(ptr(h), ptr(want), C.c_size_t(h.size))
    ref = h.view(np.float16).astype(np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    finite = np.isfinite(ref)
    assert np.array_equal(got[finite].view(np.uint32), ref[finite].view(np.uint32))
    assert np.isnan(got[~finite]).sum() == np.isnan(ref[~finite]).sum()


def test_float32_to_16_rounding(native_lib, host_oracle):
    rng = np.random.default_rng(0)
    # every half value, the midpoints between neighbours (ties), values just off the ties, range edges, subnormals
    halves = np.arange(0x7C00, dtype=np.uint16).view(np.float16).astype(np.float64)
    mids = (halves[:-1] + halves[1:]) / 2
    f = np.concatenate([halves, mids, np.nextafter(mids.astype(np.float32), np.float32(np.inf)).astype(np.float64),
                        np.nextafter(mids.astype(np.float32), np.float32(-np.inf)).astype(np.float64),
                        [65504.0, 65519.99, 65520.0, 65536.0, 1e9, 2.0 ** -24, 2.0 ** -25, 2.0 ** -26, 5.9e-8, 0.0, np.inf],
                        rng.standard_normal(20000) * 10.0 ** rng.uniform(-9, 6, 20000)]).astype(np.float32)
    f = np.concatenate([f, -f])
    got = np.empty(f.size, np.uint16)
    want = np.empty(f.size, np.uint16)
    assert native_lib.smelter_float32_to_16(ptr(f), ptr(got), f.size) == 0
    host_oracle.oracle_float32_to_16(ptr(f), ptr(want), C.c_size_t(f.size))
    with np.errstate(over="ignore"):
        ref = f.astype(np.float16).view(np.uint16)  # numpy astype: round-to-nearest-even (ONNX2MPS.py:27)
    assert np.array_equal(got, ref)
    assert np.array_equal(want, ref)
    nan = np.array([np.nan], np.float32)
    out = np.empty(1, np.uint16)
    native_lib.smelter_float32_to_16(ptr(nan), ptr(out), 1)
    assert np.isnan(out.view(np.float16)[0])


def test_conv_output_size_matches_torch_and_oracle(native_lib, host_oracle):
    import torch
    import torch.nn.functional as F

    out = C.c_int32()
    for i in (7, 14, 56, 57, 224):
        for k in (1, 3, 7, 9):
            for s in (1, 2, 3):
                for d in (1, 2):
                    for p in (0, 1, 3):
                        if i + 2 * p < d * (k - 1) + 1:
                            continue
                        assert native_lib.smelter_conv_output_size(i, k, s, d, p, p, 0, 0, C.byref(out)) == 0
                        t = F.conv2d(torch.zeros(1, 1, i, i), torch.zeros(1, 1, k, k), stride=s, padding=p, dilation=d).shape[-1]
                        assert out.value == t == host_oracle.oracle_conv_padded_size_dilated(i, k, s, d, p, p)
                        if d == 1:  # the reference formula has no dilation term; identical when dilation is 1
                            assert out.value == host_oracle.oracle_conv_padded_size(i, k, s, p, p, 0, 0)
    # transpose branch (ONNXConvolutionPadding.swift:97-103)
    assert native_lib.smelter_conv_output_size(14, 3, 2, 1, 1, 1, 1, 1, C.byref(out)) == 0
    assert out.value == host_oracle.oracle_conv_padded_size(14, 3, 2, 1, 1, 1, 1) == \
        F.conv_transpose2d(torch.zeros(1, 1, 14, 14), torch.zeros(1, 1, 3, 3), stride=2, padding=1, output_padding=1).shape[-1]


def test_pool_output_size_matches_torch_and_oracle(native_lib, host_oracle):
    import torch
    import torch.nn.functional as F

    out = C.c_int32()
    for i in (7, 14, 55, 56, 112, 113):
        for k, s, p in ((3, 2, 1), (2, 2, 0), (3, 1, 1), (7, 7, 0), (3, 2, 0)):
            assert native_lib.smelter_pool_output_size(i, k, s, p, C.byref(out)) == 0
            t = F.max_pool2d(torch.zeros(1, 1, i, i), k, s, p).shape[-1]
            assert out.value == t == host_oracle.oracle_pool_padded_size(i, k, s, p)


def _coerce(native_lib, t: op.Tensor, kind: str):
    data = t.serialize()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    n = C.c_size_t()
    if kind == "integers":
        out = (C.c_int64 * 64)()
        rc = native_lib.smelter_tensorproto_integers(buf, len(data), out, 64, C.byref(n))
    else:
        out = (C.c_float * 64)()
        rc = native_lib.smelter_tensorproto_floats(buf, len(data), out, 64, C.byref(n))
    return rc, list(out[: n.value])


def test_tensorproto_coercion_matrix(native_lib):
    """Every DataType x storage-field rule of Onnx_TensorProto+Extensions.swift:2-62."""
    I, Fl = "integers", "floats"
    # int32_data-backed types
    for dt in (op.INT32, op.INT16, op.INT8, op.UINT16, op.UINT8, op.BOOL):
        t = op.Tensor(dims=[3], data_type=dt, int32_data=[1, 0, 7])
        assert _coerce(native_lib, t, I) == (0, [1, 0, 7])
        assert _coerce(native_lib, t, Fl) == (0, [1.0, 0.0, 7.0])
    t = op.Tensor(dims=[2], data_type=op.INT32, int32_data=[-5, 2 ** 31 - 1])
    assert _coerce(native_lib, t, I) == (0, [-5, 2 ** 31 - 1])
    # int64: typed field or raw_data for integers (:11-17); floats read the typed field only (:46-47)
    t = op.Tensor(dims=[3], data_type=op.INT64, int64_data=[-1, 2048, 1 << 40])
    assert _coerce(native_lib, t, I) == (0, [-1, 2048, 1 << 40])
    assert _coerce(native_lib, t, Fl)[1][:2] == [-1.0, 2048.0]
    t = op.Tensor.from_numpy("", np.asarray([1, -1, 10], np.int64))
    assert _coerce(native_lib, t, I) == (0, [1, -1, 10])
    assert _coerce(native_lib, t, Fl) == (0, [])  # raw int64 is invisible to `.floats`, as in the reference
    # uint32/uint64 -> uint64_data
    t = op.Tensor(dims=[2], data_type=op.UINT64, uint64_data=[3, 99])
    assert _coerce(native_lib, t, I) == (0, [3, 99]) and _coerce(native_lib, t, Fl) == (0, [3.0, 99.0])
    # float: float_data or raw_data; Int(Float) truncates toward zero (:20-25)
    t = op.Tensor(dims=[4], data_type=op.FLOAT, float_data=[1.0, 1.0, 2.0, 2.9])
    assert _coerce(native_lib, t, I) == (0, [1, 1, 2, 2])
    t = op.Tensor.from_numpy("", np.asarray([-1.7, 2.5], np.float32))
    assert _coerce(native_lib, t, I) == (0, [-1, 2]) and _coerce(native_lib, t, Fl) == (0, [np.float32(-1.7), 2.5])
    # double -> double_data
    t = op.Tensor(dims=[1], data_type=op.DOUBLE, double_data=[3.75])
    assert _coerce(native_lib, t, I) == (0, [3]) and _coerce(native_lib, t, Fl) == (0, [3.75])
    # float16 -> raw_data only (:28-30, :56-57); this is how ONNX2MPS'ed shape tensors are read back
    t = op.Tensor.from_numpy("", np.asarray([1, 1, 2, 2, -1, 2048], np.float16))
    assert _coerce(native_lib, t, I) == (0, [1, 1, 2, 2, -1, 2048])
    assert _coerce(native_lib, t, Fl) == (0, [1.0, 1.0, 2.0, 2.0, -1.0, 2048.0])
    # anything else: the reference fatalErrors; the ABI returns UNSUPPORTED
    t = op.Tensor(dims=[1], data_type=op.STRING)
    assert _coerce(native_lib, t, I)[0] == 104 and _coerce(native_lib, t, Fl)[0] == 104


def test_malformed_tensorproto_is_a_parse_error(native_lib):
    bad = bytes([0x0A, 0x7F, 0x01])  # length-delimited field claiming 127 bytes
    buf = (C.c_uint8 * len(bad)).from_buffer_copy(bad)
    n = C.c_size_t()
    out = (C.c_int64 * 4)()
    assert native_lib.smelter_tensorproto_integers(buf, len(bad), out, 4, C.byref(n)) == 101
