"""Python host binding over the C ABI, mirroring Smelter's public surface name for name.

  reference (Swift)                                          here
  ---------------------------------------------------------  -------------------------------------------------
  ONNXGraph(data:configuration:)      ONNXGraph.swift:95     ONNXGraph(data, configuration=Configuration())
  ONNXGraph(contentsOf:configuration:)            :158-167   ONNXGraph.contentsOf(path, configuration)
  ONNXGraph.Configuration                         :6-36      Configuration(inputConstraint, billinearUpsamplingConfiguration, dims)
  ONNXGraph.Errors                                :38-47     Errors (exception; .case = the enum case name)
  ONNXGraph.Format                                :49-52     Format.onnx / Format.mpsFlavor
  ONNXGraph.modelFormat / outputShapes            :58,69-91  same names
  ONNXGraph.metalGraph(device:) -> MPSNNGraph     :169-193   ONNXGraph.metalGraph(device) -> NNGraph
  MPSNNGraph.encode(to:sourceImages:)   README.md:43-44      NNGraph.encode(to=stream, sourceImages=[Image]) -> Image
  MPSImage.toFloatArray()   MPSImage+Extensions.swift:9-59   Image.toFloatArray()  (NCHW order)
  MTLContext / MTLDevice                README.md:18         Context(device_index)
  Shape                       TypeDefinitions.swift:1-33     Shape(channels, width, height, depth)

This file is a test/bench convenience over include/smelter_b200.h; the engine's host logic (parse, registry,
converters, fusion, planning) is the C++ behind the ABI.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L

_ERROR_CASES = {1: "unsupportedInput", 2: "unsupportedOutput", 3: "unknownNodeOpType", 4: "noSuchOutput", 5: "graphInternalError",
                6: "insufficientInputs", 7: "inconsistentState", 8: "notEnoughAttributes", 100: "invalidArgument", 101: "parse",
                102: "cuda", 103: "nccl", 104: "unsupported"}


class Errors(Exception):
    """ONNXGraph.Errors (ONNXGraph.swift:38-47) plus the engine's own codes (>= 100)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{_ERROR_CASES.get(code, code)}: {message}")
        self.code = code
        self.case = _ERROR_CASES.get(code, str(code))
        self.message = message
        self.opType = message if code == 3 else None  # unknownNodeOpType(opType:)


def _check(rc: int) -> None:
    if rc != 0:
        raise Errors(rc, (L.lib().smelter_last_error() or b"").decode("utf-8", "replace"))


class Format:
    onnx = 0
    mpsFlavor = 1


@dataclass
class BillinearUpsampling:  # (sic) ONNXGraph.swift:12-24
    alignCorners: bool = True


@dataclass
class Configuration:
    """ONNXGraph.Configuration (ONNXGraph.swift:6-36).  inputConstraint: "none" | "lanczos" | "bilinear"."""
    inputConstraint: str = "none"
    billinearUpsamplingConfiguration: BillinearUpsampling = field(default_factory=BillinearUpsampling)
    dims: Dict[int, int] = field(default_factory=dict)
    # engine options (no reference counterpart)
    enableFusion: bool = True
    useCudaGraph: bool = True
    deferWeights: bool = False
    smShare: int = 1  # k >= 2: kernels sized for 1/k of the SMs so that encodes in flight on different streams co-run

    def _c(self) -> L.smelter_config:
        c = L.smelter_config()
        L.lib().smelter_config_default(C.byref(c))
        c.input_constraint = {"none": 0, "lanczos": 1, "bilinear": 2}[self.inputConstraint]
        c.bilinear_align_corners = int(self.billinearUpsamplingConfiguration.alignCorners)
        if len(self.dims) > 8:
            raise ValueError("at most 8 dims overrides")
        c.n_dims = len(self.dims)
        for i, (axis, value) in enumerate(sorted(self.dims.items())):
            c.dims_axis[i] = axis
            c.dims_value[i] = value
        c.enable_fusion = int(self.enableFusion)
        c.use_cuda_graph = int(self.useCudaGraph)
        c.defer_weights = int(self.deferWeights)
        c.sm_share = int(self.smShare)
        return c


@dataclass
class Shape:
    channels: int
    width: int
    height: int
    depth: int


class Context:
    """Device + stream (what MTLContext / MTLDevice were for the reference app, README.md:18)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        _check(L.lib().smelter_context_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        _check(L.lib().smelter_context_stream(self._h, C.byref(s)))
        return s.value or 0

    def synchronize(self) -> None:
        _check(L.lib().smelter_context_synchronize(self._h))

    def initNCCL(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(L.lib().smelter_context_init_nccl(self._h, buf, rank, world))

    @staticmethod
    def ncclUniqueId() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(L.lib().smelter_nccl_unique_id(buf))
        return bytes(buf)

    def l2Flush(self) -> None:
        _check(L.lib().smelter_l2_flush(self._h))

    def close(self) -> None:
        if self._h:
            L.lib().smelter_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Image:
    """NCHW fp16 device buffer — the MPSImage stand-in."""

    def __init__(self, context: Context, n: int, c: int, h: int, w: int, _handle=None, _owned=True):
        self.context = context
        self._owned = _owned
        if _handle is None:
            self._h = C.c_void_p()
            _check(L.lib().smelter_tensor_create(context._h, n, c, h, w, C.byref(self._h)))
        else:
            self._h = _handle
        self.shape = (n, c, h, w)

    @staticmethod
    def fromArray(context: Context, a: np.ndarray, stream: Optional[int] = None) -> "Image":
        """README.md:33-39 `texture(from:)` analogue: host array [N,C,H,W] (fp32 or fp16) -> device fp16."""
        if a.ndim == 3:
            a = a[None]
        img = Image(context, *a.shape)
        img.copyFrom(a, stream)
        return img

    @staticmethod
    def fromBytes(context: Context, pixels: np.ndarray, channels: int = 3, scale: Optional[Sequence[float]] = None,
                  bias: Optional[Sequence[float]] = None, stream: Optional[int] = None) -> "Image":
        """`MTLContext.texture(from:)` analogue (README.md:33-39): interleaved uint8 pixels [N,H,W,S] or [H,W,S] (RGB, RGBA ...)
        -> device fp16 [N,channels,H,W] with value = byte * scale[c] + bias[c] (default 1/255, 0)."""
        a = np.ascontiguousarray(pixels, dtype=np.uint8)
        if a.ndim == 3:
            a = a[None]
        n, h, w, s = a.shape
        img = Image(context, n, channels, h, w)
        sc = (C.c_float * channels)(*scale) if scale is not None else None
        bi = (C.c_float * channels)(*bias) if bias is not None else None
        _check(L.lib().smelter_tensor_from_u8(img._h, C.c_void_p(stream) if stream else None, a.ctypes.data_as(C.c_void_p), s, sc, bi))
        context.synchronize()  # `a` may be a temporary
        return img

    @staticmethod
    def wrap(context: Context, device_ptr: int, n: int, c: int, h: int, w: int) -> "Image":
        h_ = C.c_void_p()
        _check(L.lib().smelter_tensor_wrap(context._h, C.c_void_p(device_ptr), n, c, h, w, C.byref(h_)))
        return Image(context, n, c, h, w, _handle=h_, _owned=True)

    def copyFrom(self, a: np.ndarray, stream: Optional[int] = None) -> None:
        a = np.ascontiguousarray(a)
        st = C.c_void_p(stream) if stream else None
        if a.dtype == np.float16:
            _check(L.lib().smelter_tensor_from_half(self._h, st, a.ctypes.data_as(C.c_void_p), a.size))
        else:
            a = a.astype(np.float32, copy=False)
            _check(L.lib().smelter_tensor_from_float(self._h, st, a.ctypes.data_as(C.c_void_p), a.size))
        self._keepalive = a

    def copyFromPointer(self, host_ptr: int, count: int, stream: Optional[int] = None, half: bool = True) -> None:
        """Asynchronous host->device copy from a raw (ideally pinned) host pointer holding `count` fp16 (or fp32) values."""
        fn = L.lib().smelter_tensor_from_half if half else L.lib().smelter_tensor_from_float
        _check(fn(self._h, C.c_void_p(stream) if stream else None, C.c_void_p(host_ptr), count))

    @property
    def devicePointer(self) -> int:
        p = C.c_void_p()
        _check(L.lib().smelter_tensor_device_ptr(self._h, C.byref(p)))
        return p.value or 0

    def toFloatArray(self, stream: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """MPSImage.toFloatArray(): device fp16 -> host fp32, NCHW order; synchronises the stream."""
        if out is None:
            out = np.empty(self.shape, dtype=np.float32)
        _check(L.lib().smelter_tensor_to_float(self._h, C.c_void_p(stream) if stream else None, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def toFloatArrayMPS(self, stream: Optional[int] = None) -> np.ndarray:
        """toFloatArray() in the reference's element order (MPSImage+Extensions.swift:26-59): flat array of N x slices x H x W x 4
        (channels in groups of four, zero padded; C < 3: N x H x W x C)."""
        n, c, h, w = self.shape
        cpp = c if c < 3 else 4 * ((c + 3) // 4)
        out = np.empty(n * h * w * cpp, dtype=np.float32)
        _check(L.lib().smelter_tensor_to_float_mps(self._h, C.c_void_p(stream) if stream else None, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def toFloatArrayAsync(self, out: np.ndarray, stream: Optional[int] = None) -> np.ndarray:
        """Enqueue device fp16 -> host fp32 (NCHW) on the stream without waiting: `out` (ideally pinned) holds the values once the
        stream has passed this point — the analogue of reading an MPSImage in a command buffer's completion handler."""
        _check(L.lib().smelter_tensor_to_float_async(self._h, C.c_void_p(stream) if stream else None, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def toHalfArray(self, stream: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=np.float16)
        _check(L.lib().smelter_tensor_to_half(self._h, C.c_void_p(stream) if stream else None, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def close(self) -> None:
        if self._h and self._owned:
            L.lib().smelter_tensor_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NNGraph:
    """The compiled graph — MPSNNGraph's role (ONNXGraph.swift:185-190)."""

    def __init__(self, owner: "ONNXGraph"):
        self._owner = owner

    def encode(self, to: Optional[int] = None, sourceImages: Sequence[Image] = ()) -> Image:
        """MPSNNGraph.encode(to:sourceImages:) — enqueue only; the result is valid once the stream is synchronised
        (`Context.synchronize()` / `Image.toFloatArray()`), and until the next encode on this graph."""
        arr = (C.c_void_p * len(sourceImages))(*[img._h for img in sourceImages])
        res = C.c_void_p()
        _check(L.lib().smelter_graph_encode(self._owner._h, C.c_void_p(to) if to else None, arr, len(sourceImages), C.byref(res)))
        dims = (C.c_int32 * 4)()
        _check(L.lib().smelter_tensor_dims(res, dims))
        return Image(self._owner.context, *dims, _handle=res, _owned=False)

    def profile(self, sourceImages: Sequence[Image], iters: int = 3, stream: Optional[int] = None):
        """Per-kernel CUDA-event timing of one encode (no CUDA graph).  Returns a list of dicts, one per launched step:
        {desc, ms, flops, bytes, tensor}."""
        lib = L.lib()
        arr = (C.c_void_p * len(sourceImages))(*[img._h for img in sourceImages])
        n = C.c_int32()
        st = C.c_void_p(stream) if stream else None
        cap = 4096
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        by = (C.c_double * cap)()
        tc = (C.c_int32 * cap)()
        _check(lib.smelter_graph_profile(self._owner._h, st, arr, len(sourceImages), iters, ms, fl, by, tc, cap, C.byref(n)))
        descs = [l.split(" flops=")[0].rstrip() for l in self.planDump(sourceImages[0].shape[0]).splitlines()[1:]]
        return [{"desc": descs[i] if i < len(descs) else "", "ms": ms[i], "flops": fl[i], "bytes": by[i], "tensor": bool(tc[i])}
                for i in range(n.value)]

    def numLaunches(self, batch: int) -> int:
        n = C.c_int32()
        _check(L.lib().smelter_graph_num_launches(self._owner._h, batch, C.byref(n)))
        return n.value

    def planDump(self, batch: int) -> str:
        buf = C.create_string_buffer(1 << 18)
        _check(L.lib().smelter_graph_plan_dump(self._owner._h, batch, buf, len(buf)))
        return buf.value.decode()

    def broadcastWeights(self, root: int = 0) -> None:
        _check(L.lib().smelter_graph_broadcast_weights(self._owner._h, root))

    def weightChecksum(self):
        s, b = C.c_uint64(), C.c_uint64()
        _check(L.lib().smelter_graph_weight_checksum(self._owner._h, C.byref(s), C.byref(b)))
        return s.value, b.value

    def weightArena(self):
        p, b = C.c_void_p(), C.c_uint64()
        _check(L.lib().smelter_graph_weight_arena(self._owner._h, C.byref(p), C.byref(b)))
        return p.value or 0, b.value


class ONNXGraph:
    def __init__(self, data: bytes, configuration: Optional[Configuration] = None, context: Optional[Context] = None):
        """ONNXGraph.init(data:configuration:).  `context` may be supplied now or through metalGraph(device:)."""
        self.configuration = configuration or Configuration()
        self._data = bytes(data)
        self._h = C.c_void_p()
        self.context = context
        self._nn: Optional[NNGraph] = None
        if context is not None:
            self._create()

    @staticmethod
    def contentsOf(path: str, configuration: Optional[Configuration] = None, context: Optional[Context] = None) -> "ONNXGraph":
        with open(path, "rb") as f:
            return ONNXGraph(f.read(), configuration, context)

    def _create(self) -> None:
        cfg = self.configuration._c()
        buf = (C.c_uint8 * len(self._data)).from_buffer_copy(self._data)
        _check(L.lib().smelter_graph_create(self.context._h, buf, len(self._data), C.byref(cfg), C.byref(self._h)))

    def _need(self) -> None:
        if not self._h:
            raise Errors(7, "no device context yet: pass context= or call metalGraph(device:) first")

    @property
    def modelFormat(self) -> int:
        self._need()
        f = C.c_int32()
        _check(L.lib().smelter_graph_format(self._h, C.byref(f)))
        return f.value

    @property
    def outputShapes(self) -> List[Shape]:
        self._need()
        n = C.c_int32()
        _check(L.lib().smelter_graph_num_outputs(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            s = L.smelter_shape()
            _check(L.lib().smelter_graph_output_shape(self._h, i, C.byref(s)))
            out.append(Shape(s.channels, s.width, s.height, s.depth))
        return out

    def nodeOpTypes(self) -> List[str]:
        self._need()
        n = C.c_int32()
        _check(L.lib().smelter_graph_num_nodes(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            s = C.c_char_p()
            _check(L.lib().smelter_graph_node_op_type(self._h, i, C.byref(s)))
            out.append(s.value.decode())
        return out

    def hasConverter(self, op_type: str) -> bool:
        self._need()
        y = C.c_int32()
        _check(L.lib().smelter_graph_has_converter(self._h, op_type.encode(), C.byref(y)))
        return bool(y.value)

    def shape(self, output: str) -> Shape:
        self._need()
        s = L.smelter_shape()
        _check(L.lib().smelter_graph_shape(self._h, output.encode(), C.byref(s)))
        return Shape(s.channels, s.width, s.height, s.depth)

    def register(self, name: str, converter) -> None:
        """register(name:converter:) (ONNXGraph.swift:253-257).  `converter(graph_handle, node_index) -> status`."""
        self._need()
        if not hasattr(self, "_callbacks"):
            self._callbacks = []
        cb = L.CONVERTER_FN(lambda g, node, user: int(converter(g, node)))
        self._callbacks.append(cb)
        _check(L.lib().smelter_graph_register_converter(self._h, name.encode(), cb, None))

    def metalGraph(self, device: Optional[Context] = None) -> NNGraph:
        """metalGraph(device:): graph walk through the converter registry + compile (fusion, weight upload)."""
        if device is not None and self.context is None:
            self.context = device
            self._create()
        self._need()
        if self._nn is None:
            _check(L.lib().smelter_graph_build(self._h))
            self._nn = NNGraph(self)
        return self._nn

    def close(self) -> None:
        if self._h:
            L.lib().smelter_graph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- single-kernel drivers (tests / microbenchmarks) ------------------------------------------------------------
def run_conv(context: Context, x: Image, w_oihw: np.ndarray, bias: Optional[np.ndarray], *, stride=(1, 1), pads=(0, 0, 0, 0),
             dilation=(1, 1), groups=1, act=0, clip=(0.0, 0.0), residual: Optional[Image] = None, force_path=0, iters=1):
    """One convolution through smelter_run_conv.  x: Image NCHW; w: [Cout, Cin/g, kH, kW] (any float dtype, rounded to fp16);
    pads = (top, left, bottom, right).  Returns (Image y, kernel_ms)."""
    n, c, h, w_ = x.shape
    co, cig, kh, kw = w_oihw.shape
    p = L.smelter_conv_problem(n=n, h=h, w=w_, c_in=c, c_out=co, k_h=kh, k_w=kw, stride_h=stride[0], stride_w=stride[1],
                               dil_h=dilation[0], dil_w=dilation[1], pad_t=pads[0], pad_l=pads[1], pad_b=pads[2], pad_r=pads[3],
                               groups=groups, act=act, clip_lo=clip[0], clip_hi=clip[1], has_bias=int(bias is not None),
                               has_residual=int(residual is not None), force_path=force_path)
    oh = (h + pads[0] + pads[2] - (dilation[0] * (kh - 1) + 1)) // stride[0] + 1
    ow = (w_ + pads[1] + pads[3] - (dilation[1] * (kw - 1) + 1)) // stride[1] + 1
    y = Image(context, n, co, oh, ow)
    wh = np.ascontiguousarray(w_oihw.astype(np.float16))
    bf = np.ascontiguousarray(bias.astype(np.float32)) if bias is not None else None
    ms = C.c_float()
    _check(L.lib().smelter_run_conv(context._h, C.byref(p), C.c_void_p(x.devicePointer), wh.ctypes.data_as(C.c_void_p),
                                    bf.ctypes.data_as(C.c_void_p) if bf is not None else None,
                                    C.c_void_p(residual.devicePointer) if residual is not None else None,
                                    C.c_void_p(y.devicePointer), iters, C.byref(ms)))
    return y, ms.value


EW_OPS = {"unary": 0, "binary": 1, "scale_shift": 2, "pool": 3, "global_avgpool": 4, "softmax": 5, "upsample": 6, "pad": 7,
          "concat": 8, "instance_norm": 9, "layout_roundtrip": 10}


def run_elementwise(context: Context, op: str, x: Image, *, x2: Optional[Image] = None, p0: Optional[np.ndarray] = None,
                    p1: Optional[np.ndarray] = None, out_shape: Sequence[int], iters=1, **kw):
    n, c, h, w = x.shape
    p = L.smelter_ew_problem(op=EW_OPS[op], n=n, c=c, h=h, w=w, **kw)
    y = Image(context, *out_shape)
    a = np.ascontiguousarray(p0.astype(np.float32)) if p0 is not None else None
    b = np.ascontiguousarray(p1.astype(np.float32)) if p1 is not None else None
    ms = C.c_float()
    _check(L.lib().smelter_run_elementwise(context._h, C.byref(p), C.c_void_p(x.devicePointer),
                                           C.c_void_p(x2.devicePointer) if x2 is not None else None,
                                           a.ctypes.data_as(C.c_void_p) if a is not None else None,
                                           b.ctypes.data_as(C.c_void_p) if b is not None else None,
                                           C.c_void_p(y.devicePointer), iters, C.byref(ms)))
    return y, ms.value
