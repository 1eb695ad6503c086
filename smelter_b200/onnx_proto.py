"""Minimal ONNX protobuf codec (reader + writer) in pure Python/numpy — no `onnx`, no `google.protobuf`.

Host-side tooling only (model generation, the ONNX2MPS restatement, the oracle); the engine itself parses
models with the C++ wire reader in csrc/onnx_wire.cc.  Field numbers follow the reference's generated
schema Sources/Smelter/onnx.pb.swift:
  ModelProto :1402-1412, GraphProto :1597-1606, NodeProto :1337-1345, AttributeProto :1079-1094,
  TensorProto :1668-1683, DataType :1832-1850, ValueInfoProto :1260-1263, TypeProto.Tensor :2051-2053,
  TensorShapeProto.Dimension :1926-1929, OperatorSetIdProto (domain 1, version 2).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

# TensorProto.DataType
FLOAT, UINT8, INT8, UINT16, INT16, INT32, INT64, STRING, BOOL, FLOAT16, DOUBLE, UINT32, UINT64 = range(1, 14)
# AttributeProto.AttributeType
AT_FLOAT, AT_INT, AT_STRING, AT_TENSOR, AT_GRAPH, AT_FLOATS, AT_INTS, AT_STRINGS = 1, 2, 3, 4, 5, 6, 7, 8

NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, UINT16: np.uint16, INT16: np.int16, INT32: np.int32,
         INT64: np.int64, BOOL: np.bool_, FLOAT16: np.float16, DOUBLE: np.float64, UINT32: np.uint32, UINT64: np.uint64}
DT_OF = {np.dtype(v): k for k, v in NP_OF.items()}


# ---------------------------------------------------------------------------------------------- wire level
def _put_varint(out: bytearray, v: int) -> None:
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    shift = 0
    result = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _fields(buf: memoryview):
    """Yield (field_number, wire_type, value) — value is int for varint/fixed, memoryview for length-delimited."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
            yield fno, wt, v
        elif wt == 1:
            if pos + 8 > n:
                raise ValueError("truncated fixed64")
            yield fno, wt, bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated length-delimited field")
            yield fno, wt, buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            if pos + 4 > n:
                raise ValueError("truncated fixed32")
            yield fno, wt, bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")


def _key(out: bytearray, fno: int, wt: int) -> None:
    _put_varint(out, (fno << 3) | wt)


def _put_bytes(out: bytearray, fno: int, data: Union[bytes, bytearray, memoryview]) -> None:
    _key(out, fno, 2)
    _put_varint(out, len(data))
    out += data


def _put_str(out: bytearray, fno: int, s: str) -> None:
    _put_bytes(out, fno, s.encode("utf-8"))


def _put_int(out: bytearray, fno: int, v: int) -> None:
    _key(out, fno, 0)
    _put_varint(out, v)


# ---------------------------------------------------------------------------------------------- messages
@dataclass
class Tensor:
    name: str = ""
    dims: List[int] = field(default_factory=list)
    data_type: int = 0
    raw_data: Optional[bytes] = None
    float_data: List[float] = field(default_factory=list)
    int32_data: List[int] = field(default_factory=list)
    int64_data: List[int] = field(default_factory=list)
    double_data: List[float] = field(default_factory=list)
    uint64_data: List[int] = field(default_factory=list)

    @staticmethod
    def from_numpy(name: str, a: np.ndarray) -> "Tensor":
        a = np.ascontiguousarray(a)
        return Tensor(name=name, dims=list(a.shape), data_type=DT_OF[a.dtype], raw_data=a.tobytes())

    def numpy(self) -> np.ndarray:
        """Typed view of the payload, whatever storage field carries it."""
        dt = NP_OF[self.data_type]
        shape = tuple(self.dims)
        if self.raw_data is not None and len(self.raw_data):
            return np.frombuffer(self.raw_data, dtype=dt).reshape(shape).copy()
        if self.data_type == FLOAT and self.float_data:
            return np.asarray(self.float_data, dtype=np.float32).reshape(shape)
        if self.data_type == DOUBLE and self.double_data:
            return np.asarray(self.double_data, dtype=np.float64).reshape(shape)
        if self.data_type == INT64 and self.int64_data:
            return np.asarray(self.int64_data, dtype=np.int64).reshape(shape)
        if self.data_type in (UINT32, UINT64) and self.uint64_data:
            return np.asarray(self.uint64_data, dtype=dt).reshape(shape)
        if self.data_type == FLOAT16 and self.int32_data:  # fp16 bit patterns in int32_data
            return np.asarray(self.int32_data, dtype=np.uint16).view(np.float16).reshape(shape)
        if self.int32_data:
            return np.asarray(self.int32_data).astype(dt).reshape(shape)
        return np.zeros(shape, dtype=dt)

    def serialize(self) -> bytes:
        out = bytearray()
        if self.dims:
            packed = bytearray()
            for d in self.dims:
                _put_varint(packed, d)
            _put_bytes(out, 1, packed)
        _put_int(out, 2, self.data_type)
        if self.float_data:
            _put_bytes(out, 4, struct.pack(f"<{len(self.float_data)}f", *self.float_data))
        if self.int32_data:
            packed = bytearray()
            for v in self.int32_data:
                _put_varint(packed, v)
            _put_bytes(out, 5, packed)
        if self.int64_data:
            packed = bytearray()
            for v in self.int64_data:
                _put_varint(packed, v)
            _put_bytes(out, 7, packed)
        if self.name:
            _put_str(out, 8, self.name)
        if self.raw_data is not None:
            _put_bytes(out, 9, self.raw_data)
        if self.double_data:
            _put_bytes(out, 10, struct.pack(f"<{len(self.double_data)}d", *self.double_data))
        if self.uint64_data:
            packed = bytearray()
            for v in self.uint64_data:
                _put_varint(packed, v)
            _put_bytes(out, 11, packed)
        return bytes(out)

    @staticmethod
    def parse(buf: memoryview) -> "Tensor":
        t = Tensor()
        for fno, wt, v in _fields(buf):
            if fno == 1:
                if wt == 2:
                    pos = 0
                    while pos < len(v):
                        d, pos = _get_varint(v, pos)
                        t.dims.append(_signed(d))
                else:
                    t.dims.append(_signed(v))
            elif fno == 2:
                t.data_type = v
            elif fno == 4:
                if wt == 2:
                    t.float_data.extend(struct.unpack(f"<{len(v) // 4}f", bytes(v)))
                else:
                    t.float_data.append(struct.unpack("<f", v)[0])
            elif fno == 5:
                if wt == 2:
                    pos = 0
                    while pos < len(v):
                        d, pos = _get_varint(v, pos)
                        d = _signed(d)
                        t.int32_data.append(d)
                else:
                    t.int32_data.append(_signed(v))
            elif fno == 7:
                if wt == 2:
                    pos = 0
                    while pos < len(v):
                        d, pos = _get_varint(v, pos)
                        t.int64_data.append(_signed(d))
                else:
                    t.int64_data.append(_signed(v))
            elif fno == 8:
                t.name = bytes(v).decode("utf-8")
            elif fno == 9:
                t.raw_data = bytes(v)
            elif fno == 10:
                if wt == 2:
                    t.double_data.extend(struct.unpack(f"<{len(v) // 8}d", bytes(v)))
                else:
                    t.double_data.append(struct.unpack("<d", v)[0])
            elif fno == 11:
                if wt == 2:
                    pos = 0
                    while pos < len(v):
                        d, pos = _get_varint(v, pos)
                        t.uint64_data.append(d)
                else:
                    t.uint64_data.append(v)
        return t


@dataclass
class Attribute:
    name: str = ""
    type: int = 0
    f: float = 0.0
    i: int = 0
    s: bytes = b""
    t: Optional[Tensor] = None
    floats: List[float] = field(default_factory=list)
    ints: List[int] = field(default_factory=list)

    def serialize(self) -> bytes:
        out = bytearray()
        _put_str(out, 1, self.name)
        if self.type == AT_FLOAT:
            _key(out, 2, 5)
            out += struct.pack("<f", self.f)
        elif self.type == AT_INT:
            _put_int(out, 3, self.i)
        elif self.type == AT_STRING:
            _put_bytes(out, 4, self.s)
        elif self.type == AT_TENSOR:
            _put_bytes(out, 5, self.t.serialize())
        elif self.type == AT_FLOATS:
            _put_bytes(out, 7, struct.pack(f"<{len(self.floats)}f", *self.floats))
        elif self.type == AT_INTS:
            packed = bytearray()
            for v in self.ints:
                _put_varint(packed, v)
            _put_bytes(out, 8, packed)
        _put_int(out, 20, self.type)
        return bytes(out)

    @staticmethod
    def parse(buf: memoryview) -> "Attribute":
        a = Attribute()
        for fno, wt, v in _fields(buf):
            if fno == 1:
                a.name = bytes(v).decode("utf-8")
            elif fno == 2:
                a.f = struct.unpack("<f", v)[0]
            elif fno == 3:
                a.i = _signed(v)
            elif fno == 4:
                a.s = bytes(v)
            elif fno == 5:
                a.t = Tensor.parse(v)
            elif fno == 7:
                if wt == 2:
                    a.floats.extend(struct.unpack(f"<{len(v) // 4}f", bytes(v)))
                else:
                    a.floats.append(struct.unpack("<f", v)[0])
            elif fno == 8:
                if wt == 2:
                    pos = 0
                    while pos < len(v):
                        d, pos = _get_varint(v, pos)
                        a.ints.append(_signed(d))
                else:
                    a.ints.append(_signed(v))
            elif fno == 20:
                a.type = v
        return a


def attr(name: str, value) -> Attribute:
    """Build an attribute from a Python value (int, float, str/bytes, list of ints/floats, Tensor)."""
    if isinstance(value, Tensor):
        return Attribute(name=name, type=AT_TENSOR, t=value)
    if isinstance(value, bool):
        return Attribute(name=name, type=AT_INT, i=int(value))
    if isinstance(value, int):
        return Attribute(name=name, type=AT_INT, i=value)
    if isinstance(value, float):
        return Attribute(name=name, type=AT_FLOAT, f=value)
    if isinstance(value, str):
        return Attribute(name=name, type=AT_STRING, s=value.encode())
    if isinstance(value, bytes):
        return Attribute(name=name, type=AT_STRING, s=value)
    if isinstance(value, (list, tuple)):
        if all(isinstance(v, int) for v in value):
            return Attribute(name=name, type=AT_INTS, ints=list(value))
        return Attribute(name=name, type=AT_FLOATS, floats=[float(v) for v in value])
    raise TypeError(f"unsupported attribute value {value!r}")


@dataclass
class Node:
    op_type: str = ""
    input: List[str] = field(default_factory=list)
    output: List[str] = field(default_factory=list)
    name: str = ""
    attribute: List[Attribute] = field(default_factory=list)
    domain: str = ""

    def attr(self, name: str) -> Optional[Attribute]:
        for a in self.attribute:
            if a.name == name:
                return a
        return None

    def serialize(self) -> bytes:
        out = bytearray()
        for s in self.input:
            _put_str(out, 1, s)
        for s in self.output:
            _put_str(out, 2, s)
        if self.name:
            _put_str(out, 3, self.name)
        _put_str(out, 4, self.op_type)
        for a in self.attribute:
            _put_bytes(out, 5, a.serialize())
        if self.domain:
            _put_str(out, 7, self.domain)
        return bytes(out)

    @staticmethod
    def parse(buf: memoryview) -> "Node":
        n = Node()
        for fno, wt, v in _fields(buf):
            if fno == 1:
                n.input.append(bytes(v).decode("utf-8"))
            elif fno == 2:
                n.output.append(bytes(v).decode("utf-8"))
            elif fno == 3:
                n.name = bytes(v).decode("utf-8")
            elif fno == 4:
                n.op_type = bytes(v).decode("utf-8")
            elif fno == 5:
                n.attribute.append(Attribute.parse(v))
            elif fno == 7:
                n.domain = bytes(v).decode("utf-8")
        return n


@dataclass
class ValueInfo:
    name: str = ""
    elem_type: int = 0
    dims: Optional[List[Union[int, str]]] = None  # int = dim_value, str = dim_param

    def serialize(self) -> bytes:
        shape = bytearray()
        for d in self.dims or []:
            dim = bytearray()
            if isinstance(d, str):
                _put_str(dim, 2, d)
            else:
                _put_int(dim, 1, d)
            _put_bytes(shape, 1, dim)
        tt = bytearray()
        _put_int(tt, 1, self.elem_type)
        if self.dims is not None:
            _put_bytes(tt, 2, shape)
        tp = bytearray()
        _put_bytes(tp, 1, tt)
        out = bytearray()
        _put_str(out, 1, self.name)
        _put_bytes(out, 2, tp)
        return bytes(out)

    @staticmethod
    def parse(buf: memoryview) -> "ValueInfo":
        vi = ValueInfo()
        for fno, wt, v in _fields(buf):
            if fno == 1:
                vi.name = bytes(v).decode("utf-8")
            elif fno == 2:
                for f2, _, v2 in _fields(v):
                    if f2 != 1:
                        continue
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            vi.elem_type = v3
                        elif f3 == 2:
                            vi.dims = []
                            for f4, _, v4 in _fields(v3):
                                if f4 != 1:
                                    continue
                                d: Union[int, str] = 0
                                for f5, _, v5 in _fields(v4):
                                    if f5 == 1:
                                        d = _signed(v5)
                                    elif f5 == 2:
                                        d = bytes(v5).decode("utf-8")
                                vi.dims.append(d)
        return vi


@dataclass
class Graph:
    node: List[Node] = field(default_factory=list)
    name: str = ""
    initializer: List[Tensor] = field(default_factory=list)
    input: List[ValueInfo] = field(default_factory=list)
    output: List[ValueInfo] = field(default_factory=list)
    value_info: List[ValueInfo] = field(default_factory=list)

    def serialize(self) -> bytes:
        out = bytearray()
        for n in self.node:
            _put_bytes(out, 1, n.serialize())
        if self.name:
            _put_str(out, 2, self.name)
        for t in self.initializer:
            _put_bytes(out, 5, t.serialize())
        for v in self.input:
            _put_bytes(out, 11, v.serialize())
        for v in self.output:
            _put_bytes(out, 12, v.serialize())
        for v in self.value_info:
            _put_bytes(out, 13, v.serialize())
        return bytes(out)

    @staticmethod
    def parse(buf: memoryview) -> "Graph":
        g = Graph()
        for fno, wt, v in _fields(buf):
            if fno == 1:
                g.node.append(Node.parse(v))
            elif fno == 2:
                g.name = bytes(v).decode("utf-8")
            elif fno == 5:
                g.initializer.append(Tensor.parse(v))
            elif fno == 11:
                g.input.append(ValueInfo.parse(v))
            elif fno == 12:
                g.output.append(ValueInfo.parse(v))
            elif fno == 13:
                g.value_info.append(ValueInfo.parse(v))
        return g


@dataclass
class Model:
    ir_version: int = 4
    producer_name: str = ""
    producer_version: str = ""
    graph: Graph = field(default_factory=Graph)
    opset_import: List[Tuple[str, int]] = field(default_factory=lambda: [("", 9)])

    def serialize(self) -> bytes:
        out = bytearray()
        _put_int(out, 1, self.ir_version)
        if self.producer_name:
            _put_str(out, 2, self.producer_name)
        if self.producer_version:
            _put_str(out, 3, self.producer_version)
        _put_bytes(out, 7, self.graph.serialize())
        for domain, version in self.opset_import:
            op = bytearray()
            if domain:
                _put_str(op, 1, domain)
            _put_int(op, 2, version)
            _put_bytes(out, 8, op)
        return bytes(out)

    @staticmethod
    def parse(data: Union[bytes, bytearray, memoryview]) -> "Model":
        m = Model(opset_import=[])
        for fno, wt, v in _fields(memoryview(data)):
            if fno == 1:
                m.ir_version = v
            elif fno == 2:
                m.producer_name = bytes(v).decode("utf-8")
            elif fno == 3:
                m.producer_version = bytes(v).decode("utf-8")
            elif fno == 7:
                m.graph = Graph.parse(v)
            elif fno == 8:
                domain, version = "", 0
                for f2, _, v2 in _fields(v):
                    if f2 == 1:
                        domain = bytes(v2).decode("utf-8")
                    elif f2 == 2:
                        version = v2
                m.opset_import.append((domain, version))
        return m

    def initializers(self) -> Dict[str, Tensor]:
        return {t.name: t for t in self.graph.initializer}


def load(path: str) -> Model:
    with open(path, "rb") as f:
        return Model.parse(f.read())


def save(model: Model, path: str) -> None:
    with open(path, "wb") as f:
        f.write(model.serialize())
