// ONNX protobuf wire-format reader.  See onnx_wire.h for the reference citations.
#include "onnx_wire.h"

#include <cmath>
#include <cstring>

namespace smelter {
namespace onnx {

namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;

    bool done() const { return p >= end || !ok; }

    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (p < end && shift < 64) {
            uint8_t b = *p++;
            v |= uint64_t(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false;
        return 0;
    }
    uint32_t fixed32() {
        if (end - p < 4) { ok = false; return 0; }
        uint32_t v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    uint64_t fixed64() {
        if (end - p < 8) { ok = false; return 0; }
        uint64_t v;
        memcpy(&v, p, 8);
        p += 8;
        return v;
    }
    Reader sub() {
        uint64_t n = varint();
        if (!ok || uint64_t(end - p) < n) { ok = false; return Reader{p, p}; }
        Reader r{p, p + n};
        p += n;
        return r;
    }
    std::string_view bytes() {
        Reader r = sub();
        return std::string_view(reinterpret_cast<const char*>(r.p), size_t(r.end - r.p));
    }
    void skip(uint32_t wire) {
        switch (wire) {
            case 0: varint(); break;
            case 1: fixed64(); break;
            case 2: sub(); break;
            case 5: fixed32(); break;
            default: ok = false; break;  // groups (3,4) are not used by ONNX
        }
    }
};

template <class T>
void read_varints(Reader& r, uint32_t wire, std::vector<T>* out) {
    if (wire == 2) {  // packed
        Reader s = r.sub();
        while (!s.done()) out->push_back(T(s.varint()));
        if (!s.ok) r.ok = false;
    } else {
        out->push_back(T(r.varint()));
    }
}

void read_floats(Reader& r, uint32_t wire, std::vector<float>* out) {
    auto one = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
    if (wire == 2) {
        Reader s = r.sub();
        size_t n = size_t(s.end - s.p) / 4;
        out->reserve(out->size() + n);
        while (!s.done()) out->push_back(one(s.fixed32()));
        if (!s.ok) r.ok = false;
    } else {
        out->push_back(one(r.fixed32()));
    }
}

void read_doubles(Reader& r, uint32_t wire, std::vector<double>* out) {
    auto one = [](uint64_t u) { double f; memcpy(&f, &u, 8); return f; };
    if (wire == 2) {
        Reader s = r.sub();
        while (!s.done()) out->push_back(one(s.fixed64()));
        if (!s.ok) r.ok = false;
    } else {
        out->push_back(one(r.fixed64()));
    }
}

bool parse_tensor(Reader r, TensorProto* t) {
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        switch (field) {
            case 1: read_varints(r, wire, &t->dims); break;
            case 2: t->data_type = int32_t(r.varint()); break;
            case 4: read_floats(r, wire, &t->float_data); break;
            case 5: read_varints(r, wire, &t->int32_data); break;
            case 7: read_varints(r, wire, &t->int64_data); break;
            case 8: t->name = std::string(r.bytes()); break;
            case 9: t->raw_data = r.bytes(); break;
            case 10: read_doubles(r, wire, &t->double_data); break;
            case 11: read_varints(r, wire, &t->uint64_data); break;
            default: r.skip(wire); break;
        }
    }
    return r.ok;
}

bool parse_attribute(Reader r, AttributeProto* a) {
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        switch (field) {
            case 1: a->name = std::string(r.bytes()); break;
            case 2: { uint32_t u = r.fixed32(); memcpy(&a->f, &u, 4); break; }
            case 3: a->i = int64_t(r.varint()); break;
            case 4: a->s = r.bytes(); break;
            case 5: a->has_t = true; if (!parse_tensor(r.sub(), &a->t)) r.ok = false; break;
            case 7: read_floats(r, wire, &a->floats); break;
            case 8: read_varints(r, wire, &a->ints); break;
            case 20: a->type = int32_t(r.varint()); break;
            default: r.skip(wire); break;
        }
    }
    return r.ok;
}

bool parse_node(Reader r, NodeProto* n) {
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        switch (field) {
            case 1: n->input.emplace_back(r.bytes()); break;
            case 2: n->output.emplace_back(r.bytes()); break;
            case 3: n->name = std::string(r.bytes()); break;
            case 4: n->op_type = std::string(r.bytes()); break;
            case 5: n->attribute.emplace_back(); if (!parse_attribute(r.sub(), &n->attribute.back())) r.ok = false; break;
            case 7: n->domain = std::string(r.bytes()); break;
            default: r.skip(wire); break;
        }
    }
    return r.ok;
}

bool parse_shape(Reader r, ValueInfoProto* v) {
    v->has_shape = true;
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        if (field == 1 && wire == 2) {  // Dimension
            Reader d = r.sub();
            int64_t value = 0;
            while (!d.done()) {
                uint64_t k2 = d.varint();
                uint32_t f2 = uint32_t(k2 >> 3), w2 = uint32_t(k2 & 7);
                if (f2 == 1) value = int64_t(d.varint());
                else d.skip(w2);  // dim_param (2), denotation (3)
            }
            if (!d.ok) r.ok = false;
            v->dims.push_back(value);
        } else {
            r.skip(wire);
        }
    }
    return r.ok;
}

bool parse_value_info(Reader r, ValueInfoProto* v) {
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        if (field == 1) {
            v->name = std::string(r.bytes());
        } else if (field == 2 && wire == 2) {  // TypeProto
            Reader t = r.sub();
            while (!t.done()) {
                uint64_t k2 = t.varint();
                uint32_t f2 = uint32_t(k2 >> 3), w2 = uint32_t(k2 & 7);
                if (f2 == 1 && w2 == 2) {  // tensor_type
                    Reader tt = t.sub();
                    while (!tt.done()) {
                        uint64_t k3 = tt.varint();
                        uint32_t f3 = uint32_t(k3 >> 3), w3 = uint32_t(k3 & 7);
                        if (f3 == 1) v->elem_type = int32_t(tt.varint());
                        else if (f3 == 2 && w3 == 2) { if (!parse_shape(tt.sub(), v)) tt.ok = false; }
                        else tt.skip(w3);
                    }
                    if (!tt.ok) t.ok = false;
                } else {
                    t.skip(w2);
                }
            }
            if (!t.ok) r.ok = false;
        } else {
            r.skip(wire);
        }
    }
    return r.ok;
}

bool parse_graph(Reader r, GraphProto* g) {
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        switch (field) {
            case 1: g->node.emplace_back(); if (!parse_node(r.sub(), &g->node.back())) r.ok = false; break;
            case 2: g->name = std::string(r.bytes()); break;
            case 5: g->initializer.emplace_back(); if (!parse_tensor(r.sub(), &g->initializer.back())) r.ok = false; break;
            case 11: g->input.emplace_back(); if (!parse_value_info(r.sub(), &g->input.back())) r.ok = false; break;
            case 12: g->output.emplace_back(); if (!parse_value_info(r.sub(), &g->output.back())) r.ok = false; break;
            case 13: g->value_info.emplace_back(); if (!parse_value_info(r.sub(), &g->value_info.back())) r.ok = false; break;
            default: r.skip(wire); break;
        }
    }
    return r.ok;
}

template <class T>
void raw_array(std::string_view raw, std::vector<T>* out) {
    // Data+Extensions.swift:4-14 — reinterpret raw bytes as a typed array (count = bytes / stride).
    size_t n = raw.size() / sizeof(T);
    out->resize(n);
    if (n) memcpy(out->data(), raw.data(), n * sizeof(T));
}

}  // namespace

const AttributeProto* NodeProto::attr(const char* nm) const {
    for (const auto& a : attribute)
        if (a.name == nm) return &a;
    return nullptr;
}

bool parse_model(const uint8_t* data, size_t len, ModelProto* out, std::string* err) {
    Reader r{data, data + len};
    bool saw_graph = false;
    while (!r.done()) {
        uint64_t key = r.varint();
        uint32_t field = uint32_t(key >> 3), wire = uint32_t(key & 7);
        switch (field) {
            case 1: out->ir_version = int64_t(r.varint()); break;
            case 2: out->producer_name = std::string(r.bytes()); break;
            case 3: out->producer_version = std::string(r.bytes()); break;
            case 7: saw_graph = true; if (!parse_graph(r.sub(), &out->graph)) r.ok = false; break;
            case 8: {
                Reader o = r.sub();
                OperatorSetId id;
                while (!o.done()) {
                    uint64_t k2 = o.varint();
                    uint32_t f2 = uint32_t(k2 >> 3), w2 = uint32_t(k2 & 7);
                    if (f2 == 1) id.domain = std::string(o.bytes());
                    else if (f2 == 2) id.version = int64_t(o.varint());
                    else o.skip(w2);
                }
                if (!o.ok) r.ok = false;
                out->opset_import.push_back(id);
                break;
            }
            default: r.skip(wire); break;
        }
    }
    if (!r.ok) {
        if (err) *err = "malformed ONNX protobuf";
        return false;
    }
    (void)saw_graph;  // SwiftProtobuf accepts a model without a graph (empty default); so do we
    return true;
}

// ---- Onnx_TensorProto+Extensions.swift:2-62 ------------------------------------------------------

bool TensorProto::integers(std::vector<int64_t>* out) const {
    out->clear();
    switch (data_type) {
        case DT_INT32: case DT_INT16: case DT_INT8: case DT_UINT16: case DT_UINT8: case DT_BOOL:
            for (int32_t v : int32_data) out->push_back(v);
            return true;
        case DT_INT64:
            if (int64_data.empty()) {
                std::vector<int64_t> raw;
                raw_array(raw_data, &raw);
                *out = raw;
            } else {
                *out = int64_data;
            }
            return true;
        case DT_UINT32: case DT_UINT64:
            for (uint64_t v : uint64_data) out->push_back(int64_t(v));
            return true;
        case DT_FLOAT:
            if (float_data.empty()) {
                std::vector<float> raw;
                raw_array(raw_data, &raw);
                for (float v : raw) out->push_back(int64_t(v));  // Int(Float) truncates toward zero
            } else {
                for (float v : float_data) out->push_back(int64_t(v));
            }
            return true;
        case DT_DOUBLE:
            for (double v : double_data) out->push_back(int64_t(v));
            return true;
        case DT_FLOAT16: {
            std::vector<uint16_t> raw;
            raw_array(raw_data, &raw);
            for (uint16_t h : raw) out->push_back(int64_t(half_to_float(h)));
            return true;
        }
        default:
            return false;
    }
}

bool TensorProto::floats(std::vector<float>* out) const {
    out->clear();
    switch (data_type) {
        case DT_INT32: case DT_INT16: case DT_INT8: case DT_UINT16: case DT_UINT8: case DT_BOOL:
            for (int32_t v : int32_data) out->push_back(float(v));
            return true;
        case DT_INT64:
            for (int64_t v : int64_data) out->push_back(float(v));
            return true;
        case DT_UINT32: case DT_UINT64:
            for (uint64_t v : uint64_data) out->push_back(float(v));
            return true;
        case DT_FLOAT:
            if (float_data.empty()) raw_array(raw_data, out);
            else *out = float_data;
            return true;
        case DT_DOUBLE:
            for (double v : double_data) out->push_back(float(v));
            return true;
        case DT_FLOAT16: {
            std::vector<uint16_t> raw;
            raw_array(raw_data, &raw);
            for (uint16_t h : raw) out->push_back(half_to_float(h));
            return true;
        }
        default:
            return false;
    }
}

// ---- Float16.swift ---------------------------------------------------------------------------------

float half_to_float(uint16_t h) {
    uint32_t sign = uint32_t(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | uint32_t(127 - 15 - e) << 23 | man << 13;
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | man << 13;
    } else {
        bits = sign | (exp + 127 - 15) << 23 | man << 13;
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) {  // inf / nan
        uint32_t man = absx & 0x7fffffu;
        return uint16_t(sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0));
    }
    if (absx >= 0x477ff000u) {  // rounds to >= 65520 -> inf
        return uint16_t(sign | 0x7c00u);
    }
    if (absx < 0x38800000u) {  // subnormal half or zero
        if (absx < 0x33000000u) return uint16_t(sign);  // < 2^-25 -> 0
        int e = int(absx >> 23);
        uint32_t man = (absx & 0x7fffffu) | 0x800000u;
        int shift = 113 - e + 13;  // bits to drop
        uint32_t half_man = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_man & 1))) ++half_man;
        return uint16_t(sign | half_man);
    }
    uint32_t e = (absx >> 23) - 112;
    uint32_t man = absx & 0x7fffffu;
    uint32_t h = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;  // carries into the exponent correctly
    return uint16_t(sign | h);
}

}  // namespace onnx
}  // namespace smelter
