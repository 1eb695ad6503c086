// ONNX protobuf wire-format reader (no libprotobuf).
//
// Replaces `Onnx_ModelProto(serializedData:)` (reference Sources/Smelter/ONNXGraph.swift:96) and the
// generated schema Sources/Smelter/onnx.pb.swift.  Field numbers follow the reference's name maps:
//   ModelProto    onnx.pb.swift:1402-1412   GraphProto  :1597-1606   NodeProto :1337-1345
//   AttributeProto :1079-1094               TensorProto :1668-1683   DataType  :1832-1850
//   ValueInfoProto :1260-1263  TypeProto :1975-1977  TypeProto.Tensor :2051-2053
//   TensorShapeProto :1897-1898  Dimension :1926-1929
//
// `raw_data` / `s` payloads are string_views into the caller's buffer: zero copy.  The owner of the
// ModelProto must keep the serialized bytes alive for as long as the views are used.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

namespace smelter {
namespace onnx {

// TensorProto.DataType (onnx.pb.swift:1832-1850)
enum DataType : int32_t {
    DT_UNDEFINED = 0, DT_FLOAT = 1, DT_UINT8 = 2, DT_INT8 = 3, DT_UINT16 = 4, DT_INT16 = 5, DT_INT32 = 6,
    DT_INT64 = 7, DT_STRING = 8, DT_BOOL = 9, DT_FLOAT16 = 10, DT_DOUBLE = 11, DT_UINT32 = 12,
    DT_UINT64 = 13, DT_COMPLEX64 = 14, DT_COMPLEX128 = 15, DT_BFLOAT16 = 16,
};

struct TensorProto {
    std::vector<int64_t> dims;         // 1
    int32_t data_type = 0;             // 2
    std::vector<float> float_data;     // 4
    std::vector<int32_t> int32_data;   // 5
    std::vector<int64_t> int64_data;   // 7
    std::string name;                  // 8
    std::string_view raw_data;         // 9 (view into the model bytes)
    std::vector<double> double_data;   // 10
    std::vector<uint64_t> uint64_data; // 11

    // Onnx_TensorProto+Extensions.swift:64-66
    int64_t length() const {
        int64_t n = 1;
        for (int64_t d : dims) n *= d;
        return n;
    }
    // Onnx_TensorProto+Extensions.swift:2-34.  Returns false where the reference calls fatalError.
    bool integers(std::vector<int64_t>* out) const;
    // Onnx_TensorProto+Extensions.swift:36-62.  Returns false where the reference calls fatalError.
    bool floats(std::vector<float>* out) const;
};

struct AttributeProto {
    std::string name;              // 1
    float f = 0.f;                 // 2
    int64_t i = 0;                 // 3
    std::string_view s;            // 4
    TensorProto t;                 // 5
    bool has_t = false;
    std::vector<float> floats;     // 7
    std::vector<int64_t> ints;     // 8
    int32_t type = 0;              // 20
};

struct NodeProto {
    std::vector<std::string> input;   // 1
    std::vector<std::string> output;  // 2
    std::string name;                 // 3
    std::string op_type;              // 4
    std::vector<AttributeProto> attribute;  // 5
    std::string domain;               // 7

    const AttributeProto* attr(const char* name) const;
};

struct ValueInfoProto {
    std::string name;          // 1
    int32_t elem_type = 0;     // type(2).tensor_type(1).elem_type(1)
    std::vector<int64_t> dims; // type.tensor_type.shape(2).dim(1).dim_value(1); dim_param => 0
    bool has_shape = false;
};

struct GraphProto {
    std::vector<NodeProto> node;             // 1
    std::string name;                        // 2
    std::vector<TensorProto> initializer;    // 5
    std::vector<ValueInfoProto> input;       // 11
    std::vector<ValueInfoProto> output;      // 12
    std::vector<ValueInfoProto> value_info;  // 13
};

struct OperatorSetId {
    std::string domain;   // 1
    int64_t version = 0;  // 2
};

struct ModelProto {
    int64_t ir_version = 0;          // 1
    std::string producer_name;       // 2
    std::string producer_version;    // 3
    GraphProto graph;                // 7
    std::vector<OperatorSetId> opset_import;  // 8
};

// Decode `len` bytes at `data`.  Returns false (and fills *err) on malformed input.
bool parse_model(const uint8_t* data, size_t len, ModelProto* out, std::string* err);

// IEEE half <-> float (host).  Restates Float16.swift:17-45 / 53-77 (vImage planar conversions):
// round-to-nearest-even, subnormals, inf and NaN preserved.
float half_to_float(uint16_t h);
uint16_t float_to_half(float f);

}  // namespace onnx
}  // namespace smelter
