// Host side of the engine: the ONNXGraph mirror (symbol tables + converter registry), the filter list it
// builds, the fusion pass, the per-batch execution plan and the executor.
//
// Reference map (Sources/Smelter/):
//   ONNXGraph            ONNXGraph.swift:3-286      -> class ONNXGraph below (same tables, same five-call converter surface)
//   NodeConverter        NodeConverter.swift:3-5    -> ConverterFn
//   MPSNNFilterNode list ONNXGraph.swift:65,263-273 -> std::vector<Filter> filters
//   MPSNNGraph           ONNXGraph.swift:185-190    -> Plan (compiled per batch size) + encode()
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/smelter_b200.h"
#include "kernels/kernels.h"
#include "onnx_wire.h"

namespace smelter {

void set_last_error(const std::string& msg);
int fail(int code, const std::string& msg);  // sets the thread-local message and returns code
#define SM_CUDA(expr)                                                                                          \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) return ::smelter::fail(SMELTER_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    // NCCL (loaded lazily with dlopen; see nccl_shim.cc)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    // benchmark helper: scratch larger than L2, written by smelter_l2_flush
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    // fp32 <-> fp16 / u8 conversion staging at the boundary: one device buffer per stream of this context (uses on one stream are
    // ordered by the stream; different streams never share a buffer), grown on demand, freed with the context
    struct Staging { void* ptr = nullptr; size_t bytes = 0; };
    std::map<cudaStream_t, Staging> staging;
    std::mutex staging_mu;
};

struct Tensor {
    Context* ctx = nullptr;
    __half* ptr = nullptr;
    int n = 0, c = 0, h = 0, w = 0;
    bool owned = false;
    size_t count() const { return size_t(n) * c * h * w; }
};

struct ImageShape {
    int c = 0, h = 0, w = 0;
    bool operator==(const ImageShape& o) const { return c == o.c && h == o.h && w == o.w; }
};

enum class FilterKind {
    Conv, BatchNorm, InstanceNorm, Unary, Binary, Pool, GlobalAvgPool, Upsample, Concat, Reshape, Softmax, Pad, Alias
};

struct Filter {
    FilterKind kind;
    std::string op_type;           // ONNX op that created it (for the plan dump)
    std::vector<int> in;           // value ids
    int out = -1;                  // value id
    bool removed = false;          // fused away

    // Conv / Gemm
    int c_out = 0, c_in_g = 0, k_h = 1, k_w = 1, stride_h = 1, stride_w = 1, dil_h = 1, dil_w = 1, groups = 1;
    bool bcast = false;      // Binary: in[1] is [C, 1, 1] and is broadcast over in[0]'s pixels
    int group_expanded = 0;  // G > 1: a grouped convolution running as the dense one with block-diagonal weights (engine.cc build())
    int pads[4] = {0, 0, 0, 0};    // top, left, bottom, right
    std::vector<float> w;          // OHWI fp32 [c_out][k_h][k_w][c_in_g]
    std::vector<float> bias;       // [c_out]
    bool is_gemm = false;
    // fused epilogue (also used by BatchNorm / InstanceNorm / Binary for a trailing ReLU)
    int act = k::ACT_NONE;
    float clip_lo = 0.f, clip_hi = 0.f;
    int residual = -1;             // value id added before the activation

    // BatchNorm: scale/shift (folded from gamma,beta,mean,var,eps).  InstanceNorm: gamma/beta + eps.
    std::vector<float> p0, p1;
    float eps = 1e-5f;
    // Unary / Binary / Softmax / Upsample / Pad / Pool
    int sub = 0;                   // unary kind, binary kind, log-softmax flag, upsample mode, pad mode, is_max
    float alpha = 0.f, beta = 0.f; // unary params, pad value in alpha
    int scale_h = 1, scale_w = 1, align_corners = 1;
    int pool_pad_h = 0, pool_pad_w = 0;

    // device-side constants (offsets into the weight arena), filled by build()
    size_t w_off = 0, bias_off = 0, p0_off = 0, p1_off = 0;
    int conv_mode = 0;             // k::ConvMode, or 4 = depthwise
    // Stride-2 stem reading a graph input: run as a stride-1 convolution over the 2x2 space-to-depth image the boundary
    // conversion writes (4 * Cin channels, pitch 16): a 7x7/2 stem becomes 4 dense K blocks instead of 7 sparse ones.
    // ConvTranspose (Converters.swift:266-287): run as a stride-1 convolution with the 180-degree-flipped filter over the input with
    // stride-1 zeros inserted between pixels and a (k-1)*d - pad border (weights are already flipped, reformat_conv_weight).
    bool transposed = false;
    int tr_stride_h = 1, tr_stride_w = 1, out_pad_h = 0, out_pad_w = 0;
    bool s2d = false;
    // Narrow-output stride-1 convolution behind a Pad (TransformerNet's 9x9 32 -> 3 output layer): computed on the 2 x 2
    // space-to-depth fold of its padded input, all four output phases of a folded pixel as 4 * c_out GEMM columns -- ceil(k / 2)^2
    // taps of 4 * Cin dense channels instead of k^2 taps of Cin (engine.cc "phase-folded").  The Pad writes the fold (s2d_out).
    int phase_fold = 0;            // fold factor F (2 or 4), 0 = not folded
    // Narrow-input stride-1 convolution without padding of its own (TransformerNet's 9x9 3 -> 32 input layer): `wfold` horizontally
    // neighbouring pixels are one pixel of wfold x Cin channels -- a pure re-interpretation of the NHWC input and output buffers
    // (engine.cc "width-folded"); 0 = not folded
    int wfold = 0;
    int s2d_out = 0;               // Pad: output in F x F space-to-depth layout (F = 2 or 4)
    // nearest Upsample x2 -> reflect Pad 1 -> Conv 3x3 -> InstanceNorm (TransformerNet's decoder stages): the convolution runs on the
    // edge-padded LOW-resolution image with the four output phases of a low-resolution pixel as 4 * Cout GEMM columns (every phase
    // sees a 2 x 2 neighbourhood, tap weights that meet the same source pixel pre-summed): a quarter of the operand bytes and of the
    // multiply-adds, no upsampled tensor.  Its output [H, W, 4, Cout] is the 2H x 2W image in a permuted pixel order -- which the
    // normalisation's statistics do not see; its store un-permutes (unfold_w = W).  engine.cc "upsample-folded".
    bool upfold = false;           // Conv
    // Network-input convolution behind a Pad (TransformerNet's 9x9 3 -> 32 layer, k = 1 mod 4, at most 4 input channels): the boundary
    // conversion writes the padded image as its 4 x 4 space-to-depth fold with 4 channels per pixel (64 dense channels), the
    // convolution runs on it with ((k + 6) / 4)^2 taps and the 16 output phases as 16 * Cout GEMM columns, the instance norm behind it
    // un-permutes (unfold_f = 4).  in_fold_pad / in_fold_mode / in_fold_value: the absorbed Pad.
    bool in_fold = false;
    int in_fold_pad[4] = {0, 0, 0, 0}, in_fold_mode = 0;
    float in_fold_value = 0.f;
    int unfold_f = 2;              // InstanceNorm: fold factor of the permuted pixels (2 or 4)
    int unfold_w = 0;              // InstanceNorm behind an upsample-folded convolution: low-resolution width
    // InstanceNorm whose only reader was a reflection Pad: the norm writes the padded image itself (interior at an offset, the border
    // pixels from their mirror sources, kernels.h NormStore) -- `out` is the Pad's output, the Pad is gone.  norm_s2d = its s2d_out.
    // Pad whose input was a residual Add: the Pad kernel adds (in[0] + in[1], activation `act`) on its way, and also stores the plain
    // sum as value out2 when the Add had other readers (the next block's skip connection); out2 < 0: nobody else reads it
    bool pad_add = false;
    int out2 = -1;
    int norm_pad[4] = {0, 0, 0, 0};
    int norm_s2d = 0, norm_pad_mode = 1;  // k::PAD_REFLECT or k::PAD_EDGE; out2 >= 0: the un-padded result has another reader and is stored too
    bool norm_padded = false;
    int s2d_h = 0, s2d_w = 0;      // padded input size (even) that is folded: the s2d image is s2d_h/2 x s2d_w/2
};

struct Value {
    std::string name;
    ImageShape shape;
    int alias_of = -1;  // view of another value's buffer (Flatten/Reshape views, Dropout)
    bool is_input = false;
};

struct Plan;  // per-batch compiled executor state

class ONNXGraph;
using ConverterFn = std::function<int(ONNXGraph&, int node_index)>;

class ONNXGraph {
   public:
    ONNXGraph(Context* ctx) : ctx_(ctx) {}
    ~ONNXGraph();

    // ONNXGraph.init(data:configuration:)  ONNXGraph.swift:95-156
    int init(const uint8_t* data, size_t len, const smelter_config& cfg);
    // ONNXGraph.metalGraph(device:)        ONNXGraph.swift:169-193
    int build();
    int encode(cudaStream_t stream, const Tensor* const* sources, int n_sources, const Tensor** result);

    // ---- the converter-facing surface (ONNXGraph.swift:259-285) ----
    int output(const std::string& name) const;                       // output(name:)  -> value id or -1
    const ImageShape* shape(const std::string& name) const;          // shape(output:)
    const onnx::TensorProto* tensor(const std::string& name) const;  // tensor(name:)
    void initTensor(const std::string& name, const onnx::TensorProto* t) { tensors_[name] = t; }
    // addFilter: registers `f` and creates the named output image node(s) with `shape`
    int addFilter(Filter&& f, const ImageShape& shape, const std::vector<std::string>& outputs);
    int addAlias(int value, const ImageShape& shape, const std::vector<std::string>& outputs);
    void registerConverter(const std::string& name, ConverterFn fn) { converters_[name] = std::move(fn); }  // register(name:converter:)

    const onnx::NodeProto& node(int i) const { return model_.graph.node[size_t(i)]; }
    int num_nodes() const { return int(model_.graph.node.size()); }
    const onnx::ModelProto& model() const { return model_; }
    const smelter_config& config() const { return cfg_; }
    smelter_format format() const { return format_; }
    bool has_converter(const std::string& op) const { return converters_.count(op) != 0; }
    bool built() const { return built_; }
    Context* ctx() const { return ctx_; }
    const std::vector<Value>& values() const { return values_; }

    // One plan (activation arena, captured CUDA graph, result tensor) per (batch size, stream): encodes on different streams own
    // different arenas and may be in flight together; on one stream they are ordered by the stream.  stream == nullptr is the context's.
    int plan_for(int batch, Plan** out, cudaStream_t stream = nullptr);
    int num_launches(int batch, int* n);
    int profile(cudaStream_t stream, const Tensor* const* sources, int n_sources, int iters, std::vector<float>* ms,
                std::vector<double>* flops, std::vector<double>* bytes, std::vector<int>* is_tensor);
    int plan_dump(int batch, std::string* out);
    int broadcast_weights(int root);
    int weight_checksum(uint64_t* sum, uint64_t* bytes);
    void* weight_arena() const { return weight_arena_; }
    size_t weight_bytes() const { return weight_bytes_; }

   private:
    int initOutputs();  // ONNXGraph.swift:197-251
    void registerBuiltins();
    void fuse();
    int upload_weights();
    int consumers_of(int value) const;

    Context* ctx_;
    smelter_config cfg_{};
    smelter_format format_ = SMELTER_FORMAT_ONNX;
    std::vector<uint8_t> bytes_;     // owned copy of the model (raw_data views point in here)
    onnx::ModelProto model_;
    std::map<std::string, ConverterFn> converters_;
    std::map<std::string, const onnx::TensorProto*> tensors_;
    std::map<std::string, int> outputs_;   // name -> value id   (the MPSNNImageNode table)
    std::vector<Value> values_;            // per value: name + shape (the nodeShapes table)
    std::vector<Filter> filters_;
    std::vector<int> input_values_;        // graph inputs that are not initializers, in file order
    int output_value_ = -1;
    bool built_ = false;
    bool weights_pending_ = false;  // built with defer_weights and not yet filled by broadcast_weights()

    void* weight_arena_ = nullptr;
    size_t weight_bytes_ = 0;
    std::map<std::pair<int, uintptr_t>, std::shared_ptr<Plan>> plans_;
};

// ---- shared host helpers (also exported through the C ABI) -------------------------------------------------
// Array.reformatingConvolutionWeight  Extensions/Foundation/Array+Extensions.swift:52-93
void reformat_conv_weight(const void* src, void* dst, int elem_size, int c_out, int c_in, int k_h, int k_w, bool is_transpose);
// ONNX_ConvolutionPadding.paddedSize  Padding/ONNXConvolutionPadding.swift:91-113 (+ dilation)
int conv_output_size(int in, int k, int stride, int dil, int pad_lo, int pad_hi, int out_pad, bool is_transpose);
// PyTorchPoolPadding.paddedSize       Padding/PyTorchPoolPadding.swift:94-103
int pool_output_size(int in, int k, int stride, int pad);

// Weight packers for the conv kernels (w is OHWI fp32 [c_out][k_h][k_w][c_in]).
void pack_weights_ohwi(const float* w, int c_out, int c_in, int k_h, int k_w, int c_in_pitch, uint16_t* dst);  // [c_out][k_h*k_w][pitch]
void pack_weights_depthwise(const float* w, int c, int k_h, int k_w, int c_pitch, uint16_t* dst);               // [k_h*k_w][pitch]
int pick_conv_mode(int c_in, int c_out, int groups, int k_h, int k_w, int stride_h, int stride_w, int dil_w, const int pads[4]);

// NCCL shim (dlopen'd so the library loads on machines without NCCL / GPUs).
int nccl_unique_id(uint8_t id[128]);
int nccl_init(Context* ctx, const uint8_t id[128], int rank, int world);
int nccl_broadcast(Context* ctx, void* buf, size_t bytes, int root, cudaStream_t stream);
void nccl_destroy(Context* ctx);

}  // namespace smelter
