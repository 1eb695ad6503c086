// ONNXGraph mirror: init / graph walk / fusion / weight upload / per-batch plan / encode.
// See engine.h for the reference map.
#include "engine.h"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace smelter {

namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error_string() { return g_last_error; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ---- host helpers -------------------------------------------------------------------------------------------

void reformat_conv_weight(const void* src, void* dst, int elem_size, int c_out, int c_in, int k_h, int k_w, bool is_transpose) {
    // Index maps of Array+Extensions.swift:63-90: plain (OIHW->OHWI) and ConvTranspose (IOHW->OHWI, 180 degree flip).
    // Loop order is output-major so the writes stream; same result as the reference's scalar 4-deep loop.
    const char* s = static_cast<const char*>(src);
    char* d = static_cast<char*>(dst);
    const size_t khw = size_t(k_h) * k_w;
    for (int oc = 0; oc < c_out; ++oc)
        for (int kh = 0; kh < k_h; ++kh)
            for (int kw = 0; kw < k_w; ++kw) {
                const int dkh = is_transpose ? k_h - 1 - kh : kh;
                const int dkw = is_transpose ? k_w - 1 - kw : kw;
                char* drow = d + ((size_t(oc) * k_h + dkh) * k_w + dkw) * c_in * elem_size;
                for (int ic = 0; ic < c_in; ++ic) {
                    const size_t in_idx = is_transpose ? (size_t(ic) * c_out + oc) * khw + size_t(kh) * k_w + kw
                                                       : (size_t(oc) * c_in + ic) * khw + size_t(kh) * k_w + kw;
                    memcpy(drow + size_t(ic) * elem_size, s + in_idx * elem_size, size_t(elem_size));
                }
            }
}

int conv_output_size(int in, int k, int stride, int dil, int pad_lo, int pad_hi, int out_pad, bool is_transpose) {
    if (is_transpose) return (in - 1) * stride - pad_lo - pad_hi + dil * (k - 1) + 1 + out_pad;  // ONNXConvolutionPadding.swift:97-103
    const int num = in + pad_lo + pad_hi - (dil * (k - 1) + 1);                                   // :105-110, with the dilation term
    if (num < 0) return 0;
    return num / stride + 1;
}

int pool_output_size(int in, int k, int stride, int pad) {
    // Int(Float(in + 2p - k) / Float(stride) + 1.0)  (PyTorchPoolPadding.swift:94-103): truncation toward zero
    return int(float(in + 2 * pad - k) / float(stride) + 1.0f);
}

void pack_weights_ohwi(const float* w, int c_out, int c_in, int k_h, int k_w, int c_in_pitch, uint16_t* dst) {
    const int taps = k_h * k_w;
    for (int o = 0; o < c_out; ++o)
        for (int t = 0; t < taps; ++t) {
            const float* src = w + (size_t(o) * taps + t) * c_in;
            uint16_t* d = dst + (size_t(o) * taps + t) * c_in_pitch;
            for (int c = 0; c < c_in; ++c) d[c] = onnx::float_to_half(src[c]);
            for (int c = c_in; c < c_in_pitch; ++c) d[c] = 0;
        }
}
// Stride-2 stem folded to stride 1 over a 2x2 space-to-depth image (Filter::s2d): w is OHWI [o][r][s][c]; the packed row
// operand is [o][r2][s2][16] with channel (dy*2+dx)*c_in + c holding w[o][2*r2+dy][2*s2+dx][c] (zero outside the filter).
void pack_weights_s2d(const float* w, int c_out, int c_in, int k_h, int k_w, uint16_t* dst) {
    const int r2n = (k_h + 1) / 2, s2n = (k_w + 1) / 2;
    for (int o = 0; o < c_out; ++o)
        for (int r2 = 0; r2 < r2n; ++r2)
            for (int s2 = 0; s2 < s2n; ++s2) {
                uint16_t* d = dst + ((size_t(o) * r2n + r2) * s2n + s2) * 16;
                for (int j = 0; j < 16; ++j) d[j] = 0;
                for (int dd = 0; dd < 4; ++dd) {
                    const int r = 2 * r2 + (dd >> 1), sx = 2 * s2 + (dd & 1);
                    if (r >= k_h || sx >= k_w) continue;
                    for (int c = 0; c < c_in; ++c) d[dd * c_in + c] = onnx::float_to_half(w[((size_t(o) * k_h + r) * k_w + sx) * c_in + c]);
                }
            }
}
void pack_weights_depthwise(const float* w, int c, int k_h, int k_w, int c_pitch, uint16_t* dst) {
    const int taps = k_h * k_w;  // w: OHWI with I = 1 -> [c][taps]
    for (int t = 0; t < taps; ++t)
        for (int ch = 0; ch < c_pitch; ++ch) dst[size_t(t) * c_pitch + ch] = ch < c ? onnx::float_to_half(w[size_t(ch) * taps + t]) : uint16_t(0);
}

// Phase-folded convolution (Filter::phase_fold = F, 2 or 4): w is OHWI [co][R][S][c].  The folded problem has F^2 * c_out output
// channels (ey * F + ex) * c_out + co, taps(R) x taps(S) taps with taps(k) = (k + 2F - 2) / F, and F^2 * cp input channels
// (dy * F + dx) * cp + c:
//   W'[(ey, ex, co)][r2][s2][(dy, dx, c)] = w[co][F * r2 + dy - ey][F * s2 + dx - ex][c]   (zero outside the filter)
void pack_weights_phase(const float* w, int c_out, int c_in, int k_h, int k_w, int cp, int fold, uint16_t* dst) {
    const int r2n = (k_h + 2 * fold - 2) / fold, s2n = (k_w + 2 * fold - 2) / fold, ff = fold * fold;
    const size_t row = size_t(r2n) * s2n * ff * cp;
    for (size_t i = 0; i < size_t(ff * c_out) * row; ++i) dst[i] = 0;
    for (int ey = 0; ey < fold; ++ey)
        for (int ex = 0; ex < fold; ++ex)
            for (int co = 0; co < c_out; ++co) {
                uint16_t* d0 = dst + size_t((ey * fold + ex) * c_out + co) * row;
                for (int r2 = 0; r2 < r2n; ++r2)
                    for (int s2 = 0; s2 < s2n; ++s2)
                        for (int dy = 0; dy < fold; ++dy)
                            for (int dx = 0; dx < fold; ++dx) {
                                const int r = fold * r2 + dy - ey, sx = fold * s2 + dx - ex;
                                if (r < 0 || r >= k_h || sx < 0 || sx >= k_w) continue;
                                uint16_t* d = d0 + (size_t(r2) * s2n + s2) * ff * cp + (dy * fold + dx) * cp;
                                const float* src = w + ((size_t(co) * k_h + r) * k_w + sx) * c_in;
                                for (int c = 0; c < c_in; ++c) d[c] = onnx::float_to_half(src[c]);
                            }
            }
}

// Upsample-folded convolution (Filter::upfold): w is OHWI [co][3][3][c] of a 3x3 convolution that reads the reflect-padded (1) nearest
// x2 upsampling of an image.  Output pixel (2y + ey, 2x + ex) reads upsampled rows 2y + ey + r - 1, i.e. low-resolution rows
// y + floor((ey + r - 1) / 2): tap r of phase ey lands on tap (ey + r + 1) / 2 of a 3x3 filter over the edge-padded (1) low-resolution
// image (reflection of the upsampled border row -1 is row 1 = low-resolution row 0 = edge padding).  Taps that meet are summed in fp32:
//   W'[(ey, ex, co)][r'][s'][c] = sum over r, s with (ey + r + 1) / 2 == r', (ex + s + 1) / 2 == s' of w[co][r][s][c]
void pack_weights_upfold(const float* w, int c_out, int c_in, int cp, uint16_t* dst) {
    const size_t row = size_t(9) * cp;
    std::vector<float> acc(row);
    for (int ey = 0; ey < 2; ++ey)
        for (int ex = 0; ex < 2; ++ex)
            for (int co = 0; co < c_out; ++co) {
                std::fill(acc.begin(), acc.end(), 0.f);
                for (int r = 0; r < 3; ++r)
                    for (int sx = 0; sx < 3; ++sx) {
                        float* d = acc.data() + (size_t((ey + r + 1) / 2) * 3 + (ex + sx + 1) / 2) * cp;
                        const float* src = w + ((size_t(co) * 3 + r) * 3 + sx) * c_in;
                        for (int c = 0; c < c_in; ++c) d[c] += src[c];
                    }
                uint16_t* d0 = dst + size_t((ey * 2 + ex) * c_out + co) * row;
                for (size_t i = 0; i < row; ++i) d0[i] = onnx::float_to_half(acc[i]);
            }
}

// Width-folded convolution (Filter::wfold = F): F horizontally neighbouring pixels are one pixel of F * cp channels (dx * cp + c), the F
// outputs computed from a folded pixel are F * c_out GEMM columns (ex * c_out + co), the filter has ceil((k_w + F - 1) / F) horizontal taps:
//   W'[(ex, co)][r][s2][(dx, c)] = w[co][r][F * s2 + dx - ex][c]   (zero outside the filter)
void pack_weights_wfold(const float* w, int c_out, int c_in, int k_h, int k_w, int cp, int fold, uint16_t* dst) {
    const int s2n = (k_w + 2 * fold - 2) / fold;
    const size_t row = size_t(k_h) * s2n * fold * cp;
    for (size_t i = 0; i < size_t(fold) * c_out * row; ++i) dst[i] = 0;
    for (int ex = 0; ex < fold; ++ex)
        for (int co = 0; co < c_out; ++co) {
            uint16_t* d0 = dst + size_t(ex * c_out + co) * row;
            for (int r = 0; r < k_h; ++r)
                for (int s2 = 0; s2 < s2n; ++s2)
                    for (int dx = 0; dx < fold; ++dx) {
                        const int sx = fold * s2 + dx - ex;
                        if (sx < 0 || sx >= k_w) continue;
                        uint16_t* d = d0 + ((size_t(r) * s2n + s2) * fold + dx) * cp;
                        const float* src = w + ((size_t(co) * k_h + r) * k_w + sx) * c_in;
                        for (int c = 0; c < c_in; ++c) d[c] = onnx::float_to_half(src[c]);
                    }
        }
}

int pick_conv_mode(int c_in, int c_out, int groups, int k_h, int k_w, int stride_h, int stride_w, int dil_w, const int pads[4]) {
    if (groups != 1) {
        if (groups == c_in && groups == c_out) return 4;  // depthwise, multiplier 1 (Converters.swift:57)
        return -1;
    }
    if (c_in <= 8 && dil_w == 1 && k_w * 8 <= 256 && k_h * k_w > 1) return k::CONV_MODE_PACKED_ROW;
    // Narrow inputs (channel pitch 16..32) without padding of their own (the graph pads explicitly, e.g. reflection padding in front of
    // every TransformerNet convolution): a filter row of S taps is S x pitch contiguous NHWC elements, so one K block can span several
    // taps -- R x ceil(S x pitch / 64) dense k-blocks instead of R x S half-empty ones (9x9 on 32 channels: 45 instead of 81).
    const int pitch = (c_in + 7) / 8 * 8;
    if (pitch <= 32 && dil_w == 1 && k_w >= 3 && !pads[0] && !pads[1] && !pads[2] && !pads[3] && k_w * pitch <= 1024 &&
        k_h * ((k_w * pitch + 63) / 64) < k_h * k_w)
        return k::CONV_MODE_PACKED_ROW;
    if (k_h == 1 && k_w == 1 && stride_h == 1 && stride_w == 1 && !pads[0] && !pads[1] && !pads[2] && !pads[3]) return k::CONV_MODE_TILED;
    return k::CONV_MODE_IM2COL;
}

// ---- ONNXGraph ------------------------------------------------------------------------------------------------

// NVTX range "opType:outputName" around every plan step (the reference's only tracing hook is the node label,
// Converters.swift:931,1064).  Header-only NVTX v3: a no-op unless a tool (nsys / ncu --nvtx) is attached.  Steps that are replayed
// from the captured CUDA graph carry their range on the capture pass; SMELTER_DEBUG_SYNC=1 / useCudaGraph=0 give per-launch ranges.
struct NvtxRange {
    explicit NvtxRange(const std::string& name) { nvtxRangePushA(name.c_str()); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Step {
    std::function<cudaError_t(cudaStream_t)> run;
    std::string desc;
    double flops = 0, bytes = 0;
    int launches = 1;      // device kernels (or memset/memcpy nodes) this step enqueues
    bool tensor = false;   // the tcgen05 implicit-GEMM kernel
    bool boundary = false; // reads a caller-owned source image: launched eagerly ahead of the captured CUDA graph,
                           // so the graph never bakes in a caller pointer and is captured once per batch size
};

struct Plan {
    int batch = 0;
    std::vector<Step> steps;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<__half**> src_slots;          // where each graph input's current NCHW device pointer is read from
    std::vector<std::unique_ptr<__half*>> src_ptr_storage;
    std::vector<ImageShape> src_shapes;
    Tensor result;
    cudaGraphExec_t exec = nullptr;
    void* counters = nullptr;  // split-K arrival counters and instance-norm accumulators of every conv in the plan (zero between encodes)
    size_t counter_bytes = 0;
    std::vector<__half*> resized;  // per graph input: NCHW staging the source is resized into under .forceInputScale (lazily allocated)
    std::vector<void*> blobs;  // further device allocations owned by the plan (none today)
    ~Plan() {
        if (exec) cudaGraphExecDestroy(exec);
        if (arena) cudaFree(arena);
        if (counters) cudaFree(counters);
        for (void* b : blobs) cudaFree(b);
    }
};

ONNXGraph::~ONNXGraph() {
    plans_.clear();
    if (weight_arena_) cudaFree(weight_arena_);
}

int ONNXGraph::init(const uint8_t* data, size_t len, const smelter_config& cfg) {
    if (!data || !len) return fail(SMELTER_ERR_INVALID_ARGUMENT, "empty model data");
    bytes_.assign(data, data + len);
    std::string perr;
    if (!onnx::parse_model(bytes_.data(), bytes_.size(), &model_, &perr)) return fail(SMELTER_ERR_PARSE, perr);  // ONNXGraph.swift:96
    format_ = model_.producer_name == "ONNX2MPS" ? SMELTER_FORMAT_MPS_FLAVOR : SMELTER_FORMAT_ONNX;             // :98-103
    cfg_ = cfg;
    for (const auto& t : model_.graph.initializer) tensors_[t.name] = &t;                                         // :106-108
    registerBuiltins();                                                                                          // :110-155
    return SMELTER_OK;
}

int ONNXGraph::output(const std::string& name) const {
    auto it = outputs_.find(name);
    return it == outputs_.end() ? -1 : it->second;
}
const ImageShape* ONNXGraph::shape(const std::string& name) const {
    auto it = outputs_.find(name);
    return it == outputs_.end() ? nullptr : &values_[size_t(it->second)].shape;
}
const onnx::TensorProto* ONNXGraph::tensor(const std::string& name) const {
    auto it = tensors_.find(name);
    return it == tensors_.end() ? nullptr : it->second;
}

int ONNXGraph::addFilter(Filter&& f, const ImageShape& shape, const std::vector<std::string>& outputs) {
    if (outputs.empty()) return fail(SMELTER_ERR_INCONSISTENT_STATE, "filter without outputs");
    Value v;
    v.name = outputs[0];
    v.shape = shape;
    values_.push_back(v);
    const int id = int(values_.size()) - 1;
    f.out = id;
    filters_.push_back(std::move(f));
    for (const auto& o : outputs) outputs_[o] = id;  // ONNXGraph.swift:269-272: every output name maps to the same image
    return SMELTER_OK;
}

int ONNXGraph::addAlias(int value, const ImageShape& shape, const std::vector<std::string>& outputs) {
    if (outputs.empty()) return fail(SMELTER_ERR_INCONSISTENT_STATE, "alias without outputs");
    Value v;
    v.name = outputs[0];
    v.shape = shape;
    v.alias_of = value;
    values_.push_back(v);
    const int id = int(values_.size()) - 1;
    for (const auto& o : outputs) outputs_[o] = id;
    return SMELTER_OK;
}

// ONNXGraph.swift:197-251
int ONNXGraph::initOutputs() {
    for (const auto& vi : model_.graph.input) {
        if (tensor(vi.name)) continue;  // :199 initializers listed as inputs are not images
        std::vector<int64_t> dims = vi.dims;
        for (int i = 0; i < cfg_.n_dims && i < 8; ++i) {  // :200-202 Configuration.dims overrides by axis index
            const int axis = cfg_.dims_axis[i];
            if (axis >= 0 && size_t(axis) < dims.size()) dims[size_t(axis)] = cfg_.dims_value[i];
        }
        ImageShape s;
        if (dims.size() == 3) { s.c = int(dims[0]); s.h = int(dims[1]); s.w = int(dims[2]); }        // :205-208
        else if (dims.size() == 4) { s.c = int(dims[1]); s.h = int(dims[2]); s.w = int(dims[3]); }   // :209-212 (N dropped)
        else return fail(SMELTER_ERR_UNSUPPORTED_INPUT, "graph input '" + vi.name + "' must have rank 3 or 4");  // :213-214
        if (s.c <= 0 || s.h <= 0 || s.w <= 0) return fail(SMELTER_ERR_UNSUPPORTED_INPUT, "graph input '" + vi.name + "' has unknown dims");
        // .forceInputScale (:219-241): sources of any H x W are resized to the graph input at encode time (ONNXGraph::encode)
        Value v;
        v.name = vi.name;
        v.shape = s;
        v.is_input = true;
        values_.push_back(v);
        outputs_[vi.name] = int(values_.size()) - 1;
        input_values_.push_back(int(values_.size()) - 1);
    }
    return SMELTER_OK;
}

int ONNXGraph::consumers_of(int value) const {
    int n = 0;
    for (const auto& f : filters_) {
        if (f.removed) continue;
        for (int i : f.in) n += (i == value);
        n += (f.residual == value);
    }
    for (const auto& v : values_) n += (v.alias_of == value) * 2;  // aliased values are never fused through
    if (value == output_value_) n += 2;
    return n;
}

// Fusion (what ONNX2MPS.py's fuse_bn_into_conv does offline, ONNX2MPS.py:104-109, plus epilogue fusion that
// MPSNNGraph performs internally): Conv+BN fold, Conv(+Add)+activation, norm/add + ReLU.
void ONNXGraph::fuse() {
    auto producer = [&](int value) -> Filter* {
        for (auto& f : filters_)
            if (!f.removed && f.out == value) return &f;
        return nullptr;
    };
    auto redirect = [&](int from, int to) {  // every reader of `from` now reads `to`
        for (auto& f : filters_) {
            for (int& i : f.in) if (i == from) i = to;
            if (f.residual == from) f.residual = to;
        }
        for (auto& v : values_) if (v.alias_of == from) v.alias_of = to;
        if (output_value_ == from) output_value_ = to;
    };
    // 1. BatchNorm directly after a Conv that has no other reader and no fused epilogue yet.
    for (auto& bn : filters_) {
        if (bn.removed || bn.kind != FilterKind::BatchNorm) continue;
        Filter* conv = producer(bn.in[0]);
        if (!conv || conv->kind != FilterKind::Conv || conv->act != k::ACT_NONE || conv->residual >= 0) continue;
        if (consumers_of(conv->out) != 1) continue;
        const size_t per_out = conv->w.size() / size_t(conv->c_out);
        for (int o = 0; o < conv->c_out; ++o) {  // W' = W * s ; b' = b * s + shift
            const float s = bn.p0[size_t(o)];
            float* w = conv->w.data() + size_t(o) * per_out;
            for (size_t i = 0; i < per_out; ++i) w[i] *= s;
            conv->bias[size_t(o)] = conv->bias[size_t(o)] * s + bn.p1[size_t(o)];
        }
        bn.removed = true;
        redirect(bn.out, conv->out);
    }
    // 2. Conv -> Add(residual) : the other operand must already exist when the conv runs.
    for (size_t ai = 0; ai < filters_.size(); ++ai) {
        Filter& add = filters_[ai];
        if (add.removed || add.kind != FilterKind::Binary || add.sub != k::BIN_ADD || add.act != k::ACT_NONE || add.bcast) continue;
        for (int side = 0; side < 2; ++side) {
            Filter* conv = producer(add.in[size_t(side)]);
            const int other = add.in[size_t(1 - side)];
            if (!conv || conv->kind != FilterKind::Conv || conv->act != k::ACT_NONE || conv->residual >= 0) continue;
            if (conv->groups != 1 || consumers_of(conv->out) != 1 || other == conv->out) continue;
            // `other` must be produced before the conv in filter order (or be a graph input)
            size_t conv_idx = size_t(conv - filters_.data());
            int root = other;
            while (values_[size_t(root)].alias_of >= 0) root = values_[size_t(root)].alias_of;
            bool ok = values_[size_t(root)].is_input;
            for (size_t j = 0; j < conv_idx && !ok; ++j) ok = !filters_[j].removed && filters_[j].out == root;
            if (!ok) continue;
            conv->residual = other;
            add.removed = true;
            redirect(add.out, conv->out);
            break;
        }
    }
    // 3. trailing activation into Conv / BatchNorm / InstanceNorm / Add.
    for (auto& act : filters_) {
        if (act.removed || act.kind != FilterKind::Unary) continue;
        if (act.sub != k::UN_RELU && act.sub != k::UN_CLIP && act.sub != k::UN_SIGMOID) continue;
        Filter* p = producer(act.in[0]);
        if (!p || p->act != k::ACT_NONE || consumers_of(p->out) != 1) continue;
        if (p->kind == FilterKind::Conv) {
            p->act = act.sub == k::UN_RELU ? k::ACT_RELU : (act.sub == k::UN_CLIP ? k::ACT_CLIP : k::ACT_SIGMOID);
            p->clip_lo = act.alpha; p->clip_hi = act.beta;
        } else if ((p->kind == FilterKind::BatchNorm || p->kind == FilterKind::InstanceNorm ||
                    (p->kind == FilterKind::Binary && p->sub == k::BIN_ADD)) && act.sub == k::UN_RELU) {
            p->act = k::ACT_RELU;
        } else {
            continue;
        }
        act.removed = true;
        redirect(act.out, p->out);
    }
}

int ONNXGraph::build() {
    if (built_) return SMELTER_OK;
    int rc = initOutputs();  // ONNXGraph.swift:170
    if (rc) return rc;
    for (int i = 0; i < num_nodes(); ++i) {  // :172-176 file order = topological order
        const auto& n = node(i);
        auto it = converters_.find(n.op_type);
        if (it == converters_.end()) return fail(SMELTER_ERR_UNKNOWN_NODE_OP_TYPE, n.op_type);
        rc = it->second(*this, i);
        if (rc) return rc;
    }
    if (model_.graph.output.size() != 1) return fail(SMELTER_ERR_UNSUPPORTED_OUTPUT, "exactly one graph output is supported");  // :178-180
    output_value_ = output(model_.graph.output[0].name);
    if (output_value_ < 0) return fail(SMELTER_ERR_NO_SUCH_OUTPUT, "graph output '" + model_.graph.output[0].name + "' was not produced");  // :182-183
    if (input_values_.empty()) return fail(SMELTER_ERR_UNSUPPORTED_INPUT, "graph has no image input");
    if (cfg_.enable_fusion) fuse();
    for (auto& f : filters_) {
        if (f.removed || f.kind != FilterKind::Conv) continue;
        // Grouped convolution other than depthwise (Converters.swift:57-75 hands MPS the grouped descriptor): run as the dense
        // convolution with a block-diagonal weight tensor -- OHWI [Cout][R][S][Cin] with zeros outside each output channel's group.
        // `groups` times the arithmetic of the grouped form on the tensor cores, no kernel of its own; exact (the zeros add nothing).
        if (f.groups > 1 && !f.transposed && !(f.groups == f.c_in_g * f.groups && f.groups == f.c_out && f.c_in_g == 1)) {
            const int G = f.groups, cig = f.c_in_g, cin = cig * G, taps = f.k_h * f.k_w;
            if (f.c_out % G) return fail(SMELTER_ERR_INCONSISTENT_STATE, "output channels not divisible by group (" + std::to_string(G) + ")");
            if (f.w.size() != size_t(f.c_out) * taps * cig) return fail(SMELTER_ERR_INCONSISTENT_STATE, "grouped convolution: weight size mismatch");
            const int cog = f.c_out / G;
            std::vector<float> dense(size_t(f.c_out) * taps * cin, 0.f);
            for (int co = 0; co < f.c_out; ++co)
                for (int t = 0; t < taps; ++t)
                    std::copy_n(&f.w[(size_t(co) * taps + t) * cig], cig, &dense[(size_t(co) * taps + t) * cin + size_t(co / cog) * cig]);
            f.w.swap(dense);
            f.c_in_g = cin;
            f.group_expanded = G;
            f.groups = 1;
        }
        f.conv_mode = pick_conv_mode(f.c_in_g * f.groups, f.c_out, f.groups, f.k_h, f.k_w, f.stride_h, f.stride_w, f.dil_w, f.pads);
        if (f.transposed) f.conv_mode = (f.k_h == 1 && f.k_w == 1) ? k::CONV_MODE_TILED : k::CONV_MODE_IM2COL;  // reads a materialised image
        if (f.conv_mode < 0)
            return fail(SMELTER_ERR_UNSUPPORTED, "grouped convolution other than depthwise (groups=" + std::to_string(f.groups) + ")");
    }
    // stride-2 stems that are the only reader of a graph input run on a space-to-depth image (see Filter::s2d)
    if (!getenv("SMELTER_NO_S2D")) {
        auto root_of = [&](int v) { while (values_[size_t(v)].alias_of >= 0) v = values_[size_t(v)].alias_of; return v; };
        for (auto& f : filters_) {
            if (f.removed || f.kind != FilterKind::Conv || f.conv_mode != k::CONV_MODE_PACKED_ROW) continue;
            const int v = root_of(f.in[0]);
            if (!values_[size_t(v)].is_input || v == root_of(output_value_)) continue;
            int readers = 0;
            for (const auto& g : filters_) {
                if (g.removed) continue;
                for (int i : g.in) if (root_of(i) == v) ++readers;
                if (g.residual >= 0 && root_of(g.residual) == v) ++readers;
            }
            const int c_in = f.c_in_g * f.groups;
            if (readers != 1 || f.stride_h != 2 || f.stride_w != 2 || f.dil_h != 1 || f.dil_w != 1 || c_in > 4 || (f.k_w + 1) / 2 * 16 > 256) continue;
            const ImageShape& is = values_[size_t(v)].shape;
            int hp = is.h + f.pads[0] + f.pads[2], wp = is.w + f.pads[1] + f.pads[3];
            const int p_out = (hp - f.k_h) / 2 + 1, q_out = (wp - f.k_w) / 2 + 1;
            hp += hp & 1; wp += wp & 1;  // one more zero row / column makes the padded image foldable
            if (hp / 2 - (f.k_h + 1) / 2 + 1 != p_out || wp / 2 - (f.k_w + 1) / 2 + 1 != q_out) continue;  // folded conv must give the same output size
            f.s2d = true;
            f.s2d_h = hp; f.s2d_w = wp;
        }
    }
    // Upsample-folded convolutions (Filter::upfold, engine.h): nearest Upsample x2 -> reflect Pad 1 -> Conv 3x3 -> InstanceNorm with
    // single readers all along.  The Upsample goes, the Pad becomes an edge pad of the low-resolution image, the convolution gets
    // 4 * Cout phase columns, the normalisation un-permutes its pixels on the way out.
    if (!getenv("SMELTER_NO_UPSAMPLE_FOLD")) {
        auto root_of = [&](int v) { while (values_[size_t(v)].alias_of >= 0) v = values_[size_t(v)].alias_of; return v; };
        auto producer_of = [&](int v) -> Filter* {
            for (auto& g : filters_) if (!g.removed && g.out >= 0 && root_of(g.out) == root_of(v)) return &g;
            return nullptr;
        };
        for (auto& f : filters_) {
            if (f.removed || f.kind != FilterKind::Conv || f.is_gemm || f.transposed || f.s2d || f.groups != 1 || f.residual >= 0) continue;
            if (f.stride_h != 1 || f.stride_w != 1 || f.dil_h != 1 || f.dil_w != 1 || f.pads[0] || f.pads[1] || f.pads[2] || f.pads[3]) continue;
            if (f.k_h != 3 || f.k_w != 3 || f.c_out % 8 || 4 * f.c_out > 1024 || f.c_in_g < 16) continue;
            if (root_of(f.out) == root_of(output_value_) || consumers_of(f.out) != 1) continue;
            Filter* pad = producer_of(f.in[0]);
            if (!pad || pad->kind != FilterKind::Pad || pad->sub != k::PAD_REFLECT || consumers_of(pad->out) != 1) continue;
            if (pad->pads[0] != 1 || pad->pads[1] != 1 || pad->pads[2] != 1 || pad->pads[3] != 1) continue;
            Filter* up = producer_of(pad->in[0]);
            if (!up || up->kind != FilterKind::Upsample || up->sub != k::UP_NEAREST || up->scale_h != 2 || up->scale_w != 2 || consumers_of(up->out) != 1) continue;
            if (root_of(up->out) == root_of(output_value_) || root_of(pad->out) == root_of(output_value_)) continue;
            Filter* norm = nullptr;
            for (auto& g : filters_) if (!g.removed && g.kind == FilterKind::InstanceNorm && !g.in.empty() && root_of(g.in[0]) == root_of(f.out)) norm = &g;
            if (!norm) continue;
            const ImageShape low = values_[size_t(up->in[0])].shape;
            if (low.h < 2 || low.w < 2) continue;
            up->removed = true;
            pad->in[0] = up->in[0];
            pad->sub = k::PAD_EDGE;
            values_[size_t(pad->out)].shape.h = low.h + 2;
            values_[size_t(pad->out)].shape.w = low.w + 2;
            f.upfold = true;
            norm->unfold_w = low.w;
            std::vector<float> b4(size_t(4) * f.c_out);
            for (int ph = 0; ph < 4; ++ph) for (int co = 0; co < f.c_out; ++co) b4[size_t(ph) * f.c_out + co] = f.bias[size_t(co)];
            f.bias.swap(b4);
        }
    }
    // Phase-folded output convolutions: a stride-1, unpadded k x k convolution (k odd >= 5) with at most 8 output channels and at
    // most 32 input channels that produces the graph output and reads a Pad nobody else reads -- TransformerNet's 9x9 32 -> 3 layer,
    // 81 half-empty k-blocks per tile as im2col, 45 as packed rows, all bound by the TMA's pixel-row rate with 3 useful GEMM columns.
    // On the 2 x 2 space-to-depth fold of the padded image the four output phases of a folded pixel are 4 * c_out columns of one GEMM
    // over ceil(k / 2)^2 taps of 4 * Cin dense channels: 50 k-blocks for a quarter of the rows (12.5 instead of 45 per output pixel).
    if (!getenv("SMELTER_NO_PHASE_FOLD")) {
        auto root_of = [&](int v) { while (values_[size_t(v)].alias_of >= 0) v = values_[size_t(v)].alias_of; return v; };
        for (auto& f : filters_) {
            if (f.removed || f.kind != FilterKind::Conv || f.is_gemm || f.transposed || f.s2d || f.upfold || f.groups != 1 || f.residual >= 0) continue;
            if (f.stride_h != 1 || f.stride_w != 1 || f.dil_h != 1 || f.dil_w != 1 || f.pads[0] || f.pads[1] || f.pads[2] || f.pads[3]) continue;
            if (f.k_h != f.k_w || f.k_h < 5 || !(f.k_h & 1) || f.c_out > 8 || round_up(f.c_in_g, 8) > 32 || round_up(f.c_in_g, 8) < 16) continue;
            if (root_of(f.out) != root_of(output_value_)) continue;
            Filter* pad = nullptr;
            for (auto& g : filters_) if (!g.removed && g.out == f.in[0] && g.kind == FilterKind::Pad) pad = &g;
            if (!pad || consumers_of(pad->out) != 1) continue;
            const ImageShape& ps = values_[size_t(pad->out)].shape;
            const ImageShape& os = values_[size_t(f.out)].shape;
            if ((ps.h | ps.w) & 1 || (os.h | os.w) & 1 || (os.h == 1 && os.w == 1)) continue;
            // fold by 4 where the sizes divide: operand bytes per output pixel go with taps^2 * Cin, (k + 2F - 2) / F taps per axis --
            // 25 -> 9 for k = 9 -- and the 16 * c_out GEMM columns still fit one narrow tile
            const bool by4 = !getenv("SMELTER_PHASE_FOLD_2") && (ps.h | ps.w | os.h | os.w) % 4 == 0 && 16 * f.c_out <= 64 && f.k_h >= 7;
            const int fold = by4 ? 4 : 2, ff = fold * fold;
            f.phase_fold = fold;
            f.conv_mode = k::CONV_MODE_IM2COL;
            pad->s2d_out = fold;
            std::vector<float> b4(size_t(ff) * f.c_out);
            for (int ph = 0; ph < ff; ++ph) for (int co = 0; co < f.c_out; ++co) b4[size_t(ph) * f.c_out + co] = f.bias[size_t(co)];
            f.bias.swap(b4);
        }
    }
    // Input-folded convolutions (Filter::in_fold, engine.h): graph input -> Pad -> Conv k x k (k = 1 mod 4, stride 1) -> InstanceNorm
    // with single readers.  The Pad goes into the boundary conversion, which writes the 4 x 4 space-to-depth fold.
    if (!getenv("SMELTER_NO_INPUT_FOLD")) {
        for (auto& f : filters_) {
            if (f.removed || f.kind != FilterKind::Conv || f.is_gemm || f.transposed || f.s2d || f.phase_fold || f.upfold || f.groups != 1 || f.residual >= 0) continue;
            if (f.stride_h != 1 || f.stride_w != 1 || f.dil_h != 1 || f.dil_w != 1 || f.pads[0] || f.pads[1] || f.pads[2] || f.pads[3]) continue;
            if (f.k_h != f.k_w || f.k_h < 5 || f.k_h % 4 != 1 || f.c_in_g > 4 || f.c_out % 8 || 16 * f.c_out > 1024) continue;
            if (f.out == output_value_ || consumers_of(f.out) != 1) continue;
            Filter* pad = nullptr;
            for (auto& g : filters_) if (!g.removed && g.out == f.in[0] && g.kind == FilterKind::Pad) pad = &g;
            if (!pad || pad->s2d_out || pad->pad_add || consumers_of(pad->out) != 1) continue;
            const int v = pad->in[0];
            if (!values_[size_t(v)].is_input || consumers_of(v) != 1 || v == output_value_) continue;
            Filter* norm = nullptr;
            for (auto& g : filters_) if (!g.removed && g.kind == FilterKind::InstanceNorm && !g.in.empty() && g.in[0] == f.out) norm = &g;
            if (!norm || norm->unfold_w) continue;
            const ImageShape& ps = values_[size_t(pad->out)].shape;
            const ImageShape& os = values_[size_t(f.out)].shape;
            if ((ps.h | ps.w | os.h | os.w) % 4) continue;
            f.in_fold = true;
            for (int i = 0; i < 4; ++i) f.in_fold_pad[i] = pad->pads[i];
            f.in_fold_mode = pad->sub;
            f.in_fold_value = pad->alpha;
            f.in[0] = v;
            f.conv_mode = k::CONV_MODE_IM2COL;
            pad->removed = true;
            norm->unfold_w = os.w / 4;
            norm->unfold_f = 4;
            std::vector<float> bf(size_t(16) * f.c_out);
            for (int ph = 0; ph < 16; ++ph) for (int co = 0; co < f.c_out; ++co) bf[size_t(ph) * f.c_out + co] = f.bias[size_t(co)];
            f.bias.swap(bf);
        }
    }
    // Width-folded input convolutions: a stride-1, unpadded convolution on at most 32 input channels whose packed filter rows leave
    // most of every 128-byte TMA pixel row empty (9x9 on 3 channels: 18 k-blocks per 128 output pixels, 72 of 128 bytes used in one
    // half of them and 16 in the other) reads and writes the SAME NHWC buffers re-interpreted with F = 64 / pitch neighbouring pixels
    // as one: [H, W / F, 64] -> [P, Q / F, F * Cout], ceil((k_w + F - 1) / F) horizontal taps of 64 dense channels.  Pixel rows per output
    // pixel: k_h * ceil(k_w * pitch / 64) -> k_h * taps / F (18 -> 2.25 for TransformerNet's input layer); the zero weights it multiplies
    // cost tensor time the layer has to spare.  No kernel and no layout change: weights, bias and the problem's dimensions only.
    if (!getenv("SMELTER_NO_WIDTH_FOLD")) {
        for (auto& f : filters_) {
            if (f.removed || f.kind != FilterKind::Conv || f.is_gemm || f.transposed || f.s2d || f.phase_fold || f.upfold || f.in_fold || f.groups != 1 || f.residual >= 0) continue;
            if (f.stride_h != 1 || f.stride_w != 1 || f.dil_h != 1 || f.dil_w != 1 || f.pads[0] || f.pads[1] || f.pads[2] || f.pads[3]) continue;
            const int cp = round_up(f.c_in_g, 8);
            if (cp > 32 || f.k_w < 3 || f.c_out % 8) continue;
            const int fold = 64 / cp;  // 8, 4 (pitch 16) or 2 (pitch 24 / 32: 48 / 64 channels per folded pixel)
            const ImageShape& is = values_[size_t(f.in[0])].shape;
            const ImageShape& os = values_[size_t(f.out)].shape;
            if (round_up(is.c, 8) != cp || is.w % fold || os.w % fold || fold * f.c_out > 512) continue;
            const int taps = (f.k_w + 2 * fold - 2) / fold;
            const int rows_before = f.k_h * ((f.k_w * cp + 63) / 64) * fold, rows_after = f.k_h * taps * ((fold * cp + 63) / 64);
            if (rows_after * 2 > rows_before) continue;  // worth it only with at least half of the pixel rows gone
            f.wfold = fold;
            f.conv_mode = k::CONV_MODE_IM2COL;
            std::vector<float> bf(size_t(fold) * f.c_out);
            for (int ex = 0; ex < fold; ++ex) for (int co = 0; co < f.c_out; ++co) bf[size_t(ex) * f.c_out + co] = f.bias[size_t(co)];
            f.bias.swap(bf);
        }
    }
    // A residual Add in front of a Pad runs inside the Pad kernel (Filter::pad_add): one launch and one read of the sum less.  The
    // plain sum is stored too while it has other readers, all of which must come after the Pad in execution order.
    if (!getenv("SMELTER_NO_PAD_ADD")) {
        for (size_t pi = 0; pi < filters_.size(); ++pi) {
            Filter& pd = filters_[pi];
            if (pd.removed || pd.kind != FilterKind::Pad || pd.s2d_out || pd.pad_add || pd.in.size() != 1) continue;
            Filter* ad = nullptr;
            for (auto& g : filters_) if (!g.removed && g.out == pd.in[0] && g.kind == FilterKind::Binary && g.sub == k::BIN_ADD) ad = &g;
            if (!ad || ad->in.size() != 2 || ad->out == output_value_ || (ad->act != k::ACT_NONE && ad->act != k::ACT_RELU)) continue;
            const ImageShape& sa = values_[size_t(ad->in[0])].shape;
            const ImageShape& sb = values_[size_t(ad->in[1])].shape;
            const ImageShape& so = values_[size_t(ad->out)].shape;
            auto same = [](const ImageShape& p, const ImageShape& q) { return p.c == q.c && p.h == q.h && p.w == q.w; };
            if (!same(sa, so) || !same(sb, so)) continue;  // no broadcasting
            bool ok = true;
            int others = 0;
            for (size_t gi = 0; gi < filters_.size(); ++gi) {
                const Filter& g = filters_[gi];
                if (g.removed || gi == pi) continue;
                bool reads = g.residual == ad->out;
                for (int i : g.in) reads = reads || i == ad->out;
                if (reads) { ++others; ok = ok && gi > pi; }
            }
            for (const auto& v : values_) ok = ok && v.alias_of != ad->out;
            if (!ok) continue;
            pd.in = {ad->in[0], ad->in[1]};
            pd.act = ad->act;
            pd.out2 = others ? ad->out : -1;
            pd.pad_add = true;
            ad->removed = true;
        }
    }
    // Reflection pads behind an instance norm are written by the norm (Filter::norm_pad): one launch and one round trip of the
    // tensor less per Pad.  The apply pass stores the interior at its padded position; the border pixels (a few hundred to a few
    // thousand) are extra tasks that re-read their mirror source -- not a per-pixel test on the latency-bound apply loop.
    if (!getenv("SMELTER_NO_NORM_PAD")) {
        for (size_t pi = 0; pi < filters_.size(); ++pi) {
            Filter& pd = filters_[pi];
            if (pd.removed || pd.kind != FilterKind::Pad || (pd.sub != k::PAD_REFLECT && pd.sub != k::PAD_EDGE) || pd.pad_add) continue;
            Filter* nm = nullptr;
            for (auto& g : filters_) if (!g.removed && g.out == pd.in[0] && g.kind == FilterKind::InstanceNorm) nm = &g;
            if (!nm || nm->norm_padded || nm->out == output_value_) continue;
            // other readers of the norm's result (a skip connection) get the plain image as a second store; they must run after the Pad
            bool ok = true;
            int others = 0;
            for (size_t gi = 0; gi < filters_.size(); ++gi) {
                const Filter& g = filters_[gi];
                if (g.removed || gi == pi) continue;
                bool reads = g.residual == nm->out;
                for (int i : g.in) reads = reads || i == nm->out;
                if (reads) { ++others; ok = ok && gi > pi; }
            }
            for (const auto& v : values_) ok = ok && v.alias_of != nm->out;
            if (!ok || others > 1) continue;
            const ImageShape& s = values_[size_t(nm->out)].shape;
            if (pd.sub == k::PAD_REFLECT && (pd.pads[0] >= s.h || pd.pads[2] >= s.h || pd.pads[1] >= s.w || pd.pads[3] >= s.w)) continue;
            for (int i = 0; i < 4; ++i) nm->norm_pad[i] = pd.pads[i];
            nm->norm_s2d = pd.s2d_out;
            nm->norm_pad_mode = pd.sub;
            nm->norm_padded = true;
            nm->out2 = others ? nm->out : -1;
            nm->out = pd.out;
            pd.removed = true;
        }
    }
    rc = upload_weights();  // MPSNNGraph(device:resultImage:) pulls weights from the data sources (:185-190)
    if (rc) return fail(SMELTER_ERR_GRAPH_INTERNAL, "weight upload failed: " + last_error_string());
    built_ = true;
    // host copies of the weights and the model bytes are no longer needed
    for (auto& f : filters_) { std::vector<float>().swap(f.w); }
    return SMELTER_OK;
}

int ONNXGraph::upload_weights() {
    SM_CUDA(cudaSetDevice(ctx_->device));
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t total = 0;
    for (auto& f : filters_) {
        if (f.removed) continue;
        if (f.kind == FilterKind::Conv) {
            const int c_in = f.c_in_g * f.groups;
            size_t wbytes;
            if (f.conv_mode == 4) wbytes = size_t(f.k_h) * f.k_w * round_up(f.c_out, 8) * 2;
            else if (f.s2d) wbytes = size_t(f.c_out) * ((f.k_h + 1) / 2) * ((f.k_w + 1) / 2) * 16 * 2;
            else if (f.phase_fold) {
                const int F = f.phase_fold;
                wbytes = size_t(F * F * f.c_out) * ((f.k_h + 2 * F - 2) / F) * ((f.k_w + 2 * F - 2) / F) * F * F * round_up(c_in, 8) * 2;
            }
            else if (f.upfold) wbytes = size_t(4 * f.c_out) * 9 * round_up(c_in, 8) * 2;
            else if (f.in_fold) wbytes = size_t(16 * f.c_out) * ((f.k_h + 6) / 4) * ((f.k_w + 6) / 4) * 64 * 2;
            else if (f.wfold) wbytes = size_t(f.wfold) * f.c_out * f.k_h * ((f.k_w + 2 * f.wfold - 2) / f.wfold) * f.wfold * round_up(c_in, 8) * 2;
            else wbytes = size_t(f.c_out) * f.k_h * f.k_w * round_up(c_in, 8) * 2;
            f.w_off = total; total = align(total + wbytes);
            f.bias_off = total; total = align(total + size_t(round_up(std::max<int>(f.c_out, int(f.bias.size())), 256)) * 4);
        } else if (f.kind == FilterKind::BatchNorm || f.kind == FilterKind::InstanceNorm) {
            const size_t n = size_t(round_up(int(f.p0.size()), 8)) * 4;
            f.p0_off = total; total = align(total + n);
            f.p1_off = total; total = align(total + n);
        }
    }
    weight_bytes_ = std::max<size_t>(total, 256);
    std::vector<uint8_t> host(weight_bytes_, 0);
    for (auto& f : filters_) {
        if (f.removed) continue;
        if (f.kind == FilterKind::Conv) {
            const int c_in = f.c_in_g * f.groups;
            uint16_t* w = reinterpret_cast<uint16_t*>(host.data() + f.w_off);
            if (f.conv_mode == 4) pack_weights_depthwise(f.w.data(), f.c_out, f.k_h, f.k_w, round_up(f.c_out, 8), w);
            else if (f.s2d) pack_weights_s2d(f.w.data(), f.c_out, c_in, f.k_h, f.k_w, w);
            else if (f.phase_fold) pack_weights_phase(f.w.data(), f.c_out, c_in, f.k_h, f.k_w, round_up(c_in, 8), f.phase_fold, w);
            else if (f.in_fold) pack_weights_phase(f.w.data(), f.c_out, c_in, f.k_h, f.k_w, 4, 4, w);
            else if (f.upfold) pack_weights_upfold(f.w.data(), f.c_out, c_in, round_up(c_in, 8), w);
            else if (f.wfold) pack_weights_wfold(f.w.data(), f.c_out, c_in, f.k_h, f.k_w, round_up(c_in, 8), f.wfold, w);
            else pack_weights_ohwi(f.w.data(), f.c_out, c_in, f.k_h, f.k_w, round_up(c_in, 8), w);
            memcpy(host.data() + f.bias_off, f.bias.data(), f.bias.size() * 4);
        } else if (f.kind == FilterKind::BatchNorm || f.kind == FilterKind::InstanceNorm) {
            memcpy(host.data() + f.p0_off, f.p0.data(), f.p0.size() * 4);
            memcpy(host.data() + f.p1_off, f.p1.data(), f.p1.size() * 4);
        }
    }
    SM_CUDA(cudaMalloc(&weight_arena_, weight_bytes_));
    if (cfg_.defer_weights) {
        SM_CUDA(cudaMemset(weight_arena_, 0, weight_bytes_));  // filled by broadcast_weights()
        weights_pending_ = true;   // encode refuses to run on the zero arena
    } else {
        SM_CUDA(cudaMemcpy(weight_arena_, host.data(), weight_bytes_, cudaMemcpyHostToDevice));
    }
    return SMELTER_OK;
}

// ---- per-batch plan ---------------------------------------------------------------------------------------------

namespace {

struct ArenaAlloc {
    struct Block { size_t off, size; };
    std::vector<Block> free_list;
    size_t top = 0;
    size_t peak = 0;  // high-water mark: `top` shrinks again when trailing blocks are released
    static size_t align(size_t v) { return (v + 1023) & ~size_t(1023); }
    size_t alloc(size_t bytes) {
        bytes = align(std::max<size_t>(bytes, 16));
        int best = -1;
        for (size_t i = 0; i < free_list.size(); ++i)
            if (free_list[i].size >= bytes && (best < 0 || free_list[i].size < free_list[size_t(best)].size)) best = int(i);
        if (best >= 0) {
            Block b = free_list[size_t(best)];
            free_list.erase(free_list.begin() + best);
            if (b.size > bytes) free_list.push_back({b.off + bytes, b.size - bytes});
            return b.off;
        }
        const size_t off = top;
        top += bytes;
        if (top > peak) peak = top;
        return off;
    }
    void release(size_t off, size_t bytes) {
        bytes = align(std::max<size_t>(bytes, 16));
        free_list.push_back({off, bytes});
        // coalesce
        std::sort(free_list.begin(), free_list.end(), [](const Block& a, const Block& b) { return a.off < b.off; });
        std::vector<Block> merged;
        for (const Block& b : free_list) {
            if (!merged.empty() && merged.back().off + merged.back().size == b.off) merged.back().size += b.size;
            else merged.push_back(b);
        }
        free_list.swap(merged);
        if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
            top = free_list.back().off;
            free_list.pop_back();
        }
    }
};

}  // namespace

int ONNXGraph::plan_for(int batch, Plan** out, cudaStream_t stream) {
    if (!built_) return fail(SMELTER_ERR_INCONSISTENT_STATE, "graph not built");
    const std::pair<int, uintptr_t> key(batch, (stream == nullptr || stream == ctx_->stream) ? uintptr_t(0) : reinterpret_cast<uintptr_t>(stream));
    auto it = plans_.find(key);
    if (it != plans_.end()) { *out = it->second.get(); return SMELTER_OK; }
    {   // every stream that encodes this batch size owns an activation arena: bound the number of them
        int same_batch = 0;
        for (const auto& kv : plans_) same_batch += kv.first.first == batch;
        if (same_batch >= 8) return fail(SMELTER_ERR_INCONSISTENT_STATE, "more than 8 streams encode batch " + std::to_string(batch) + " on one graph");
    }
    SM_CUDA(cudaSetDevice(ctx_->device));
    auto plan = std::make_unique<Plan>();
    plan->batch = batch;
    const int N = batch;
    // Configuration.smShare: this graph's kernels are planned and launched for a 1/k share of the SMs (an even number: CTA pairs)
    const int num_sms = cfg_.sm_share >= 2 ? std::max(2, (ctx_->num_sms / cfg_.sm_share) & ~1) : ctx_->num_sms;
    const char* wbase = static_cast<const char*>(weight_arena_);

    auto root_of = [&](int v) { while (values_[size_t(v)].alias_of >= 0) v = values_[size_t(v)].alias_of; return v; };
    auto pitch_of = [&](int v) { return round_up(values_[size_t(v)].shape.c, 8); };
    auto bytes_of = [&](int v) { const ImageShape& s = values_[size_t(v)].shape; return size_t(N) * s.h * s.w * round_up(s.c, 8) * 2; };

    const int out_root = root_of(output_value_);

    // shape part of the tensor-core conv problem of filter f (pointers are bound later)
    auto conv_problem = [&](const Filter& f, const std::vector<const Filter*>& stem_map) {
        const ImageShape is = values_[size_t(f.in[0])].shape;
        const ImageShape osz = values_[size_t(f.out)].shape;
        k::ConvTcProblem q{};
        q.mode = f.conv_mode;
        q.n = N; q.h = is.h; q.w = is.w;
        q.c_in = f.c_in_g; q.c_in_pitch = round_up(is.c, 8);
        if (f.is_gemm) { q.h = q.w = 1; q.c_in_pitch = round_up(is.c * is.h * is.w, 8); }
        q.c_out = f.c_out; q.c_out_pitch = round_up(osz.c, 8);
        q.k_h = f.k_h; q.k_w = f.k_w; q.stride_h = f.stride_h; q.stride_w = f.stride_w; q.dil_h = f.dil_h; q.dil_w = f.dil_w;
        q.pad_t = f.pads[0]; q.pad_l = f.pads[1]; q.pad_b = f.pads[2]; q.pad_r = f.pads[3];
        q.act = f.act; q.clip_lo = f.clip_lo; q.clip_hi = f.clip_hi;
        if (f.transposed) {  // stride-1 convolution over the zero-stuffed, bordered image
            q.h = (is.h - 1) * f.tr_stride_h + 1 + f.pads[0] + f.pads[2];
            q.w = (is.w - 1) * f.tr_stride_w + 1 + f.pads[1] + f.pads[3];
            q.pad_t = q.pad_l = q.pad_b = q.pad_r = 0;
        } else if (f.in_fold) {  // [N, Hp/4, Wp/4, 16 x 4] -> [N, P/4, Q/4, 16 Cout], ((k + 6) / 4)^2 taps
            q.h = (is.h + f.in_fold_pad[0] + f.in_fold_pad[2]) / 4; q.w = (is.w + f.in_fold_pad[1] + f.in_fold_pad[3]) / 4;
            q.c_in = 64; q.c_in_pitch = 64;
            q.c_out = 16 * f.c_out; q.c_out_pitch = 16 * f.c_out;
            q.k_h = (f.k_h + 6) / 4; q.k_w = (f.k_w + 6) / 4;
        } else if (f.upfold) {  // [N, H + 2, W + 2, Cin] -> [N, H, W, 4 Cout]: the 2H x 2W output with its pixels in (y, x, phase) order
            q.c_out = 4 * f.c_out; q.c_out_pitch = 4 * f.c_out;
        } else if (f.wfold) {  // the same buffers with wfold neighbouring pixels as one: [N, H, W/F, F Cin] -> [N, P, Q/F, F Cout]
            q.w = is.w / f.wfold;
            q.c_in = f.wfold * round_up(is.c, 8); q.c_in_pitch = f.wfold * round_up(is.c, 8);
            q.c_out = f.wfold * f.c_out; q.c_out_pitch = f.wfold * f.c_out;
            q.k_w = (f.k_w + 2 * f.wfold - 2) / f.wfold;
        } else if (f.phase_fold) {  // the folded problem: [N, H/F, W/F, F^2 Cin] -> [N, P/F, Q/F, F^2 Cout], ((k + 2F - 2) / F)^2 taps
            const int F = f.phase_fold;
            q.h = is.h / F; q.w = is.w / F;
            q.c_in = F * F * round_up(is.c, 8); q.c_in_pitch = F * F * round_up(is.c, 8);
            q.c_out = F * F * f.c_out; q.c_out_pitch = round_up(F * F * f.c_out, 8);
            q.k_h = (f.k_h + 2 * F - 2) / F; q.k_w = (f.k_w + 2 * F - 2) / F;
        } else if (f.s2d) {  // stride-1 convolution over the folded image the boundary conversion writes
            q.h = f.s2d_h / 2; q.w = f.s2d_w / 2;
            q.c_in_pitch = 16;
            q.k_h = (f.k_h + 1) / 2; q.k_w = (f.k_w + 1) / 2;
            q.stride_h = q.stride_w = 1;
            q.pad_t = q.pad_l = q.pad_b = q.pad_r = 0;
        } else if (f.conv_mode == k::CONV_MODE_PACKED_ROW && (f.pads[0] || f.pads[1] || f.pads[2] || f.pads[3])) {
            q.h = is.h + f.pads[0] + f.pads[2];  // packed-row convolutions read a materialised zero-padded image
            q.w = is.w + f.pads[1] + f.pads[3];
            q.pad_t = q.pad_l = q.pad_b = q.pad_r = 0;
        }
        (void)stem_map;
        return q;
    };
    // ---- projection shortcuts folded into the consuming convolution's k-loop ----
    // A residual block whose skip path is a 1x1 convolution (ResNet "downsample"): conv3(a) + conv_ds(x) is one GEMM over the
    // concatenated K axis, so the shortcut runs as extra k-blocks of conv3's launch (kernels/conv_pair.cu) -- no launch of its
    // own, and its output never exists in HBM.  side_of[conv3] = index of the absorbed shortcut filter.
    std::vector<int> side_of(filters_.size(), -1);
    std::vector<char> absorbed(filters_.size(), 0);
    auto side_fields = [&](k::ConvTcProblem& q, const Filter& g) {
        const ImageShape gs = values_[size_t(g.in[0])].shape;
        q.side_h = gs.h; q.side_w = gs.w;
        q.side_c_in = g.c_in_g; q.side_c_in_pitch = round_up(gs.c, 8);
        q.side_stride_h = g.stride_h; q.side_stride_w = g.stride_w;
    };
    std::vector<int> reads(values_.size(), 0), producer(values_.size(), -1);
    for (size_t fi = 0; fi < filters_.size(); ++fi) {
        const Filter& f = filters_[fi];
        if (f.removed) continue;
        for (int i : f.in) ++reads[size_t(root_of(i))];
        if (f.residual >= 0) ++reads[size_t(root_of(f.residual))];
        if (f.out >= 0 && values_[size_t(f.out)].alias_of < 0) producer[size_t(f.out)] = int(fi);
    }
    auto plain_tc_conv = [&](const Filter& f) {
        return f.kind == FilterKind::Conv && !f.is_gemm && !f.transposed && !f.s2d && !f.phase_fold && !f.wfold && (f.conv_mode == k::CONV_MODE_TILED || f.conv_mode == k::CONV_MODE_IM2COL);
    };
    if (!getenv("SMELTER_NO_SIDE")) {
        for (size_t fi = 0; fi < filters_.size(); ++fi) {
            const Filter& f = filters_[fi];
            if (f.removed || !plain_tc_conv(f) || f.residual < 0) continue;
            const int r = root_of(f.residual);
            const int gi = producer[size_t(r)];
            if (gi < 0 || gi >= int(fi) || r == out_root || reads[size_t(r)] != 1) continue;
            const Filter& g = filters_[size_t(gi)];
            if (!plain_tc_conv(g) || absorbed[size_t(gi)] || side_of[size_t(gi)] >= 0 || g.k_h != 1 || g.k_w != 1 || g.pads[0] || g.pads[1] || g.pads[2] ||
                g.pads[3] || g.act != k::ACT_NONE || g.residual >= 0 || g.c_out != f.c_out)
                continue;
            k::ConvTcProblem q = conv_problem(f, {});
            side_fields(q, g);
            q.side_x = reinterpret_cast<const __half*>(uintptr_t(16));  // shape query only: any non-null pointer
            if (!k::conv_tc_side_supported(q, num_sms)) continue;
            side_of[fi] = gi;
            absorbed[size_t(gi)] = 1;
        }
    }

    // last use of every root value (filter index); the output value lives forever.  An absorbed shortcut's input is read by the
    // convolution that absorbed it.
    std::vector<int> last_use(values_.size(), -1);
    for (size_t fi = 0; fi < filters_.size(); ++fi) {
        const Filter& f = filters_[fi];
        if (f.removed || absorbed[fi]) continue;
        for (int i : f.in) last_use[size_t(root_of(i))] = int(fi);
        if (f.residual >= 0 && side_of[fi] < 0) last_use[size_t(root_of(f.residual))] = int(fi);
        if (side_of[fi] >= 0) last_use[size_t(root_of(filters_[size_t(side_of[fi])].in[0]))] = int(fi);
    }
    last_use[size_t(out_root)] = int(filters_.size()) + 1;

    // ---- offsets (two passes: first compute offsets, then allocate, then bind pointers) ----
    ArenaAlloc arena;
    std::vector<size_t> off(values_.size(), size_t(-1));
    struct Scratch { size_t off, bytes; };
    std::vector<Scratch> scratch(filters_.size(), Scratch{size_t(-1), 0});   // per-filter temporary (padded input copy, IN partials, NCHW staging)
    std::vector<Scratch> scratch2(filters_.size(), Scratch{size_t(-1), 0});   // split-K fp32 workspace
    std::vector<size_t> counter_off(filters_.size(), size_t(-1));
    size_t counter_total = 0;

    // A graph input whose only reader is a packed-row (stem) convolution is converted straight into the zero-padded
    // NHWC image that kernel wants: the boundary conversion materialises the padding, no separate pad pass.
    std::vector<const Filter*> stem_of(values_.size(), nullptr);
    for (int v : input_values_) {
        const Filter* only = nullptr;
        int readers = 0;
        for (const Filter& f : filters_) {
            if (f.removed) continue;
            for (int i : f.in) if (root_of(i) == v) { ++readers; only = &f; }
            if (f.residual >= 0 && root_of(f.residual) == v) ++readers;
        }
        if (readers == 1 && v != out_root && only->kind == FilterKind::Conv && only->conv_mode == k::CONV_MODE_PACKED_ROW && root_of(only->in[0]) == v &&
            (only->s2d || only->pads[0] || only->pads[1] || only->pads[2] || only->pads[3]))
            stem_of[size_t(v)] = only;
        if (readers == 1 && only->kind == FilterKind::Conv && only->in_fold && only->in[0] == v) stem_of[size_t(v)] = only;
    }
    auto input_bytes = [&](int v) {
        const Filter* f = stem_of[size_t(v)];
        if (!f) return bytes_of(v);
        const ImageShape& s = values_[size_t(v)].shape;
        if (f->s2d) return size_t(N) * (f->s2d_h / 2) * (f->s2d_w / 2) * 16 * 2;
        if (f->in_fold) return size_t(N) * ((s.h + f->in_fold_pad[0] + f->in_fold_pad[2]) / 4) * ((s.w + f->in_fold_pad[1] + f->in_fold_pad[3]) / 4) * 64 * 2;
        return size_t(N) * (s.h + f->pads[0] + f->pads[2]) * (s.w + f->pads[1] + f->pads[3]) * round_up(s.c, 8) * 2;
    };
    // ---- instance-norm statistics accumulated by the producing convolution's epilogue ----
    // Conv -> InstanceNormalization (the convolution's only reader): the two-CTA kernel adds the column sums / sums of squares of
    // every output tile to per-image fp64 accumulators (ConvKernelParams::stats), and the norm is ONE pass over the tensor
    // (k::instance_norm_from_stats) instead of statistics + apply.  Phase-column convolutions (upsample-, input-, width-folded)
    // keep one accumulator per (phase, channel) column; the norm adds the phases up.  SMELTER_NO_CONV_STATS=1 turns it off.
    std::vector<size_t> stats_off(filters_.size(), size_t(-1));  // by convolution, into plan->counters: [N][replicas][c_out_pitch][2] fp64, then the arrival counter
    std::vector<int> stats_conv(filters_.size(), -1);            // by norm: the convolution that supplies its statistics
    std::vector<int> stats_phases(filters_.size(), 1);
    if (!getenv("SMELTER_NO_CONV_STATS")) {
        for (size_t fi = 0; fi < filters_.size(); ++fi) {
            const Filter& f = filters_[fi];
            if (f.removed || absorbed[fi] || f.kind != FilterKind::InstanceNorm || f.sub > 1 || f.in.empty()) continue;
            const int src = root_of(f.in[0]);
            const int ci = producer[size_t(src)];
            if (ci < 0 || reads[size_t(src)] != 1 || src == out_root) continue;
            const Filter& c = filters_[size_t(ci)];
            if (c.removed || absorbed[size_t(ci)] || c.kind != FilterKind::Conv || c.is_gemm || c.transposed || c.conv_mode == 4 || c.phase_fold || c.residual >= 0 ||
                side_of[size_t(ci)] >= 0 || c.act != k::ACT_NONE)
                continue;
            const int phases = c.in_fold ? 16 : c.upfold ? 4 : c.wfold ? c.wfold : 1;
            if (phases > 1 && c.c_out % 8) continue;
            const k::ConvTcProblem q = conv_problem(c, stem_of);
            if (q.c_out_pitch != phases * round_up(values_[size_t(f.in[0])].shape.c, 8) || !k::conv_tc_stats_supported(q, num_sms)) continue;
            stats_conv[fi] = ci;
            stats_phases[fi] = phases * k::conv_tc_stats_reps(q);  // replicas are more column groups to the norm
            stats_off[size_t(ci)] = counter_total;
            counter_total += (size_t(N) * k::conv_tc_stats_reps(q) * q.c_out_pitch * 2 * sizeof(double) + sizeof(unsigned int) + 255) & ~size_t(255);
        }
    }

    // ---- norm tail: InstanceNorm -> Add(skip) (-> ReLU) -> Pad, the end of a residual block ----
    // The one-pass norm (statistics from its convolution) also reads the skip tensor and stores the padded sum, and the plain sum
    // while the next block's skip connection reads it: the Add+Pad kernel (Filter::pad_add) runs inside the norm's launch.  Decided
    // per plan because the one-pass form is; the Pad must be the next live filter.  SMELTER_NO_NORM_TAIL=1 turns it off.
    std::vector<int> tail_of(filters_.size(), -1);
    if (!getenv("SMELTER_NO_NORM_TAIL")) {
        for (size_t fi = 0; fi < filters_.size(); ++fi) {
            const Filter& f = filters_[fi];
            if (stats_conv[fi] < 0 || f.norm_padded || f.unfold_w || f.out2 >= 0) continue;
            const int nr = root_of(f.out);
            if (nr == out_root || reads[size_t(nr)] != 1) continue;
            size_t pi = fi + 1;
            while (pi < filters_.size() && (filters_[pi].removed || absorbed[pi] || filters_[pi].kind == FilterKind::Alias)) ++pi;
            if (pi >= filters_.size()) continue;
            const Filter& pd = filters_[pi];
            if (pd.kind != FilterKind::Pad || !pd.pad_add || pd.s2d_out || (pd.sub != k::PAD_REFLECT && pd.sub != k::PAD_EDGE) || pd.in.size() != 2) continue;
            const int a = root_of(pd.in[0]), b = root_of(pd.in[1]);
            if ((a == nr) == (b == nr)) continue;
            const ImageShape& sh = values_[size_t(f.out)].shape;
            if (pd.sub == k::PAD_REFLECT && (pd.pads[0] >= sh.h || pd.pads[2] >= sh.h || pd.pads[1] >= sh.w || pd.pads[3] >= sh.w)) continue;
            const int skip = a == nr ? b : a;
            tail_of[fi] = int(pi);
            absorbed[pi] = 1;
            if (last_use[size_t(skip)] == int(pi)) last_use[size_t(skip)] = int(fi);
        }
    }

    // graph inputs: NHWC copies produced by the boundary conversion
    for (int v : input_values_) off[size_t(v)] = arena.alloc(input_bytes(v));
    for (size_t fi = 0; fi < filters_.size(); ++fi) {
        const Filter& f = filters_[fi];
        if (f.removed || absorbed[fi]) continue;
        // temporaries first (live only during this filter)
        if (f.kind == FilterKind::Conv && f.conv_mode == k::CONV_MODE_PACKED_ROW && (f.pads[0] || f.pads[1] || f.pads[2] || f.pads[3]) &&
            stem_of[size_t(root_of(f.in[0]))] != &f) {
            const ImageShape& s = values_[size_t(f.in[0])].shape;
            scratch[fi].bytes = size_t(N) * (s.h + f.pads[0] + f.pads[2]) * (s.w + f.pads[1] + f.pads[3]) * 8 * 2;
            scratch[fi].off = arena.alloc(scratch[fi].bytes);
        }
        if (f.kind == FilterKind::Conv && f.transposed) {
            const k::ConvTcProblem tq = conv_problem(f, stem_of);
            scratch[fi].bytes = size_t(N) * tq.h * tq.w * tq.c_in_pitch * 2;
            scratch[fi].off = arena.alloc(scratch[fi].bytes);
        }
        if (f.kind == FilterKind::Conv && f.conv_mode != 4 && side_of[fi] < 0) {
            const k::ConvTcPlanInfo info = k::conv_tc_plan(conv_problem(f, stem_of), num_sms);
            if (info.splits > 1) {
                scratch2[fi].bytes = info.ws_bytes;
                scratch2[fi].off = arena.alloc(scratch2[fi].bytes);
                counter_off[fi] = counter_total;
                counter_total += (info.counter_bytes + 255) & ~size_t(255);
            }
        }
        if (f.kind == FilterKind::InstanceNorm) {
            const ImageShape& s = values_[size_t(f.in[0])].shape;
            scratch[fi].bytes = k::instance_norm_scratch_floats(N, s.h * s.w, round_up(s.c, 8)) * sizeof(float);
            scratch[fi].off = arena.alloc(scratch[fi].bytes);
        }
        if (f.kind == FilterKind::Reshape) {
            const ImageShape& a = values_[size_t(f.in[0])].shape;
            const ImageShape& b = values_[size_t(f.out)].shape;
            const bool view = a.h == 1 && a.w == 1 && b.h == 1 && b.w == 1;
            if (!view) {
                scratch[fi].bytes = size_t(N) * a.c * a.h * a.w * 2;  // dense NCHW staging
                scratch[fi].off = arena.alloc(scratch[fi].bytes);
            }
        }
        if (tail_of[fi] >= 0) {  // the norm writes the absorbed Pad's outputs; its own result never exists
            const Filter& pd = filters_[size_t(tail_of[fi])];
            off[size_t(pd.out)] = arena.alloc(bytes_of(pd.out));
            if (pd.out2 >= 0) off[size_t(pd.out2)] = arena.alloc(bytes_of(pd.out2));
        } else {
            off[size_t(f.out)] = arena.alloc(bytes_of(f.out));
            if (f.out2 >= 0) off[size_t(f.out2)] = arena.alloc(bytes_of(f.out2));
        }
        if (scratch[fi].off != size_t(-1)) arena.release(scratch[fi].off, scratch[fi].bytes);
        if (scratch2[fi].off != size_t(-1)) arena.release(scratch2[fi].off, scratch2[fi].bytes);
        // release inputs whose last use is this filter
        std::vector<int> roots;
        for (int i : f.in) roots.push_back(root_of(i));
        if (f.residual >= 0 && side_of[fi] < 0) roots.push_back(root_of(f.residual));
        if (side_of[fi] >= 0) roots.push_back(root_of(filters_[size_t(side_of[fi])].in[0]));
        if (tail_of[fi] >= 0) for (int i : filters_[size_t(tail_of[fi])].in) roots.push_back(root_of(i));
        std::sort(roots.begin(), roots.end());
        roots.erase(std::unique(roots.begin(), roots.end()), roots.end());
        for (int r : roots)
            if (last_use[size_t(r)] == int(fi) && off[size_t(r)] != size_t(-1)) {
                const size_t rb = values_[size_t(r)].is_input ? input_bytes(r) : bytes_of(r);
                arena.release(off[size_t(r)], rb);
            }
    }
    // result tensor (NCHW) unless the output value is already NCHW-compatible
    const ImageShape& os = values_[size_t(output_value_)].shape;
    const bool result_is_view = os.h == 1 && os.w == 1 && os.c % 8 == 0;
    size_t result_off = size_t(-1);
    if (!result_is_view) result_off = arena.alloc(size_t(N) * os.c * os.h * os.w * 2);

    plan->arena_bytes = std::max<size_t>(arena.peak, 1024);
    SM_CUDA(cudaMalloc(&plan->arena, plan->arena_bytes));
    if (counter_total) {
        SM_CUDA(cudaMalloc(&plan->counters, counter_total));
        SM_CUDA(cudaMemset(plan->counters, 0, counter_total));
        plan->counter_bytes = counter_total;
    }
    char* abase = static_cast<char*>(plan->arena);
    // values without a buffer (outputs of convolutions that run inside their consumer's launch) have no address
    auto ptr_of = [&](int v) -> __half* {
        const size_t o = off[size_t(root_of(v))];
        return o == size_t(-1) ? nullptr : reinterpret_cast<__half*>(abase + o);
    };

    // ---- steps ----
    auto add_step = [&](std::string desc, std::function<cudaError_t(cudaStream_t)> fn, double flops = 0, double bytes = 0) {
        Step st;
        st.desc = std::move(desc);
        st.run = std::move(fn);
        st.flops = flops;
        st.bytes = bytes;
        plan->steps.push_back(std::move(st));
    };
    // boundary: NCHW source -> NHWC (the source pointer is read at launch time through a stable slot)
    for (int v : input_values_) {
        plan->src_ptr_storage.push_back(std::make_unique<__half*>(nullptr));
        __half** slot = plan->src_ptr_storage.back().get();
        plan->src_slots.push_back(slot);
        const ImageShape s = values_[size_t(v)].shape;
        plan->src_shapes.push_back(s);
        __half* dst = ptr_of(v);
        const int cp = pitch_of(v);
        int pt = 0, pl = 0, pb = 0, pr = 0;
        if (const Filter* sf = stem_of[size_t(v)]) { pt = sf->pads[0]; pl = sf->pads[1]; pb = sf->pads[2]; pr = sf->pads[3]; }
        if (const Filter* sf = stem_of[size_t(v)]; sf && sf->in_fold) {
            const int h4 = (s.h + sf->in_fold_pad[0] + sf->in_fold_pad[2]) / 4, w4 = (s.w + sf->in_fold_pad[1] + sf->in_fold_pad[3]) / 4;
            const int fpt = sf->in_fold_pad[0], fpl = sf->in_fold_pad[1], mode = sf->in_fold_mode;
            const float value = sf->in_fold_value;
            add_step("nchw_to_s2d4+pad " + values_[size_t(v)].name,
                     [=](cudaStream_t st) { return k::nchw_to_s2d4(*slot, dst, N, s.c, s.h, s.w, fpt, fpl, h4, w4, mode, value, st); }, 0,
                     double(N) * (double(s.h) * s.w * s.c + double(h4) * w4 * 64) * 2);
            plan->steps.back().boundary = true;
            continue;
        }
        if (const Filter* sf = stem_of[size_t(v)]; sf && sf->s2d) {
            const int h2 = sf->s2d_h / 2, w2 = sf->s2d_w / 2;
            add_step("nchw_to_s2d+pad0 " + values_[size_t(v)].name,
                     [=](cudaStream_t st) { return k::nchw_to_s2d(*slot, dst, N, s.c, s.h, s.w, pt, pl, h2, w2, st); }, 0,
                     double(N) * (double(s.h) * s.w * s.c + double(h2) * w2 * 16) * 2);
            plan->steps.back().boundary = true;
            continue;
        }
        add_step(std::string(pt || pl || pb || pr ? "nchw_to_nhwc+pad0 " : "nchw_to_nhwc ") + values_[size_t(v)].name,
                 [=](cudaStream_t st) { return k::nchw_to_nhwc(*slot, dst, N, s.c, s.h, s.w, cp, pt, pl, pb, pr, st); }, 0,
                 double(N) * (double(s.h) * s.w * s.c + double(s.h + pt + pb) * (s.w + pl + pr) * cp) * 2);
        plan->steps.back().boundary = true;
    }

    for (size_t fi = 0; fi < filters_.size(); ++fi) {
        const Filter& f = filters_[fi];
        if (f.removed || absorbed[fi]) continue;
        const ImageShape is = values_[size_t(f.in[0])].shape;
        const ImageShape osz = values_[size_t(f.out)].shape;
        const int icp = round_up(is.c, 8), ocp = round_up(osz.c, 8);
        const __half* x = ptr_of(f.in[0]);
        __half* y = ptr_of(f.out);
        const std::string name = f.op_type + " " + values_[size_t(f.out)].name;
        const double io_bytes = double(N) * (double(is.h) * is.w * icp + double(osz.h) * osz.w * ocp) * 2;
        switch (f.kind) {
            case FilterKind::Conv: {
                const float* bias = reinterpret_cast<const float*>(wbase + f.bias_off);
                const __half* w = reinterpret_cast<const __half*>(wbase + f.w_off);
                const Filter* side = side_of[fi] >= 0 ? &filters_[size_t(side_of[fi])] : nullptr;
                const __half* res = f.residual >= 0 && !side ? ptr_of(f.residual) : nullptr;
                std::string suffix = f.act == k::ACT_RELU ? "+relu" : f.act == k::ACT_CLIP ? "+clip" : f.act == k::ACT_SIGMOID ? "+sigmoid" : "";
                if (res) suffix = "+add" + suffix;
                if (f.group_expanded) suffix = "/groups" + std::to_string(f.group_expanded) + "-as-dense" + suffix;
                if (side) suffix = "+conv1x1(" + values_[size_t(side->in[0])].name + ")" + suffix;
                const double flops = f.transposed ? 2.0 * N * is.h * is.w * double(f.c_out) * f.c_in_g * f.k_h * f.k_w
                                                  : 2.0 * N * osz.h * osz.w * double(f.c_out) * f.c_in_g * f.k_h * f.k_w;
                if (f.conv_mode == 4) {
                    const Filter* fp = &f;
                    add_step("depthwise" + suffix + " " + name, [=](cudaStream_t st) {
                        return k::depthwise_conv(x, w, bias, y, N, is.h, is.w, icp, osz.h, osz.w, fp->k_h, fp->k_w, fp->stride_h, fp->stride_w, fp->dil_h,
                                                 fp->dil_w, fp->pads[0], fp->pads[1], fp->act, fp->clip_lo, fp->clip_hi, st);
                    }, flops, io_bytes);
                    break;
                }
                k::ConvTcProblem q = conv_problem(f, stem_of);
                q.x = x; q.w_packed = w; q.bias = bias; q.residual = res; q.y = y;
                double side_bytes = 0, side_flops = 0;
                if (side) {
                    side_fields(q, *side);
                    q.side_x = ptr_of(side->in[0]);
                    q.side_w_packed = reinterpret_cast<const __half*>(wbase + side->w_off);
                    q.side_bias = reinterpret_cast<const float*>(wbase + side->bias_off);
                    side_flops = 2.0 * N * osz.h * osz.w * double(f.c_out) * q.side_c_in;
                    side_bytes = double(N) * q.side_h * q.side_w * q.side_c_in_pitch * 2 / (q.side_stride_h * q.side_stride_w) + double(f.c_out) * q.side_c_in * 2;
                }
                // L2 eviction priority (measured on ResNet-50 / batch 32, same-box A/B: ~1 % of the step): a residual operand read for
                // the last time is marked evict_first so it does not push the freshly written block output out of L2.  (Writing
                // skip tensors evict_last as well measured slightly worse.)
                if (res && last_use[size_t(root_of(f.residual))] == int(fi)) q.l2_hints |= 2;
                if (scratch2[fi].off != size_t(-1)) {
                    q.split_ws = reinterpret_cast<float*>(abase + scratch2[fi].off);
                    q.split_counters = reinterpret_cast<unsigned int*>(static_cast<char*>(plan->counters) + counter_off[fi]);
                }
                if (stats_off[fi] != size_t(-1)) {
                    q.stats = reinterpret_cast<double*>(static_cast<char*>(plan->counters) + stats_off[fi]);
                    suffix += "+stats";
                }
                if (f.transposed) {
                    __half* stuffed = reinterpret_cast<__half*>(abase + scratch[fi].off);
                    const Filter* fp = &f;
                    const int hz = q.h, wz = q.w;
                    add_step("zero_stuff " + name, [=](cudaStream_t st) {
                        return k::zero_stuff2d(x, stuffed, N, is.h, is.w, icp, hz, wz, fp->tr_stride_h, fp->tr_stride_w, fp->pads[0], fp->pads[1], st);
                    }, 0, double(scratch[fi].bytes) + double(N) * is.h * is.w * icp * 2);
                    q.x = stuffed;
                } else if (f.conv_mode == k::CONV_MODE_PACKED_ROW && stem_of[size_t(root_of(f.in[0]))] == &f) {
                    // the boundary conversion already wrote the padded image into this conv's input buffer
                } else if (f.conv_mode == k::CONV_MODE_PACKED_ROW && scratch[fi].off != size_t(-1)) {
                    // materialise the zero padding so one K block can span a whole filter row
                    __half* padded = reinterpret_cast<__half*>(abase + scratch[fi].off);
                    const Filter* fp = &f;
                    add_step("pad0 " + name, [=](cudaStream_t st) {
                        return k::pad2d(x, padded, N, is.h, is.w, 8, fp->pads[0], fp->pads[1], fp->pads[2], fp->pads[3], k::PAD_CONSTANT, 0.f, st);
                    }, 0, double(scratch[fi].bytes) + double(N) * is.h * is.w * 16);
                    q.x = padded;
                }
                auto L = std::make_shared<k::ConvTcLaunch>();
                std::string cerr;
                if (!k::conv_tc_prepare(L.get(), q, num_sms, &cerr)) return fail(SMELTER_ERR_GRAPH_INTERNAL, name + ": " + cerr);
                L->balanced_grid = cfg_.sm_share >= 2 ? 1 : 0;
                const char* mode_name = f.in_fold ? "im2col/input-fold4" : f.upfold ? (f.conv_mode == k::CONV_MODE_PACKED_ROW ? "rows/upsample-fold" : "im2col/upsample-fold") : f.wfold ? "im2col/width-fold" : f.phase_fold == 4 ? "im2col/phase-fold4" : f.phase_fold ? "im2col/phase-fold" : f.conv_mode == k::CONV_MODE_TILED ? "tiled" : f.conv_mode == k::CONV_MODE_IM2COL ? "im2col" : f.s2d ? "rows/s2d" : "rows";
                add_step(std::string(L->pair ? "conv_pair[" : "conv_igemm[") + mode_name + ",bn" + std::to_string(L->block_n) + (L->splits > 1 ? ",k/" + std::to_string(L->splits) : "") +
                             "]" + suffix + " " + name,
                         [L](cudaStream_t st) { return k::conv_tc_launch(*L, st); }, flops + side_flops,
                         io_bytes + double(f.c_out) * f.c_in_g * f.k_h * f.k_w * 2 + (res ? double(N) * osz.h * osz.w * ocp * 2 : 0) + side_bytes);
                plan->steps.back().tensor = true;
                break;
            }
            case FilterKind::BatchNorm: {
                const float* sc = reinterpret_cast<const float*>(wbase + f.p0_off);
                const float* sh = reinterpret_cast<const float*>(wbase + f.p1_off);
                const int act = f.act;
                add_step("scale_shift " + name, [=](cudaStream_t st) { return k::scale_shift(x, y, size_t(N) * is.h * is.w, icp, sc, sh, act, st); }, 0, io_bytes);
                break;
            }
            case FilterKind::InstanceNorm: {
                const float* ga = reinterpret_cast<const float*>(wbase + f.p0_off);
                const float* be = reinterpret_cast<const float*>(wbase + f.p1_off);
                float* partials = reinterpret_cast<float*>(abase + scratch[fi].off);
                const int act = f.act;
                const float eps = f.eps;
                const int group_size = f.sub > 0 ? f.sub : 1, channels = is.c;
                k::NormStore store;
                store.unfold_w = f.unfold_w;
                store.unfold_f = f.unfold_f;
                store.h = is.h; store.w = is.w;
                if (f.norm_padded) {
                    store.pad_t = f.norm_pad[0]; store.pad_l = f.norm_pad[1]; store.pad_b = f.norm_pad[2]; store.pad_r = f.norm_pad[3];
                    store.s2d = f.norm_s2d; store.pad_mode = f.norm_pad_mode;
                    store.plain = f.out2 >= 0 ? ptr_of(f.out2) : nullptr;
                }
                std::string what = group_size > 1 ? "group_norm" : "instance_norm";
                if (f.unfold_w) what += "+unfold";
                if (f.norm_padded) what += f.norm_s2d ? "+pad+s2d" : f.out2 >= 0 ? "+pad+plain" : "+pad";
                if (stats_conv[fi] >= 0) {  // statistics from the producing convolution: one pass
                    double* stats = reinterpret_cast<double*>(static_cast<char*>(plan->counters) + stats_off[size_t(stats_conv[fi])]);
                    const int phases = stats_phases[fi];
                    unsigned int* counter = reinterpret_cast<unsigned int*>(stats + size_t(N) * phases * icp * 2);
                    if (tail_of[fi] >= 0) {
                        const Filter& pd = filters_[size_t(tail_of[fi])];
                        const int skip = root_of(pd.in[0]) == root_of(f.out) ? pd.in[1] : pd.in[0];
                        const __half* res = ptr_of(skip);
                        __half* yp = ptr_of(pd.out);
                        store.pad_t = pd.pads[0]; store.pad_l = pd.pads[1]; store.pad_b = pd.pads[2]; store.pad_r = pd.pads[3];
                        store.s2d = 0; store.pad_mode = pd.sub;
                        store.plain = pd.out2 >= 0 ? ptr_of(pd.out2) : nullptr;
                        const int act2 = pd.act;
                        const ImageShape ps = values_[size_t(pd.out)].shape;
                        add_step(what + "+add+pad" + (pd.out2 >= 0 ? "+sum" : "") + "<-stats " + f.op_type + " " + values_[size_t(pd.out)].name,
                                 [=](cudaStream_t st) { return k::instance_norm_from_stats(x, yp, N, is.h * is.w, icp, ga, be, eps, act, stats, counter, phases, st, &store, res, act2); }, 0,
                                 double(N) * (2.0 * is.h * is.w + double(ps.h) * ps.w + (pd.out2 >= 0 ? double(is.h) * is.w : 0.0)) * icp * 2);
                        break;
                    }
                    add_step(what + "<-stats " + name,
                             [=](cudaStream_t st) { return k::instance_norm_from_stats(x, y, N, is.h * is.w, icp, ga, be, eps, act, stats, counter, phases, st, &store); }, 0,
                             io_bytes);
                    break;
                }
                add_step(what + " " + name,
                         [=](cudaStream_t st) { return k::instance_norm(x, y, N, is.h * is.w, icp, ga, be, eps, act, partials, st, group_size, channels, &store); }, 0,
                         io_bytes);  // algorithmic bytes: one read + one write (the second pass finds its images in L2)
                plan->steps.back().launches = k::instance_norm_launches(N, is.h * is.w, icp, group_size);
                break;
            }
            case FilterKind::Unary: {
                const int kind = f.sub;
                const float a = f.alpha, b = f.beta;
                add_step("unary " + name, [=](cudaStream_t st) { return k::unary(x, y, size_t(N) * is.h * is.w * icp, kind, a, b, st, is.c, icp); }, 0, io_bytes);
                break;
            }
            case FilterKind::Binary: {
                const __half* x2 = ptr_of(f.in[1]);
                const int kind = f.sub, act = f.act;
                const size_t bcast = f.bcast ? size_t(is.h) * is.w : 0;
                add_step(std::string(f.bcast ? "binary/broadcast " : "binary ") + name,
                         [=](cudaStream_t st) { return k::binary(x, x2, y, size_t(N) * is.h * is.w * icp, kind, act, st, is.c, icp, bcast); }, 0,
                         io_bytes + double(N) * (f.bcast ? 1 : is.h * is.w) * icp * 2);
                break;
            }
            case FilterKind::Pool: {
                const Filter* fp = &f;
                add_step("pool " + name, [=](cudaStream_t st) {
                    return k::pool2d(x, y, N, is.h, is.w, icp, osz.h, osz.w, fp->k_h, fp->k_w, fp->stride_h, fp->stride_w, fp->pool_pad_h, fp->pool_pad_w,
                                     fp->sub, st);
                }, 0, io_bytes);
                break;
            }
            case FilterKind::GlobalAvgPool:
                add_step("global_avgpool " + name, [=](cudaStream_t st) { return k::global_avgpool(x, y, N, is.h * is.w, icp, st); }, 0, io_bytes);
                break;
            case FilterKind::Upsample: {
                const Filter* fp = &f;
                add_step("upsample " + name, [=](cudaStream_t st) {
                    return k::upsample2d(x, y, N, is.h, is.w, icp, fp->scale_h, fp->scale_w, fp->sub, fp->align_corners, st);
                }, 0, io_bytes);
                break;
            }
            case FilterKind::Concat: {
                int c_off = 0;
                for (int v : f.in) {
                    const ImageShape s = values_[size_t(v)].shape;
                    const __half* src = ptr_of(v);
                    const int scp = round_up(s.c, 8), co = c_off;
                    add_step("concat " + name, [=](cudaStream_t st) { return k::concat_channels(src, y, size_t(N) * s.h * s.w, s.c, scp, ocp, co, st); }, 0,
                             double(N) * s.h * s.w * s.c * 4);
                    c_off += s.c;
                }
                if (ocp != osz.c) {
                    // padded channels of the destination stay uninitialised otherwise: clear once per encode
                    // (cheap; only when the concatenated channel count is not a multiple of 8)
                    const size_t bytes = bytes_of(f.out);
                    Step clear;
                    clear.desc = "memset " + name;
                    clear.run = [=](cudaStream_t st) { return cudaMemsetAsync(y, 0, bytes, st); };
                    plan->steps.insert(plan->steps.end() - long(f.in.size()), std::move(clear));
                }
                break;
            }
            case FilterKind::Reshape: {
                if (scratch[fi].off == size_t(-1)) {
                    // [N,C,1,1] -> [N,C',1,1] with C == C': same bytes.  Copy keeps the planner simple when the
                    // value could not be aliased (pitches equal because both are round_up(c, 8)).
                    const size_t bytes = bytes_of(f.out);
                    add_step("reshape(copy) " + name, [=](cudaStream_t st) { return cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice, st); }, 0, 2.0 * bytes);
                } else {
                    __half* staging = reinterpret_cast<__half*>(abase + scratch[fi].off);
                    add_step("reshape(nhwc->nchw) " + name, [=](cudaStream_t st) { return k::nhwc_to_nchw(x, staging, N, is.c, is.h, is.w, icp, long(is.c) * is.h * is.w, st); }, 0, io_bytes);
                    add_step("reshape(nchw->nhwc) " + name, [=](cudaStream_t st) { return k::nchw_to_nhwc(staging, y, N, osz.c, osz.h, osz.w, ocp, 0, 0, 0, 0, st); }, 0, io_bytes);
                }
                break;
            }
            case FilterKind::Softmax: {
                const int lg = f.sub;
                add_step("softmax " + name, [=](cudaStream_t st) { return k::softmax_rows(x, y, size_t(N) * is.h * is.w, is.c, icp, lg, st); }, 0, io_bytes);
                break;
            }
            case FilterKind::Pad: {
                const Filter* fp = &f;
                const __half* x2 = f.pad_add ? ptr_of(f.in[1]) : nullptr;
                __half* y2 = f.pad_add && f.out2 >= 0 ? ptr_of(f.out2) : nullptr;
                const int act = f.pad_add ? f.act : int(k::ACT_NONE);
                add_step(std::string(f.pad_add ? (y2 ? "add+pad+sum " : "add+pad ") : f.s2d_out ? "pad+s2d " : "pad ") + name, [=](cudaStream_t st) {
                    return k::pad2d(x, y, N, is.h, is.w, icp, fp->pads[0], fp->pads[1], fp->pads[2], fp->pads[3], fp->sub, fp->alpha, st, fp->s2d_out, x2, y2, act);
                }, 0, io_bytes + (x2 ? double(N) * is.h * is.w * icp * 2 * (y2 ? 2 : 1) : 0));
                break;
            }
            case FilterKind::Alias:
                break;
        }
    }
    // boundary: result as NCHW
    plan->result.ctx = ctx_;
    plan->result.n = N; plan->result.c = os.c; plan->result.h = os.h; plan->result.w = os.w;
    plan->result.owned = false;
    if (result_is_view) {
        plan->result.ptr = ptr_of(output_value_);
    } else {
        __half* dst = reinterpret_cast<__half*>(abase + result_off);
        const __half* src = ptr_of(output_value_);
        const int cp = round_up(os.c, 8);
        const ImageShape o2 = os;
        int folded = 0;  // produced by a phase-folded convolution: [N, H/F, W/F, F^2 C] with the F^2 phases as channel blocks
        for (const Filter& pf : filters_) if (!pf.removed && pf.phase_fold && root_of(pf.out) == out_root) folded = pf.phase_fold;
        if (folded) {
            const int F = folded, cpf = round_up(F * F * os.c, 8);
            add_step("phase_to_nchw " + values_[size_t(output_value_)].name,
                     [=](cudaStream_t st) { return k::phase_to_nchw(src, dst, N, o2.c, o2.h / F, o2.w / F, cpf, long(o2.c) * o2.h * o2.w, st, F); }, 0,
                     double(N) * o2.h * o2.w * o2.c * 2 + double(N) * (o2.h / F) * (o2.w / F) * cpf * 2);
        } else
        add_step("nhwc_to_nchw " + values_[size_t(output_value_)].name,
                 [=](cudaStream_t st) { return k::nhwc_to_nchw(src, dst, N, o2.c, o2.h, o2.w, cp, long(o2.c) * o2.h * o2.w, st); }, 0,
                 double(N) * o2.h * o2.w * (o2.c + cp) * 2);
        plan->result.ptr = dst;
    }
    *out = plan.get();
    plans_[key] = std::move(plan);
    return SMELTER_OK;
}

int ONNXGraph::encode(cudaStream_t stream, const Tensor* const* sources, int n_sources, const Tensor** result) {
    if (!built_) return fail(SMELTER_ERR_INCONSISTENT_STATE, "encode before build");
    if (weights_pending_) return fail(SMELTER_ERR_INCONSISTENT_STATE, "weights were deferred (deferWeights) and never broadcast: call broadcastWeights first");
    if (!sources || n_sources != int(input_values_.size()))
        return fail(SMELTER_ERR_INSUFFICIENT_INPUTS, "graph expects " + std::to_string(input_values_.size()) + " source image(s)");
    for (int i = 0; i < n_sources; ++i)
        if (!sources[i] || !sources[i]->ptr) return fail(SMELTER_ERR_INVALID_ARGUMENT, "null source");
    const int batch = sources[0]->n;
    if (batch <= 0) return fail(SMELTER_ERR_INVALID_ARGUMENT, "empty batch");
    Plan* plan = nullptr;
    int rc = plan_for(batch, &plan, stream);
    if (rc) return rc;
    for (int i = 0; i < n_sources; ++i) {
        const ImageShape& s = plan->src_shapes[size_t(i)];
        const Tensor* t = sources[i];
        if (cfg_.input_constraint != SMELTER_INPUT_NONE && t->n == batch && t->c == s.c && (t->h != s.h || t->w != s.w) && t->h > 0 && t->w > 0) {
            // MPSNNLanczosScaleNode / MPSNNBilinearScaleNode in front of the graph (ONNXGraph.swift:219-241): eager, like the boundary step
            if (plan->resized.size() < size_t(n_sources)) plan->resized.resize(size_t(n_sources), nullptr);
            if (!plan->resized[size_t(i)]) {
                void* buf = nullptr;
                SM_CUDA(cudaSetDevice(ctx_->device));
                SM_CUDA(cudaMalloc(&buf, size_t(batch) * s.c * s.h * s.w * 2));
                plan->blobs.push_back(buf);
                plan->resized[size_t(i)] = static_cast<__half*>(buf);
            }
            cudaError_t e = k::resize_planes(t->ptr, plan->resized[size_t(i)], batch * s.c, t->h, t->w, s.h, s.w,
                                             cfg_.input_constraint == SMELTER_INPUT_FORCE_SCALE_LANCZOS ? 1 : 0, stream ? stream : ctx_->stream);
            if (e != cudaSuccess) return fail(SMELTER_ERR_CUDA, std::string("input resize: ") + cudaGetErrorString(e));
            *plan->src_slots[size_t(i)] = plan->resized[size_t(i)];
            continue;
        }
        if (t->n != batch || t->c != s.c || t->h != s.h || t->w != s.w)
            return fail(SMELTER_ERR_UNSUPPORTED_INPUT, "source " + std::to_string(i) + " is [" + std::to_string(t->n) + "," + std::to_string(t->c) + "," +
                                                           std::to_string(t->h) + "," + std::to_string(t->w) + "], graph input is [N," + std::to_string(s.c) + "," +
                                                           std::to_string(s.h) + "," + std::to_string(s.w) + "]");
        *plan->src_slots[size_t(i)] = t->ptr;
    }
    if (!stream) stream = ctx_->stream;
    SM_CUDA(cudaSetDevice(ctx_->device));
    // perf experiments (instrumented builds): SMELTER_CHAIN_TIMELINE=n prints the launch-chain stamps of the n-th encode of the process
    static const int chain_at = getenv("SMELTER_CHAIN_TIMELINE") ? atoi(getenv("SMELTER_CHAIN_TIMELINE")) : 0;
    static int chain_calls = 0;
    const bool chain_now = chain_at > 0 && ++chain_calls == chain_at;
    if (chain_now) { cudaStreamSynchronize(stream); k::conv_tc_chain_reset(); }
    struct ChainDump {
        bool on; cudaStream_t s;
        ~ChainDump() { if (on) { cudaStreamSynchronize(s); k::conv_tc_chain_dump(); } }
    } chain_dump{chain_now, stream};
    if (cfg_.use_cuda_graph) {
        for (auto& st : plan->steps) {  // source-image conversions run eagerly: their pointers change per call
            if (!st.boundary) continue;
            NvtxRange range(st.desc);
            cudaError_t e = st.run(stream);
            if (e != cudaSuccess) return fail(SMELTER_ERR_CUDA, "launch failed at '" + st.desc + "': " + cudaGetErrorString(e));
        }
        if (!plan->exec) {
            cudaGraph_t graph = nullptr;
            SM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            cudaError_t e = cudaSuccess;
            std::string where;
            for (auto& st : plan->steps) {
                if (st.boundary) continue;
                NvtxRange range(st.desc);
                e = st.run(stream);
                if (e != cudaSuccess) { where = st.desc; break; }
            }
            cudaError_t e2 = cudaStreamEndCapture(stream, &graph);
            if (e != cudaSuccess || e2 != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                return fail(SMELTER_ERR_CUDA, "capture failed at '" + where + "': " + cudaGetErrorString(e != cudaSuccess ? e : e2));
            }
            e = cudaGraphInstantiate(&plan->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(SMELTER_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
        SM_CUDA(cudaGraphLaunch(plan->exec, stream));
    } else {
        static const bool debug_sync = getenv("SMELTER_DEBUG_SYNC") != nullptr;  // fault isolation: sync after every kernel
        for (auto& st : plan->steps) {
            NvtxRange range(st.desc);
            cudaError_t e = st.run(stream);
            if (e == cudaSuccess && debug_sync) e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) {
                // an encode cut short leaves counters / accumulators that the kernels behind the failed step would have put back to zero
                if (plan->counters) { cudaStreamSynchronize(stream); cudaMemset(plan->counters, 0, plan->counter_bytes); }
                return fail(SMELTER_ERR_CUDA, "launch failed at '" + st.desc + "': " + cudaGetErrorString(e));
            }
        }
    }
    *result = &plan->result;
    return SMELTER_OK;
}

int ONNXGraph::num_launches(int batch, int* n) {
    Plan* plan = nullptr;
    int rc = plan_for(batch, &plan);
    if (rc) return rc;
    int count = 0;
    for (auto& st : plan->steps) count += st.launches;
    *n = count;
    return SMELTER_OK;
}

int ONNXGraph::profile(cudaStream_t stream, const Tensor* const* sources, int n_sources, int iters, std::vector<float>* ms,
                       std::vector<double>* flops, std::vector<double>* bytes, std::vector<int>* is_tensor) {
    if (!built_) return fail(SMELTER_ERR_INCONSISTENT_STATE, "profile before build");
    if (!sources || n_sources != int(input_values_.size())) return fail(SMELTER_ERR_INSUFFICIENT_INPUTS, "wrong number of sources");
    Plan* plan = nullptr;
    int rc = plan_for(sources[0]->n, &plan);
    if (rc) return rc;
    for (int i = 0; i < n_sources; ++i) {
        const ImageShape& s = plan->src_shapes[size_t(i)];
        const Tensor* t = sources[i];
        if (!t || !t->ptr || t->n != plan->batch || t->c != s.c || t->h != s.h || t->w != s.w) return fail(SMELTER_ERR_UNSUPPORTED_INPUT, "source shape mismatch");
        *plan->src_slots[size_t(i)] = t->ptr;
    }
    if (!stream) stream = ctx_->stream;
    SM_CUDA(cudaSetDevice(ctx_->device));
    const size_t n = plan->steps.size();
    std::vector<cudaEvent_t> ev(2 * n, nullptr);
    auto cleanup = [&]() { for (auto e : ev) if (e) cudaEventDestroy(e); };
    for (auto& e : ev) {
        cudaError_t ce = cudaEventCreate(&e);
        if (ce != cudaSuccess) { cleanup(); return fail(SMELTER_ERR_CUDA, cudaGetErrorString(ce)); }
    }
    ms->assign(n, 0.f);
    flops->resize(n); bytes->resize(n); is_tensor->resize(n);
    if (iters < 1) iters = 1;
    // In-situ timing: the whole plan is captured into one CUDA graph with an external event-record node before and
    // after every kernel, so each duration is measured under graph replay (no host launch gaps between kernels).
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    {
        cudaError_t ce = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
        std::string where;
        for (size_t i = 0; i < n && ce == cudaSuccess; ++i) {
            if (plan->steps[i].boundary) continue;
            ce = cudaEventRecordWithFlags(ev[2 * i], stream, cudaEventRecordExternal);
            if (ce == cudaSuccess) ce = plan->steps[i].run(stream);
            if (ce == cudaSuccess) ce = cudaEventRecordWithFlags(ev[2 * i + 1], stream, cudaEventRecordExternal);
            if (ce != cudaSuccess) where = plan->steps[i].desc;
        }
        cudaError_t ce2 = cudaStreamEndCapture(stream, &graph);
        if (ce == cudaSuccess) ce = ce2;
        if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            cleanup();
            return fail(SMELTER_ERR_CUDA, "profile capture failed" + (where.empty() ? std::string() : " at '" + where + "'") + ": " + cudaGetErrorString(ce));
        }
    }
    for (int it = 0; it < iters; ++it) {
        cudaError_t ce = cudaSuccess;
        for (size_t i = 0; i < n && ce == cudaSuccess; ++i) {
            if (!plan->steps[i].boundary) continue;
            cudaEventRecord(ev[2 * i], stream);
            ce = plan->steps[i].run(stream);
            cudaEventRecord(ev[2 * i + 1], stream);
        }
        if (ce == cudaSuccess) ce = cudaGraphLaunch(exec, stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
        if (ce != cudaSuccess) { cudaGraphExecDestroy(exec); cleanup(); return fail(SMELTER_ERR_CUDA, std::string("profile run failed: ") + cudaGetErrorString(ce)); }
        for (size_t i = 0; i < n; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
            (*ms)[i] += t / float(iters);
        }
    }
    cudaGraphExecDestroy(exec);
    for (size_t i = 0; i < n; ++i) {
        (*flops)[i] = plan->steps[i].flops;
        (*bytes)[i] = plan->steps[i].bytes;
        (*is_tensor)[i] = plan->steps[i].tensor ? 1 : 0;
    }
    cleanup();
    return SMELTER_OK;
}

int ONNXGraph::plan_dump(int batch, std::string* out) {
    Plan* plan = nullptr;
    int rc = plan_for(batch, &plan);
    if (rc) return rc;
    char line[512];
    snprintf(line, sizeof line, "# batch %d, %zu steps, activation arena %.1f MiB, weight arena %.1f MiB\n", batch, plan->steps.size(),
             plan->arena_bytes / 1048576.0, weight_bytes_ / 1048576.0);
    *out = line;
    for (auto& st : plan->steps) {
        snprintf(line, sizeof line, "%-90s flops=%.4g bytes=%.4g\n", st.desc.c_str(), st.flops, st.bytes);
        *out += line;
    }
    return SMELTER_OK;
}

int ONNXGraph::broadcast_weights(int root) {
    if (!built_) return fail(SMELTER_ERR_INCONSISTENT_STATE, "graph not built");
    if (weights_pending_ && !ctx_->nccl_comm)
        return fail(SMELTER_ERR_INCONSISTENT_STATE, "graph was built with deferWeights but the context has no NCCL communicator: nobody can fill the weight arena");
    const int rc = nccl_broadcast(ctx_, weight_arena_, weight_bytes_, root, ctx_->stream);
    if (rc == SMELTER_OK) weights_pending_ = false;
    return rc;
}

int ONNXGraph::weight_checksum(uint64_t* sum, uint64_t* bytes) {
    if (!built_) return fail(SMELTER_ERR_INCONSISTENT_STATE, "graph not built");
    SM_CUDA(cudaSetDevice(ctx_->device));
    unsigned long long* d = nullptr;
    SM_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    cudaError_t e = k::checksum64(weight_arena_, weight_bytes_, d, ctx_->stream);
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, ctx_->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx_->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SMELTER_ERR_CUDA, cudaGetErrorString(e));
    *sum = h;
    *bytes = weight_bytes_;
    return SMELTER_OK;
}

}  // namespace smelter
