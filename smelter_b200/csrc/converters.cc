// Per-op converters: one function per ONNX op, each turning a NodeProto into a Filter through the same
// five-call surface the reference's converters use (output(name:), shape(output:), tensor(name:),
// initTensor, addFilter — ONNXGraph.swift:259-285).  Error behaviour follows the reference's `guard`s
// (Sources/Smelter/Converters.swift; line ranges cited per converter).  Where the reference's behaviour is an
// MPS quirk and ONNX semantics differ, ONNX wins and the difference is noted.
#include <cmath>
#include <cstring>

#include "engine.h"

namespace smelter {

namespace {

using onnx::AttributeProto;
using onnx::NodeProto;
using onnx::TensorProto;

int err(int code, const NodeProto& n, const char* what) {
    return fail(code, n.op_type + " '" + (n.output.empty() ? n.name : n.output[0]) + "': " + what);
}

// ConvWeightArray decode (Converters.swift:82-89): fp32 via `.floats`, fp16 via raw_data; anything else is
// `.invalid` in the reference (silently wrong) — here an error.
bool weight_floats(const TensorProto& t, std::vector<float>* out) {
    if (t.data_type != onnx::DT_FLOAT && t.data_type != onnx::DT_FLOAT16) return false;
    return t.floats(out);
}

// ---- Conv / Gemm : ConvolutionConverter, Converters.swift:187-338 ------------------------------------------
int convert_conv(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    if (node.input.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "no input");
    const int input = g.output(node.input[0]);
    const ImageShape* in_shape = g.shape(node.input[0]);
    if (input < 0 || !in_shape) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");  // :189-191
    const TensorProto* weight = node.input.size() > 1 ? g.tensor(node.input[1]) : nullptr;
    if (!weight) return err(SMELTER_ERR_INSUFFICIENT_INPUTS, node, "weight initializer missing");  // :193-194
    const TensorProto* bias = node.input.size() > 2 ? g.tensor(node.input[2]) : nullptr;            // :196-199

    const bool is_gemm = node.op_type == "Gemm";
    const bool is_transpose = node.op_type == "ConvTranspose";  // :266-287
    bool have_k = false, have_d = false, have_s = false;
    Filter f;
    f.kind = FilterKind::Conv;
    f.op_type = node.op_type;
    f.is_gemm = is_gemm;
    float alpha = 1.f, beta = 1.f;
    int trans_a = 0, trans_b = 0;
    std::string auto_pad = "NOTSET";
    for (const AttributeProto& a : node.attribute) {  // :208-225
        if (a.name == "dilations" && a.ints.size() >= 2) { f.dil_h = int(a.ints[0]); f.dil_w = int(a.ints[1]); have_d = true; }
        else if (a.name == "strides" && a.ints.size() >= 2) { f.stride_h = int(a.ints[0]); f.stride_w = int(a.ints[1]); have_s = true; }
        else if (a.name == "group") f.groups = int(a.i);
        else if (a.name == "pads" && a.ints.size() >= 4) { for (int i = 0; i < 4; ++i) f.pads[i] = int(a.ints[size_t(i)]); }
        else if (a.name == "kernel_shape" && a.ints.size() >= 2) { f.k_h = int(a.ints[0]); f.k_w = int(a.ints[1]); have_k = true; }
        else if (a.name == "auto_pad") auto_pad = std::string(a.s);
        else if (a.name == "output_padding" && a.ints.size() >= 2) { f.out_pad_h = int(a.ints[0]); f.out_pad_w = int(a.ints[1]); }  // :220-221
        else if (a.name == "alpha") alpha = a.f;
        else if (a.name == "beta") beta = a.f;
        else if (a.name == "transA") trans_a = int(a.i);
        else if (a.name == "transB") trans_b = int(a.i);
    }
    if (is_gemm) {  // :228-232
        f.k_h = f.k_w = 1; f.dil_h = f.dil_w = 1; f.stride_h = f.stride_w = 1;
        have_k = have_d = have_s = true;
    }
    if (!have_k || !have_d || !have_s) return err(SMELTER_ERR_NOT_ENOUGH_ATTRIBUTES, node, "kernel_shape/dilations/strides required");  // :234-237

    std::vector<float> wv;
    if (!weight_floats(*weight, &wv)) return err(SMELTER_ERR_UNSUPPORTED, node, "weights must be FLOAT or FLOAT16");
    const bool mps = g.format() == SMELTER_FORMAT_MPS_FLAVOR;
    if (is_gemm) {
        // Reference: W = [out, in], alpha/beta/transA/transB ignored.  Here ONNX semantics:
        // Y = alpha * A' * B' + beta * C; transA is not expressible on an image input.
        if (weight->dims.size() != 2) return err(SMELTER_ERR_UNSUPPORTED, node, "Gemm weight must be rank 2");
        if (trans_a) return err(SMELTER_ERR_UNSUPPORTED, node, "Gemm transA=1 is not supported");
        const int d0 = int(weight->dims[0]), d1 = int(weight->dims[1]);
        if (size_t(d0) * d1 != wv.size()) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "weight size mismatch");
        if (trans_b) { f.c_out = d0; f.c_in_g = d1; f.w = std::move(wv); }
        else {
            f.c_out = d1; f.c_in_g = d0;
            f.w.resize(wv.size());
            for (int i = 0; i < d0; ++i)
                for (int o = 0; o < d1; ++o) f.w[size_t(o) * d0 + i] = wv[size_t(i) * d1 + o];
        }
        if (alpha != 1.f) for (float& v : f.w) v *= alpha;
    } else {
        if (weight->dims.size() != 4) return err(SMELTER_ERR_UNSUPPORTED, node, "Conv weight must be rank 4 (2-D convolution)");
        int o, i, kh, kw;
        if (mps) { o = int(weight->dims[0]); kh = int(weight->dims[1]); kw = int(weight->dims[2]); i = int(weight->dims[3]); }  // :40-44
        else if (is_transpose) { i = int(weight->dims[0]); o = int(weight->dims[1]); kh = int(weight->dims[2]); kw = int(weight->dims[3]); }  // :48-50
        else { o = int(weight->dims[0]); i = int(weight->dims[1]); kh = int(weight->dims[2]); kw = int(weight->dims[3]); }      // :46-54
        if (kh != f.k_h || kw != f.k_w) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "kernel_shape disagrees with the weight dims");
        if (size_t(o) * i * kh * kw != wv.size()) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "weight size mismatch");
        f.c_out = o; f.c_in_g = i;
        if (mps) f.w = std::move(wv);  // already OHWI (ONNX2MPS.py:75), the reference skips the re-layout too (:91)
        else {
            f.w.resize(wv.size());
            reformat_conv_weight(wv.data(), f.w.data(), 4, o, i, kh, kw, is_transpose);  // :94-120 (ConvTranspose: IOHW -> OHWI + 180 degree flip)
        }
    }
    if (in_shape->c != f.c_in_g * f.groups) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "input channels disagree with the weight dims");
    // Gemm runs as a 1x1 fully-connected layer over [N, C, 1, 1] (:228-232).  An un-flattened image (H*W > 1) would need the NCHW
    // flatten order the weights were trained with; Flatten / Reshape produce exactly that [N, C*H*W, 1, 1] value, so demand it
    // instead of silently reading NHWC memory in the wrong order.
    if (is_gemm && in_shape->h * in_shape->w != 1)
        return err(SMELTER_ERR_UNSUPPORTED, node, "Gemm input must be flattened to [N, C, 1, 1] (insert Flatten / Reshape)");
    f.bias.assign(size_t(f.c_out), 0.f);
    if (bias) {  // :126-135 bias always widened to fp32
        std::vector<float> bv;
        if (!weight_floats(*bias, &bv) || int(bv.size()) != f.c_out) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "bias must be FLOAT/FLOAT16 of length Cout");
        for (int i = 0; i < f.c_out; ++i) f.bias[size_t(i)] = bv[size_t(i)] * (is_gemm ? beta : 1.f);
    }
    if (is_transpose) {
        // executed as a stride-1 convolution over the zero-stuffed input (engine.cc); ONNX output size incl. dilation and output_padding
        if (f.groups != 1) return err(SMELTER_ERR_UNSUPPORTED, node, "grouped ConvTranspose");
        if (auto_pad != "NOTSET" && auto_pad != "VALID") return err(SMELTER_ERR_UNSUPPORTED, node, "ConvTranspose auto_pad");
        const int lo_h = f.dil_h * (f.k_h - 1) - f.pads[0], hi_h = f.dil_h * (f.k_h - 1) - f.pads[2] + f.out_pad_h;
        const int lo_w = f.dil_w * (f.k_w - 1) - f.pads[1], hi_w = f.dil_w * (f.k_w - 1) - f.pads[3] + f.out_pad_w;
        if (lo_h < 0 || hi_h < 0 || lo_w < 0 || hi_w < 0) return err(SMELTER_ERR_UNSUPPORTED, node, "ConvTranspose pads larger than the dilated kernel");
        ImageShape out;
        out.c = f.c_out;
        out.h = conv_output_size(in_shape->h, f.k_h, f.stride_h, f.dil_h, f.pads[0], f.pads[2], f.out_pad_h, true);  // ONNXConvolutionPadding.swift:97-103
        out.w = conv_output_size(in_shape->w, f.k_w, f.stride_w, f.dil_w, f.pads[1], f.pads[3], f.out_pad_w, true);
        if (out.h <= 0 || out.w <= 0) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "empty output");
        f.transposed = true;
        f.tr_stride_h = f.stride_h; f.tr_stride_w = f.stride_w;
        f.stride_h = f.stride_w = 1;  // the convolution that actually runs
        f.pads[0] = lo_h; f.pads[1] = lo_w; f.pads[2] = hi_h; f.pads[3] = hi_w;  // border of the stuffed image
        f.in = {input};
        return g.addFilter(std::move(f), out, node.output);
    }
    if (!is_gemm && auto_pad != "NOTSET" && auto_pad != "VALID") {
        // reference ignores auto_pad; ONNX semantics honoured for SAME_*
        for (int d = 0; d < 2; ++d) {
            const int in = d == 0 ? in_shape->h : in_shape->w, k = d == 0 ? f.k_h : f.k_w, s = d == 0 ? f.stride_h : f.stride_w,
                      dl = d == 0 ? f.dil_h : f.dil_w;
            const int out = (in + s - 1) / s;
            int total = (out - 1) * s + dl * (k - 1) + 1 - in;
            if (total < 0) total = 0;
            const int lo = auto_pad == "SAME_UPPER" ? total / 2 : total - total / 2;
            f.pads[d] = lo; f.pads[d + 2] = total - lo;
        }
    }
    ImageShape out;
    out.c = f.c_out;
    if (is_gemm) {  // :323-330 FC output is 1x1
        out.h = out.w = 1;
    } else {       // :311-322 paddedSize
        out.h = conv_output_size(in_shape->h, f.k_h, f.stride_h, f.dil_h, f.pads[0], f.pads[2], 0, false);
        out.w = conv_output_size(in_shape->w, f.k_w, f.stride_w, f.dil_w, f.pads[1], f.pads[3], 0, false);
        if (out.h <= 0 || out.w <= 0) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "empty output");
    }
    f.in = {input};
    return g.addFilter(std::move(f), out, node.output);  // :332-336
}

// ---- unary family -----------------------------------------------------------------------------------------
int add_unary(ONNXGraph& g, const NodeProto& node, int kind, float alpha, float beta) {
    if (node.input.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "no input");
    const int input = g.output(node.input[0]);
    const ImageShape* s = g.shape(node.input[0]);
    if (input < 0 || !s) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");
    Filter f;
    f.kind = FilterKind::Unary;
    f.op_type = node.op_type;
    f.sub = kind; f.alpha = alpha; f.beta = beta;
    f.in = {input};
    return g.addFilter(std::move(f), *s, node.output);
}
int convert_relu(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_RELU, 0, 0); }        // :342-359
int convert_sigmoid(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_SIGMOID, 0, 0); }  // :466-476
int convert_tanh(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_TANH, 0, 0); }        // :1124-1139
int convert_abs(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_ABS, 0, 0); }          // :1056-1071
int convert_exp(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_EXP, 0, 0); }          // :411-428
int convert_log(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_LOG, 0, 0); }          // :1142-1157
int convert_softplus(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_SOFTPLUS, 0, 0); }  // :1090-1105
int convert_softsign(ONNXGraph& g, int ni) { return add_unary(g, g.node(ni), k::UN_SOFTSIGN, 0, 0); }  // :1107-1122
int convert_elu(ONNXGraph& g, int ni) {  // :386-408 — alpha attribute required
    const NodeProto& node = g.node(ni);
    const AttributeProto* a = node.attr("alpha");
    if (!a) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "alpha attribute required");
    return add_unary(g, node, k::UN_ELU, a->f, 0);
}
int convert_hard_sigmoid(ONNXGraph& g, int ni) {  // :1073-1088 (reference ignores alpha/beta; ONNX defaults honoured)
    const NodeProto& node = g.node(ni);
    float alpha = 0.2f, beta = 0.5f;
    if (const AttributeProto* a = node.attr("alpha")) alpha = a->f;
    if (const AttributeProto* b = node.attr("beta")) beta = b->f;
    return add_unary(g, node, k::UN_HARD_SIGMOID, alpha, beta);
}
int convert_prelu(ONNXGraph& g, int ni) {  // :361-384 — a single scalar slope read from raw_data
    const NodeProto& node = g.node(ni);
    const TensorProto* a = node.input.size() >= 2 ? g.tensor(node.input[1]) : nullptr;
    if (!a) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "slope initializer missing");
    std::vector<float> v;
    if (!a->floats(&v) || v.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "slope initializer unreadable");
    return add_unary(g, node, k::UN_LEAKY_RELU, v[0], 0);
}
// Clip: NOT in the reference's registry (ONNXGraph.swift:110-155) — documented extension, required by
// MobileNetV2's ReLU6 (SURVEY §2.2 "Ops the configs need that the reference does NOT have").
int convert_clip(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    float lo = -3.402823466e38f, hi = 3.402823466e38f;
    if (const AttributeProto* a = node.attr("min")) lo = a->f;
    if (const AttributeProto* a = node.attr("max")) hi = a->f;
    for (size_t i = 1; i < node.input.size() && i < 3; ++i) {  // opset >= 11: min/max as initializer inputs
        if (node.input[i].empty()) continue;
        const TensorProto* t = g.tensor(node.input[i]);
        std::vector<float> v;
        if (!t || !t->floats(&v) || v.empty()) return err(SMELTER_ERR_INSUFFICIENT_INPUTS, node, "min/max must be initializers");
        (i == 1 ? lo : hi) = v[0];
    }
    return add_unary(g, node, k::UN_CLIP, lo, hi);
}

// ---- binary family: both operands must be image nodes, no broadcasting (Converters.swift:430-464, 1177-1211)
int add_binary(ONNXGraph& g, const NodeProto& node, int kind) {
    if (node.input.size() < 2) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "two inputs required");
    const int a = g.output(node.input[0]), b = g.output(node.input[1]);
    const ImageShape* sa = g.shape(node.input[0]);
    const ImageShape* sb = g.shape(node.input[1]);
    if (a < 0 || b < 0 || !sa || !sb) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "both operands must be image nodes");
    Filter f;
    f.kind = FilterKind::Binary;
    f.op_type = node.op_type;
    f.sub = kind;
    f.in = {a, b};
    if (!(*sa == *sb)) {
        // One pixel per image against a whole image ([N, C, 1, 1] x [N, C, H, W], the gate of a squeeze-and-excitation block): MPS
        // arithmetic nodes broadcast a 1 x 1 source (Converters.swift:430-464 passes the images straight through).  Nothing else.
        const bool b_pixel = sb->c == sa->c && sb->h == 1 && sb->w == 1;
        const bool a_pixel = sa->c == sb->c && sa->h == 1 && sa->w == 1 && (kind == k::BIN_ADD || kind == k::BIN_MUL);
        if (!b_pixel && !a_pixel) return err(SMELTER_ERR_UNSUPPORTED, node, "operand shapes differ (only a [C,1,1] operand is broadcast)");
        f.bcast = true;
        if (!b_pixel) { f.in = {b, a}; return g.addFilter(std::move(f), *sb, node.output); }
    }
    return g.addFilter(std::move(f), *sa, node.output);
}
int convert_add(ONNXGraph& g, int ni) { return add_binary(g, g.node(ni), k::BIN_ADD); }
int convert_sub(ONNXGraph& g, int ni) { return add_binary(g, g.node(ni), k::BIN_SUB); }
int convert_mul(ONNXGraph& g, int ni) { return add_binary(g, g.node(ni), k::BIN_MUL); }
int convert_div(ONNXGraph& g, int ni) { return add_binary(g, g.node(ni), k::BIN_DIV); }

// ---- Upsample: UpsampleConverter, Converters.swift:478-552 --------------------------------------------------
int convert_upsample(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    if (node.input.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "no input");
    const int input = g.output(node.input[0]);
    const ImageShape* s = g.shape(node.input[0]);
    if (input < 0 || !s) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");
    std::string mode;
    bool have_mode = false, have_scales = false;
    int sh = 1, sw = 1;
    for (const AttributeProto& a : node.attribute) {
        if (a.name == "mode") { mode = std::string(a.s); have_mode = true; }
        else if (a.name == "scales" && a.floats.size() >= 4) { sh = int(a.floats[2]); sw = int(a.floats[3]); have_scales = true; }  // :498-499 truncation
    }
    if (!have_scales) {  // :505-514 opset 9: scales tensor through `.integers`
        const TensorProto* t = node.input.size() >= 2 ? g.tensor(node.input[1]) : nullptr;
        std::vector<int64_t> iv;
        if (!t || t->dims.empty() || t->dims[0] != 4 || !t->integers(&iv) || iv.size() < 4)
            return err(SMELTER_ERR_NOT_ENOUGH_ATTRIBUTES, node, "scales missing");
        sh = int(iv[2]); sw = int(iv[3]);
    }
    if (!have_mode) return err(SMELTER_ERR_NOT_ENOUGH_ATTRIBUTES, node, "mode missing");  // :516-518
    if (sh < 1 || sw < 1) return err(SMELTER_ERR_UNSUPPORTED, node, "scale factors must be >= 1");
    Filter f;
    f.kind = FilterKind::Upsample;
    f.op_type = node.op_type;
    if (mode == "nearest") f.sub = k::UP_NEAREST;
    else if (mode == "bilinear" || mode == "linear") f.sub = k::UP_BILINEAR;
    else return fail(SMELTER_ERR_UNKNOWN_NODE_OP_TYPE, node.op_type);  // :537-538
    f.scale_h = sh; f.scale_w = sw;
    f.align_corners = g.config().bilinear_align_corners;  // ONNXGraph.swift:118-120
    f.in = {input};
    return g.addFilter(std::move(f), ImageShape{s->c, s->h * sh, s->w * sw}, node.output);  // :541-550
}

// ---- Concat: ConcatConverter, Converters.swift:554-574 ------------------------------------------------------
// Reference: `axis` ignored, output depth hard-coded to 2x the first input.  Here: ONNX semantics,
// channel axis only, output channels = sum of inputs.
int convert_concat(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    Filter f;
    f.kind = FilterKind::Concat;
    f.op_type = node.op_type;
    ImageShape out{0, 0, 0};
    for (size_t i = 0; i < node.input.size(); ++i) {
        const int v = g.output(node.input[i]);
        const ImageShape* s = g.shape(node.input[i]);
        if (v < 0 || !s) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");
        if (i == 0) { out.h = s->h; out.w = s->w; }
        else if (s->h != out.h || s->w != out.w) return err(SMELTER_ERR_UNSUPPORTED, node, "spatial sizes differ");
        out.c += s->c;
        f.in.push_back(v);
    }
    if (f.in.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "no inputs");
    if (const AttributeProto* a = node.attr("axis")) {
        if (a->i != 1 && a->i != -3) return err(SMELTER_ERR_UNSUPPORTED, node, "only channel concatenation (axis=1) is supported");
    }
    return g.addFilter(std::move(f), out, node.output);
}

// ---- pools: Converters.swift:578-695 --------------------------------------------------------------------------
int convert_global_avgpool(ONNXGraph& g, int ni) {  // :578-605
    const NodeProto& node = g.node(ni);
    if (node.input.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "no input");
    const int input = g.output(node.input[0]);
    const ImageShape* s = g.shape(node.input[0]);
    if (input < 0 || !s) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");
    Filter f;
    f.kind = FilterKind::GlobalAvgPool;
    f.op_type = node.op_type;
    f.in = {input};
    return g.addFilter(std::move(f), ImageShape{s->c, 1, 1}, node.output);
}
// ReduceMean over the spatial axes: NOT in the reference's registry -- how newer torch exporters write x.mean([2, 3]) (MNASNet, the
// squeeze of squeeze-and-excitation blocks).  axes must be {2, 3} (attribute, or initializer input from opset 18); keepdims either way.
int convert_reduce_mean(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    std::vector<int64_t> axes;
    if (const AttributeProto* a = node.attr("axes")) axes.assign(a->ints.begin(), a->ints.end());
    else if (node.input.size() >= 2) { const TensorProto* t = g.tensor(node.input[1]); if (!t || !t->integers(&axes)) axes.clear(); }
    for (auto& v : axes) if (v < 0) v += 4;
    std::sort(axes.begin(), axes.end());
    if (axes != std::vector<int64_t>{2, 3}) return err(SMELTER_ERR_UNSUPPORTED, node, "only ReduceMean over axes {2, 3} (global average pooling)");
    return convert_global_avgpool(g, ni);
}
int convert_pool(ONNXGraph& g, int ni) {  // AveragePool :607-650, MaxPool :652-695
    const NodeProto& node = g.node(ni);
    const AttributeProto* ks = node.attr("kernel_shape");
    const AttributeProto* pads = node.attr("pads");
    const AttributeProto* strides = node.attr("strides");
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    // The reference guards on all three attributes at once and throws noSuchOutput (:654-661).  Here `pads` and `strides` fall back
    // to their ONNX defaults (0 and 1) when absent -- torch's exporter writes AdaptiveAvgPool2d without `pads` (VGG, AlexNet) -- and
    // only a missing input or kernel_shape is the reference's error (DESIGN.md "ONNX semantics win").
    if (input < 0 || !s || !ks || ks->ints.size() < 2 || (pads && pads->ints.size() < 2) || (strides && strides->ints.size() < 2))
        return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and a kernel_shape attribute (pads / strides default to 0 / 1)");
    Filter f;
    f.kind = FilterKind::Pool;
    f.op_type = node.op_type;
    f.sub = node.op_type == "MaxPool" ? 1 : 0;
    f.k_h = int(ks->ints[0]); f.k_w = int(ks->ints[1]);
    f.stride_h = strides ? int(strides->ints[0]) : 1; f.stride_w = strides ? int(strides->ints[1]) : 1;
    f.pool_pad_h = pads ? int(pads->ints[0]) : 0; f.pool_pad_w = pads ? int(pads->ints[1]) : 0;  // symmetric from pads[0..1]
    if (f.k_h < 1 || f.k_w < 1 || f.stride_h < 1 || f.stride_w < 1) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "bad kernel/stride");
    ImageShape out{s->c, pool_output_size(s->h, f.k_h, f.stride_h, f.pool_pad_h), pool_output_size(s->w, f.k_w, f.stride_w, f.pool_pad_w)};
    if (out.h <= 0 || out.w <= 0) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "empty output");
    f.in = {input};
    return g.addFilter(std::move(f), out, node.output);
}

// ---- Softmax / LogSoftmax: Converters.swift:697-714, 1213-1231 — explicit axis == 1 required ------------------
int convert_softmax(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    const AttributeProto* axis = node.attr("axis");
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    if (input < 0 || !s || !axis || axis->i != 1) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and axis == 1");
    Filter f;
    f.kind = FilterKind::Softmax;
    f.op_type = node.op_type;
    f.sub = node.op_type == "LogSoftmax" ? 1 : 0;
    f.in = {input};
    return g.addFilter(std::move(f), *s, node.output);
}

// ---- Constant: ConstantConverter, Converters.swift:716-727 ------------------------------------------------------
int convert_constant(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    const AttributeProto* v = node.attr("value");
    if (node.output.empty() || !v || !v->has_t) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "value tensor missing");
    g.initTensor(node.output[0], &v->t);
    return SMELTER_OK;
}

// ---- BatchNormalization: Converters.swift:730-827 (epsilon honoured) ----------------------------------
int convert_batchnorm(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    if (node.input.empty()) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "no input");
    const int input = g.output(node.input[0]);
    const ImageShape* s = g.shape(node.input[0]);
    if (input < 0 || !s) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");
    if (node.input.size() < 5) return err(SMELTER_ERR_INSUFFICIENT_INPUTS, node, "gamma, beta, mean, variance required");
    const TensorProto* t[4];
    std::vector<float> v[4];
    for (int i = 0; i < 4; ++i) {
        t[i] = g.tensor(node.input[size_t(i) + 1]);
        if (!t[i] || !t[i]->floats(&v[i]) || int(v[i].size()) != s->c) return err(SMELTER_ERR_INSUFFICIENT_INPUTS, node, "parameter initializer missing or wrong length");
    }
    float eps = 1e-5f;
    if (const AttributeProto* a = node.attr("epsilon")) eps = a->f;
    Filter f;
    f.kind = FilterKind::BatchNorm;
    f.op_type = node.op_type;
    f.eps = eps;
    f.p0.resize(size_t(s->c));
    f.p1.resize(size_t(s->c));
    for (int c = 0; c < s->c; ++c) {  // y = gamma (x - mean) / sqrt(var + eps) + beta
        const float sc = v[0][size_t(c)] / std::sqrt(v[3][size_t(c)] + eps);
        f.p0[size_t(c)] = sc;
        f.p1[size_t(c)] = v[1][size_t(c)] - v[2][size_t(c)] * sc;
    }
    f.in = {input};
    return g.addFilter(std::move(f), *s, node.output);
}

// ---- InstanceNormalization: Converters.swift:992-1054 (epsilon honoured) -----------------------------------------
int convert_instancenorm(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    const TensorProto* gamma = node.input.size() >= 3 ? g.tensor(node.input[1]) : nullptr;
    const TensorProto* beta = node.input.size() >= 3 ? g.tensor(node.input[2]) : nullptr;
    Filter f;
    if (input < 0 || !s || !gamma || !beta || !gamma->floats(&f.p0) || !beta->floats(&f.p1))
        return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and gamma/beta initializers");  // :994-999
    if (int(f.p0.size()) != s->c || int(f.p1.size()) != s->c) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "gamma/beta length != channels");
    f.kind = FilterKind::InstanceNorm;
    f.op_type = node.op_type;
    f.eps = 1e-5f;
    if (const AttributeProto* a = node.attr("epsilon")) f.eps = a->f;
    f.in = {input};
    return g.addFilter(std::move(f), *s, node.output);
}

// ---- custom_group_norm: GroupNormConverter, Converters.swift:1273-1300 — inputs X, groups (int tensor), gamma, beta ----------
// (the op the reference's own exporter emits for torch.nn.GroupNorm; MPSCNNGroupNormalizationNode, epsilon = MPS default)
int convert_groupnorm(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    const TensorProto* groups = node.input.size() >= 4 ? g.tensor(node.input[1]) : nullptr;
    const TensorProto* gamma = node.input.size() >= 4 ? g.tensor(node.input[2]) : nullptr;
    const TensorProto* beta = node.input.size() >= 4 ? g.tensor(node.input[3]) : nullptr;
    Filter f;
    std::vector<int64_t> gi;
    if (input < 0 || !s || !groups || !gamma || !beta || !groups->integers(&gi) || gi.empty() || !gamma->floats(&f.p0) || !beta->floats(&f.p1))
        return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and groups/gamma/beta initializers");  // :1275-1281
    const int ng = int(gi[0]);
    if (int(f.p0.size()) != s->c || int(f.p1.size()) != s->c) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "gamma/beta length != channels");
    if (ng < 1 || s->c % ng != 0) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "channels must be divisible by groups");
    f.kind = FilterKind::InstanceNorm;
    f.op_type = node.op_type;
    f.sub = s->c / ng;  // channels per group (InstanceNormalization leaves this 0 = one channel per group)
    f.eps = 1e-5f;
    if (const AttributeProto* a = node.attr("epsilon")) f.eps = a->f;
    f.in = {input};
    return g.addFilter(std::move(f), *s, node.output);
}

// ---- Pow: PowConverter, Converters.swift:1160-1175 ------------------------------------------------------------------------
// The reference builds MPSCNNNeuronPowerNode with its default parameters and never reads the exponent.  Here the
// ONNX meaning: the exponent must be a one-element initializer (no tensor-valued exponents on an image path).
int convert_pow(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    float e = 1.f;
    if (node.input.size() >= 2) {
        const TensorProto* t = g.tensor(node.input[1]);
        std::vector<float> v;
        if (!t || !t->floats(&v) || v.size() != 1) return err(SMELTER_ERR_UNSUPPORTED, node, "exponent must be a one-element initializer");
        e = v[0];
    }
    return add_unary(g, node, k::UN_POW, e, 0);
}

// ---- Reshape / Flatten: Converters.swift:830-915 -------------------------------------------------------------------
// Reference reads shape[0..2] as (W,H,C) and has no batch axis.  Here: ONNX semantics on NCHW
// with the batch axis carried implicitly — the target must keep dim 0 (N, 0 or -1 resolving to N).
int convert_reshape(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    const TensorProto* st = node.input.size() >= 2 ? g.tensor(node.input[1]) : nullptr;
    std::vector<int64_t> dims;
    if (input < 0 || !s || !st || !st->integers(&dims)) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and a shape initializer");  // :832-836
    if (dims.size() < 2 || dims.size() > 4) return err(SMELTER_ERR_UNSUPPORTED, node, "target rank must be 2..4");
    const int64_t total = int64_t(s->c) * s->h * s->w;
    const int64_t in_dims[4] = {-1 /*batch*/, s->c, s->h, s->w};
    int64_t tgt[3] = {1, 1, 1};
    int infer = -1;
    int64_t known = 1;
    for (size_t i = 1; i < dims.size(); ++i) {
        int64_t d = dims[i];
        if (d == 0) d = in_dims[i];
        if (d == -1) { if (infer >= 0) return err(SMELTER_ERR_UNSUPPORTED, node, "two inferred dims"); infer = int(i) - 1; d = 1; }
        else known *= d;
        tgt[i - 1] = d;
    }
    if (dims[0] == -1 && infer < 0) { /* batch inferred: fine */ }
    if (infer >= 0) {
        if (known == 0 || total % known) return err(SMELTER_ERR_INCONSISTENT_STATE, node, "cannot infer dim");
        tgt[infer] = total / known;
    } else if (known != total) {
        return err(SMELTER_ERR_UNSUPPORTED, node, "reshape would change the batch axis");
    }
    const ImageShape out{int(tgt[0]), int(tgt[1]), int(tgt[2])};
    if (s->h == 1 && s->w == 1 && out.h == 1 && out.w == 1) return g.addAlias(input, out, node.output);  // same bytes: a view
    Filter f;
    f.kind = FilterKind::Reshape;
    f.op_type = node.op_type;
    f.in = {input};
    return g.addFilter(std::move(f), out, node.output);
}
int convert_flatten(ONNXGraph& g, int ni) {  // :879-915 — axis attribute required; only axis 1 (reference: fatalError otherwise)
    const NodeProto& node = g.node(ni);
    const AttributeProto* axis = node.attr("axis");
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    if (node.attribute.empty() || input < 0 || !s || !axis) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and an axis attribute");
    if (axis->i != 1) return err(SMELTER_ERR_UNSUPPORTED, node, "other axes are not supported");
    if (s->h == 1 && s->w == 1) return g.addAlias(input, ImageShape{s->c, 1, 1}, node.output);  // [N,C,1,1] -> [N,C]: a view
    Filter f;
    f.kind = FilterKind::Reshape;
    f.op_type = node.op_type;
    f.in = {input};
    return g.addFilter(std::move(f), ImageShape{s->c * s->h * s->w, 1, 1}, node.output);
}

// ---- Dropout: Converters.swift:918-939.  The reference passes `ratio` as MPS keepProbability;
// at inference Dropout is the identity, which is what this engine does.  Identity is an extension.
int convert_identity(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    // torch's exporter de-duplicates equal initializers (all-zero biases, ...) through Identity nodes: an alias of a weight tensor
    if (node.op_type == "Identity" && !node.input.empty() && !node.output.empty() && g.output(node.input[0]) < 0)
        if (const onnx::TensorProto* t = g.tensor(node.input[0])) { g.initTensor(node.output[0], t); return SMELTER_OK; }
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    if (input < 0 || !s) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "input is not an image node");
    if (node.op_type == "Dropout" && (node.attribute.empty() || !node.attr("ratio")))
        return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "ratio attribute required");  // :920-925
    std::vector<std::string> outs(node.output.begin(), node.output.begin() + (node.output.empty() ? 0 : 1));
    return g.addAlias(input, *s, outs);
}

// ---- Pad: PaddingConverter, Converters.swift:942-989 (opset <= 10 attribute pads; `value` honoured) -------------
int convert_pad(ONNXGraph& g, int ni) {
    const NodeProto& node = g.node(ni);
    const int input = node.input.empty() ? -1 : g.output(node.input[0]);
    const ImageShape* s = node.input.empty() ? nullptr : g.shape(node.input[0]);
    const AttributeProto* pads = node.attr("pads");
    if (input < 0 || !s || !pads || pads->ints.size() < 8) return err(SMELTER_ERR_NO_SUCH_OUTPUT, node, "needs an image input and an 8-long pads attribute");
    std::string mode = "constant";
    if (const AttributeProto* m = node.attr("mode")) mode = std::string(m->s);
    Filter f;
    f.kind = FilterKind::Pad;
    f.op_type = node.op_type;
    if (mode == "constant") f.sub = k::PAD_CONSTANT;
    else if (mode == "reflect") f.sub = k::PAD_REFLECT;
    else if (mode == "edge") f.sub = k::PAD_EDGE;
    else return err(SMELTER_ERR_INCONSISTENT_STATE, node, "unknown mode");  // :968-969
    if (const AttributeProto* v = node.attr("value")) f.alpha = v->f;
    const auto& p = pads->ints;  // NCHW begin x4, end x4
    if (p[0] || p[4]) return err(SMELTER_ERR_UNSUPPORTED, node, "batch padding");
    if (p[1] || p[5]) return err(SMELTER_ERR_UNSUPPORTED, node, "channel padding");
    f.pads[0] = int(p[2]); f.pads[1] = int(p[3]); f.pads[2] = int(p[6]); f.pads[3] = int(p[7]);
    for (int i = 0; i < 4; ++i)
        if (f.pads[i] < 0) return err(SMELTER_ERR_UNSUPPORTED, node, "negative pads");
    if (f.sub == k::PAD_REFLECT && (f.pads[0] >= s->h || f.pads[2] >= s->h || f.pads[1] >= s->w || f.pads[3] >= s->w))
        return err(SMELTER_ERR_UNSUPPORTED, node, "reflect pad must be smaller than the image");
    f.in = {input};
    return g.addFilter(std::move(f), ImageShape{s->c, s->h + f.pads[0] + f.pads[2], s->w + f.pads[1] + f.pads[3]}, node.output);  // :978-987
}

}  // namespace

// ONNXGraph.swift:110-155: all 36 entries of the reference's registry, plus Clip and Identity.  Anything else is
// unknownNodeOpType, exactly like an op the reference lacks.
void ONNXGraph::registerBuiltins() {
    registerConverter("Conv", convert_conv);
    registerConverter("Gemm", convert_conv);
    registerConverter("Relu", convert_relu);
    registerConverter("Elu", convert_elu);
    registerConverter("Add", convert_add);
    registerConverter("Sub", convert_sub);
    registerConverter("Sigmoid", convert_sigmoid);
    registerConverter("Upsample", convert_upsample);
    registerConverter("HardSigmoid", convert_hard_sigmoid);
    registerConverter("Concat", convert_concat);
    registerConverter("AveragePool", convert_pool);
    registerConverter("MaxPool", convert_pool);
    registerConverter("Softmax", convert_softmax);
    registerConverter("LogSoftmax", convert_softmax);
    registerConverter("Constant", convert_constant);
    registerConverter("Mul", convert_mul);
    registerConverter("Div", convert_div);
    registerConverter("GlobalAveragePool", convert_global_avgpool);
    registerConverter("Abs", convert_abs);
    registerConverter("Softplus", convert_softplus);
    registerConverter("Softsign", convert_softsign);
    registerConverter("Tanh", convert_tanh);
    registerConverter("PRelu", convert_prelu);
    registerConverter("BatchNormalization", convert_batchnorm);
    registerConverter("Dropout", convert_identity);
    registerConverter("InstanceNormalization", convert_instancenorm);
    registerConverter("custom_group_norm", convert_groupnorm);
    registerConverter("Pow", convert_pow);
    registerConverter("ConvTranspose", convert_conv);
    registerConverter("Log", convert_log);
    registerConverter("Exp", convert_exp);
    registerConverter("Reshape", convert_reshape);
    registerConverter("Flatten", convert_flatten);
    registerConverter("Pad", convert_pad);
    // extensions (not in the reference registry)
    registerConverter("Clip", convert_clip);
    registerConverter("Identity", convert_identity);
    registerConverter("ReduceMean", convert_reduce_mean);
}

}  // namespace smelter
