// extern "C" surface declared in include/smelter_b200.h.  Thin: argument checks, then the C++ engine.
#include <cmath>
#include <cstring>
#include <new>

#include "engine.h"

namespace smelter {
const std::string& last_error_string();
}

using namespace smelter;

struct smelter_context { Context c; };
struct smelter_graph { ONNXGraph* g; };
struct smelter_tensor { Tensor t; };

#define ARG(cond)                                                                      \
    do {                                                                               \
        if (!(cond)) return fail(SMELTER_ERR_INVALID_ARGUMENT, "invalid argument: " #cond); \
    } while (0)

extern "C" {

const char* smelter_last_error(void) { return last_error_string().c_str(); }
int32_t smelter_abi_version(void) { return SMELTER_B200_ABI_VERSION; }

void smelter_config_default(smelter_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->input_constraint = SMELTER_INPUT_NONE;   // ONNXGraph.swift:28
    cfg->bilinear_align_corners = 1;              // ONNXGraph.swift:20
    cfg->n_dims = 0;                              // ONNXGraph.swift:30
    cfg->enable_fusion = 1;
    cfg->use_cuda_graph = 1;
    cfg->defer_weights = 0;
    cfg->sm_share = 1;
}

// ---- context ----------------------------------------------------------------------------------------------
int32_t smelter_context_create(int32_t device, void* cuda_stream, smelter_context** out) {
    ARG(out);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SMELTER_ERR_CUDA, std::string("no CUDA device available (this engine has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(SMELTER_ERR_INVALID_ARGUMENT, "device index out of range");
    SM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SMELTER_ERR_CUDA, "device " + std::to_string(device) + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                          "; this library contains sm_100a code only");
    auto* ctx = new (std::nothrow) smelter_context();
    if (!ctx) return fail(SMELTER_ERR_GRAPH_INTERNAL, "out of memory");
    ctx->c.device = device;
    ctx->c.num_sms = prop.multiProcessorCount;
    // SMELTER_SM_LIMIT=n: plan and launch every kernel of this context's graphs for n SMs only (experiments with several encodes in
    // flight, each confined to a share of the chip so that their kernels co-run instead of time-slicing whole-chip grids)
    if (const char* lim = getenv("SMELTER_SM_LIMIT")) {
        const int n = atoi(lim);
        if (n >= 2 && n < ctx->c.num_sms) ctx->c.num_sms = n & ~1;
    }
    if (cuda_stream) {
        ctx->c.stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        e = cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(SMELTER_ERR_CUDA, cudaGetErrorString(e)); }
        ctx->c.own_stream = true;
    }
    *out = ctx;
    return SMELTER_OK;
}
int32_t smelter_context_destroy(smelter_context* ctx) {
    if (!ctx) return SMELTER_OK;
    nccl_destroy(&ctx->c);
    if (ctx->c.flush_buf) cudaFree(ctx->c.flush_buf);
    for (auto& kv : ctx->c.staging) if (kv.second.ptr) cudaFree(kv.second.ptr);
    if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
    return SMELTER_OK;
}
int32_t smelter_context_stream(smelter_context* ctx, void** cuda_stream) {
    ARG(ctx && cuda_stream);
    *cuda_stream = ctx->c.stream;
    return SMELTER_OK;
}
int32_t smelter_context_synchronize(smelter_context* ctx) {
    ARG(ctx);
    SM_CUDA(cudaSetDevice(ctx->c.device));
    SM_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return SMELTER_OK;
}
int32_t smelter_nccl_unique_id(uint8_t id[128]) { ARG(id); return nccl_unique_id(id); }
int32_t smelter_context_init_nccl(smelter_context* ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    ARG(ctx && id && world >= 1 && rank >= 0 && rank < world);
    return nccl_init(&ctx->c, id, rank, world);
}

// ---- tensors ------------------------------------------------------------------------------------------------
int32_t smelter_tensor_create(smelter_context* ctx, int32_t n, int32_t c, int32_t h, int32_t w, smelter_tensor** out) {
    ARG(ctx && out && n > 0 && c > 0 && h > 0 && w > 0);
    SM_CUDA(cudaSetDevice(ctx->c.device));
    auto* t = new (std::nothrow) smelter_tensor();
    if (!t) return fail(SMELTER_ERR_GRAPH_INTERNAL, "out of memory");
    t->t.ctx = &ctx->c; t->t.n = n; t->t.c = c; t->t.h = h; t->t.w = w; t->t.owned = true;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&t->t.ptr), t->t.count() * 2);
    if (e != cudaSuccess) { delete t; return fail(SMELTER_ERR_CUDA, cudaGetErrorString(e)); }
    *out = t;
    return SMELTER_OK;
}
int32_t smelter_tensor_wrap(smelter_context* ctx, void* ptr, int32_t n, int32_t c, int32_t h, int32_t w, smelter_tensor** out) {
    ARG(ctx && out && ptr && n > 0 && c > 0 && h > 0 && w > 0);
    auto* t = new (std::nothrow) smelter_tensor();
    if (!t) return fail(SMELTER_ERR_GRAPH_INTERNAL, "out of memory");
    t->t.ctx = &ctx->c; t->t.ptr = static_cast<__half*>(ptr); t->t.n = n; t->t.c = c; t->t.h = h; t->t.w = w; t->t.owned = false;
    *out = t;
    return SMELTER_OK;
}
int32_t smelter_tensor_destroy(smelter_tensor* t) {
    if (!t) return SMELTER_OK;
    if (t->t.owned && t->t.ptr) cudaFree(t->t.ptr);
    delete t;
    return SMELTER_OK;
}
int32_t smelter_tensor_dims(const smelter_tensor* t, int32_t dims[4]) {
    ARG(t && dims);
    dims[0] = t->t.n; dims[1] = t->t.c; dims[2] = t->t.h; dims[3] = t->t.w;
    return SMELTER_OK;
}
int32_t smelter_tensor_device_ptr(const smelter_tensor* t, void** ptr) {
    ARG(t && ptr);
    *ptr = t->t.ptr;
    return SMELTER_OK;
}

namespace {
// staging buffer for fp32<->fp16 / u8 conversion at the boundary: per (context, stream), see Context::staging
int ensure_staging(Context* ctx, cudaStream_t s, size_t bytes, void** out) {
    std::lock_guard<std::mutex> lock(ctx->staging_mu);
    Context::Staging& st = ctx->staging[s];
    if (st.bytes < bytes) {
        if (st.ptr) cudaFree(st.ptr);  // synchronises: nothing in flight still reads the old buffer
        st = Context::Staging();
        SM_CUDA(cudaMalloc(&st.ptr, bytes));
        st.bytes = bytes;
    }
    *out = st.ptr;
    return SMELTER_OK;
}
}  // namespace

int32_t smelter_tensor_from_float(smelter_tensor* t, void* cuda_stream, const float* host, size_t count) {
    ARG(t && host && count == t->t.count());
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->t.ctx->stream;
    SM_CUDA(cudaSetDevice(t->t.ctx->device));
    void* stage = nullptr;
    int rc = ensure_staging(t->t.ctx, s, count * 4, &stage);
    if (rc) return rc;
    SM_CUDA(cudaMemcpyAsync(stage, host, count * 4, cudaMemcpyHostToDevice, s));
    SM_CUDA(k::f32_to_f16(static_cast<const float*>(stage), t->t.ptr, count, s));
    return SMELTER_OK;
}
int32_t smelter_tensor_from_half(smelter_tensor* t, void* cuda_stream, const uint16_t* host, size_t count) {
    ARG(t && host && count == t->t.count());
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->t.ctx->stream;
    SM_CUDA(cudaSetDevice(t->t.ctx->device));
    SM_CUDA(cudaMemcpyAsync(t->t.ptr, host, count * 2, cudaMemcpyHostToDevice, s));
    return SMELTER_OK;
}
int32_t smelter_tensor_from_u8(smelter_tensor* t, void* cuda_stream, const uint8_t* host, int32_t src_channels, const float* scale, const float* bias) {
    ARG(t && host && t->t.c >= 1 && t->t.c <= 4 && src_channels >= t->t.c && src_channels <= 4);
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->t.ctx->stream;
    SM_CUDA(cudaSetDevice(t->t.ctx->device));
    const size_t hw = size_t(t->t.h) * t->t.w, bytes = size_t(t->t.n) * hw * src_channels;
    void* stage = nullptr;
    int rc = ensure_staging(t->t.ctx, s, bytes, &stage);
    if (rc) return rc;
    float sc4[4], b4[4];
    for (int i = 0; i < 4; ++i) { sc4[i] = scale && i < t->t.c ? scale[i] : 1.f / 255.f; b4[i] = bias && i < t->t.c ? bias[i] : 0.f; }
    SM_CUDA(cudaMemcpyAsync(stage, host, bytes, cudaMemcpyHostToDevice, s));
    SM_CUDA(k::u8_to_nchw_f16(static_cast<const uint8_t*>(stage), t->t.ptr, t->t.n, t->t.c, hw, src_channels, sc4, b4, s));
    return SMELTER_OK;
}
int32_t smelter_tensor_to_float(const smelter_tensor* t, void* cuda_stream, float* host, size_t capacity) {
    ARG(t && host && capacity >= t->t.count());
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->t.ctx->stream;
    SM_CUDA(cudaSetDevice(t->t.ctx->device));
    const size_t count = t->t.count();
    void* stage = nullptr;
    int rc = ensure_staging(t->t.ctx, s, count * 4, &stage);
    if (rc) return rc;
    SM_CUDA(k::f16_to_f32(t->t.ptr, static_cast<float*>(stage), count, s));
    SM_CUDA(cudaMemcpyAsync(host, stage, count * 4, cudaMemcpyDeviceToHost, s));
    SM_CUDA(cudaStreamSynchronize(s));
    return SMELTER_OK;
}
int32_t smelter_tensor_to_float_mps(const smelter_tensor* t, void* cuda_stream, float* host, size_t capacity) {
    ARG(t && host);
    const int N = t->t.n, Cc = t->t.c, H = t->t.h, W = t->t.w;
    const int comps = Cc < 3 ? Cc : 4, slices = (Cc + 3) / 4, cpp = Cc < 3 ? Cc : slices * 4;  // MPSImage+Extensions.swift:27-36
    ARG(capacity >= size_t(N) * H * W * cpp);
    std::vector<float> nchw(t->t.count());
    int rc = smelter_tensor_to_float(t, cuda_stream, nchw.data(), nchw.size());
    if (rc) return rc;
    const size_t hw = size_t(H) * W;
    for (int n = 0; n < N; ++n)
        for (int s = 0; s < slices; ++s) {
            float* dst = host + (size_t(n) * slices + s) * hw * comps;  // :49-57 slice i = n * numSlices + s
            for (size_t p = 0; p < hw; ++p)
                for (int j = 0; j < comps; ++j) {
                    const int c = s * 4 + j;
                    dst[p * comps + j] = c < Cc ? nchw[(size_t(n) * Cc + c) * hw + p] : 0.f;
                }
        }
    return SMELTER_OK;
}
int32_t smelter_tensor_to_float_async(const smelter_tensor* t, void* cuda_stream, float* host, size_t capacity) {
    ARG(t && host && capacity >= t->t.count());
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->t.ctx->stream;
    SM_CUDA(cudaSetDevice(t->t.ctx->device));
    const size_t count = t->t.count();
    void* stage = nullptr;
    int rc = ensure_staging(t->t.ctx, s, count * 4, &stage);  // per stream: uses on the same stream are ordered by the stream
    if (rc) return rc;
    SM_CUDA(k::f16_to_f32(t->t.ptr, static_cast<float*>(stage), count, s));
    SM_CUDA(cudaMemcpyAsync(host, stage, count * 4, cudaMemcpyDeviceToHost, s));
    return SMELTER_OK;
}
int32_t smelter_tensor_to_half(const smelter_tensor* t, void* cuda_stream, uint16_t* host, size_t capacity) {
    ARG(t && host && capacity >= t->t.count());
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->t.ctx->stream;
    SM_CUDA(cudaSetDevice(t->t.ctx->device));
    SM_CUDA(cudaMemcpyAsync(host, t->t.ptr, t->t.count() * 2, cudaMemcpyDeviceToHost, s));
    SM_CUDA(cudaStreamSynchronize(s));
    return SMELTER_OK;
}

// ---- graph ------------------------------------------------------------------------------------------------------
int32_t smelter_graph_create(smelter_context* ctx, const uint8_t* onnx, size_t len, const smelter_config* cfg, smelter_graph** out) {
    ARG(ctx && out);
    smelter_config c;
    if (cfg) c = *cfg; else smelter_config_default(&c);
    auto* g = new (std::nothrow) smelter_graph();
    if (!g) return fail(SMELTER_ERR_GRAPH_INTERNAL, "out of memory");
    g->g = new ONNXGraph(&ctx->c);
    int rc = g->g->init(onnx, len, c);
    if (rc) { delete g->g; delete g; return rc; }
    *out = g;
    return SMELTER_OK;
}
int32_t smelter_graph_build(smelter_graph* g) { ARG(g); return g->g->build(); }
int32_t smelter_graph_destroy(smelter_graph* g) {
    if (!g) return SMELTER_OK;
    delete g->g;
    delete g;
    return SMELTER_OK;
}
int32_t smelter_graph_format(const smelter_graph* g, int32_t* format) { ARG(g && format); *format = g->g->format(); return SMELTER_OK; }
int32_t smelter_graph_num_outputs(const smelter_graph* g, int32_t* n) {
    ARG(g && n);
    // outputShapes (ONNXGraph.swift:69-91) compactMaps graph outputs of rank 3 or 4
    int count = 0;
    for (const auto& o : g->g->model().graph.output) count += (o.dims.size() == 3 || o.dims.size() == 4);
    *n = count;
    return SMELTER_OK;
}
int32_t smelter_graph_output_shape(const smelter_graph* g, int32_t idx, smelter_shape* shape) {
    ARG(g && shape && idx >= 0);
    int count = 0;
    for (const auto& o : g->g->model().graph.output) {
        if (o.dims.size() != 3 && o.dims.size() != 4) continue;
        if (count++ != idx) continue;
        const size_t b = o.dims.size() == 4 ? 1 : 0;
        // ONNXGraph.swift:73-87 (note the reference's width/height naming: dim[b+1] -> width, dim[b+2] -> height)
        shape->channels = int32_t(o.dims[b]);
        shape->width = int32_t(o.dims[b + 1]);
        shape->height = int32_t(o.dims[b + 2]);
        shape->depth = 1;
        return SMELTER_OK;
    }
    return fail(SMELTER_ERR_INVALID_ARGUMENT, "output index out of range");
}
int32_t smelter_graph_num_nodes(const smelter_graph* g, int32_t* n) { ARG(g && n); *n = g->g->num_nodes(); return SMELTER_OK; }
int32_t smelter_graph_node_op_type(const smelter_graph* g, int32_t idx, const char** op_type) {
    ARG(g && op_type && idx >= 0 && idx < g->g->num_nodes());
    *op_type = g->g->node(idx).op_type.c_str();
    return SMELTER_OK;
}
int32_t smelter_graph_has_converter(const smelter_graph* g, const char* op_type, int32_t* yes) {
    ARG(g && op_type && yes);
    *yes = g->g->has_converter(op_type) ? 1 : 0;
    return SMELTER_OK;
}
int32_t smelter_graph_num_launches(smelter_graph* g, int32_t batch, int32_t* n) { ARG(g && n && batch > 0); int v = 0; int rc = g->g->num_launches(batch, &v); *n = v; return rc; }
int32_t smelter_graph_plan_dump(smelter_graph* g, int32_t batch, char* buf, size_t cap) {
    ARG(g && buf && cap > 0 && batch > 0);
    std::string s;
    int rc = g->g->plan_dump(batch, &s);
    if (rc) return rc;
    const size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
    return SMELTER_OK;
}
int32_t smelter_graph_encode(smelter_graph* g, void* cuda_stream, const smelter_tensor* const* sources, int32_t n_sources,
                             const smelter_tensor** result) {
    ARG(g && sources && result && n_sources > 0 && n_sources <= 16);
    const Tensor* src[16];
    for (int i = 0; i < n_sources; ++i) { ARG(sources[i]); src[i] = &sources[i]->t; }
    const Tensor* res = nullptr;
    int rc = g->g->encode(static_cast<cudaStream_t>(cuda_stream), src, n_sources, &res);
    if (rc) return rc;
    // smelter_tensor is a standard-layout wrapper whose only member is the Tensor
    *result = reinterpret_cast<const smelter_tensor*>(res);
    return SMELTER_OK;
}
int32_t smelter_graph_profile(smelter_graph* g, void* cuda_stream, const smelter_tensor* const* sources, int32_t n_sources, int32_t iters,
                              float* ms, double* flops, double* bytes, int32_t* is_tensor, int32_t cap, int32_t* n_steps) {
    ARG(g && sources && n_steps && n_sources > 0 && n_sources <= 16 && cap >= 0);
    const Tensor* src[16];
    for (int i = 0; i < n_sources; ++i) { ARG(sources[i]); src[i] = &sources[i]->t; }
    std::vector<float> v_ms;
    std::vector<double> v_fl, v_by;
    std::vector<int> v_tc;
    int rc = g->g->profile(static_cast<cudaStream_t>(cuda_stream), src, n_sources, iters, &v_ms, &v_fl, &v_by, &v_tc);
    if (rc) return rc;
    *n_steps = int32_t(v_ms.size());
    for (int i = 0; i < cap && size_t(i) < v_ms.size(); ++i) {
        if (ms) ms[i] = v_ms[size_t(i)];
        if (flops) flops[i] = v_fl[size_t(i)];
        if (bytes) bytes[i] = v_by[size_t(i)];
        if (is_tensor) is_tensor[i] = v_tc[size_t(i)];
    }
    return SMELTER_OK;
}
int32_t smelter_graph_broadcast_weights(smelter_graph* g, int32_t root) { ARG(g); return g->g->broadcast_weights(root); }
int32_t smelter_graph_weight_checksum(smelter_graph* g, uint64_t* checksum, uint64_t* bytes) {
    ARG(g && checksum && bytes);
    return g->g->weight_checksum(checksum, bytes);
}

int32_t smelter_graph_weight_arena(smelter_graph* g, void** device_ptr, uint64_t* bytes) {
    ARG(g && device_ptr && bytes);
    if (!g->g->built()) return fail(SMELTER_ERR_INCONSISTENT_STATE, "graph not built");
    *device_ptr = g->g->weight_arena();
    *bytes = g->g->weight_bytes();
    return SMELTER_OK;
}

// ---- symbol tables + builder ---------------------------------------------------------------------------------------
int32_t smelter_graph_has_output(const smelter_graph* g, const char* name, int32_t* yes) { ARG(g && name && yes); *yes = g->g->output(name) >= 0; return SMELTER_OK; }
int32_t smelter_graph_shape(const smelter_graph* g, const char* name, smelter_shape* shape) {
    ARG(g && name && shape);
    const ImageShape* s = g->g->shape(name);
    if (!s) return fail(SMELTER_ERR_NO_SUCH_OUTPUT, std::string("no image node named '") + name + "'");
    // the converters' convention (ONNXGraph.swift:243-248): channels = 1, depth = feature channels
    shape->channels = 1; shape->width = s->w; shape->height = s->h; shape->depth = s->c;
    return SMELTER_OK;
}
int32_t smelter_graph_has_tensor(const smelter_graph* g, const char* name, int32_t* yes) { ARG(g && name && yes); *yes = g->g->tensor(name) != nullptr; return SMELTER_OK; }

namespace {
int find_input(ONNXGraph* g, const char* name, int* value, ImageShape* shape) {
    const int v = g->output(name);
    const ImageShape* s = g->shape(name);
    if (v < 0 || !s) return fail(SMELTER_ERR_NO_SUCH_OUTPUT, std::string("no image node named '") + name + "'");
    *value = v;
    *shape = *s;
    return SMELTER_OK;
}
int not_built(ONNXGraph* g) { return g->built() ? fail(SMELTER_ERR_INCONSISTENT_STATE, "graph already built") : SMELTER_OK; }
}  // namespace

int32_t smelter_add_conv(smelter_graph* g, const char* in_name, const smelter_conv_desc* d, const void* weights, const void* bias,
                         const char* out_name) {
    ARG(g && in_name && d && weights && out_name);
    ARG(d->c_out > 0 && d->c_in_per_group > 0 && d->k_h > 0 && d->k_w > 0 && d->groups > 0 && d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0);
    ARG(d->weight_dtype == SMELTER_F32 || d->weight_dtype == SMELTER_F16);
    int rc = not_built(g->g);
    if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s);
    if (rc) return rc;
    if (s.c != d->c_in_per_group * d->groups && !(d->is_gemm && s.c * s.h * s.w == d->c_in_per_group))
        return fail(SMELTER_ERR_INCONSISTENT_STATE, "input channels disagree with the weight dims");
    Filter f;
    f.kind = FilterKind::Conv;
    f.op_type = d->is_gemm ? "Gemm" : "Conv";
    f.is_gemm = d->is_gemm != 0;
    f.c_out = d->c_out; f.c_in_g = d->c_in_per_group; f.k_h = d->k_h; f.k_w = d->k_w;
    f.stride_h = d->stride_h; f.stride_w = d->stride_w; f.dil_h = d->dil_h; f.dil_w = d->dil_w; f.groups = d->groups;
    for (int i = 0; i < 4; ++i) f.pads[i] = d->pads[i];
    const size_t count = size_t(d->c_out) * d->c_in_per_group * d->k_h * d->k_w;
    std::vector<float> wf(count);
    if (d->weight_dtype == SMELTER_F32) memcpy(wf.data(), weights, count * 4);
    else for (size_t i = 0; i < count; ++i) wf[i] = onnx::half_to_float(static_cast<const uint16_t*>(weights)[i]);
    if (d->weight_layout == SMELTER_OHWI || (d->k_h == 1 && d->k_w == 1)) f.w = std::move(wf);
    else { f.w.resize(count); reformat_conv_weight(wf.data(), f.w.data(), 4, d->c_out, d->c_in_per_group, d->k_h, d->k_w, false); }
    f.bias.assign(size_t(d->c_out), 0.f);
    if (bias) {
        if (d->bias_dtype == SMELTER_F16) for (int i = 0; i < d->c_out; ++i) f.bias[size_t(i)] = onnx::half_to_float(static_cast<const uint16_t*>(bias)[i]);
        else memcpy(f.bias.data(), bias, size_t(d->c_out) * 4);
    }
    ImageShape out{d->c_out, 1, 1};
    if (!d->is_gemm) {
        out.h = conv_output_size(s.h, d->k_h, d->stride_h, d->dil_h, d->pads[0], d->pads[2], 0, false);
        out.w = conv_output_size(s.w, d->k_w, d->stride_w, d->dil_w, d->pads[1], d->pads[3], 0, false);
        if (out.h <= 0 || out.w <= 0) return fail(SMELTER_ERR_INCONSISTENT_STATE, "empty output");
    }
    f.in = {in};
    return g->g->addFilter(std::move(f), out, {out_name});
}

int32_t smelter_add_batchnorm(smelter_graph* g, const char* in_name, int32_t channels, const float* gamma, const float* beta, const float* mean,
                              const float* var, float epsilon, const char* out_name) {
    ARG(g && in_name && gamma && beta && mean && var && out_name);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    if (channels != s.c) return fail(SMELTER_ERR_INSUFFICIENT_INPUTS, "parameter length != channels");
    Filter f;
    f.kind = FilterKind::BatchNorm; f.op_type = "BatchNormalization"; f.eps = epsilon;
    f.p0.resize(size_t(channels)); f.p1.resize(size_t(channels));
    for (int c = 0; c < channels; ++c) {
        const float sc = gamma[c] / std::sqrt(var[c] + epsilon);
        f.p0[size_t(c)] = sc; f.p1[size_t(c)] = beta[c] - mean[c] * sc;
    }
    f.in = {in};
    return g->g->addFilter(std::move(f), s, {out_name});
}
int32_t smelter_add_instancenorm(smelter_graph* g, const char* in_name, int32_t channels, const float* gamma, const float* beta, float epsilon,
                                 const char* out_name) {
    ARG(g && in_name && gamma && beta && out_name);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    if (channels != s.c) return fail(SMELTER_ERR_INCONSISTENT_STATE, "gamma/beta length != channels");
    Filter f;
    f.kind = FilterKind::InstanceNorm; f.op_type = "InstanceNormalization"; f.eps = epsilon;
    f.p0.assign(gamma, gamma + channels); f.p1.assign(beta, beta + channels);
    f.in = {in};
    return g->g->addFilter(std::move(f), s, {out_name});
}
int32_t smelter_add_unary(smelter_graph* g, const char* in_name, int32_t kind, float alpha, float beta, const char* out_name) {
    ARG(g && in_name && out_name && kind >= 0 && kind <= SMELTER_UNARY_IDENTITY);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    Filter f;
    f.kind = FilterKind::Unary; f.op_type = "Unary"; f.sub = kind; f.alpha = alpha; f.beta = beta; f.in = {in};
    return g->g->addFilter(std::move(f), s, {out_name});
}
int32_t smelter_add_binary(smelter_graph* g, const char* a_name, const char* b_name, int32_t kind, const char* out_name) {
    ARG(g && a_name && b_name && out_name && kind >= 0 && kind <= SMELTER_BIN_DIV);
    int rc = not_built(g->g); if (rc) return rc;
    int a, b; ImageShape sa, sb;
    rc = find_input(g->g, a_name, &a, &sa); if (rc) return rc;
    rc = find_input(g->g, b_name, &b, &sb); if (rc) return rc;
    if (!(sa == sb)) return fail(SMELTER_ERR_UNSUPPORTED, "operand shapes differ (no broadcasting)");
    Filter f;
    f.kind = FilterKind::Binary; f.op_type = "Binary"; f.sub = kind; f.in = {a, b};
    return g->g->addFilter(std::move(f), sa, {out_name});
}
int32_t smelter_add_pool(smelter_graph* g, const char* in_name, int32_t is_max, int32_t k_h, int32_t k_w, int32_t stride_h, int32_t stride_w,
                         int32_t pad_h, int32_t pad_w, const char* out_name) {
    ARG(g && in_name && out_name && k_h > 0 && k_w > 0 && stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    Filter f;
    f.kind = FilterKind::Pool; f.op_type = is_max ? "MaxPool" : "AveragePool"; f.sub = is_max ? 1 : 0;
    f.k_h = k_h; f.k_w = k_w; f.stride_h = stride_h; f.stride_w = stride_w; f.pool_pad_h = pad_h; f.pool_pad_w = pad_w;
    ImageShape out{s.c, pool_output_size(s.h, k_h, stride_h, pad_h), pool_output_size(s.w, k_w, stride_w, pad_w)};
    if (out.h <= 0 || out.w <= 0) return fail(SMELTER_ERR_INCONSISTENT_STATE, "empty output");
    f.in = {in};
    return g->g->addFilter(std::move(f), out, {out_name});
}
int32_t smelter_add_global_avgpool(smelter_graph* g, const char* in_name, const char* out_name) {
    ARG(g && in_name && out_name);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    Filter f;
    f.kind = FilterKind::GlobalAvgPool; f.op_type = "GlobalAveragePool"; f.in = {in};
    return g->g->addFilter(std::move(f), ImageShape{s.c, 1, 1}, {out_name});
}
int32_t smelter_add_upsample(smelter_graph* g, const char* in_name, int32_t mode, int32_t scale_h, int32_t scale_w, int32_t align_corners,
                             const char* out_name) {
    ARG(g && in_name && out_name && scale_h >= 1 && scale_w >= 1 && (mode == SMELTER_UPSAMPLE_NEAREST || mode == SMELTER_UPSAMPLE_BILINEAR));
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    Filter f;
    f.kind = FilterKind::Upsample; f.op_type = "Upsample"; f.sub = mode; f.scale_h = scale_h; f.scale_w = scale_w; f.align_corners = align_corners;
    f.in = {in};
    return g->g->addFilter(std::move(f), ImageShape{s.c, s.h * scale_h, s.w * scale_w}, {out_name});
}
int32_t smelter_add_concat(smelter_graph* g, const char* const* in_names, int32_t n_inputs, const char* out_name) {
    ARG(g && in_names && n_inputs > 0 && out_name);
    int rc = not_built(g->g); if (rc) return rc;
    Filter f;
    f.kind = FilterKind::Concat; f.op_type = "Concat";
    ImageShape out{0, 0, 0};
    for (int i = 0; i < n_inputs; ++i) {
        ARG(in_names[i]);
        int in; ImageShape s;
        rc = find_input(g->g, in_names[i], &in, &s); if (rc) return rc;
        if (i == 0) { out.h = s.h; out.w = s.w; }
        else if (s.h != out.h || s.w != out.w) return fail(SMELTER_ERR_UNSUPPORTED, "spatial sizes differ");
        out.c += s.c;
        f.in.push_back(in);
    }
    return g->g->addFilter(std::move(f), out, {out_name});
}
int32_t smelter_add_reshape(smelter_graph* g, const char* in_name, int32_t c, int32_t h, int32_t w, const char* out_name) {
    ARG(g && in_name && out_name && c > 0 && h > 0 && w > 0);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    if (long(c) * h * w != long(s.c) * s.h * s.w) return fail(SMELTER_ERR_INCONSISTENT_STATE, "element count changes");
    if (s.h == 1 && s.w == 1 && h == 1 && w == 1) return g->g->addAlias(in, ImageShape{c, 1, 1}, {out_name});
    Filter f;
    f.kind = FilterKind::Reshape; f.op_type = "Reshape"; f.in = {in};
    return g->g->addFilter(std::move(f), ImageShape{c, h, w}, {out_name});
}
int32_t smelter_add_softmax(smelter_graph* g, const char* in_name, int32_t log_softmax, const char* out_name) {
    ARG(g && in_name && out_name);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    Filter f;
    f.kind = FilterKind::Softmax; f.op_type = log_softmax ? "LogSoftmax" : "Softmax"; f.sub = log_softmax ? 1 : 0; f.in = {in};
    return g->g->addFilter(std::move(f), s, {out_name});
}
int32_t smelter_add_pad(smelter_graph* g, const char* in_name, int32_t mode, const int32_t p[8], float value, const char* out_name) {
    ARG(g && in_name && p && out_name && mode >= SMELTER_PAD_CONSTANT && mode <= SMELTER_PAD_EDGE);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    if (p[0] || p[4] || p[1] || p[5]) return fail(SMELTER_ERR_UNSUPPORTED, "batch/channel padding");
    Filter f;
    f.kind = FilterKind::Pad; f.op_type = "Pad"; f.sub = mode; f.alpha = value;
    f.pads[0] = p[2]; f.pads[1] = p[3]; f.pads[2] = p[6]; f.pads[3] = p[7];
    for (int i = 0; i < 4; ++i) if (f.pads[i] < 0) return fail(SMELTER_ERR_UNSUPPORTED, "negative pads");
    if (mode == SMELTER_PAD_REFLECT && (f.pads[0] >= s.h || f.pads[2] >= s.h || f.pads[1] >= s.w || f.pads[3] >= s.w))
        return fail(SMELTER_ERR_UNSUPPORTED, "reflect pad must be smaller than the image");
    f.in = {in};
    return g->g->addFilter(std::move(f), ImageShape{s.c, s.h + f.pads[0] + f.pads[2], s.w + f.pads[1] + f.pads[3]}, {out_name});
}
int32_t smelter_add_alias(smelter_graph* g, const char* in_name, const char* out_name) {
    ARG(g && in_name && out_name);
    int rc = not_built(g->g); if (rc) return rc;
    int in; ImageShape s;
    rc = find_input(g->g, in_name, &in, &s); if (rc) return rc;
    return g->g->addAlias(in, s, {out_name});
}

int32_t smelter_graph_register_converter(smelter_graph* g, const char* op_type, smelter_converter_fn fn, void* user) {
    ARG(g && op_type && fn);
    smelter_graph* handle = g;
    g->g->registerConverter(op_type, [handle, fn, user](ONNXGraph&, int node) { return int(fn(handle, node, user)); });
    return SMELTER_OK;
}
#define NODE_ARG() ARG(g && node >= 0 && node < g->g->num_nodes())
int32_t smelter_node_num_inputs(const smelter_graph* g, int32_t node, int32_t* n) { NODE_ARG(); ARG(n); *n = int32_t(g->g->node(node).input.size()); return SMELTER_OK; }
int32_t smelter_node_input(const smelter_graph* g, int32_t node, int32_t i, const char** name) {
    NODE_ARG(); ARG(name && i >= 0 && size_t(i) < g->g->node(node).input.size());
    *name = g->g->node(node).input[size_t(i)].c_str();
    return SMELTER_OK;
}
int32_t smelter_node_num_outputs(const smelter_graph* g, int32_t node, int32_t* n) { NODE_ARG(); ARG(n); *n = int32_t(g->g->node(node).output.size()); return SMELTER_OK; }
int32_t smelter_node_output(const smelter_graph* g, int32_t node, int32_t i, const char** name) {
    NODE_ARG(); ARG(name && i >= 0 && size_t(i) < g->g->node(node).output.size());
    *name = g->g->node(node).output[size_t(i)].c_str();
    return SMELTER_OK;
}
int32_t smelter_node_attr_int(const smelter_graph* g, int32_t node, const char* attr, int64_t* v, int32_t* found) {
    NODE_ARG(); ARG(attr && v && found);
    const auto* a = g->g->node(node).attr(attr);
    *found = a != nullptr;
    if (a) *v = a->i;
    return SMELTER_OK;
}
int32_t smelter_node_attr_float(const smelter_graph* g, int32_t node, const char* attr, float* v, int32_t* found) {
    NODE_ARG(); ARG(attr && v && found);
    const auto* a = g->g->node(node).attr(attr);
    *found = a != nullptr;
    if (a) *v = a->f;
    return SMELTER_OK;
}
int32_t smelter_node_attr_ints(const smelter_graph* g, int32_t node, const char* attr, int64_t* v, int32_t cap, int32_t* n) {
    NODE_ARG(); ARG(attr && n && cap >= 0 && (v || cap == 0));
    const auto* a = g->g->node(node).attr(attr);
    if (!a) { *n = -1; return SMELTER_OK; }
    *n = int32_t(a->ints.size());
    for (int i = 0; i < cap && size_t(i) < a->ints.size(); ++i) v[i] = a->ints[size_t(i)];
    return SMELTER_OK;
}

// ---- host utilities ---------------------------------------------------------------------------------------------------
int32_t smelter_reformat_conv_weight(const void* src, void* dst, int32_t elem_size, int32_t c_out, int32_t c_in, int32_t k_h, int32_t k_w,
                                     int32_t is_transpose) {
    ARG(src && dst && src != dst && (elem_size == 2 || elem_size == 4) && c_out > 0 && c_in > 0 && k_h > 0 && k_w > 0);
    reformat_conv_weight(src, dst, elem_size, c_out, c_in, k_h, k_w, is_transpose != 0);
    return SMELTER_OK;
}
int32_t smelter_float16_to_32(const uint16_t* src, float* dst, size_t n) {
    ARG((src && dst) || n == 0);
    for (size_t i = 0; i < n; ++i) dst[i] = onnx::half_to_float(src[i]);
    return SMELTER_OK;
}
int32_t smelter_float32_to_16(const float* src, uint16_t* dst, size_t n) {
    ARG((src && dst) || n == 0);
    for (size_t i = 0; i < n; ++i) dst[i] = onnx::float_to_half(src[i]);
    return SMELTER_OK;
}
int32_t smelter_conv_output_size(int32_t in, int32_t k, int32_t stride, int32_t dil, int32_t pad_lo, int32_t pad_hi, int32_t out_pad,
                                 int32_t is_transpose, int32_t* out) {
    ARG(out && stride > 0 && dil > 0 && k > 0);
    *out = conv_output_size(in, k, stride, dil, pad_lo, pad_hi, out_pad, is_transpose != 0);
    return SMELTER_OK;
}
int32_t smelter_pool_output_size(int32_t in, int32_t k, int32_t stride, int32_t pad, int32_t* out) {
    ARG(out && stride > 0 && k > 0);
    *out = pool_output_size(in, k, stride, pad);
    return SMELTER_OK;
}
namespace {
int parse_single_tensor(const uint8_t* data, size_t len, onnx::ModelProto* holder, const onnx::TensorProto** t) {
    // Wrap the TensorProto as GraphProto.initializer (field 5) inside ModelProto.graph (field 7) so the one
    // public parser entry point can be reused.
    std::string err;
    std::vector<uint8_t> buf;
    auto put_varint = [&](uint64_t v) { while (v >= 0x80) { buf.push_back(uint8_t(v) | 0x80); v >>= 7; } buf.push_back(uint8_t(v)); };
    std::vector<uint8_t> graph;
    {
        std::vector<uint8_t> saved;
        saved.swap(buf);
        buf.push_back((5 << 3) | 2);
        put_varint(len);
        buf.insert(buf.end(), data, data + len);
        graph.swap(buf);
        buf.swap(saved);
    }
    buf.push_back((7 << 3) | 2);
    put_varint(graph.size());
    buf.insert(buf.end(), graph.begin(), graph.end());
    static thread_local std::vector<uint8_t> keep;  // raw_data views point here until the next call
    keep.swap(buf);
    if (!onnx::parse_model(keep.data(), keep.size(), holder, &err) || holder->graph.initializer.size() != 1) return fail(SMELTER_ERR_PARSE, "malformed TensorProto");
    *t = &holder->graph.initializer[0];
    return SMELTER_OK;
}
}  // namespace
int32_t smelter_tensorproto_integers(const uint8_t* tensor_proto, size_t len, int64_t* out, size_t cap, size_t* n) {
    ARG(tensor_proto && n && (out || cap == 0));
    onnx::ModelProto m;
    const onnx::TensorProto* t = nullptr;
    int rc = parse_single_tensor(tensor_proto, len, &m, &t);
    if (rc) return rc;
    std::vector<int64_t> v;
    if (!t->integers(&v)) return fail(SMELTER_ERR_UNSUPPORTED, "Unsupported conversion rule");  // the reference fatalErrors here
    *n = v.size();
    for (size_t i = 0; i < cap && i < v.size(); ++i) out[i] = v[i];
    return SMELTER_OK;
}
int32_t smelter_tensorproto_floats(const uint8_t* tensor_proto, size_t len, float* out, size_t cap, size_t* n) {
    ARG(tensor_proto && n && (out || cap == 0));
    onnx::ModelProto m;
    const onnx::TensorProto* t = nullptr;
    int rc = parse_single_tensor(tensor_proto, len, &m, &t);
    if (rc) return rc;
    std::vector<float> v;
    if (!t->floats(&v)) return fail(SMELTER_ERR_UNSUPPORTED, "Unsupported conversion rule");
    *n = v.size();
    for (size_t i = 0; i < cap && i < v.size(); ++i) out[i] = v[i];
    return SMELTER_OK;
}

// ---- single-kernel driver for tests / microbenchmarks -------------------------------------------------------------------
int32_t smelter_run_conv(smelter_context* ctx, const smelter_conv_problem* p, const void* x, const void* w, const float* bias, const void* residual,
                         void* y, int32_t iters, float* kernel_ms) {
    ARG(ctx && p && x && w && y);
    ARG(p->n > 0 && p->h > 0 && p->w > 0 && p->c_in > 0 && p->c_out > 0 && p->k_h > 0 && p->k_w > 0 && p->groups > 0);
    ARG(!p->has_bias || bias);
    ARG(!p->has_residual || residual);
    SM_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.stream;
    const int c_in_g = p->c_in / p->groups;
    const int pads[4] = {p->pad_t, p->pad_l, p->pad_b, p->pad_r};
    int mode = p->force_path ? p->force_path : pick_conv_mode(p->c_in, p->c_out, p->groups, p->k_h, p->k_w, p->stride_h, p->stride_w, p->dil_w, pads);
    if (mode < 0) return fail(SMELTER_ERR_UNSUPPORTED, "grouped convolution other than depthwise");
    const int P = conv_output_size(p->h, p->k_h, p->stride_h, p->dil_h, p->pad_t, p->pad_b, 0, false);
    const int Q = conv_output_size(p->w, p->k_w, p->stride_w, p->dil_w, p->pad_l, p->pad_r, 0, false);
    if (P <= 0 || Q <= 0) return fail(SMELTER_ERR_INCONSISTENT_STATE, "empty output");
    const int icp = round_up(p->c_in, 8), ocp = round_up(p->c_out, 8);

    // weights: OIHW fp16 host -> OHWI fp32 -> packed
    const size_t wcount = size_t(p->c_out) * c_in_g * p->k_h * p->k_w;
    std::vector<float> w_oihw(wcount), w_ohwi(wcount);
    for (size_t i = 0; i < wcount; ++i) w_oihw[i] = onnx::half_to_float(static_cast<const uint16_t*>(w)[i]);
    reformat_conv_weight(w_oihw.data(), w_ohwi.data(), 4, p->c_out, c_in_g, p->k_h, p->k_w, false);
    std::vector<uint16_t> packed;
    if (mode == 4) { packed.resize(size_t(p->k_h) * p->k_w * ocp); pack_weights_depthwise(w_ohwi.data(), p->c_out, p->k_h, p->k_w, ocp, packed.data()); }
    else { packed.resize(size_t(p->c_out) * p->k_h * p->k_w * icp); pack_weights_ohwi(w_ohwi.data(), p->c_out, p->c_in, p->k_h, p->k_w, icp, packed.data()); }
    std::vector<float> bias_pad(size_t(round_up(p->c_out, 256)), 0.f);
    if (p->has_bias) memcpy(bias_pad.data(), bias, size_t(p->c_out) * 4);

    const bool materialise = mode == k::CONV_MODE_PACKED_ROW;
    const int hp = materialise ? p->h + p->pad_t + p->pad_b : p->h;
    const int wp = materialise ? p->w + p->pad_l + p->pad_r : p->w;
    struct Bufs {
        void *w = nullptr, *b = nullptr, *xi = nullptr, *yo = nullptr, *res = nullptr, *ws = nullptr, *cnt = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Bufs() { cudaFree(w); cudaFree(b); cudaFree(xi); cudaFree(yo); cudaFree(res); cudaFree(ws); cudaFree(cnt); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    } B;
    SM_CUDA(cudaMalloc(&B.w, packed.size() * 2));
    SM_CUDA(cudaMalloc(&B.b, bias_pad.size() * 4));
    SM_CUDA(cudaMalloc(&B.xi, size_t(p->n) * hp * wp * icp * 2));
    SM_CUDA(cudaMalloc(&B.yo, size_t(p->n) * P * Q * ocp * 2));
    SM_CUDA(cudaMemcpyAsync(B.w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, s));
    SM_CUDA(cudaMemcpyAsync(B.b, bias_pad.data(), bias_pad.size() * 4, cudaMemcpyHostToDevice, s));
    SM_CUDA(k::nchw_to_nhwc(static_cast<const __half*>(x), static_cast<__half*>(B.xi), p->n, p->c_in, p->h, p->w, icp, materialise ? p->pad_t : 0,
                            materialise ? p->pad_l : 0, materialise ? p->pad_b : 0, materialise ? p->pad_r : 0, s));
    if (p->has_residual) {
        SM_CUDA(cudaMalloc(&B.res, size_t(p->n) * P * Q * ocp * 2));
        SM_CUDA(k::nchw_to_nhwc(static_cast<const __half*>(residual), static_cast<__half*>(B.res), p->n, p->c_out, P, Q, ocp, 0, 0, 0, 0, s));
    }
    SM_CUDA(cudaMemsetAsync(B.yo, 0xff, size_t(p->n) * P * Q * ocp * 2, s));  // NaN pattern: unwritten outputs are caught
    SM_CUDA(cudaEventCreate(&B.e0));
    SM_CUDA(cudaEventCreate(&B.e1));
    if (iters < 1) iters = 1;
    if (mode == 4) {
        if (p->has_residual) return fail(SMELTER_ERR_UNSUPPORTED, "depthwise kernel has no residual input");
        SM_CUDA(cudaEventRecord(B.e0, s));
        for (int it = 0; it < iters; ++it)
            SM_CUDA(k::depthwise_conv(static_cast<const __half*>(B.xi), static_cast<const __half*>(B.w), static_cast<const float*>(B.b), static_cast<__half*>(B.yo),
                                      p->n, p->h, p->w, icp, P, Q, p->k_h, p->k_w, p->stride_h, p->stride_w, p->dil_h, p->dil_w, p->pad_t, p->pad_l, p->act,
                                      p->clip_lo, p->clip_hi, s));
        SM_CUDA(cudaEventRecord(B.e1, s));
    } else {
        k::ConvTcProblem q{};
        q.mode = mode;
        q.n = p->n; q.h = hp; q.w = wp; q.c_in = p->c_in; q.c_in_pitch = icp; q.c_out = p->c_out; q.c_out_pitch = ocp;
        q.k_h = p->k_h; q.k_w = p->k_w; q.stride_h = p->stride_h; q.stride_w = p->stride_w; q.dil_h = p->dil_h; q.dil_w = p->dil_w;
        if (!materialise) { q.pad_t = p->pad_t; q.pad_l = p->pad_l; q.pad_b = p->pad_b; q.pad_r = p->pad_r; }
        q.x = static_cast<const __half*>(B.xi); q.w_packed = static_cast<const __half*>(B.w); q.bias = static_cast<const float*>(B.b);
        q.residual = static_cast<const __half*>(B.res); q.y = static_cast<__half*>(B.yo);
        q.act = p->act; q.clip_lo = p->clip_lo; q.clip_hi = p->clip_hi;
        const k::ConvTcPlanInfo info = k::conv_tc_plan(q, ctx->c.num_sms);
        if (info.splits > 1) {
            SM_CUDA(cudaMalloc(&B.ws, info.ws_bytes));
            SM_CUDA(cudaMalloc(&B.cnt, info.counter_bytes));
            SM_CUDA(cudaMemsetAsync(B.cnt, 0, info.counter_bytes, s));
            q.split_ws = static_cast<float*>(B.ws);
            q.split_counters = static_cast<unsigned int*>(B.cnt);
        }
        k::ConvTcLaunch L;
        std::string cerr;
        if (!k::conv_tc_prepare(&L, q, ctx->c.num_sms, &cerr)) return fail(SMELTER_ERR_GRAPH_INTERNAL, cerr);
        SM_CUDA(cudaEventRecord(B.e0, s));
        for (int it = 0; it < iters; ++it) SM_CUDA(k::conv_tc_launch(L, s));
        SM_CUDA(cudaEventRecord(B.e1, s));
        SM_CUDA(cudaStreamSynchronize(s));
        k::conv_tc_dump_timeline(L);
    }
    SM_CUDA(k::nhwc_to_nchw(static_cast<const __half*>(B.yo), static_cast<__half*>(y), p->n, p->c_out, P, Q, ocp, long(p->c_out) * P * Q, s));
    SM_CUDA(cudaStreamSynchronize(s));
    if (kernel_ms) {
        float ms = 0.f;
        SM_CUDA(cudaEventElapsedTime(&ms, B.e0, B.e1));
        *kernel_ms = ms / float(iters);
    }
    return SMELTER_OK;
}

int32_t smelter_tma_probe(smelter_context* ctx, int32_t mode, int32_t c, int32_t w, int32_t h, int32_t n, int32_t stages, int32_t iters, int32_t grid,
                          int32_t distinct, float* ms) {
    ARG(ctx && ms && c >= 8 && c % 8 == 0 && w > 0 && h > 0 && n > 0 && stages >= 1 && stages <= 12 && iters > 0 && grid > 0);
    SM_CUDA(cudaSetDevice(ctx->c.device));
    void* x = nullptr;
    const size_t bytes = size_t(n) * h * w * c * 2;
    SM_CUDA(cudaMalloc(&x, bytes));
    cudaMemsetAsync(x, 0, bytes, ctx->c.stream);
    std::string err;
    int rc;
    if (mode >= 4) {  // mode 4: distinct = K slabs per instruction; mode 5: distinct = issuing warps
        rc = k::tma_probe3(mode, c, long(n) * h * w, distinct, stages, iters, grid, static_cast<const __half*>(x), ctx->c.stream, ms, &err);
    } else if (mode >= 2) {  // mode 2: distinct = box_c * 1000 + box_r (no swizzle); mode 3: distinct = cluster size (multicast)
        rc = k::tma_probe2(mode, c, long(n) * h * w, mode == 2 ? distinct / 1000 : 64, mode == 2 ? distinct % 1000 : 128, mode == 3 ? distinct : 1, stages, iters,
                           grid, static_cast<const __half*>(x), ctx->c.stream, ms, &err);
    } else
    rc = k::tma_probe(mode, c, w, h, n, stages, iters, grid, distinct, static_cast<const __half*>(x), ctx->c.stream, ms, &err);
    cudaFree(x);
    if (rc) return fail(SMELTER_ERR_CUDA, err);
    return SMELTER_OK;
}

int32_t smelter_l2_flush(smelter_context* ctx) {
    ARG(ctx);
    SM_CUDA(cudaSetDevice(ctx->c.device));
    if (!ctx->c.flush_buf) {
        ctx->c.flush_bytes = size_t(256) << 20;  // 2x the 126 MB L2
        SM_CUDA(cudaMalloc(&ctx->c.flush_buf, ctx->c.flush_bytes));
    }
    SM_CUDA(cudaMemsetAsync(ctx->c.flush_buf, 0, ctx->c.flush_bytes, ctx->c.stream));
    return SMELTER_OK;
}

int32_t smelter_run_elementwise(smelter_context* ctx, const smelter_ew_problem* p, const void* x, const void* x2, const float* p0,
                                const float* p1, void* y, int32_t iters, float* kernel_ms) {
    ARG(ctx && p && x && y);
    ARG(p->n > 0 && p->c > 0 && p->h > 0 && p->w > 0);
    ARG(p->op >= SMELTER_EW_UNARY && p->op <= SMELTER_EW_LAYOUT_ROUNDTRIP);
    SM_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.stream;
    const int N = p->n, Cc = p->c, H = p->h, W = p->w;
    const int cp = round_up(Cc, 8);
    int oc = Cc, oh = H, ow = W;
    switch (p->op) {
        case SMELTER_EW_BINARY: ARG(x2); break;
        case SMELTER_EW_SCALE_SHIFT: case SMELTER_EW_INSTANCE_NORM: ARG(p0 && p1); break;
        case SMELTER_EW_POOL:
            ARG(p->k_h > 0 && p->k_w > 0 && p->stride_h > 0 && p->stride_w > 0);
            oh = pool_output_size(H, p->k_h, p->stride_h, p->pad_h);
            ow = pool_output_size(W, p->k_w, p->stride_w, p->pad_w);
            break;
        case SMELTER_EW_GLOBAL_AVGPOOL: oh = ow = 1; break;
        case SMELTER_EW_UPSAMPLE: ARG(p->scale_h >= 1 && p->scale_w >= 1); oh = H * p->scale_h; ow = W * p->scale_w; break;
        case SMELTER_EW_PAD: oh = H + p->pad_h + p->pad_b; ow = W + p->pad_w + p->pad_r; break;
        case SMELTER_EW_CONCAT: ARG(x2 && p->c2 > 0); oc = Cc + p->c2; break;
        default: break;
    }
    if (oh <= 0 || ow <= 0) return fail(SMELTER_ERR_INCONSISTENT_STATE, "empty output");
    const int ocp = round_up(oc, 8);
    struct Bufs {
        void *xi = nullptr, *x2i = nullptr, *yo = nullptr, *q0 = nullptr, *q1 = nullptr, *scratch = nullptr, *stats = nullptr;
        void *raw_x2 = nullptr, *raw_y = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Bufs() { cudaFree(xi); cudaFree(raw_x2); cudaFree(raw_y); cudaFree(q0); cudaFree(q1); cudaFree(scratch); cudaFree(stats); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    } B;
    // The operands of a streaming kernel advance through their buffers in lock-step.  cudaMalloc hands out 2 MiB-aligned blocks, so
    // equal offsets of input and output would keep landing on the same DRAM channels / banks (measured: the two-input add fell
    // from 98 % to 14-32 % of the copy rate depending on what the allocator had handed out before).  The engine's arena places
    // tensors at arbitrary 256-byte offsets; this harness staggers its buffers by odd multiples of 4 KiB to measure the same thing.
    // SMELTER_EW_STAGGER = k: offsets of k x 641 x 4 KiB between the buffers (default 1); the benchmark tool tries several placements.
    const char* stg = getenv("SMELTER_EW_STAGGER");
    const size_t kStagger = size_t(stg ? std::max(0, atoi(stg)) : 1) * 641 * 4096;
    const size_t in_elems = size_t(N) * H * W * cp, out_elems = size_t(N) * oh * ow * ocp;
    SM_CUDA(cudaMalloc(&B.xi, in_elems * 2));
    SM_CUDA(cudaMalloc(&B.raw_y, out_elems * 2 + kStagger));
    B.yo = static_cast<char*>(B.raw_y) + kStagger;
    SM_CUDA(cudaMemsetAsync(B.yo, 0xff, out_elems * 2, s));
    SM_CUDA(k::nchw_to_nhwc(static_cast<const __half*>(x), static_cast<__half*>(B.xi), N, Cc, H, W, cp, 0, 0, 0, 0, s));
    int c2p = 0;
    if (x2) {
        const int c2 = p->op == SMELTER_EW_CONCAT ? p->c2 : Cc;
        c2p = round_up(c2, 8);
        SM_CUDA(cudaMalloc(&B.raw_x2, size_t(N) * H * W * c2p * 2 + 2 * kStagger));
        B.x2i = static_cast<char*>(B.raw_x2) + 2 * kStagger;
        SM_CUDA(k::nchw_to_nhwc(static_cast<const __half*>(x2), static_cast<__half*>(B.x2i), N, c2, H, W, c2p, 0, 0, 0, 0, s));
    }
    if (p0 && p1) {
        std::vector<float> a(size_t(cp), 0.f), b(size_t(cp), 0.f);
        memcpy(a.data(), p0, size_t(Cc) * 4);
        memcpy(b.data(), p1, size_t(Cc) * 4);
        SM_CUDA(cudaMalloc(&B.q0, size_t(cp) * 4));
        SM_CUDA(cudaMalloc(&B.q1, size_t(cp) * 4));
        SM_CUDA(cudaMemcpy(B.q0, a.data(), size_t(cp) * 4, cudaMemcpyHostToDevice));
        SM_CUDA(cudaMemcpy(B.q1, b.data(), size_t(cp) * 4, cudaMemcpyHostToDevice));
    }
    if (p->op == SMELTER_EW_INSTANCE_NORM)
        SM_CUDA(cudaMalloc(&B.scratch, k::instance_norm_scratch_floats(N, H * W, cp) * sizeof(float)));
    if (p->op == SMELTER_EW_CONCAT && ocp != oc) SM_CUDA(cudaMemsetAsync(B.yo, 0, out_elems * 2, s));
    const __half* xi = static_cast<const __half*>(B.xi);
    const __half* x2i = static_cast<const __half*>(B.x2i);
    __half* yo = static_cast<__half*>(B.yo);
    auto launch = [&]() -> cudaError_t {
        switch (p->op) {
            case SMELTER_EW_UNARY: return k::unary(xi, yo, in_elems, p->sub, p->alpha, p->beta, s, Cc, cp);
            case SMELTER_EW_BINARY: return k::binary(xi, x2i, yo, in_elems, p->sub, p->act, s, Cc, cp);
            case SMELTER_EW_SCALE_SHIFT:
                return k::scale_shift(xi, yo, size_t(N) * H * W, cp, static_cast<const float*>(B.q0), static_cast<const float*>(B.q1), p->act, s);
            case SMELTER_EW_POOL: return k::pool2d(xi, yo, N, H, W, cp, oh, ow, p->k_h, p->k_w, p->stride_h, p->stride_w, p->pad_h, p->pad_w, p->sub, s);
            case SMELTER_EW_GLOBAL_AVGPOOL: return k::global_avgpool(xi, yo, N, H * W, cp, s);
            case SMELTER_EW_SOFTMAX: return k::softmax_rows(xi, yo, size_t(N) * H * W, Cc, cp, p->sub, s);
            case SMELTER_EW_UPSAMPLE: return k::upsample2d(xi, yo, N, H, W, cp, p->scale_h, p->scale_w, p->sub, p->align_corners, s);
            case SMELTER_EW_PAD: return k::pad2d(xi, yo, N, H, W, cp, p->pad_h, p->pad_w, p->pad_b, p->pad_r, p->sub, p->alpha, s);
            case SMELTER_EW_CONCAT: {
                cudaError_t e = k::concat_channels(xi, yo, size_t(N) * H * W, Cc, cp, ocp, 0, s);
                if (e != cudaSuccess) return e;
                return k::concat_channels(x2i, yo, size_t(N) * H * W, p->c2, c2p, ocp, Cc, s);
            }
            case SMELTER_EW_INSTANCE_NORM:
                if (p->sub == 1)  // the one-pass form behind a convolution that supplied the statistics (computed once, untimed, below)
                    return k::instance_norm_from_stats(xi, yo, N, H * W, cp, static_cast<const float*>(B.q0), static_cast<const float*>(B.q1), p->alpha, p->act,
                                                       static_cast<double*>(B.stats), nullptr, 1, s);
                return k::instance_norm(xi, yo, N, H * W, cp, static_cast<const float*>(B.q0), static_cast<const float*>(B.q1), p->alpha, p->act,
                                        static_cast<float*>(B.scratch), s);
            default: return cudaMemcpyAsync(yo, xi, in_elems * 2, cudaMemcpyDeviceToDevice, s);
        }
    };
    if (p->op == SMELTER_EW_INSTANCE_NORM && p->sub == 1) {
        SM_CUDA(cudaMalloc(&B.stats, size_t(N) * cp * 2 * sizeof(double)));
        SM_CUDA(k::instance_norm_stats_f64(xi, N, H * W, cp, static_cast<float*>(B.scratch), static_cast<double*>(B.stats), s));
    }
    SM_CUDA(cudaEventCreate(&B.e0));
    SM_CUDA(cudaEventCreate(&B.e1));
    if (iters < 1) iters = 1;
    SM_CUDA(cudaEventRecord(B.e0, s));
    for (int it = 0; it < iters; ++it) SM_CUDA(launch());
    SM_CUDA(cudaEventRecord(B.e1, s));
    SM_CUDA(k::nhwc_to_nchw(yo, static_cast<__half*>(y), N, oc, oh, ow, ocp, long(oc) * oh * ow, s));
    SM_CUDA(cudaStreamSynchronize(s));
    if (kernel_ms) {
        float ms = 0.f;
        SM_CUDA(cudaEventElapsedTime(&ms, B.e0, B.e1));
        *kernel_ms = ms / float(iters);
    }
    return SMELTER_OK;
}

}  // extern "C"
