/* smelter_b200 — C ABI of the B200-native ONNX graph inference engine.
 *
 * This header is the drop-in boundary for Smelter's inference path.  Every entry point
 * names the reference interface it replaces (paths relative to the Smelter repository).  All arguments are
 * plain pointers / sizes / PODs; there are no C++ or torch types in any signature.  Every function returns
 * an int32 status (smelter_status); none aborts.  There is no CPU fallback anywhere behind this ABI: a
 * build without a usable sm_100a device fails with SMELTER_ERR_CUDA.
 *
 * Layout at the boundary: activations are dense NCHW fp16 device buffers (what `MPSImage` was for the
 * reference); internally the engine runs NHWC with channels padded to a multiple of 8.
 */
#ifndef SMELTER_B200_H_
#define SMELTER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMELTER_B200_ABI_VERSION 2

/* ---- status codes -----------------------------------------------------------------------------------
 * 1..8 are ONNXGraph.Errors in declaration order (Sources/Smelter/ONNXGraph.swift:38-47). */
typedef enum smelter_status {
    SMELTER_OK = 0,
    SMELTER_ERR_UNSUPPORTED_INPUT = 1,
    SMELTER_ERR_UNSUPPORTED_OUTPUT = 2,
    SMELTER_ERR_UNKNOWN_NODE_OP_TYPE = 3, /* smelter_last_error() carries the op type */
    SMELTER_ERR_NO_SUCH_OUTPUT = 4,
    SMELTER_ERR_GRAPH_INTERNAL = 5,
    SMELTER_ERR_INSUFFICIENT_INPUTS = 6,
    SMELTER_ERR_INCONSISTENT_STATE = 7,
    SMELTER_ERR_NOT_ENOUGH_ATTRIBUTES = 8,
    SMELTER_ERR_INVALID_ARGUMENT = 100,
    SMELTER_ERR_PARSE = 101, /* malformed protobuf (SwiftProtobuf decoding error in the reference) */
    SMELTER_ERR_CUDA = 102,
    SMELTER_ERR_NCCL = 103,
    SMELTER_ERR_UNSUPPORTED = 104 /* attribute combination outside the engine (never a CPU fallback) */
} smelter_status;

/* ONNXGraph.Format (ONNXGraph.swift:49-52) */
typedef enum smelter_format { SMELTER_FORMAT_ONNX = 0, SMELTER_FORMAT_MPS_FLAVOR = 1 } smelter_format;

/* ONNXGraph.Configuration (ONNXGraph.swift:6-36). */
typedef enum smelter_input_constraint {
    SMELTER_INPUT_NONE = 0,                 /* .none */
    SMELTER_INPUT_FORCE_SCALE_LANCZOS = 1,  /* .forceInputScale(.lanczos): sources of any H x W are resampled (Lanczos-3) to the graph input */
    SMELTER_INPUT_FORCE_SCALE_BILINEAR = 2  /* .forceInputScale(.bilinear): same with bilinear interpolation (half-pixel centres) */
} smelter_input_constraint;

typedef struct smelter_config {
    int32_t input_constraint;        /* smelter_input_constraint */
    int32_t bilinear_align_corners;  /* Configuration.BillinearUpsampling.alignCorners; reference default 1 */
    int32_t n_dims;                  /* number of (axis,value) overrides below (Configuration.dims) */
    int32_t dims_axis[8];
    int64_t dims_value[8];
    /* engine options (no reference counterpart) */
    int32_t enable_fusion;           /* 1 (default): fold BN, fuse bias/activation/residual into Conv epilogues */
    int32_t use_cuda_graph;          /* 1 (default): replay encode() from a captured CUDA graph */
    int32_t defer_weights;           /* 1: build() leaves the device weight arena zero-filled; the caller must fill it with
                                        smelter_graph_broadcast_weights (non-root ranks of a multi-GPU job). default 0 */
    int32_t sm_share;                /* k >= 2: plan and launch every kernel of this graph for 1/k of the SMs, so that encodes in
                                        flight on different streams (the reference's asynchronous command buffers) co-run on
                                        disjoint parts of the chip instead of taking turns with whole-chip grids; one encode
                                        is slower, k + 1 in flight are faster.  default 1 (whole chip).  ABI version 2 */
} smelter_config;

/* Shape (Sources/Smelter/TypeDefinitions.swift:1-33) */
typedef struct smelter_shape {
    int32_t channels, width, height, depth;
} smelter_shape;

typedef struct smelter_context smelter_context; /* replaces MTLDevice/MTLContext: device + default stream (+ NCCL comm) */
typedef struct smelter_graph smelter_graph;     /* replaces ONNXGraph + the MPSNNGraph it builds */
typedef struct smelter_tensor smelter_tensor;   /* replaces MPSImage: NCHW fp16 device buffer */

/* Thread-local message of the last failing call on this thread. */
const char* smelter_last_error(void);
int32_t smelter_abi_version(void);

/* Fill *cfg with Configuration.init defaults (ONNXGraph.swift:27-35) and engine defaults. */
void smelter_config_default(smelter_config* cfg);

/* ---- context ------------------------------------------------------------------------------------------
 * README.md:18 `MTLContext()` → device + stream.  `stream` may be NULL (the context creates one). */
int32_t smelter_context_create(int32_t device, void* cuda_stream, smelter_context** out);
int32_t smelter_context_destroy(smelter_context* ctx);
int32_t smelter_context_stream(smelter_context* ctx, void** cuda_stream);
int32_t smelter_context_synchronize(smelter_context* ctx);

/* Multi-GPU (new; the reference is single-device): one process per GPU.  Rank 0 obtains a 128-byte NCCL
 * unique id, the host distributes it, every rank calls init.  Used only for the weight-arena broadcast. */
int32_t smelter_nccl_unique_id(uint8_t id[128]);
int32_t smelter_context_init_nccl(smelter_context* ctx, const uint8_t id[128], int32_t rank, int32_t world);

/* ---- tensors (MPSImage) ------------------------------------------------------------------------------- */
int32_t smelter_tensor_create(smelter_context* ctx, int32_t n, int32_t c, int32_t h, int32_t w, smelter_tensor** out);
int32_t smelter_tensor_wrap(smelter_context* ctx, void* device_ptr_fp16_nchw, int32_t n, int32_t c, int32_t h, int32_t w,
                            smelter_tensor** out); /* borrows device memory */
int32_t smelter_tensor_destroy(smelter_tensor* t);
int32_t smelter_tensor_dims(const smelter_tensor* t, int32_t dims[4]);
int32_t smelter_tensor_device_ptr(const smelter_tensor* t, void** ptr);
/* README.md:33-39 `texture(from:)` analogue: host fp32 NCHW → device fp16 NCHW (async on `cuda_stream`;
 * `host` should be pinned for true asynchrony). */
int32_t smelter_tensor_from_float(smelter_tensor* t, void* cuda_stream, const float* host, size_t count);
int32_t smelter_tensor_from_half(smelter_tensor* t, void* cuda_stream, const uint16_t* host, size_t count);
/* The step in front of the path in real use (README.md:33-39, `MTLContext.texture(from: CGImage)`): interleaved 8-bit pixels
 * [N][H][W][src_channels] (RGB, RGBA, grey ...) -> device fp16 NCHW, value = byte * scale[c] + bias[c] for the tensor's first
 * C <= src_channels channels (scale / bias: C floats each, NULL = 1/255 and 0).  Asynchronous on `cuda_stream`. */
int32_t smelter_tensor_from_u8(smelter_tensor* t, void* cuda_stream, const uint8_t* host, int32_t src_channels, const float* scale,
                               const float* bias);
/* MPSImage.toFloatArray() (Extensions/Foundation/MPSImage+Extensions.swift:9-59): device fp16 → host fp32,
 * NCHW order (the reference returns MPS slice order).  Synchronises the stream. */
int32_t smelter_tensor_to_float(const smelter_tensor* t, void* cuda_stream, float* host, size_t capacity);
/* toFloatArray() in the reference's own element order (MPSImage+Extensions.swift:26-59): per image, slices of four channels,
 * each slice [H][W][4] (for C < 3: one slice [H][W][C]); C is padded with zeros to a multiple of 4 when C >= 3.  `capacity` must be
 * at least N * H * W * (C < 3 ? C : 4 * ceil(C / 4)).  Synchronises the stream. */
int32_t smelter_tensor_to_float_mps(const smelter_tensor* t, void* cuda_stream, float* host, size_t capacity);
/* Same, but only enqueues the conversion and the device→host copy on the stream (the Metal analogue: read the image in a
 * command buffer's completion handler instead of after waitUntilCompleted()).  `host` must stay valid — and should be pinned —
 * until the stream reaches this point; the tensor may be overwritten by a later encode() on the same stream. */
int32_t smelter_tensor_to_float_async(const smelter_tensor* t, void* cuda_stream, float* host, size_t capacity);
int32_t smelter_tensor_to_half(const smelter_tensor* t, void* cuda_stream, uint16_t* host, size_t capacity);

/* ---- graph construction --------------------------------------------------------------------------------
 * ONNXGraph.init(data:configuration:) (ONNXGraph.swift:95-156): parse, detect the ONNX2MPS flavour, index
 * initializers, register converters.  `onnx` is borrowed for the duration of the call only. */
int32_t smelter_graph_create(smelter_context* ctx, const uint8_t* onnx, size_t len, const smelter_config* cfg,
                             smelter_graph** out);
/* ONNXGraph.metalGraph(device:) (ONNXGraph.swift:169-193): initOutputs, walk nodes in file order through the
 * converter registry, require exactly one graph output, then compile (fusion, weight packing + upload,
 * activation arena).  Replaces MPSNNGraph(device:resultImage:resultImageIsNeeded:). */
int32_t smelter_graph_build(smelter_graph* g);
int32_t smelter_graph_destroy(smelter_graph* g);

/* Introspection: ONNXGraph.modelFormat (:58), outputShapes (:69-91), converter registry (:110-155). */
int32_t smelter_graph_format(const smelter_graph* g, int32_t* format);
int32_t smelter_graph_num_outputs(const smelter_graph* g, int32_t* n);
int32_t smelter_graph_output_shape(const smelter_graph* g, int32_t idx, smelter_shape* shape);
int32_t smelter_graph_num_nodes(const smelter_graph* g, int32_t* n);
int32_t smelter_graph_node_op_type(const smelter_graph* g, int32_t idx, const char** op_type);
int32_t smelter_graph_has_converter(const smelter_graph* g, const char* op_type, int32_t* yes);
/* Engine introspection (after build): number of device kernels one encode() launches for batch `n`,
 * and a human-readable plan dump (one line per launched kernel). */
int32_t smelter_graph_num_launches(smelter_graph* g, int32_t batch, int32_t* n);
int32_t smelter_graph_plan_dump(smelter_graph* g, int32_t batch, char* buf, size_t cap);
/* Per-kernel timing of one encode: runs the plan for sources' batch WITHOUT the CUDA graph, every kernel launch
 * bracketed by a CUDA-event pair on `cuda_stream`, `iters` passes; ms[i] = mean duration of plan step i,
 * flops[i] / bytes[i] = its algorithmic work (conv: 2*M*N*K; others: 2 B x elements read + written), is_tensor[i] = 1
 * for the tcgen05 conv/gemm kernel.  *n_steps receives the step count (call with cap = 0 to size the arrays). */
int32_t smelter_graph_profile(smelter_graph* g, void* cuda_stream, const smelter_tensor* const* sources, int32_t n_sources,
                              int32_t iters, float* ms, double* flops, double* bytes, int32_t* is_tensor, int32_t cap,
                              int32_t* n_steps);

/* ---- inference ------------------------------------------------------------------------------------------
 * MPSNNGraph.encode(to:sourceImages:) (README.md:43-44): enqueue only, no synchronisation.  `*result` is
 * owned by the graph and valid until the next encode on this graph with the same `cuda_stream` (NULL = the context's).
 * Like Metal command buffers, encodes on DIFFERENT streams may be in flight together: each (batch size, stream) owns its
 * activation arena and captured CUDA graph (at most 8 streams per batch size); on one stream they run in order.
 * Batch = leading dim of sources[0]. */
int32_t smelter_graph_encode(smelter_graph* g, void* cuda_stream, const smelter_tensor* const* sources, int32_t n_sources,
                             const smelter_tensor** result);

/* Multi-GPU weight replica: ncclBroadcast of the packed weight arena from `root`, and a
 * 64-bit checksum of the arena (computed on the device) to prove replicas are identical. */
int32_t smelter_graph_broadcast_weights(smelter_graph* g, int32_t root);
int32_t smelter_graph_weight_checksum(smelter_graph* g, uint64_t* checksum, uint64_t* bytes);
/* The packed weight arena itself (device pointer + size), so a host that already owns a communicator
 * (e.g. torch.distributed under torchrun) can broadcast it instead of smelter_graph_broadcast_weights. */
int32_t smelter_graph_weight_arena(smelter_graph* g, void** device_ptr, uint64_t* bytes);

/* ---- fine-grained builder (what a Swift NodeConverter would call) ----------------------------------------
 * Mirrors the five-call surface converters use on ONNXGraph (ONNXGraph.swift:259-285): output(name:),
 * shape(output:), tensor(name:), initTensor, addFilter.  The C++ converter table inside the library calls
 * exactly these; a host-language converter can replace any of them via smelter_graph_register_converter. */
typedef enum smelter_act {
    SMELTER_ACT_NONE = 0, SMELTER_ACT_RELU = 1, SMELTER_ACT_CLIP = 2, SMELTER_ACT_SIGMOID = 3
} smelter_act;
typedef enum smelter_dtype { SMELTER_F32 = 1, SMELTER_F16 = 10 } smelter_dtype; /* TensorProto.DataType values */
typedef enum smelter_weight_layout { SMELTER_OIHW = 0, SMELTER_OHWI = 1 } smelter_weight_layout;
typedef enum smelter_unary {
    SMELTER_UNARY_RELU = 0, SMELTER_UNARY_SIGMOID = 1, SMELTER_UNARY_CLIP = 2, SMELTER_UNARY_TANH = 3,
    SMELTER_UNARY_ABS = 4, SMELTER_UNARY_EXP = 5, SMELTER_UNARY_LOG = 6, SMELTER_UNARY_ELU = 7,
    SMELTER_UNARY_LEAKY_RELU = 8, SMELTER_UNARY_HARD_SIGMOID = 9, SMELTER_UNARY_SOFTPLUS = 10,
    SMELTER_UNARY_SOFTSIGN = 11, SMELTER_UNARY_IDENTITY = 12
} smelter_unary;
typedef enum smelter_binary { SMELTER_BIN_ADD = 0, SMELTER_BIN_SUB = 1, SMELTER_BIN_MUL = 2, SMELTER_BIN_DIV = 3 } smelter_binary;
typedef enum smelter_pad_mode { SMELTER_PAD_CONSTANT = 0, SMELTER_PAD_REFLECT = 1, SMELTER_PAD_EDGE = 2 } smelter_pad_mode;
typedef enum smelter_upsample_mode { SMELTER_UPSAMPLE_NEAREST = 0, SMELTER_UPSAMPLE_BILINEAR = 1 } smelter_upsample_mode;

typedef struct smelter_conv_desc { /* ConvolutionConverter.convert, Converters.swift:188-337 */
    int32_t c_out, c_in_per_group, k_h, k_w;
    int32_t stride_h, stride_w, dil_h, dil_w, groups;
    int32_t pads[4];            /* top, left, bottom, right (Pads, ONNXConvolutionPadding.swift:6) */
    int32_t weight_dtype;       /* smelter_dtype */
    int32_t weight_layout;      /* smelter_weight_layout: OHWI when the model is ONNX2MPS-flavoured */
    int32_t bias_dtype;         /* smelter_dtype, ignored when bias == NULL */
    int32_t is_gemm;            /* Gemm as 1x1 fully connected (Converters.swift:228-232, 288-302) */
} smelter_conv_desc;

/* Symbol tables. */
int32_t smelter_graph_has_output(const smelter_graph* g, const char* name, int32_t* yes);          /* output(name:) */
int32_t smelter_graph_shape(const smelter_graph* g, const char* name, smelter_shape* shape);       /* shape(output:) */
int32_t smelter_graph_has_tensor(const smelter_graph* g, const char* name, int32_t* yes);          /* tensor(name:) */

/* addFilter family.  `out_name` becomes a new image node; inputs must exist (else NO_SUCH_OUTPUT).
 * Host weight pointers are read during the call only (copied / repacked). */
int32_t smelter_add_conv(smelter_graph* g, const char* in_name, const smelter_conv_desc* d, const void* weights,
                         const void* bias, const char* out_name);
int32_t smelter_add_batchnorm(smelter_graph* g, const char* in_name, int32_t channels, const float* gamma, const float* beta,
                              const float* mean, const float* var, float epsilon, const char* out_name);
int32_t smelter_add_instancenorm(smelter_graph* g, const char* in_name, int32_t channels, const float* gamma,
                                 const float* beta, float epsilon, const char* out_name);
int32_t smelter_add_unary(smelter_graph* g, const char* in_name, int32_t kind, float alpha, float beta, const char* out_name);
int32_t smelter_add_binary(smelter_graph* g, const char* a_name, const char* b_name, int32_t kind, const char* out_name);
int32_t smelter_add_pool(smelter_graph* g, const char* in_name, int32_t is_max, int32_t k_h, int32_t k_w, int32_t stride_h,
                         int32_t stride_w, int32_t pad_h, int32_t pad_w, const char* out_name);
int32_t smelter_add_global_avgpool(smelter_graph* g, const char* in_name, const char* out_name);
int32_t smelter_add_upsample(smelter_graph* g, const char* in_name, int32_t mode, int32_t scale_h, int32_t scale_w,
                             int32_t align_corners, const char* out_name);
int32_t smelter_add_concat(smelter_graph* g, const char* const* in_names, int32_t n_inputs, const char* out_name);
int32_t smelter_add_reshape(smelter_graph* g, const char* in_name, int32_t c, int32_t h, int32_t w, const char* out_name);
int32_t smelter_add_softmax(smelter_graph* g, const char* in_name, int32_t log_softmax, const char* out_name);
int32_t smelter_add_pad(smelter_graph* g, const char* in_name, int32_t mode, const int32_t pads_nchw[8], float value,
                        const char* out_name);
int32_t smelter_add_alias(smelter_graph* g, const char* in_name, const char* out_name); /* Dropout/Identity at inference */

/* NodeConverter protocol (NodeConverter.swift:3-5) for host-language plugins: `fn` is called with the node
 * index when the walk reaches a node whose op_type matches; it uses the builder calls above. */
typedef int32_t (*smelter_converter_fn)(smelter_graph* g, int32_t node_index, void* user);
int32_t smelter_graph_register_converter(smelter_graph* g, const char* op_type, smelter_converter_fn fn, void* user);
/* Node access for plugins. */
int32_t smelter_node_num_inputs(const smelter_graph* g, int32_t node, int32_t* n);
int32_t smelter_node_input(const smelter_graph* g, int32_t node, int32_t i, const char** name);
int32_t smelter_node_num_outputs(const smelter_graph* g, int32_t node, int32_t* n);
int32_t smelter_node_output(const smelter_graph* g, int32_t node, int32_t i, const char** name);
int32_t smelter_node_attr_int(const smelter_graph* g, int32_t node, const char* attr, int64_t* v, int32_t* found);
int32_t smelter_node_attr_float(const smelter_graph* g, int32_t node, const char* attr, float* v, int32_t* found);
int32_t smelter_node_attr_ints(const smelter_graph* g, int32_t node, const char* attr, int64_t* v, int32_t cap, int32_t* n);

/* ---- host utilities on the path (exported so they can be tested against the reference's index maps) ------
 * Array.reformatingConvolutionWeight (Extensions/Foundation/Array+Extensions.swift:52-93): OIHW→OHWI, and for
 * ConvTranspose IOHW→OHWI with a 180° spatial flip.  `elem_size` 2 or 4. */
int32_t smelter_reformat_conv_weight(const void* src, void* dst, int32_t elem_size, int32_t c_out, int32_t c_in, int32_t k_h,
                                     int32_t k_w, int32_t is_transpose);
/* Float16.swift:17-45 / 53-77 */
int32_t smelter_float16_to_32(const uint16_t* src, float* dst, size_t n);
int32_t smelter_float32_to_16(const float* src, uint16_t* dst, size_t n);
/* ONNX_ConvolutionPadding.paddedSize (Padding/ONNXConvolutionPadding.swift:91-113; dilation honoured)
 * and PyTorchPoolPadding.paddedSize (Padding/PyTorchPoolPadding.swift:94-103). */
int32_t smelter_conv_output_size(int32_t in, int32_t k, int32_t stride, int32_t dil, int32_t pad_lo, int32_t pad_hi,
                                 int32_t out_pad, int32_t is_transpose, int32_t* out);
int32_t smelter_pool_output_size(int32_t in, int32_t k, int32_t stride, int32_t pad, int32_t* out);
/* TensorProto coercions (Onnx_TensorProto+Extensions.swift:2-62) on a serialized TensorProto. */
int32_t smelter_tensorproto_integers(const uint8_t* tensor_proto, size_t len, int64_t* out, size_t cap, size_t* n);
int32_t smelter_tensorproto_floats(const uint8_t* tensor_proto, size_t len, float* out, size_t cap, size_t* n);

/* ---- kernel-level entry points (tests / benchmarks drive single kernels through these) -------------------- */
typedef struct smelter_conv_problem {
    int32_t n, h, w, c_in, c_out, k_h, k_w, stride_h, stride_w, dil_h, dil_w, pad_t, pad_l, pad_b, pad_r, groups;
    int32_t act;       /* smelter_act */
    float clip_lo, clip_hi;
    int32_t has_bias, has_residual;
    int32_t force_path; /* 0 auto, 1 tiled-TMA 1x1, 2 im2col-TMA, 3 packed-row small-Cin, 4 depthwise */
} smelter_conv_problem;
/* x: NCHW fp16 device, w: OIHW fp16 HOST, bias: fp32 HOST, residual: NCHW fp16 device (same shape as y), y: NCHW fp16
 * device.  Runs layout conversion + the conv kernel + conversion back; *kernel_ms (optional) receives the
 * CUDA-event time of `iters` back-to-back launches of the conv kernel alone, divided by iters. */
int32_t smelter_run_conv(smelter_context* ctx, const smelter_conv_problem* p, const void* x, const void* w, const float* bias,
                         const void* residual, void* y, int32_t iters, float* kernel_ms);


/* Single non-GEMM kernels (everything that is HBM-bound on the path).  x / x2 / y are NCHW fp16 DEVICE buffers,
 * p0 / p1 are fp32 HOST arrays of length c (scale/shift for SCALE_SHIFT, gamma/beta for INSTANCE_NORM).  The
 * driver converts to the engine's NHWC layout, launches the kernel `iters` times between two CUDA events on the
 * context's stream, and converts the last result back.  *kernel_ms = event time / iters (kernel only). */
typedef enum smelter_ew_op {
    SMELTER_EW_UNARY = 0,          /* sub = smelter_unary, alpha/beta = parameters           Converters.swift:342-476, 1056-1175 */
    SMELTER_EW_BINARY = 1,         /* sub = smelter_binary, act = fused ReLU                 :430-464, 1177-1211 */
    SMELTER_EW_SCALE_SHIFT = 2,    /* un-fused BatchNormalization, act = fused ReLU          :797-827 */
    SMELTER_EW_POOL = 3,           /* sub = is_max; k/stride/pad_h/pad_w                     :607-695 */
    SMELTER_EW_GLOBAL_AVGPOOL = 4, /*                                                        :578-605 */
    SMELTER_EW_SOFTMAX = 5,        /* sub = log-softmax flag; channel axis                   :697-714, 1213-1231 */
    SMELTER_EW_UPSAMPLE = 6,       /* sub = smelter_upsample_mode; scale_h/w; align_corners  :478-552 */
    SMELTER_EW_PAD = 7,            /* sub = smelter_pad_mode; pad_h=top pad_w=left pad_b pad_r; alpha = value  :942-989 */
    SMELTER_EW_CONCAT = 8,         /* x (c channels) ++ x2 (c2 channels)                     :554-574 */
    SMELTER_EW_INSTANCE_NORM = 9,  /* alpha = epsilon, act = fused ReLU; sub = 1: the one-pass form whose statistics the producing
                                      convolution supplies (here computed once, outside the timed launches)   :992-1054 */
    SMELTER_EW_LAYOUT_ROUNDTRIP = 10 /* NCHW -> NHWC -> NCHW boundary conversion only        MPSImage+Extensions.swift:26-59 */
} smelter_ew_op;
typedef struct smelter_ew_problem {
    int32_t op, n, c, h, w, c2, sub, act;
    int32_t k_h, k_w, stride_h, stride_w, pad_h, pad_w, pad_b, pad_r;
    int32_t scale_h, scale_w, align_corners;
    float alpha, beta;
} smelter_ew_problem;
int32_t smelter_run_elementwise(smelter_context* ctx, const smelter_ew_problem* p, const void* x, const void* x2, const float* p0,
                                const float* p1, void* y, int32_t iters, float* kernel_ms);
/* Benchmark only: sustained TMA load rate of [128 pixel x 64 channel] fp16 boxes from an NHWC tensor [n,h,w,c] (zero-filled
 * scratch), `stages` loads in flight per CTA, `iters` loads per CTA, `grid` CTAs; mode 0 = 2-D tiled, 1 = im2col (3x3 pad 1). */
int32_t smelter_tma_probe(smelter_context* ctx, int32_t mode, int32_t c, int32_t w, int32_t h, int32_t n, int32_t stages, int32_t iters,
                          int32_t grid, int32_t distinct, float* ms);
/* Overwrite a buffer larger than L2 (126 MB) on the context's stream: benchmarks call it between timed iterations. */
int32_t smelter_l2_flush(smelter_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SMELTER_B200_H_ */
