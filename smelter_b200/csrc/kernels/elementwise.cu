// Elementwise / data-movement kernels on NHWC fp16 (128-bit vectors = 8 channels per thread access).
//
// These replace the MPS nodes the reference's converters create (Sources/Smelter/Converters.swift):
//   unary   MPSCNNNeuron{ReLU,Sigmoid,TanH,Absolute,Exponential,Logarithm,ELU,HardSigmoid,SoftPlus,SoftSign}Node  :342-476, :1056-1175
//   binary  MPSNN{Addition,Subtraction,Multiplication,Division}Node                                                :430-464, :1177-1211
//   scale_shift  MPSCNNBatchNormalizationNode (un-fused)                                                           :797-827
//   upsample2d   MPSCNNUpsampling{Nearest,Bilinear}Node                                                            :478-552
//   pad2d        MPSNNPadNode                                                                                      :942-989
//   concat       MPSNNConcatenationNode                                                                            :554-574
//   nchw<->nhwc  boundary layout conversion (MPSImage texture upload / MPSImage+Extensions.swift:26-59 read-back)
// All are HBM-bound: one read + one write of each element, grid-stride with grids in multiples of the SM count.
#include "kernels.h"

namespace smelter {
namespace k {

namespace {

constexpr int kThreads = 256;
constexpr int kSMs = 148;

inline int grid_for(size_t work_items, int per_block = kThreads) {
    size_t blocks = (work_items + per_block - 1) / per_block;
    size_t cap = size_t(kSMs) * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return int(blocks);
}

struct alignas(16) Half8 {
    __half2 v[4];
};

__device__ __forceinline__ Half8 ld8(const __half* p) {
    Half8 r;
    *reinterpret_cast<uint4*>(&r) = __ldg(reinterpret_cast<const uint4*>(p));
    return r;
}
__device__ __forceinline__ void st8(__half* p, const Half8& v) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v); }

__device__ __forceinline__ void unpack(const Half8& h, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __half22float2(h.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ Half8 pack(const float (&f)[8]) {
    Half8 h;
#pragma unroll
    for (int i = 0; i < 4; ++i) h.v[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return h;
}

__device__ __forceinline__ float unary_op(float x, int kind, float a, float b) {
    switch (kind) {
        case UN_RELU: return fmaxf(x, 0.f);
        case UN_SIGMOID: return 1.f / (1.f + __expf(-x));
        case UN_CLIP: return fminf(fmaxf(x, a), b);
        case UN_TANH: return tanhf(x);
        case UN_ABS: return fabsf(x);
        case UN_EXP: return __expf(x);
        case UN_LOG: return __logf(x);
        case UN_ELU: return x > 0.f ? x : a * (__expf(x) - 1.f);
        case UN_LEAKY_RELU: return x > 0.f ? x : a * x;
        case UN_HARD_SIGMOID: return fminf(fmaxf(a * x + b, 0.f), 1.f);
        case UN_SOFTPLUS: return x > 20.f ? x : log1pf(__expf(x));
        case UN_SOFTSIGN: return x / (1.f + fabsf(x));
        default: return x;
    }
}

__global__ void __launch_bounds__(kThreads) unary_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t n8, int kind,
                                                        float a, float b) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
        float f[8];
        unpack(ld8(x + i * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = unary_op(f[j], kind, a, b);
        st8(y + i * 8, pack(f));
    }
}

__global__ void __launch_bounds__(kThreads) binary_kernel(const __half* __restrict__ pa, const __half* __restrict__ pb,
                                                         __half* __restrict__ y, size_t n8, int kind, int act) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
        float a[8], b[8];
        unpack(ld8(pa + i * 8), a);
        unpack(ld8(pb + i * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float r;
            switch (kind) {
                case BIN_ADD: r = a[j] + b[j]; break;
                case BIN_SUB: r = a[j] - b[j]; break;
                case BIN_MUL: r = a[j] * b[j]; break;
                default: r = a[j] / b[j]; break;
            }
            a[j] = act == ACT_RELU ? fmaxf(r, 0.f) : r;
        }
        st8(y + i * 8, pack(a));
    }
}

__global__ void __launch_bounds__(kThreads) scale_shift_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t n8, int cp8,
                                                              const float* __restrict__ scale, const float* __restrict__ shift, int act) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n8; i += size_t(gridDim.x) * blockDim.x) {
        const int c = int(i % cp8) * 8;
        float f[8];
        unpack(ld8(x + i * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float r = f[j] * __ldg(scale + c + j) + __ldg(shift + c + j);
            f[j] = act == ACT_RELU ? fmaxf(r, 0.f) : r;
        }
        st8(y + i * 8, pack(f));
    }
}

// One thread per destination pixel; reads are coalesced along W within each channel plane.
__global__ void __launch_bounds__(kThreads) nchw_to_nhwc_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int n, int c,
                                                               int h, int w, int cp, int pt, int pl, int hp, int wp) {
    const size_t total = size_t(n) * hp * wp;
    const size_t plane = size_t(h) * w;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int x = int(i % wp);
        const int y = int((i / wp) % hp);
        const int img = int(i / (size_t(wp) * hp));
        const int sx = x - pl, sy = y - pt;
        const bool inside = sx >= 0 && sx < w && sy >= 0 && sy < h;
        const __half* sp = src + size_t(img) * c * plane + size_t(inside ? sy : 0) * w + (inside ? sx : 0);
        __half* dp = dst + i * cp;
        for (int c0 = 0; c0 < cp; c0 += 8) {
            Half8 v;
            __half* hv = reinterpret_cast<__half*>(&v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = c0 + j;
                hv[j] = (inside && ch < c) ? sp[size_t(ch) * plane] : __float2half(0.f);
            }
            st8(dp + c0, v);
        }
    }
}

// One thread per (pixel, 8-channel group), pixel fastest so plane writes are coalesced along W.
__global__ void __launch_bounds__(kThreads) nhwc_to_nchw_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int n, int c,
                                                               int hw, int cp, long dst_image_pitch) {
    const int groups = cp / 8;
    const size_t total = size_t(n) * groups * hw;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int pix = int(i % hw);
        const int g = int((i / hw) % groups);
        const int img = int(i / (size_t(hw) * groups));
        Half8 v = ld8(src + (size_t(img) * hw + pix) * cp + g * 8);
        const __half* hv = reinterpret_cast<const __half*>(&v);
        __half* dp = dst + size_t(img) * dst_image_pitch + pix;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = g * 8 + j;
            if (ch < c) dp[size_t(ch) * hw] = hv[j];
        }
    }
}

__global__ void __launch_bounds__(kThreads) upsample_nearest_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h,
                                                                   int w, int cp8, int sh, int sw) {
    const int ho = h * sh, wo = w * sw;
    const size_t total = size_t(n) * ho * wo * cp8;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int g = int(i % cp8);
        size_t pix = i / cp8;
        const int ox = int(pix % wo);
        const int oy = int((pix / wo) % ho);
        const int img = int(pix / (size_t(wo) * ho));
        const size_t sidx = ((size_t(img) * h + oy / sh) * w + ox / sw) * cp8 + g;
        st8(y + i * 8, ld8(x + sidx * 8));
    }
}

// ONNX Upsample 'linear': align_corners=1 maps corners to corners ((in-1)/(out-1)); align_corners=0 is the
// asymmetric opset-9 rule (src = dst / scale).  The reference forwards Configuration.alignCorners to MPS
// (Converters.swift:529-536).
__global__ void __launch_bounds__(kThreads) upsample_bilinear_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h,
                                                                    int w, int cp8, int sh, int sw, int align) {
    const int ho = h * sh, wo = w * sw;
    const size_t total = size_t(n) * ho * wo * cp8;
    const float ry = align ? (ho > 1 ? float(h - 1) / float(ho - 1) : 0.f) : 1.f / float(sh);
    const float rx = align ? (wo > 1 ? float(w - 1) / float(wo - 1) : 0.f) : 1.f / float(sw);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int g = int(i % cp8);
        size_t pix = i / cp8;
        const int ox = int(pix % wo);
        const int oy = int((pix / wo) % ho);
        const int img = int(pix / (size_t(wo) * ho));
        const float fy = oy * ry, fx = ox * rx;
        int y0 = min(int(fy), h - 1), x0 = min(int(fx), w - 1);
        const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
        const float wy = fy - y0, wx = fx - x0;
        const __half* base = x + size_t(img) * h * w * cp8 * 8 + g * 8;
        float a[8], b[8], c[8], d[8];
        unpack(ld8(base + (size_t(y0) * w + x0) * cp8 * 8), a);
        unpack(ld8(base + (size_t(y0) * w + x1) * cp8 * 8), b);
        unpack(ld8(base + (size_t(y1) * w + x0) * cp8 * 8), c);
        unpack(ld8(base + (size_t(y1) * w + x1) * cp8 * 8), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float top = a[j] + (b[j] - a[j]) * wx;
            const float bot = c[j] + (d[j] - c[j]) * wx;
            a[j] = top + (bot - top) * wy;
        }
        st8(y + i * 8, pack(a));
    }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // ONNX 'reflect' (no edge repeat); pads < n assumed
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(kThreads) pad2d_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int cp8,
                                                        int pt, int pl, int ho, int wo, int mode, float value) {
    const size_t total = size_t(n) * ho * wo * cp8;
    Half8 fill;
#pragma unroll
    for (int j = 0; j < 4; ++j) fill.v[j] = __floats2half2_rn(value, value);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int g = int(i % cp8);
        size_t pix = i / cp8;
        const int ox = int(pix % wo);
        const int oy = int((pix / wo) % ho);
        const int img = int(pix / (size_t(wo) * ho));
        int sx = ox - pl, sy = oy - pt;
        bool inside = sx >= 0 && sx < w && sy >= 0 && sy < h;
        if (mode == PAD_REFLECT) {
            sx = reflect_idx(sx, w); sy = reflect_idx(sy, h); inside = true;
        } else if (mode == PAD_EDGE) {
            sx = min(max(sx, 0), w - 1); sy = min(max(sy, 0), h - 1); inside = true;
        }
        if (inside) st8(y + i * 8, ld8(x + (((size_t(img) * h + sy) * w + sx) * cp8 + g) * 8));
        else st8(y + i * 8, fill);
    }
}

__global__ void __launch_bounds__(kThreads) concat_vec_kernel(const __half* __restrict__ src, __half* __restrict__ dst, size_t pixels,
                                                             int src8, int dst_pitch, int c_off) {
    const size_t total = pixels * src8;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int g = int(i % src8);
        const size_t pix = i / src8;
        st8(dst + pix * dst_pitch + c_off + g * 8, ld8(src + i * 8));
    }
}
__global__ void __launch_bounds__(kThreads) concat_scalar_kernel(const __half* __restrict__ src, __half* __restrict__ dst, size_t pixels,
                                                                int c_src, int src_pitch, int dst_pitch, int c_off) {
    const size_t total = pixels * c_src;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int ch = int(i % c_src);
        const size_t pix = i / c_src;
        dst[pix * dst_pitch + c_off + ch] = src[pix * src_pitch + ch];
    }
}

__global__ void __launch_bounds__(kThreads) f32_to_f16_kernel(const float* __restrict__ s, __half* __restrict__ d, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) d[i] = __float2half_rn(s[i]);
}
__global__ void __launch_bounds__(kThreads) f16_to_f32_kernel(const __half* __restrict__ s, float* __restrict__ d, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) d[i] = __half2float(s[i]);
}

__global__ void __launch_bounds__(kThreads) checksum_kernel(const uint32_t* __restrict__ p, size_t n_words, unsigned long long* out) {
    unsigned long long acc = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n_words; i += size_t(gridDim.x) * blockDim.x) {
        acc += (unsigned long long)(p[i]) * (2654435761ull + 2ull * (unsigned long long)(i & 0xffffffffull));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace

cudaError_t unary(const __half* x, __half* y, size_t n_elems, int kind, float alpha, float beta, cudaStream_t s) {
    const size_t n8 = n_elems / 8;
    unary_kernel<<<grid_for(n8), kThreads, 0, s>>>(x, y, n8, kind, alpha, beta);
    return cudaGetLastError();
}
cudaError_t binary(const __half* a, const __half* b, __half* y, size_t n_elems, int kind, int act, cudaStream_t s) {
    const size_t n8 = n_elems / 8;
    binary_kernel<<<grid_for(n8), kThreads, 0, s>>>(a, b, y, n8, kind, act);
    return cudaGetLastError();
}
cudaError_t scale_shift(const __half* x, __half* y, size_t pixels, int cp, const float* scale, const float* shift, int act, cudaStream_t s) {
    const size_t n8 = pixels * (cp / 8);
    scale_shift_kernel<<<grid_for(n8), kThreads, 0, s>>>(x, y, n8, cp / 8, scale, shift, act);
    return cudaGetLastError();
}
cudaError_t nchw_to_nhwc(const __half* src, __half* dst, int n, int c, int h, int w, int cp, int pad_t, int pad_l, int pad_b, int pad_r,
                         cudaStream_t s) {
    const int hp = h + pad_t + pad_b, wp = w + pad_l + pad_r;
    nchw_to_nhwc_kernel<<<grid_for(size_t(n) * hp * wp), kThreads, 0, s>>>(src, dst, n, c, h, w, cp, pad_t, pad_l, hp, wp);
    return cudaGetLastError();
}
cudaError_t nhwc_to_nchw(const __half* src, __half* dst, int n, int c, int h, int w, int cp, long dst_image_pitch, cudaStream_t s) {
    nhwc_to_nchw_kernel<<<grid_for(size_t(n) * (cp / 8) * h * w), kThreads, 0, s>>>(src, dst, n, c, h * w, cp, dst_image_pitch);
    return cudaGetLastError();
}
cudaError_t upsample2d(const __half* x, __half* y, int n, int h, int w, int cp, int scale_h, int scale_w, int mode, int align_corners,
                       cudaStream_t s) {
    const size_t total = size_t(n) * h * scale_h * w * scale_w * (cp / 8);
    if (mode == UP_NEAREST) upsample_nearest_kernel<<<grid_for(total), kThreads, 0, s>>>(x, y, n, h, w, cp / 8, scale_h, scale_w);
    else upsample_bilinear_kernel<<<grid_for(total), kThreads, 0, s>>>(x, y, n, h, w, cp / 8, scale_h, scale_w, align_corners);
    return cudaGetLastError();
}
cudaError_t pad2d(const __half* x, __half* y, int n, int h, int w, int cp, int pt, int pl, int pb, int pr, int mode, float value,
                  cudaStream_t s) {
    const int ho = h + pt + pb, wo = w + pl + pr;
    pad2d_kernel<<<grid_for(size_t(n) * ho * wo * (cp / 8)), kThreads, 0, s>>>(x, y, n, h, w, cp / 8, pt, pl, ho, wo, mode, value);
    return cudaGetLastError();
}
cudaError_t concat_channels(const __half* src, __half* dst, size_t pixels, int c_src, int c_src_pitch, int c_dst_pitch, int c_off,
                            cudaStream_t s) {
    if (c_off % 8 == 0 && c_src % 8 == 0) {
        concat_vec_kernel<<<grid_for(pixels * (c_src / 8)), kThreads, 0, s>>>(src, dst, pixels, c_src / 8, c_dst_pitch, c_off);
        // vector path assumes src pitch == c_src (true when c_src % 8 == 0)
        (void)c_src_pitch;
    } else {
        concat_scalar_kernel<<<grid_for(pixels * c_src), kThreads, 0, s>>>(src, dst, pixels, c_src, c_src_pitch, c_dst_pitch, c_off);
    }
    return cudaGetLastError();
}
cudaError_t f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t s) {
    f32_to_f16_kernel<<<grid_for(n), kThreads, 0, s>>>(src, dst, n);
    return cudaGetLastError();
}
cudaError_t f16_to_f32(const __half* src, float* dst, size_t n, cudaStream_t s) {
    f16_to_f32_kernel<<<grid_for(n), kThreads, 0, s>>>(src, dst, n);
    return cudaGetLastError();
}
cudaError_t checksum64(const void* p, size_t bytes, unsigned long long* out_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    const size_t n_words = bytes / 4;
    checksum_kernel<<<grid_for(n_words), kThreads, 0, s>>>(reinterpret_cast<const uint32_t*>(p), n_words, out_dev);
    return cudaGetLastError();
}

}  // namespace k
}  // namespace smelter
