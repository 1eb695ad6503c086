// Window / reduction kernels on NHWC fp16: pooling, global average pool, channel softmax, instance norm,
// depthwise convolution.  Replaces (Sources/Smelter/Converters.swift):
//   pool2d          MPSCNNPooling{Max,Average}Node + PyTorchPoolPadding     :607-695, Padding/PyTorchPoolPadding.swift
//   global_avgpool  MPSCNNPoolingAverageNode + GlobalPoolPadding            :578-605
//   softmax_rows    MPSCNN{SoftMax,LogSoftMax}Node                          :697-714, :1213-1231
//   instance_norm   MPSCNNInstanceNormalizationNode                         :992-1054
//   depthwise_conv  MPSCNNConvolutionNode with MPSCNNDepthWiseConvolutionDescriptor  :57-66
// All HBM-bound; fp32 math, fp16 storage; warp-shuffle reductions.
#include <algorithm>
#include <cfloat>

#include "kernels.h"
#include "pdl.cuh"

#include <cooperative_groups.h>

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
__device__ __forceinline__ float act_op(float v, int act, float lo, float hi) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_CLIP) return fminf(fmaxf(v, lo), hi);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

// ---- pooling: one thread per (output pixel, 8 channels); floor mode; average divides by the full window
//      (count_include_pad=1, the MPS/PyTorch default the reference relies on, Converters.swift:609-616).
__global__ void __launch_bounds__(kThreads) pool2d_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int cp8,
                                                         int p, int q, int kh, int kw, int sh, int sw, int ph, int pw, int is_max) {
    pdl_prologue();
    // grid.y walks output rows (image, op); threads walk (oq, channel group) of that row
    const int op = blockIdx.y % p;
    const int img = blockIdx.y / p;
    const int row_items = q * cp8;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < row_items; it += gridDim.x * blockDim.x) {
        const int g = it % cp8;
        const int oq = it / cp8;
        const size_t i = (size_t(blockIdx.y) * q + oq) * cp8 + g;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = is_max ? -FLT_MAX : 0.f;
        if (kh == 3 && kw == 3) {
            // the common 3x3 window: issue all nine 128-bit loads before the first use (clamped address + validity flag)
            Half8 v[9];
            bool ok[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int iy = op * sh - ph + t / 3, ix = oq * sw - pw + t % 3;
                ok[t] = iy >= 0 && iy < h && ix >= 0 && ix < w;
                v[t] = ld8(x + (((size_t(img) * h + (ok[t] ? iy : 0)) * w + (ok[t] ? ix : 0)) * cp8 + g) * 8);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                if (!ok[t]) continue;
                float f[8];
                unpack(v[t], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = is_max ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
            }
        } else
        for (int r = 0; r < kh; ++r) {
            const int iy = op * sh - ph + r;
            if (iy < 0 || iy >= h) continue;
            for (int s = 0; s < kw; ++s) {
                const int ix = oq * sw - pw + s;
                if (ix < 0 || ix >= w) continue;
                float f[8];
                unpack(ld8(x + (((size_t(img) * h + iy) * w + ix) * cp8 + g) * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = is_max ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
            }
        }
        if (!is_max) {
            const float inv = 1.f / float(kh * kw);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= inv;
        }
        st8(y + i * 8, pack(acc));
    }
}

// 3x3 max pooling (the ResNet stem pool): nine 128-bit loads in flight per thread, packed half2 maxima (max is exact in fp16,
// so no conversion is needed; the generic kernel spends more issue slots on cvt than on the compare).
__global__ void __launch_bounds__(kThreads) pool_max3x3_kernel(const __half* __restrict__ x, __half* __restrict__ y, int h, int w, int cp8, int p, int q,
                                                              int sh, int sw, int ph, int pw) {
    pdl_prologue();
    const int op = blockIdx.y % p;
    const int img = blockIdx.y / p;
    const int row_items = q * cp8;
    const __half2 neg = __float2half2_rn(-65504.f);
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < row_items; it += gridDim.x * blockDim.x) {
        const int g = it % cp8;
        const int oq = it / cp8;
        Half8 v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int iy = op * sh - ph + t / 3, ix = oq * sw - pw + t % 3;
            const bool ok = iy >= 0 && iy < h && ix >= 0 && ix < w;
            if (ok) {
                v[t] = ld8(x + (((size_t(img) * h + iy) * w + ix) * cp8 + g) * 8);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[t].v[j] = neg;
            }
        }
        Half8 m = v[0];
#pragma unroll
        for (int t = 1; t < 9; ++t) {
#pragma unroll
            for (int j = 0; j < 4; ++j) m.v[j] = __hmax2(m.v[j], v[t].v[j]);
        }
        st8(y + ((size_t(blockIdx.y) * q + oq) * cp8 + g) * 8, m);
    }
}

// Two output rows per thread (SH = vertical stride 1 or 2): the SH + 3 input rows both windows cover are loaded once (3 * (SH + 3)
// 128-bit loads in flight instead of 18), reduced along W first, then along H for each of the two outputs.  blockIdx.y = (image, row pair).
template <int SH>
__global__ void __launch_bounds__(kThreads) pool_max3x3_rows2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int h, int w, int cp8, int p,
                                                                    int q, int sw, int ph, int pw) {
    pdl_prologue();
    constexpr int R = SH + 3;
    const int pairs = (p + 1) / 2;
    const int op0 = (blockIdx.y % pairs) * 2;
    const int img = blockIdx.y / pairs;
    const int row_items = q * cp8;
    const __half2 neg = __float2half2_rn(-65504.f);
    const int iy0 = op0 * SH - ph;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < row_items; it += gridDim.x * blockDim.x) {
        const int g = it % cp8;
        const int oq = it / cp8;
        const int ix0 = oq * sw - pw;
        Half8 v[R][3];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int iy = iy0 + r;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int ix = ix0 + c;
                if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
                    v[r][c] = ld8(x + (((size_t(img) * h + iy) * w + ix) * cp8 + g) * 8);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[r][c].v[j] = neg;
                }
            }
        }
        Half8 hm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) hm[r].v[j] = __hmax2(__hmax2(v[r][0].v[j], v[r][1].v[j]), v[r][2].v[j]);
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            if (op0 + o >= p) break;
            Half8 m;
#pragma unroll
            for (int j = 0; j < 4; ++j) m.v[j] = __hmax2(__hmax2(hm[o * SH].v[j], hm[o * SH + 1].v[j]), hm[o * SH + 2].v[j]);
            st8(y + (((size_t(img) * p + op0 + o) * q + oq) * cp8 + g) * 8, m);
        }
    }
}

// ---- global average pool: block = (n, 8-channel group chunk); threads split the pixels, shuffle + smem reduce.
// grid = (cp8 groups, n); 256 threads over pixels.
__global__ void __launch_bounds__(kThreads) global_avgpool_kernel(const __half* __restrict__ x, __half* __restrict__ y, int hw, int cp8) {
    pdl_prologue();
    const int g = blockIdx.x;
    const int img = blockIdx.y;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __half* base = x + size_t(img) * hw * cp8 * 8 + g * 8;
    for (int pix = threadIdx.x; pix < hw; pix += blockDim.x) {
        float f[8];
        unpack(ld8(base + size_t(pix) * cp8 * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    __shared__ float red[kThreads / 32][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x >> 5][j] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float s = 0.f;
        for (int wv = 0; wv < blockDim.x / 32; ++wv) s += red[wv][threadIdx.x];
        y[size_t(img) * cp8 * 8 + g * 8 + threadIdx.x] = __float2half_rn(s / float(hw));
    }
}
// Many-channel / few-pixel variant (e.g. 7x7x2048): one thread per (image, 8 channels), loops over the pixels.
__global__ void __launch_bounds__(kThreads) global_avgpool_small_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int hw,
                                                                       int cp8) {
    const size_t total = size_t(n) * cp8;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int g = int(i % cp8);
        const int img = int(i / cp8);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const __half* base = x + size_t(img) * hw * cp8 * 8 + g * 8;
        for (int pix = 0; pix < hw; ++pix) {
            float f[8];
            unpack(ld8(base + size_t(pix) * cp8 * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
        const float inv = 1.f / float(hw);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv;
        st8(y + i * 8, pack(acc));
    }
}

// Many (image, channel group) pairs, few pixels: one thread per pair, consecutive threads = consecutive 16-byte vectors of a pixel
// (every warp load is 512 contiguous bytes), seven pixels in flight per thread.
__global__ void __launch_bounds__(kThreads) global_avgpool_rows_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int hw, int cp8) {
    pdl_prologue();
    constexpr int U = 7;
    const size_t total = size_t(n) * cp8;
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const int g = int(i % cp8);
    const int img = int(i / cp8);
    const size_t pitch = size_t(cp8) * 8;
    const __half* base = x + size_t(img) * hw * pitch + g * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int pix = 0; pix < hw; pix += U) {
        Half8 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (pix + u < hw) v[u] = ld8(base + size_t(pix + u) * pitch);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (pix + u < hw) {
                float f[8];
                unpack(v[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
    }
    const float inv = 1.f / float(hw);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    st8(y + i * 8, pack(acc));
}

// ---- softmax over the channel axis of each pixel: one warp per row; rows of up to 1024 channels are read ONCE as
//      128-bit vectors (up to four per lane, kept in registers across the max / sum / normalise passes).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// One warp per row, the row read once (up to four 128-bit vectors per lane, all loads issued before the first use).  Lean on issue
// slots, which is what bounds this kernel before HBM does: the row maximum is taken on packed half2 (exact), the exponential is
// one FFMA + one MUFU (ex2 of x * log2e - max * log2e), padded lanes are pushed to -65504 once so they contribute exp() = 0.
__global__ void __launch_bounds__(kThreads, 4) softmax_vec_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t rows, int c, int cp,
                                                              int log_softmax) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5;
    const size_t nwarps = (size_t(gridDim.x) * blockDim.x) >> 5;
    const int nvec = cp / 8;
    const int tail = c & 7;  // valid lanes of the last vector (0 = all eight)
    constexpr float kLog2e = 1.4426950408889634f;
    const __half lowest = __float2half_rn(-65504.f);
    for (size_t row = warp; row < rows; row += nwarps) {
        const __half* xr = x + row * cp;
        __half* yr = y + row * cp;
        Half8 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (lane + 32 * u < nvec) raw[u] = ld8(xr + (lane + 32 * u) * 8);
        __half2 m2 = __half2half2(lowest);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = lane + 32 * u;
            if (v < nvec) {
                if (tail && v == nvec - 1) {  // padded lanes do not take part
                    __half* hv = reinterpret_cast<__half*>(&raw[u]);
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j >= tail) hv[j] = lowest;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) m2 = __hmax2(m2, raw[u].v[j]);
            }
        }
        float mx = fmaxf(__low2float(m2), __high2float(m2));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float mxl = mx * kLog2e;
        float f[4][8];
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (lane + 32 * u < nvec) {
                unpack(raw[u], f[u]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float e = ex2_approx(fmaf(f[u][j], kLog2e, -mxl));
                    sum += e;
                    if (!log_softmax) f[u][j] = e;
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        const float lse = mx + __logf(sum);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = lane + 32 * u;
            if (v < nvec) {
                float o8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o8[j] = log_softmax ? f[u][j] - lse : f[u][j] * inv;
                if (tail && v == nvec - 1) {
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j >= tail) o8[j] = 0.f;
                }
                st8(yr + v * 8, pack(o8));
            }
        }
    }
}

// ---- general fallback (rows longer than 1024 channels): one warp per row, three passes.
__global__ void __launch_bounds__(kThreads) softmax_kernel(const __half* __restrict__ x, __half* __restrict__ y, size_t rows, int c, int cp,
                                                          int log_softmax) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5;
    const size_t nwarps = (size_t(gridDim.x) * blockDim.x) >> 5;
    for (size_t row = warp; row < rows; row += nwarps) {
        const __half* xr = x + row * cp;
        __half* yr = y + row * cp;
        float mx = -FLT_MAX;
        for (int i = lane; i < c; i += 32) mx = fmaxf(mx, __half2float(xr[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int i = lane; i < c; i += 32) sum += __expf(__half2float(xr[i]) - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        const float lse = mx + __logf(sum);
        for (int i = lane; i < cp; i += 32) {
            float v = 0.f;
            if (i < c) {
                const float xv = __half2float(xr[i]);
                v = log_softmax ? xv - lse : __expf(xv - mx) * inv;
            }
            yr[i] = __float2half_rn(v);
        }
    }
}

// ---- instance norm, pass 1: partial (sum, sumsq) per (image, split, channel).  Block = (split, image);
//      thread t owns channel group t % cp8 and pixel lane t / cp8.  Deterministic: fixed reduction order.
__global__ void __launch_bounds__(kThreads) inorm_stats_kernel(const __half* __restrict__ x, float* __restrict__ partials, int hw, int cp8,
                                                              int splits) {
    pdl_prologue();
    extern __shared__ float sm[];  // [lanes][cp8*8][2]
    const int split = blockIdx.x, img = blockIdx.y;
    const int lanes = blockDim.x / cp8;  // host guarantees cp8 <= blockDim.x
    const int g = threadIdx.x % cp8;
    const int pl = threadIdx.x / cp8;
    const int per = (hw + splits - 1) / splits;
    const int p0 = split * per, p1 = min(hw, p0 + per);
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (pl < lanes) {
        const __half* base = x + size_t(img) * hw * cp8 * 8 + g * 8;
        for (int pix = p0 + pl; pix < p1; pix += 4 * lanes) {  // four independent loads in flight per thread
            Half8 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pix + u * lanes < p1) v[u] = ld8(base + size_t(pix + u * lanes) * cp8 * 8);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pix + u * lanes < p1) {
                    float f[8];
                    unpack(v[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] += f[j] * f[j]; }
                }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sm[(pl * cp8 * 8 + g * 8 + j) * 2] = s1[j];
            sm[(pl * cp8 * 8 + g * 8 + j) * 2 + 1] = s2[j];
        }
    }
    __syncthreads();
    const int cp = cp8 * 8;
    for (int ch = threadIdx.x; ch < cp; ch += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int l = 0; l < lanes; ++l) { a += sm[(l * cp + ch) * 2]; b += sm[(l * cp + ch) * 2 + 1]; }
        float* o = partials + ((size_t(img) * splits + split) * cp + ch) * 2;
        o[0] = a; o[1] = b;
    }
}
// pass 1b: per (image, channel) scale = rstd * gamma and shift = beta - mean * scale from the split partials, once (one block per
// image; fixed summation order, fp64).  `parts` lanes of a warp share a channel: each sums every parts-th split, a shuffle tree adds
// them up -- with one thread per channel walking all 256 splits of a 512 x 512 image serially this kernel took 26 us (ncu), four times
// the statistics and apply passes next to it.  (It used to be recomputed by every block of pass 2 before that.)
__global__ void __launch_bounds__(kThreads) inorm_finalize_kernel(const float* __restrict__ partials, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float* __restrict__ params, int hw, int cp, int splits,
                                                                 float eps, int group_size, int channels, int parts) {
    pdl_prologue();
    const int img = blockIdx.x;
    const int part = threadIdx.x % parts;
    const int per_pass = blockDim.x / parts;
    for (int ch0 = 0; ch0 < cp; ch0 += per_pass) {  // every thread of the block takes part in every pass (warp shuffles below)
        const int ch = ch0 + threadIdx.x / parts;
        double a = 0.0, b = 0.0;
        // statistics are shared by the `group_size` channels of ch's group (custom_group_norm, Converters.swift:1273-1300);
        // group_size == 1 is InstanceNormalization
        int c_lo = 0, c_hi = 0;
        if (ch < cp) {
            c_lo = ch < channels ? (ch / group_size) * group_size : ch;
            c_hi = ch < channels ? min(c_lo + group_size, channels) : ch + 1;
            for (int cc = c_lo; cc < c_hi; ++cc)
                for (int sp0 = part; sp0 < splits; sp0 += 8 * parts) {  // eight independent loads in flight, then the adds (fixed order)
                    float2 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int sp = sp0 + u * parts;
                        v[u] = sp < splits ? __ldg(reinterpret_cast<const float2*>(partials + ((size_t(img) * splits + sp) * cp + cc) * 2)) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) { a += v[u].x; b += v[u].y; }
                }
        }
        for (int o = 1; o < parts; o <<= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (ch < cp && part == 0) {
            const double cnt = double(hw) * (c_hi - c_lo);
            const double mean = a / cnt;
            double var = b / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = float(1.0 / sqrt(var + double(eps)));
            const float sc = rstd * gamma[ch];
            params[(size_t(img) * cp + ch) * 2] = sc;
            params[(size_t(img) * cp + ch) * 2 + 1] = beta[ch] - float(mean) * sc;
        }
    }
}
// ---- where a normalised pixel goes (kernels.h NormStore) ----
struct NormDst {
    int unfold_w, unfold_f, h, w, pt, pl, ho, wo, s2d, edge;  // ho == 0: no padding; unfold_f: 2 or 4; edge: clamp instead of reflect
};
__device__ __forceinline__ int reflect_at(int i, int n) {  // ONNX 'reflect' (no edge repeat), pads < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}
// pixel p = (y * w + x) * F^2 + ey * F + ex of a phase-column convolution's output (upsample-folded: F = 2, input-folded: F = 4) ->
// row-major pixel (F y + ey, F x + ex) of the F h x F w image
__device__ __forceinline__ unsigned unfold_pixel(unsigned p, unsigned w, unsigned f) {
    const unsigned sh = f == 4u ? 2u : 1u, ph = p & (f * f - 1u), t = p >> (2u * sh), yy = t / w, xx = t - yy * w;
    return (f * yy + (ph >> sh)) * (f * w) + f * xx + (ph & (f - 1u));
}
// 16-byte vector index (vector 0 of the pixel) of padded pixel (oy, ox)
__device__ __forceinline__ size_t padded_vec(const NormDst& d, unsigned oy, unsigned ox, unsigned cp8) {
    if (!d.s2d) return (size_t(oy) * unsigned(d.wo) + ox) * cp8;
    const unsigned sh = d.s2d == 4 ? 2u : 1u, sm = unsigned(d.s2d) - 1u;
    return ((size_t(oy >> sh) * (unsigned(d.wo) >> sh) + (ox >> sh)) * unsigned(d.s2d * d.s2d) + ((oy & sm) * unsigned(d.s2d) + (ox & sm))) * cp8;
}
// ... of source pixel p (storage order of x); q = its row-major index in the un-padded image
__device__ __forceinline__ size_t norm_dst_vec(const NormDst& d, unsigned p, unsigned cp8, unsigned& q) {
    q = d.unfold_w ? unfold_pixel(p, unsigned(d.unfold_w), unsigned(d.unfold_f)) : p;
    if (!d.ho) return size_t(q) * cp8;
    const unsigned yy = q / unsigned(d.w), xx = q - yy * unsigned(d.w);
    return padded_vec(d, yy + unsigned(d.pt), xx + unsigned(d.pl), cp8);
}
// border pixel j (0 <= j < ho * wo - h * w: top rows, bottom rows, then the left / right columns of the rows between) -> its
// position and the storage index of the pixel it mirrors
__device__ __forceinline__ void ring_pixel(const NormDst& d, unsigned j, unsigned& oy, unsigned& ox, unsigned& src) {
    const unsigned wo = unsigned(d.wo), top = unsigned(d.pt) * wo, bottom = unsigned(d.ho - d.h - d.pt) * wo;
    if (j < top) {
        oy = j / wo; ox = j - oy * wo;
    } else if (j < top + bottom) {
        const unsigned t = j - top, r = t / wo;
        oy = unsigned(d.pt + d.h) + r; ox = t - r * wo;
    } else {
        const unsigned t = j - top - bottom, side = unsigned(d.wo - d.w), r = t / side, kk = t - r * side;
        oy = unsigned(d.pt) + r; ox = kk < unsigned(d.pl) ? kk : unsigned(d.w) + kk;
    }
    const unsigned sy = unsigned(d.edge ? min(max(int(oy) - d.pt, 0), d.h - 1) : reflect_at(int(oy) - d.pt, d.h));
    const unsigned sx = unsigned(d.edge ? min(max(int(ox) - d.pl, 0), d.w - 1) : reflect_at(int(ox) - d.pl, d.w));
    if (d.unfold_w) {
        const unsigned f = unsigned(d.unfold_f), sh = f == 4u ? 2u : 1u;
        src = ((sy >> sh) * unsigned(d.unfold_w) + (sx >> sh)) * (f * f) + (sy & (f - 1u)) * f + (sx & (f - 1u));
    } else {
        src = sy * unsigned(d.w) + sx;
    }
}
// pass 2: y = act(x * scale + shift).  Blocks past `chunks` write the reflection border (NormStore::pad_*): one task per (border
// pixel, 8 channels), source pixel re-read.
// STATS: the statistics come from the producing convolution's epilogue (ConvKernelParams::stats: fp64 {sum, sum of squares} per
// (image, phase, channel), `phases` = 1, or F^2 column groups behind a phase-column convolution): every block derives scale / shift
// of its image in shared memory, the last block to finish puts the accumulators and the arrival counter back to zero.
struct NormStats {
    double* stats;
    const float* gamma;
    const float* beta;
    unsigned int* counter;
    float eps;
    int phases;
    double inv_hw;
};
// (STATS without EARLY: two blocks per SM -- at three the compiler spills part of v[] around the prologue, which makes the prologue wait
// for the loads)
// RES: a residual tensor (plain row-major pixels, the shape of x) is added behind the norm's own activation, `act2` follows the sum --
// the Add (-> ReLU) -> Pad tail of a residual block (engine.cc "norm tail"); four vectors of each operand in flight per thread.
// EARLY (STATS only): the prologue runs before the loads are issued, nothing is live across it and three blocks fit an SM -- the form
// for tensors large enough to be bandwidth-bound; small ones keep the loads in flight behind the prologue (latency-bound).
// UU: vectors per thread (0 = 8, or 4 with a residual).  Small tensors take 4 so that the grid is about two blocks per SM: one image
// of TransformerNet's 128 x 128 x 128 layers is 128 blocks at 8 vectors per thread -- less than one per SM (instance_norm_from_stats()).
template <bool STATS, bool RES, bool EARLY = false, int UU = 0>
__global__ void __launch_bounds__(kThreads, (STATS && !EARLY) ? 2 : 3) inorm_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ params,
                                                              int hw, int cp8, int act, NormDst dst, int chunks, __half* __restrict__ y_plain, NormStats ns,
                                                              int cpb, const __half* __restrict__ res, int act2) {
    pdl_prologue();
    const int img = blockIdx.y;
    extern __shared__ float sm_params[];  // STATS: [cp][2]
    const float* sm = STATS ? sm_params : params + size_t(img) * cp8 * 16;  // [cp][2] scale, shift of this image
    // streaming part: block-tiled, eight 128-bit loads in flight per thread; when kThreads % cp8 == 0 a thread meets the same 8
    // channels in every vector it handles and keeps their scale / shift in registers
    constexpr int U = UU ? UU : (RES ? 4 : 8);
    const size_t n8 = size_t(hw) * cp8;
    const bool plain = !dst.unfold_w && !dst.ho;
    const __half* rb = RES ? res + size_t(img) * n8 * 8 : nullptr;
    // a streaming block takes `cpb` consecutive chunks (STATS: the scale / shift prologue is paid once per block)
    const int sblocks = (chunks + cpb - 1) / cpb;
    const bool ring = int(blockIdx.x) >= sblocks;
    const size_t tasks = ring ? size_t(dst.ho * dst.wo - dst.h * dst.w) * cp8 : n8;
    const __half* xb = x + size_t(img) * n8 * 8;
    __half* yb = y + size_t(img) * (dst.ho ? size_t(dst.ho) * dst.wo * cp8 : n8) * 8;
    const int c_lo = ring ? int(blockIdx.x) - sblocks : int(blockIdx.x) * cpb;
    const int c_hi = ring ? c_lo + 1 : min(chunks, c_lo + cpb);
    const bool fixed = (kThreads % cp8) == 0;
    float sc[8], sh[8];
    auto norm_prologue = [&]() {
        // `parts` threads share a channel, each adding every parts-th column group; fixed order within the block
        __shared__ double2 red[kThreads];
        const int cp = cp8 * 8;
        const int parts = (cp <= kThreads && kThreads % cp == 0) ? min(kThreads / cp, ns.phases) : 1;
        for (int ch0 = 0; ch0 < cp; ch0 += kThreads / parts) {
            const int ch = ch0 + int(threadIdx.x) % (kThreads / parts), part = int(threadIdx.x) / (kThreads / parts);
            double a = 0.0, b = 0.0;
            if (ch < cp && part < parts) {
#pragma unroll 4
                for (int ph = part; ph < ns.phases; ph += parts) {
                    const double2 t = __ldcg(reinterpret_cast<const double2*>(ns.stats + ((size_t(img) * ns.phases + ph) * cp + ch) * 2));
                    a += t.x; b += t.y;
                }
            }
            if (parts > 1) {
                red[threadIdx.x] = make_double2(a, b);
                __syncthreads();
                if (part == 0) for (int q = 1; q < parts; ++q) { a += red[threadIdx.x + q * (kThreads / parts)].x; b += red[threadIdx.x + q * (kThreads / parts)].y; }
            }
            if (ch < cp && part == 0) {
                const double inv = ns.inv_hw, mean = a * inv;
                const float var = fmaxf(float(fma(b, inv, -mean * mean)), 0.f);  // the cancellation is done in fp64, the root in fp32
                const float rstd = rsqrtf(var + ns.eps);
                const float scl = rstd * ns.gamma[ch];
                sm_params[ch * 2] = scl;
                sm_params[ch * 2 + 1] = ns.beta[ch] - float(mean) * scl;
            }
        }
        __syncthreads();
    };
    if (STATS && EARLY) norm_prologue();
  for (int chunk = c_lo; chunk < c_hi; ++chunk) {
    const size_t base = size_t(chunk) * (kThreads * U) + threadIdx.x;
    // (32-bit store offsets -- vector indices inside one image -- and three blocks per SM: with 64-bit offsets the kernel took 146
    // registers, one block per SM, and a block's 32 KiB in flight did not cover its prologue)
    Half8 v[U], r[RES ? U : 1];
    unsigned o[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t i = base + u * kThreads;
        if (i >= tasks) break;
        if (!ring) {
            v[u] = ld8(xb + i * 8);
            if (RES) r[u] = ld8(rb + i * 8);
            if (plain) {
                o[u] = unsigned(i);
            } else {
                const unsigned pix = unsigned(i / size_t(cp8));
                unsigned q;
                o[u] = unsigned(norm_dst_vec(dst, pix, unsigned(cp8), q) + (i - size_t(pix) * cp8));
            }
        } else {
            const unsigned j = unsigned(i / size_t(cp8)), g = unsigned(i - size_t(j) * cp8);
            unsigned oy, ox, src;
            ring_pixel(dst, j, oy, ox, src);
            v[u] = ld8(xb + (size_t(src) * cp8 + g) * 8);
            if (RES) r[u] = ld8(rb + (size_t(src) * cp8 + g) * 8);
            o[u] = unsigned(padded_vec(dst, oy, ox, unsigned(cp8)) + g);
        }
    }
    if (STATS && !EARLY && chunk == c_lo) norm_prologue();  // (the loads above are in flight)
    if (fixed && chunk == c_lo) {
        const int c0 = int(threadIdx.x % unsigned(cp8)) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = sm[(c0 + j) * 2]; sh[j] = sm[(c0 + j) * 2 + 1]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t i = base + u * kThreads;
        if (i >= tasks) break;
        if (!fixed) {
            const int c0 = int(i % size_t(cp8)) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { sc[j] = sm[(c0 + j) * 2]; sh[j] = sm[(c0 + j) * 2 + 1]; }
        }
        float f[8];
        unpack(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(f[j], sc[j], sh[j]);
            f[j] = act == ACT_RELU ? fmaxf(t, 0.f) : t;
        }
        if (RES) {
            float rf[8];
            unpack(r[u], rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = f[j] + rf[j];
                f[j] = act2 == ACT_RELU ? fmaxf(t, 0.f) : t;
            }
        }
        st8(yb + size_t(o[u]) * 8, pack(f));
        if (y_plain && !ring) {  // the un-padded copy for the other readers (position worked out here: rare, and eight registers less)
            const unsigned pix = unsigned(i / size_t(cp8));
            const unsigned q = dst.unfold_w ? unfold_pixel(pix, unsigned(dst.unfold_w), unsigned(dst.unfold_f)) : pix;
            st8(y_plain + (size_t(img) * n8 + size_t(q) * cp8 + (i - size_t(pix) * cp8)) * 8, pack(f));
        }
    }
  }
    if (STATS && ns.counter) {
        __shared__ unsigned int s_last;
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(ns.counter, 1u) == gridDim.x * gridDim.y - 1u;
        }
        __syncthreads();
        if (s_last) {  // every block has read its statistics: zero for the next encode's convolution
            const size_t total = size_t(gridDim.y) * ns.phases * cp8 * 16;
            for (size_t i = threadIdx.x; i < total; i += kThreads) ns.stats[i] = 0.0;
            if (threadIdx.x == 0) *ns.counter = 0u;
        }
    }
}

// ---- instance norm in ONE launch (thread-block clusters + distributed shared memory).  A cluster of `csz` CTAs owns one
//      (image, 16-channel slab): every CTA reduces sum / sum of squares over its share of the pixels, publishes them in shared memory,
//      reads its peers' through the cluster window after one cluster barrier (fixed order: deterministic), and normalises its pixels,
//      which it finds in L2 again.  Against the three-launch form (partials -> finalize -> apply) that is two kernel boundaries and the
//      partial / parameter round trips less: the three passes of a 128 x 128 x 128 image took ~16 us, most of it hand-over latency.
//      Thread t: vector (t & 1) of the slab (8 channels), pixel lane t >> 1; 32 contiguous bytes per pixel.
__global__ void __launch_bounds__(kThreads) inorm_cluster_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int hw, int cp8, float eps, int act, int group_size,
                                                                int channels, NormDst dst, __half* __restrict__ y_plain) {
    pdl_prologue();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int csz = int(cluster.num_blocks());
    const int rank = int(cluster.block_rank());
    const int slab = int(blockIdx.x) / csz, img = blockIdx.y;
    const int g0 = slab * 2;
    const int ng = min(2, cp8 - g0);
    const int v = threadIdx.x & 1, pl = threadIdx.x >> 1;
    constexpr int kLanes = kThreads / 2;
    constexpr int kU = 8;
    const int per = (hw + csz - 1) / csz;
    const int p0 = rank * per, p1 = min(hw, p0 + per);
    __shared__ float warp_part[kThreads / 32][2][8][2];
    __shared__ double cta_part[2][8][2];
    __shared__ float scsh[2][8][2];
    const bool live = v < ng;
    const __half* xb = x + (size_t(img) * hw * cp8 + g0 + v) * 8;
    __half* yb = y + (size_t(img) * (dst.ho ? size_t(dst.ho) * dst.wo : size_t(hw)) * cp8 + g0 + v) * 8;
    const bool plain = !dst.unfold_w && !dst.ho;
    float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        for (int pix = p0 + pl; pix < p1; pix += kU * kLanes) {  // kU independent loads in flight per thread
            Half8 t[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (pix + u * kLanes < p1) t[u] = ld8(xb + size_t(pix + u * kLanes) * cp8 * 8);
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (pix + u * kLanes < p1) {
                    float f[8];
                    unpack(t[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] += f[j] * f[j]; }
                }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // lanes of equal parity hold the same channels
#pragma unroll
        for (int o = 2; o < 32; o <<= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
    }
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { warp_part[wid][lane][j][0] = s1[j]; warp_part[wid][lane][j][1] = s2[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int tv = threadIdx.x >> 4, tj = (threadIdx.x >> 1) & 7, tk = threadIdx.x & 1;
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) a += double(warp_part[w][tv][tj][tk]);
        cta_part[tv][tj][tk] = a;
    }
    cluster.sync();  // every CTA's partial sums are in its shared memory
    if (threadIdx.x < 32) {
        const int tv = threadIdx.x >> 4, tj = (threadIdx.x >> 1) & 7, tk = threadIdx.x & 1;
        // all peers' values are requested before any is added: a dependent chain of distributed-shared-memory round trips (one per
        // peer) was most of this kernel's time
        double pv[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) pv[r] = r < csz ? cluster.map_shared_rank(&cta_part[0][0][0], r)[(tv * 8 + tj) * 2 + tk] : 0.0;
        double a = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) a += pv[r];
        // custom_group_norm: statistics pooled over the group_size channels of a group (Converters.swift:1273-1300).  group_size is a
        // power of two <= 16 here (host check), groups are aligned in the slab, consecutive channels sit two lanes apart.
        double pooled = a;
        for (int o = 2; o < 2 * group_size; o <<= 1) pooled += __shfl_xor_sync(0xffffffffu, pooled, o);
        const double a_other = __shfl_xor_sync(0xffffffffu, a, 1);       // tk == 0 lanes: the sum of squares next to their sum
        const double p_other = __shfl_xor_sync(0xffffffffu, pooled, 1);
        const int ch = (g0 + tv) * 8 + tj;
        if (tk == 0 && tv < ng) {
            const bool real = ch < channels;   // padding lanes of the channel pitch keep statistics of their own (all zero)
            const int c_lo = real ? (ch / group_size) * group_size : ch;
            const int c_hi = real ? min(c_lo + group_size, channels) : ch + 1;
            const double cnt = double(hw) * (c_hi - c_lo);
            const double mean = (real ? pooled : a) / cnt;
            double var = (real ? p_other : a_other) / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = float(1.0 / sqrt(var + double(eps)));
            const float sc = rstd * gamma[ch];
            scsh[tv][tj][0] = sc;
            scsh[tv][tj][1] = beta[ch] - float(mean) * sc;
        }
    }
    cluster.sync();  // peers have read this CTA's partials (nobody exits early); scsh is visible to the CTA
    if (live) {
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scsh[v][j][0]; sh[j] = scsh[v][j][1]; }
        // the reflection border (NormStore::pad_*): its pixels dealt round-robin over the cluster's threads, at most a few each
        const unsigned nring = dst.ho ? unsigned(dst.ho * dst.wo - dst.h * dst.w) : 0u;
        const unsigned ring_j = unsigned(rank * kLanes + pl);
        Half8 ring_v;
        size_t ring_o = 0;
        if (ring_j < nring) {
            unsigned oy, ox, src;
            ring_pixel(dst, ring_j, oy, ox, src);
            ring_v = ld8(xb + size_t(src) * cp8 * 8);
            ring_o = padded_vec(dst, oy, ox, unsigned(cp8));
        }
        for (int pix = p0 + pl; pix < p1; pix += kU * kLanes) {
            Half8 t[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (pix + u * kLanes < p1) t[u] = ld8(xb + size_t(pix + u * kLanes) * cp8 * 8);
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (pix + u * kLanes < p1) {
                    float f[8];
                    unpack(t[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float r = fmaf(f[j], sc[j], sh[j]);
                        f[j] = act == ACT_RELU ? fmaxf(r, 0.f) : r;
                    }
                    unsigned q = unsigned(pix + u * kLanes);
                    const size_t o = plain ? size_t(q) * cp8 : norm_dst_vec(dst, q, unsigned(cp8), q);
                    st8(yb + o * 8, pack(f));
                    if (y_plain) st8(y_plain + ((size_t(img) * hw + q) * cp8 + g0 + v) * 8, pack(f));
                }
        }
        for (unsigned j = ring_j; j < nring; j += unsigned(csz * kLanes)) {  // the first one was requested before the apply loop
            if (j != ring_j) {
                unsigned oy, ox, src;
                ring_pixel(dst, j, oy, ox, src);
                ring_v = ld8(xb + size_t(src) * cp8 * 8);
                ring_o = padded_vec(dst, oy, ox, unsigned(cp8));
            }
            float f[8];
            unpack(ring_v, f);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float r = fmaf(f[jj], sc[jj], sh[jj]);
                f[jj] = act == ACT_RELU ? fmaxf(r, 0.f) : r;
            }
            st8(yb + ring_o * 8, pack(f));
        }
    }
}

// ---- depthwise conv: one thread per (output pixel, 8 channels); weights [kh*kw][cp].
__global__ void __launch_bounds__(kThreads) depthwise_kernel(const __half* __restrict__ x, const __half* __restrict__ wt,
                                                            const float* __restrict__ bias, __half* __restrict__ y, int n, int h, int w,
                                                            int cp8, int p, int q, int kh, int kw, int sh, int sw, int dh, int dw, int pt,
                                                            int pl, int act, float lo, float hi) {
    // grid.y walks output rows (image, op); threads walk (oq, channel group) of that row
    pdl_prologue();
    const int op = blockIdx.y % p;
    const int img = blockIdx.y / p;
    const int row_items = q * cp8;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < row_items; it += gridDim.x * blockDim.x) {
        const int g = it % cp8;
        const int oq = it / cp8;
        const size_t i = (size_t(blockIdx.y) * q + oq) * cp8 + g;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + g * 8 + j);
        if (kh == 3 && kw == 3) {
            // the 3 x 3 case (every MobileNetV2 layer): all nine input vectors and nine weight vectors are requested before the first
            // multiply -- with the loads inside the tap loop they were nine dependent L2 round trips, ~5 us for a kernel that moves a
            // few hundred KB (batch 1: as long as the tensor-core convolutions around it)
            Half8 xv[9], wv9[9];
            bool ok[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int iy = op * sh - pt + (t / 3) * dh, ix = oq * sw - pl + (t % 3) * dw;
                ok[t] = iy >= 0 && iy < h && ix >= 0 && ix < w;
                if (ok[t]) xv[t] = ld8(x + (((size_t(img) * h + iy) * w + ix) * cp8 + g) * 8);
                wv9[t] = ld8(wt + (size_t(t) * cp8 + g) * 8);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (ok[t]) {
                    float f[8], wf[8];
                    unpack(xv[t], f);
                    unpack(wv9[t], wf);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wf[j], acc[j]);
                }
        } else
        for (int r = 0; r < kh; ++r) {
            const int iy = op * sh - pt + r * dh;
            if (iy < 0 || iy >= h) continue;
            for (int s = 0; s < kw; ++s) {
                const int ix = oq * sw - pl + s * dw;
                if (ix < 0 || ix >= w) continue;
                float f[8], wv[8];
                unpack(ld8(x + (((size_t(img) * h + iy) * w + ix) * cp8 + g) * 8), f);
                unpack(ld8(wt + (size_t(r * kw + s) * cp8 + g) * 8), wv);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], wv[j], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = act_op(acc[j], act, lo, hi);
        st8(y + i * 8, pack(acc));
    }
}

}  // namespace

cudaError_t pool2d(const __half* x, __half* y, int n, int h, int w, int cp, int p, int q, int kh, int kw, int sh, int sw, int ph, int pw,
                   int is_max, cudaStream_t s) {
    // grid.y = (image, output row) is limited to 65535: larger batches go in slices of whole images
    if (size_t(n) * p > 65535) {
        const int per = std::max(1, 65535 / p);
        if (p > 65535) return cudaErrorInvalidValue;
        for (int i0 = 0; i0 < n; i0 += per) {
            const int gn = std::min(per, n - i0);
            cudaError_t e = pool2d(x + size_t(i0) * h * w * cp, y + size_t(i0) * p * q * cp, gn, h, w, cp, p, q, kh, kw, sh, sw, ph, pw, is_max, s);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const int row_items = q * (cp / 8);
    dim3 grid(unsigned((row_items + kThreads - 1) / kThreads), unsigned(n * p));
    if (kh == 3 && kw == 3 && is_max && (sh == 1 || sh == 2) && p >= 2) {
        dim3 grid2(grid.x, unsigned(n * ((p + 1) / 2)));
        if (sh == 2) return launch_pdl(pool_max3x3_rows2_kernel<2>, grid2, dim3(kThreads), s, x, y, h, w, cp / 8, p, q, sw, ph, pw);
        return launch_pdl(pool_max3x3_rows2_kernel<1>, grid2, dim3(kThreads), s, x, y, h, w, cp / 8, p, q, sw, ph, pw);
    }
    if (kh == 3 && kw == 3 && is_max) return launch_pdl(pool_max3x3_kernel, grid, dim3(kThreads), s, x, y, h, w, cp / 8, p, q, sh, sw, ph, pw);
    else (void)launch_pdl(pool2d_kernel, dim3(grid), dim3(kThreads), s, x, y, n, h, w, cp / 8, p, q, kh, kw, sh, sw, ph, pw, is_max);
    return cudaGetLastError();
}

// Few-pixel variant with more parallelism (e.g. 7x7x2048 x 32 images): 8 lanes share one (image, 8-channel group), each
// summing every 8th pixel, then a 3-step shuffle reduction.  Consecutive 8-lane teams read consecutive 16-byte vectors.
__global__ void __launch_bounds__(kThreads) global_avgpool_team_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int hw,
                                                                      int cp8) {
    pdl_prologue();
    const size_t total = size_t(n) * cp8;
    const int sub = threadIdx.x & 7;
    for (size_t i = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 3; i < ((total + 31) & ~size_t(31)); i += (size_t(gridDim.x) * blockDim.x) >> 3) {
        const bool live = i < total;
        const int g = live ? int(i % cp8) : 0;
        const int img = live ? int(i / cp8) : 0;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const __half* base = x + size_t(img) * hw * cp8 * 8 + g * 8;
        if (live)
            for (int pix = sub; pix < hw; pix += 32) {  // four independent loads in flight per lane
                Half8 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (pix + 8 * u < hw) v[u] = ld8(base + size_t(pix + 8 * u) * cp8 * 8);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (pix + 8 * u < hw) {
                        float f[8];
                        unpack(v[u], f);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] += f[j];
                    }
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
        }
        if (live && sub == 0) {
            const float inv = 1.f / float(hw);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= inv;
            st8(y + i * 8, pack(acc));
        }
    }
}

cudaError_t global_avgpool(const __half* x, __half* y, int n, int hw, int cp, cudaStream_t s) {
    if (hw <= 256 && size_t(n) * (cp / 8) >= size_t(kSMs) * kThreads * 2) {
        const size_t total = size_t(n) * (cp / 8);
        return launch_pdl(global_avgpool_rows_kernel, dim3(unsigned((total + kThreads - 1) / kThreads)), dim3(kThreads), s, x, y, n, hw, cp / 8);
    }
    if (hw <= 64 || size_t(n) * (cp / 8) >= size_t(kSMs) * 64) {
        const size_t teams = size_t(n) * (cp / 8);
        return launch_pdl(global_avgpool_team_kernel, dim3(unsigned(grid_for(teams * 8))), dim3(kThreads), s, x, y, n, hw, cp / 8);
    } else {
        dim3 grid(cp / 8, n);
        (void)launch_pdl(global_avgpool_kernel, dim3(grid), dim3(kThreads), s, x, y, hw, cp / 8);
    }
    return cudaGetLastError();
}

cudaError_t softmax_rows(const __half* x, __half* y, size_t rows, int c, int cp, int log_softmax, cudaStream_t s) {
    if (cp <= 1024) (void)launch_pdl(softmax_vec_kernel, dim3(grid_for(rows * 32)), dim3(kThreads), s, x, y, rows, c, cp, log_softmax);
    else
    (void)launch_pdl(softmax_kernel, dim3(grid_for(rows * 32)), dim3(kThreads), s, x, y, rows, c, cp, log_softmax);
    return cudaGetLastError();
}

int instance_norm_splits(int hw, int cp) {
    (void)cp;
    int splits = (hw + 1023) / 1024;
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
    return splits;
}

namespace {
int inorm_group(int n, int hw, int cp) {
    // images per (stats, apply) launch pair.  Groups of <= 48 MB (second pass served by L2) were measured: 16 launches of two images are
    // each latency-bound and the whole operator got 2x slower, so all images go in one pair and large batches re-read x from HBM.
    (void)hw; (void)cp;
    return n;
}
}  // namespace
// One-launch cluster form (inorm_cluster_kernel): group statistics must pool inside a 16-channel slab, and a CTA's share of the
// pixels (32 bytes each) should stay small enough to come back from L2 on the second pass.  Returns the cluster size, 0 = three-launch form.
// Largest cluster the device schedules for inorm_cluster_kernel: 16 CTAs (non-portable size, one cluster per GPC) where the driver
// allows it, else the portable 8.  Asked once.
int inorm_max_cluster();
int inorm_cluster_size(int hw, int group_size, size_t tensor_bytes) {
    // both passes of a cluster stream 32-byte slabs of 128-byte lines: fine out of L2, 2x slower than the three launches from HBM
    // (64 x 128 x 128 x 128, 537 MB: 0.279 ms against 0.143 ms)
    if (tensor_bytes > (size_t(48) << 20)) return 0;
    const bool off = getenv("SMELTER_NO_CLUSTER_NORM") != nullptr;  // read on every call: tests switch it inside one process
    if (off || group_size < 1 || group_size > 16 || (group_size & (group_size - 1))) return 0;
    int csz = inorm_max_cluster();
    while (csz > 1 && hw / csz < 512) csz >>= 1;
    // measured (TransformerNet, one image): 32-64 KiB per CTA (128 x 128 pixels) 16 -> 10 us against the three-launch form, but 128 KiB
    // per CTA (256^2 x 64 in 16-CTA clusters) ~20 against ~17 us (the whole encode 0.404 -> 0.397 ms with those two norms in three
    // launches) and 256 KiB / 1 MiB per CTA 17 -> 27 us / 34 -> 77 us -- too few CTAs stream the image; those keep the three launches.
    // 1024-thread CTAs do not change that (256^2 x 64: 24 us either way, 512^2 x 32: 52 us against 30 us for three launches): a slab
    // is 32 bytes of every 128-byte line, so the passes are bound by the lines a CTA touches, not by the bytes it keeps in flight.
    if (size_t(hw) / csz * 32 > (size_t(64) << 10)) return 0;
    return csz;
}
int instance_norm_launches(int n, int hw, int cp, int group_size) {
    if (inorm_cluster_size(hw, group_size, size_t(n) * hw * cp * 2)) return 1;
    const int group = inorm_group(n, hw, cp);
    return 3 * ((n + group - 1) / group);
}
size_t instance_norm_scratch_floats(int n, int hw, int cp) {  // split partials + per-(image, channel) scale / shift
    return size_t(n) * instance_norm_splits(hw, cp) * cp * 2 + size_t(n) * cp * 2;
}

int inorm_max_cluster() {
    static const int best = [] {
        if (getenv("SMELTER_CLUSTER_NORM_8")) return 8;
        if (cudaFuncSetAttribute(inorm_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 8; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(16);
        cfg.blockDim = dim3(kThreads);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 16;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, inorm_cluster_kernel, &cfg) != cudaSuccess || clusters < 1) { cudaGetLastError(); return 8; }
        return 16;
    }();
    return best;
}

namespace {
cudaError_t norm_dst_of(const NormStore* store, const __half* x, const __half* y, int hw, NormDst* out) {
    NormDst dst{};
    if (store) {
        if (x == y && (store->unfold_w || store->padded())) return cudaErrorInvalidValue;
        if (store->unfold_f != 2 && store->unfold_f != 4) return cudaErrorInvalidValue;
        if (store->unfold_w < 0 || (store->unfold_w && hw % (store->unfold_f * store->unfold_f * store->unfold_w))) return cudaErrorInvalidValue;
        dst.unfold_w = store->unfold_w;
        dst.unfold_f = store->unfold_f;
        if (store->padded()) {
            if (store->h <= 0 || store->w <= 0 || size_t(store->h) * store->w != size_t(hw)) return cudaErrorInvalidValue;
            if (store->unfold_w && store->unfold_f * store->unfold_w != store->w) return cudaErrorInvalidValue;
            if (store->pad_mode == PAD_REFLECT && (store->pad_t >= store->h || store->pad_b >= store->h || store->pad_l >= store->w || store->pad_r >= store->w))
                return cudaErrorInvalidValue;
            dst.h = store->h; dst.w = store->w; dst.pt = store->pad_t; dst.pl = store->pad_l;
            dst.ho = store->h + store->pad_t + store->pad_b; dst.wo = store->w + store->pad_l + store->pad_r;
            dst.s2d = store->s2d;
            dst.edge = store->pad_mode == PAD_EDGE;
            if (store->pad_mode != PAD_EDGE && store->pad_mode != PAD_REFLECT) return cudaErrorInvalidValue;
            if (dst.s2d && ((dst.s2d != 2 && dst.s2d != 4) || dst.ho % dst.s2d || dst.wo % dst.s2d)) return cudaErrorInvalidValue;
        }
    }
    *out = dst;
    return cudaSuccess;
}
}  // namespace

cudaError_t instance_norm_from_stats(const __half* x, __half* y, int n, int hw, int cp, const float* gamma, const float* beta, float eps, int act,
                                     double* stats, unsigned int* counter, int phases, cudaStream_t s, const NormStore* store, const __half* res, int act2) {
    NormDst dst{};
    if (cudaError_t e = norm_dst_of(store, x, y, hw, &dst); e != cudaSuccess) return e;
    __half* y_plain = store && store->padded() ? store->plain : nullptr;
    const int cp8 = cp / 8;
    if (cp8 > kThreads || phases < 1 || n > 65535) return cudaErrorInvalidValue;
    const size_t n8 = size_t(hw) * cp8;
    if ((dst.ho ? size_t(dst.ho) * dst.wo * cp8 : n8) >> 32) return cudaErrorInvalidValue;  // 32-bit vector offsets inside an image
    if (res && (dst.unfold_w || (act2 != ACT_NONE && act2 != ACT_RELU))) return cudaErrorInvalidValue;  // the residual is read in x's pixel order
    // vectors per thread: 8 (4 with a residual), or 4 where 8 would give fewer than two blocks per SM.  Measured on TransformerNet's
    // 128 x 128 x 128 norms (one image, event-bracketed): 8 -> 13.6 us (128 blocks), 4 -> 11.5 us (256), 2 -> 13.9 us (512: every block
    // pays the prologue); the residual form 4 -> 11.7, 2 -> 15.8 us.  The encode 0.297 -> 0.289 ms.
    const int vpt = (res || ((n8 + size_t(kThreads) * 8 - 1) / (size_t(kThreads) * 8)) * size_t(n) < size_t(2) * 148) ? 4 : 8;
    const size_t per_block = size_t(kThreads) * vpt;
    const int chunks = int(std::max<size_t>(1, (n8 + per_block - 1) / per_block));
    const size_t ring8 = dst.ho ? size_t(dst.ho * dst.wo - dst.h * dst.w) * cp8 : 0;
    const int ring_chunks = int((ring8 + per_block - 1) / per_block);
    NormStats ns{stats, gamma, beta, counter, eps, phases, 1.0 / double(hw)};
    // chunks per block: measured with up to 8 (the prologue paid once per 256 KiB): slower (64 x 128 x 128 x 128: 0.122 -> 0.134 ms), one it
    // is.  The one-trip chunk loop stays in the kernel all the same: the build without it (same-box A/B of the two libraries, twice)
    // runs TransformerNet at 0.316 instead of 0.298 ms and batch 8 at 1.47 instead of 1.42 ms -- the compiler schedules the
    // straight-line form differently (121 / 78 instead of 128 / 80 registers), and the measured form wins.
    const int cpb = 1;
    // several rounds of blocks: bandwidth-bound.  (Threshold swept on TransformerNet: from 3 blocks per SM on, batch 1 0.289 -> 0.294 ms;
    // never, batch 8 1.42 -> 1.46 ms.)
    const bool early = !res && vpt == 8 && size_t(chunks) * n >= size_t(148) * 3 * 4;
    const dim3 grid(chunks + ring_chunks, n);
    const size_t smem = size_t(cp) * 2 * sizeof(float);
    const float* none = nullptr;
#define SMELTER_NORM_LAUNCH(...) (void)launch_pdl_smem(inorm_apply_kernel<__VA_ARGS__>, grid, dim3(kThreads), smem, s, x, y, none, hw, cp8, act, dst, chunks, y_plain, ns, cpb, res, act2)
    if (early) SMELTER_NORM_LAUNCH(true, false, true);
    else if (res) SMELTER_NORM_LAUNCH(true, true);
    else if (vpt == 8) SMELTER_NORM_LAUNCH(true, false);
    else SMELTER_NORM_LAUNCH(true, false, false, 4);
#undef SMELTER_NORM_LAUNCH
    return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(kThreads) inorm_partials_to_stats_kernel(const float* __restrict__ partials, double* __restrict__ stats, int cp, int splits) {
    const int img = blockIdx.x;
    for (int ch = threadIdx.x; ch < cp; ch += kThreads) {
        double a = 0.0, b = 0.0;
        for (int sp = 0; sp < splits; ++sp) {
            a += partials[((size_t(img) * splits + sp) * cp + ch) * 2];
            b += partials[((size_t(img) * splits + sp) * cp + ch) * 2 + 1];
        }
        stats[(size_t(img) * cp + ch) * 2] = a;
        stats[(size_t(img) * cp + ch) * 2 + 1] = b;
    }
}
}  // namespace
cudaError_t instance_norm_stats_f64(const __half* x, int n, int hw, int cp, float* partials, double* stats, cudaStream_t s) {
    const int cp8 = cp / 8;
    if (cp8 > kThreads || n > 65535) return cudaErrorInvalidValue;
    const int splits = instance_norm_splits(hw, cp);
    const int lanes = kThreads / cp8;
    (void)launch_pdl_smem(inorm_stats_kernel, dim3(splits, n), dim3(kThreads), size_t(lanes) * cp * 2 * sizeof(float), s, x, partials, hw, cp8, splits);
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
    inorm_partials_to_stats_kernel<<<n, kThreads, 0, s>>>(partials, stats, cp, splits);
    return cudaGetLastError();
}

cudaError_t instance_norm(const __half* x, __half* y, int n, int hw, int cp, const float* gamma, const float* beta, float eps, int act,
                          float* partials, cudaStream_t s, int group_size, int channels, const NormStore* store) {
    if (group_size < 1) group_size = 1;
    if (channels <= 0) channels = cp;
    NormDst dst{};
    if (cudaError_t e = norm_dst_of(store, x, y, hw, &dst); e != cudaSuccess) return e;
    __half* y_plain = store && store->padded() ? store->plain : nullptr;
    const int cp8 = cp / 8;
    if (cp8 > kThreads) return cudaErrorInvalidValue;  // > 2048 channels: not on any supported model
    if (const int csz = inorm_cluster_size(hw, group_size, size_t(n) * hw * cp * 2); csz > 0 && n <= 65535) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned((cp8 + 1) / 2 * csz), unsigned(n));
        cfg.blockDim = dim3(kThreads);
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = unsigned(csz);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        return cudaLaunchKernelEx(&cfg, inorm_cluster_kernel, x, y, gamma, beta, hw, cp8, eps, act, group_size, channels, dst, y_plain);
    }
    if ((dst.ho ? size_t(dst.ho) * dst.wo * cp8 : size_t(hw) * cp8) >> 32) return cudaErrorInvalidValue;  // 32-bit vector offsets inside an image
    const int splits = instance_norm_splits(hw, cp);
    const int lanes = kThreads / cp8;
    const size_t smem1 = size_t(lanes) * cp * 2 * sizeof(float);
    // lanes * cp <= 2048 floats x 2 => at most 16 KB: no opt-in needed
    const size_t n8 = size_t(hw) * cp8;
    // Both passes read x.  Images are handled in groups of at most ~48 MB so that the second pass finds its group in the 126 MB L2:
    // DRAM sees one read and one write of the tensor (algorithmic traffic) whatever the batch size.
    const int group = inorm_group(n, hw, cp);
    for (int i0 = 0; i0 < n; i0 += group) {
        const int gn = std::min(group, n - i0);
        const __half* xg = x + size_t(i0) * n8 * 8;
        __half* yg = y + size_t(i0) * (dst.ho ? size_t(dst.ho) * dst.wo * cp8 : n8) * 8;
        float* pg = partials + size_t(i0) * splits * cp * 2;
        (void)launch_pdl_smem(inorm_stats_kernel, dim3(dim3(splits, gn)), dim3(kThreads), smem1, s, xg, pg, hw, cp8, splits);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const int chunks = int(std::max<size_t>(1, (n8 + size_t(kThreads) * 8 - 1) / (size_t(kThreads) * 8)));  // one block per 2048 vectors
        float* params = pg + size_t(gn) * splits * cp * 2;  // after this group's partials (see instance_norm_scratch_floats)
        int parts = 1;  // lanes per channel: a power of two, at most a warp, no more than the splits there are
        while (parts < 32 && parts * 2 <= splits && cp * parts * 2 <= kThreads * 4) parts <<= 1;
        (void)launch_pdl(inorm_finalize_kernel, dim3(gn), dim3(kThreads), s, pg, gamma, beta, params, hw, cp, splits, eps, group_size, channels, parts);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const size_t ring8 = dst.ho ? size_t(dst.ho * dst.wo - dst.h * dst.w) * cp8 : 0;
        const int ring_chunks = int((ring8 + size_t(kThreads) * 8 - 1) / (size_t(kThreads) * 8));
        (void)launch_pdl(inorm_apply_kernel<false, false>, dim3(dim3(chunks + ring_chunks, gn)), dim3(kThreads), s, xg, yg, params, hw, cp8, act, dst, chunks, y_plain ? y_plain + size_t(i0) * n8 * 8 : nullptr, NormStats{}, 1,
                         static_cast<const __half*>(nullptr), 0);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t depthwise_conv(const __half* x, const __half* w, const float* bias, __half* y, int n, int h, int wd, int cp, int p, int q, int kh,
                           int kw, int sh, int sw, int dh, int dw, int pt, int pl, int act, float lo, float hi, cudaStream_t s) {
    if (size_t(n) * p > 65535) return cudaErrorInvalidValue;  // grid.y limit (batch x output rows)
    const int row_items = q * (cp / 8);
    dim3 grid(unsigned((row_items + kThreads - 1) / kThreads), unsigned(n * p));
    return launch_pdl(depthwise_kernel, grid, dim3(kThreads), s, x, w, bias, y, n, h, wd, cp / 8, p, q, kh, kw, sh, sw, dh, dw, pt, pl, act, lo, hi);
}

}  // namespace k
}  // namespace smelter
