/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library.  The engine (smelter_b200/csrc) never links or calls it.
 *
 * Plain-C restatement of the byte/integer host loops on Smelter's inference path, each function citing the
 * reference lines it follows (paths relative to the Smelter repository, Sources/Smelter/...):
 *
 *   oracle_reformat_conv_weight   Extensions/Foundation/Array+Extensions.swift:52-93 (4-deep scalar loop; the
 *                                 same index arithmetic, loop order oc/ic/kh/kw)
 *   oracle_float16_to_32 / _32_to_16   Float16.swift:17-45 / 53-77 (vImageConvert_Planar16FtoPlanarF and the
 *                                 inverse: IEEE binary16 <-> binary32, round-to-nearest-even, subnormals kept)
 *   oracle_conv_padded_size       Padding/ONNXConvolutionPadding.swift:91-113 (exact reference formula, NO dilation
 *                                 term, transpose branch included) and oracle_conv_padded_size_dilated (ONNX spec)
 *   oracle_pool_padded_size       Padding/PyTorchPoolPadding.swift:94-103 (Int(Float(..)/Float(..) + 1.0))
 *   oracle_onnx2mps_swizzle       /ONNX2MPS.py:54-67 (numpy transpose [0,2,3,1] / [1,2,3,0] + [:, ::-1, ::-1, :])
 *
 * PARITY STATUS: **parity unpinned** — the reference ships no tests, fixtures or golden vectors (SURVEY.md §4,
 * §8c) and cannot be compiled here (Swift + Apple frameworks).  This restatement is pinned only against
 * numpy (transpose / astype(float16)) in tests/test_oracle.py, with the generated vectors under tests/golden/.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* Array.reformatingConvolutionWeight — Array+Extensions.swift:52-93 */
void oracle_reformat_conv_weight(const void* src, void* dst, int elem_size, int output_channels, int input_channels,
                                 int kernel_height, int kernel_width, int is_transpose) {
    const char* s = (const char*)src;
    char* d = (char*)dst;
    for (int oc = 0; oc < output_channels; ++oc) {
        for (int ic = 0; ic < input_channels; ++ic) {
            for (int kh = 0; kh < kernel_height; ++kh) {
                for (int kw = 0; kw < kernel_width; ++kw) {
                    size_t input_idx, output_idx;
                    if (is_transpose) { /* :70-76 */
                        input_idx = (size_t)ic * output_channels * kernel_height * kernel_width +
                                    (size_t)oc * kernel_height * kernel_width + (size_t)kh * kernel_width + kw;
                        output_idx = (size_t)oc * kernel_height * kernel_width * input_channels +
                                     (size_t)(kernel_height - 1 - kh) * kernel_width * input_channels +
                                     (size_t)(kernel_width - 1 - kw) * input_channels + ic;
                    } else { /* :77-84 */
                        input_idx = (size_t)oc * input_channels * kernel_height * kernel_width +
                                    (size_t)ic * kernel_height * kernel_width + (size_t)kh * kernel_width + kw;
                        output_idx = (size_t)oc * input_channels * kernel_height * kernel_width +
                                     (size_t)kh * kernel_width * input_channels + (size_t)kw * input_channels + ic;
                    }
                    memcpy(d + output_idx * elem_size, s + input_idx * elem_size, (size_t)elem_size); /* :86 */
                }
            }
        }
    }
}

/* ONNX2MPS.py:54-67 — generic 4-D transpose by `perm` followed by the optional spatial flip of axes 1,2 */
void oracle_onnx2mps_swizzle(const void* src, void* dst, int elem_size, const int dims[4], const int perm[4], int flip) {
    int od[4];
    for (int i = 0; i < 4; ++i) od[i] = dims[perm[i]];
    size_t in_stride[4];
    in_stride[3] = 1;
    for (int i = 2; i >= 0; --i) in_stride[i] = in_stride[i + 1] * (size_t)dims[i + 1];
    const char* s = (const char*)src;
    char* d = (char*)dst;
    for (int a = 0; a < od[0]; ++a)
        for (int b = 0; b < od[1]; ++b)
            for (int c = 0; c < od[2]; ++c)
                for (int e = 0; e < od[3]; ++e) {
                    const int sb = flip ? od[1] - 1 - b : b;
                    const int sc = flip ? od[2] - 1 - c : c;
                    const int idx[4] = {a, sb, sc, e};
                    size_t in = 0;
                    for (int i = 0; i < 4; ++i) in += (size_t)idx[i] * in_stride[perm[i]];
                    const size_t out = (((size_t)a * od[1] + b) * od[2] + c) * od[3] + e;
                    memcpy(d + out * elem_size, s + in * elem_size, (size_t)elem_size);
                }
}

/* float16to32 — Float16.swift:17-45 */
static float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    int exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            exp = 1;
            while (!(man & 0x400u)) { man <<= 1; --exp; }
            man &= 0x3ffu;
            bits = sign | (uint32_t)(exp + 112) << 23 | man << 13;
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | man << 13;
    } else {
        bits = sign | (uint32_t)(exp + 112) << 23 | man << 13;
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
void oracle_float16_to_32(const uint16_t* src, float* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = half_to_float(src[i]);
}

/* float32to16 — Float16.swift:53-77; round-to-nearest-even done on the 64-bit widened significand */
static uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7e00u | ((ax & 0x7fffffu) >> 13)); /* NaN (quiet) */
    if (ax == 0x7f800000u) return (uint16_t)(sign | 0x7c00u);
    const int e = (int)(ax >> 23) - 127;
    if (e > 15) return (uint16_t)(sign | 0x7c00u);
    uint64_t sig = (ax & 0x7fffffu);
    int drop; /* number of low bits of the 24-bit significand to round away */
    uint32_t base;
    if (e >= -14) {
        sig |= 0x800000u;
        drop = 13;
        base = (uint32_t)(e + 14) << 10; /* exponent field minus the implicit one carried in sig */
    } else {
        if (e < -25) return sign;
        sig |= 0x800000u;
        drop = 13 + (-14 - e);
        base = 0;
    }
    uint64_t q = sig >> drop;
    const uint64_t rem = sig & (((uint64_t)1 << drop) - 1);
    const uint64_t half = (uint64_t)1 << (drop - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    uint32_t h = (e >= -14) ? base + (uint32_t)q : (uint32_t)q; /* q carries the implicit bit: base+q = (e+15)<<10 | mant */
    if (h >= 0x7c00u) h = 0x7c00u;
    return (uint16_t)(sign | h);
}
void oracle_float32_to_16(const float* src, uint16_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = float_to_half(src[i]);
}

/* ONNX_ConvolutionPadding.paddedSize — ONNXConvolutionPadding.swift:91-113 (reference formula: no dilation) */
int oracle_conv_padded_size(int input, int kernel, int stride, int pad_lo, int pad_hi, int output_padding, int is_transpose) {
    if (is_transpose) return (input - 1) * stride - pad_lo - pad_hi + kernel + output_padding; /* :97-103 */
    return (input + pad_lo + pad_hi - kernel) / stride + 1;                                      /* :105-110 */
}
/* ONNX operator spec (what the engine follows): effective kernel = dilation*(k-1)+1 */
int oracle_conv_padded_size_dilated(int input, int kernel, int stride, int dilation, int pad_lo, int pad_hi) {
    const int num = input + pad_lo + pad_hi - (dilation * (kernel - 1) + 1);
    if (num < 0) return 0;
    return num / stride + 1;
}
/* PyTorchPoolPadding.paddedSize — PyTorchPoolPadding.swift:94-103 */
int oracle_pool_padded_size(int input, int kernel, int stride, int padding) {
    return (int)((float)(input + 2 * padding - kernel) / (float)stride + 1.0f);
}
