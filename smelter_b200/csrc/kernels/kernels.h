// Launchers of the non-GEMM kernels.  All activations are NHWC fp16 with the channel pitch `cp` a multiple
// of 8 (one 128-bit vector = 8 channels); padded channels carry zeros / don't-care values.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "conv_igemm.h"

namespace smelter {
namespace k {

enum UnaryKind : int {
    UN_RELU = 0, UN_SIGMOID = 1, UN_CLIP = 2, UN_TANH = 3, UN_ABS = 4, UN_EXP = 5, UN_LOG = 6, UN_ELU = 7,
    UN_LEAKY_RELU = 8, UN_HARD_SIGMOID = 9, UN_SOFTPLUS = 10, UN_SOFTSIGN = 11, UN_IDENTITY = 12, UN_POW = 13
};
enum BinaryKind : int { BIN_ADD = 0, BIN_SUB = 1, BIN_MUL = 2, BIN_DIV = 3 };
enum PadMode : int { PAD_CONSTANT = 0, PAD_REFLECT = 1, PAD_EDGE = 2 };
enum UpsampleMode : int { UP_NEAREST = 0, UP_BILINEAR = 1 };

// Boundary layout conversion.  dst is [N, H+pt+pb, W+pl+pr, cp] with zero borders and zero padded channels.
// NCHW (c <= 4) -> zero-padded, 2x2 space-to-depth NHWC with 16-channel pixels: out[n][y][x][(dy*2+dx)*c + ch] = in[n][ch][2y+dy-pad_t][2x+dx-pad_l]
cudaError_t nchw_to_s2d(const __half* src, __half* dst, int n, int c, int h, int w, int pad_t, int pad_l, int h2, int w2, cudaStream_t s);
// NCHW source (c <= 4) -> the 4 x 4 space-to-depth fold [n, h4, w4, 16 x 4] of its padded image (PadMode `mode`, pads top / left; the
// bottom / right pads follow from h4, w4): channel (dy * 4 + dx) * 4 + ch of folded pixel (y4, x4) = padded pixel (4 y4 + dy, 4 x4 + dx).
// The input layout of an input-folded convolution (engine.h Filter::in_fold).
cudaError_t nchw_to_s2d4(const __half* src, __half* dst, int n, int c, int h, int w, int pad_t, int pad_l, int h4, int w4, int mode, float value,
                         cudaStream_t s);
cudaError_t nchw_to_nhwc(const __half* src, __half* dst, int n, int c, int h, int w, int cp, int pad_t, int pad_l, int pad_b,
                         int pad_r, cudaStream_t s);
// dst[n * dst_image_pitch + (c*H + h)*W + w] = src[n,h,w,c]
cudaError_t nhwc_to_nchw(const __half* src, __half* dst, int n, int c, int h, int w, int cp, long dst_image_pitch, cudaStream_t s);
// result of a phase-folded convolution [n, p2, q2, cp] (channel (ey * F + ex) * c + co, F = fold = 2 or 4) -> NCHW [n, c, F * p2, F * q2]
cudaError_t phase_to_nchw(const __half* src, __half* dst, int n, int c, int p2, int q2, int cp, long dst_image_pitch, cudaStream_t s, int fold = 2);

// ConvTranspose input: y[n][lo_h + s_h*i][lo_w + s_w*j][c] = x[n][i][j][c], zero elsewhere; y is [n][hz][wz][cp]
cudaError_t zero_stuff2d(const __half* x, __half* y, int n, int h, int w, int cp, int hz, int wz, int stride_h, int stride_w, int lo_h, int lo_w,
                         cudaStream_t s);
// Configuration.inputConstraint = .forceInputScale (ONNXGraph.swift:219-241): planar NCHW fp16 [planes][hs][ws] -> [planes][hd][wd].
// mode 0 = bilinear, 1 = Lanczos-3; half-pixel centres, taps clamped to the image edge, weights normalised.
cudaError_t resize_planes(const __half* src, __half* dst, int planes, int hs, int ws, int hd, int wd, int mode, cudaStream_t s);
// `c` logical channels at pitch `cp` (0 = dense, no padded lanes): the padded lanes of a pixel are written as zeros, never f(0)
cudaError_t unary(const __half* x, __half* y, size_t n_elems, int kind, float alpha, float beta, cudaStream_t s, int c = 0, int cp = 0);
// bcast_pixels = H * W of a: b holds one pixel per image ([N, C, 1, 1]) and is broadcast over a's pixels
cudaError_t binary(const __half* a, const __half* b, __half* y, size_t n_elems, int kind, int act, cudaStream_t s, int c = 0, int cp = 0,
                   size_t bcast_pixels = 0);
// y = act(x * scale[c] + shift[c])  (un-fused BatchNormalization)
cudaError_t scale_shift(const __half* x, __half* y, size_t pixels, int cp, const float* scale, const float* shift, int act,
                        cudaStream_t s);
cudaError_t pool2d(const __half* x, __half* y, int n, int h, int w, int cp, int p, int q, int kh, int kw, int sh, int sw, int ph,
                   int pw, int is_max, cudaStream_t s);
cudaError_t global_avgpool(const __half* x, __half* y, int n, int hw, int cp, cudaStream_t s);
cudaError_t softmax_rows(const __half* x, __half* y, size_t rows, int c, int cp, int log_softmax, cudaStream_t s);
cudaError_t upsample2d(const __half* x, __half* y, int n, int h, int w, int cp, int scale_h, int scale_w, int mode, int align_corners,
                       cudaStream_t s);
// s2d_out = F (2 or 4): the padded image is written as its F x F space-to-depth fold [n, ho / F, wo / F, F * F * cp] (F divides ho
// and wo): the input layout of a phase-folded convolution
// x2 != nullptr: the padded tensor is act(x + x2) (the residual Add in front of the Pad); y_plain != nullptr also stores the un-padded sum
cudaError_t pad2d(const __half* x, __half* y, int n, int h, int w, int cp, int pt, int pl, int pb, int pr, int mode, float value,
                  cudaStream_t s, int s2d_out = 0, const __half* x2 = nullptr, __half* y_plain = nullptr, int act = 0);
// Instance normalisation: deterministic three-kernel scheme (split statistics, per-image scale / shift, apply).
// `partials` holds instance_norm_scratch_floats(n, hw, cp) floats.
size_t instance_norm_scratch_floats(int n, int hw, int cp);
int instance_norm_splits(int hw, int cp);
int instance_norm_launches(int n, int hw, int cp, int group_size = 1);  // kernels instance_norm() enqueues: 1 (cluster form) or 3
// group_size = channels that share one mean / variance: 1 = InstanceNormalization, C / groups = group normalisation
// Where instance_norm() stores a normalised pixel (default: where it came from).  Statistics do not depend on pixel order, so the
// kernel can undo a permutation and write into a larger image on its way out.  x != y whenever anything is set.
struct NormStore {
    int unfold_w = 0;  // W > 0: x holds a 2H x 2W image as [H, W, 2 x 2 phases] pixels (hw = 4 H W, the output of an upsample-folded
                       // convolution, engine.h Filter::upfold); the store goes to plain row-major pixels
    int unfold_f = 2;  // fold factor F of those pixels: [H, W, F x F phases] of an F H x F W image (2, or 4 behind an input-folded convolution)
    int h = 0, w = 0;  // image size as stored (after un-folding, h * w == hw); needed with padding
    int pad_t = 0, pad_l = 0, pad_b = 0, pad_r = 0;  // the Pad behind the norm, written by the norm: y is the padded image
    int pad_mode = PAD_REFLECT;                      // PAD_REFLECT or PAD_EDGE
    __half* plain = nullptr;  // with padding: the un-padded result is stored here as well (it has other readers)
    int s2d = 0;       // ... in pad2d's F x F space-to-depth layout (F = 2 or 4)
    bool padded() const { return pad_t || pad_l || pad_b || pad_r || s2d; }
};
cudaError_t instance_norm(const __half* x, __half* y, int n, int hw, int cp, const float* gamma, const float* beta, float eps,
                          int act, float* partials, cudaStream_t s, int group_size = 1, int channels = 0, const NormStore* store = nullptr);
// The same normalisation behind a convolution that accumulated the statistics in its epilogue (ConvTcProblem::stats, conv_igemm.h):
// ONE launch, x read once.  stats: [n][phases][cp] fp64 {sum, sum of squares} (phases = F^2 behind a phase-column convolution whose
// columns are phase * cp + channel, else 1); the kernel leaves stats and *counter zero for the next encode.
cudaError_t instance_norm_from_stats(const __half* x, __half* y, int n, int hw, int cp, const float* gamma, const float* beta, float eps, int act,
                                     double* stats, unsigned int* counter, int phases, cudaStream_t s, const NormStore* store = nullptr,
                                     const __half* res = nullptr, int act2 = 0);  // res: y = act2(act(norm(x)) + res), res in x's (plain) pixel order
// counter == nullptr: the accumulators are left as they are (the caller owns them).
// The accumulators of a tensor that no convolution epilogue produced (benchmarks / tests of the one-pass norm alone): [n][cp][2]
cudaError_t instance_norm_stats_f64(const __half* x, int n, int hw, int cp, float* partials, double* stats, cudaStream_t s);
// copy `c_src_pitch` channels of every pixel of src into dst at channel offset c_off
cudaError_t concat_channels(const __half* src, __half* dst, size_t pixels, int c_src, int c_src_pitch, int c_dst_pitch, int c_off,
                            cudaStream_t s);
// Depthwise convolution (groups == C, multiplier 1).  w: [kh*kw][cp] fp16, bias fp32 [cp].
cudaError_t depthwise_conv(const __half* x, const __half* w, const float* bias, __half* y, int n, int h, int wd, int cp, int p, int q,
                           int kh, int kw, int sh, int sw, int dh, int dw, int pt, int pl, int act, float lo, float hi, cudaStream_t s);

cudaError_t f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t s);
cudaError_t f16_to_f32(const __half* src, float* dst, size_t n, cudaStream_t s);
// interleaved u8 [n][hw][sc] -> planar fp16 [n][c][hw], dst = byte * scale[ch] + bias[ch] (scale/bias passed by value, c <= 4)
cudaError_t u8_to_nchw_f16(const uint8_t* src, __half* dst, int n, int c, size_t hw, int sc, const float* scale4, const float* bias4, cudaStream_t s);
// 64-bit order-independent checksum (sum of 32-bit words with a position-mixing multiplier), result on device.
cudaError_t checksum64(const void* p, size_t bytes, unsigned long long* out_dev, cudaStream_t s);

}  // namespace k
}  // namespace smelter
