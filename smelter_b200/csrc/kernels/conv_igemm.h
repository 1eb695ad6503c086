// Host interface of the tcgen05 implicit-GEMM convolution (see conv_igemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>

namespace smelter {
namespace k {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_CLIP = 2, ACT_SIGMOID = 3 };

enum ConvMode : int {
    CONV_MODE_TILED = 1,       // 1x1, stride 1, no padding: A is the activation matrix [M, Cin_pad]
    CONV_MODE_IM2COL = 2,      // general: TMA im2col over NHWC
    CONV_MODE_PACKED_ROW = 3,  // Cin pitch 8: one K block = one filter row (S taps x 8 channels), padding materialised
};

struct ConvKernelParams {
    int M;                 // output pixels N*P*Q
    int out_pitch;         // elements per output pixel (Cout rounded up to 8)
    int num_m_tiles, num_n_tiles;
    int num_taps, taps_w, kblocks_per_tap;
    int P, Q, PQ;
    int stride_h, stride_w, dil_h, dil_w;
    int corner_h, corner_w;  // im2col lower corner (= -pad)
    int mode;
    const float* bias;       // fp32, zero padded to a multiple of 256 entries
    int has_residual;        // an NHWC tensor with the output's shape (tm_res) is added before the activation
    int act;
    float clip_lo, clip_hi;
    int splits;              // k-splits per output tile (1 = none); work items = tiles * splits
    int kb_per_split;        // k-blocks per split (the last split may have fewer, never zero)
    float* ws;               // split-K workspace: [tiles * splits][128][BLOCK_N] fp32 partial accumulators
    unsigned int* counters;  // split-K: one arrival counter per tile, zero between launches
    unsigned long long* timeline;  // perf experiments only (env SMELTER_CONV_TIMELINE): CTA 0 writes %globaltimer stamps here
    unsigned long long* chain;     // perf experiments only (env SMELTER_CHAIN_TIMELINE, instrumented builds): per-launch {min, max} over CTAs of six stamps
    int debug_flags;         // perf experiments only (env SMELTER_CONV_DEBUG): 16 = producers skip the TMA loads (results wrong)
    int use_pdl;             // the launch carries the programmatic-serialization attribute: call griddepcontrol.wait
    int l2_hints;            // 1 = output stores evict_last (tensor re-read by later layers), 2 = residual loads evict_first (last use)
    // Projection shortcut folded into the k-loop (conv_pair.cu only): after the main k-blocks, `side_kb` more k-blocks read a
    // second activation tensor (tm_a2: 1x1, stride side_stride_*, no padding, same P x Q) against a second weight matrix (tm_b2).
    int side_kb;             // 0 = none
    int side_mode;           // CONV_MODE_TILED (stride 1) or CONV_MODE_IM2COL
    int side_stride_h, side_stride_w;
    const float* bias2;      // the shortcut's bias, added to `bias` in the epilogue (nullptr = none)
    // Statistics for the instance norm behind the layer (conv_pair.cu only, no residual): every 128-row tile adds the column sums and
    // sums of squares of its fp16 outputs to stats[image][column] = {sum, sum of squares} (fp64 reductions; zero before the launch).
    // Layout [image][replica][column][2]: tile t of an image adds to replica t mod stats_reps (a power of two), the norm adds them up.
    double* stats;           // nullptr = none
    int stats_rows;          // output rows per image (a multiple of 128: no tile straddles two images)
    int stats_reps;
};

struct ConvTcProblem {
    int mode;
    int n, h, w;            // input (for PACKED_ROW: the physically padded input)
    int c_in;               // logical input channels (flop accounting only)
    int c_in_pitch;         // elements per input pixel (multiple of 8)
    int c_out, c_out_pitch;
    int k_h, k_w, stride_h, stride_w, dil_h, dil_w;
    int pad_t, pad_l, pad_b, pad_r;
    const __half* x;        // NHWC
    const __half* w_packed; // [Cout][taps][kc]
    const float* bias;
    const __half* residual;
    __half* y;              // NHWC, pitch c_out_pitch
    int act;
    float clip_lo, clip_hi;
    int block_n;            // 0 = auto
    // Optional projection shortcut y += conv1x1(side_x) + side_bias, accumulated in the same TMEM tile ahead of the activation
    // (the 1x1 "downsample" convolution of a residual block: no separate launch, no round trip of its output through HBM).
    const __half* side_x;   // NHWC [n, side_h, side_w, side_c_in_pitch]; nullptr = none
    const __half* side_w_packed;  // [Cout][1][side_c_in_pitch]
    const float* side_bias;
    int side_h, side_w, side_c_in, side_c_in_pitch, side_stride_h, side_stride_w;
    int l2_hints;           // ConvKernelParams::l2_hints
    int pair;               // 1 = two-CTA clusters with cta_group::2 MMAs (conv_pair.cu); 0 = single-CTA kernel
    int splits;             // 0 = auto (conv_tc_plan), 1 = no split-K
    float* split_ws;        // workspace of conv_tc_plan().ws_bytes when splits > 1
    unsigned int* split_counters;  // conv_tc_plan().counter_bytes, zero-initialised once; the kernel leaves them zero
    double* stats;          // ConvKernelParams::stats: [n][conv_tc_stats_reps()][c_out_pitch][2], zero before the launch (see conv_tc_stats_supported)
};

struct ConvTcPlanInfo {
    int block_n, splits;
    size_t ws_bytes, counter_bytes;
};
// Tile width and k-split choice for a problem (pure function of the shapes).
ConvTcPlanInfo conv_tc_plan(const ConvTcProblem& q, int num_sms);
// Whether a problem with a projection shortcut (side_*) can run as one launch: the two-CTA kernel without split-K must be the
// choice for the problem both with and without the extra k-blocks.
bool conv_tc_side_supported(const ConvTcProblem& q, int num_sms);
// Whether the layer can accumulate per-image column statistics in its epilogue (ConvTcProblem::stats): the two-CTA kernel without
// split-K, residual or activation, output rows per image a multiple of 128.
bool conv_tc_stats_supported(const ConvTcProblem& q, int num_sms);
// Accumulator replicas per image for that: enough that few tiles meet on one address, few enough that the norm's prologue stays short
int conv_tc_stats_reps(const ConvTcProblem& q);

struct ConvTcLaunch {
    CUtensorMap tm_a, tm_b, tm_out, tm_res;
    CUtensorMap tm_a2, tm_b2;  // projection shortcut operands (p.side_kb > 0)
    ConvKernelParams p;
    int block_n;
    int splits;
    int grid;
    int use_pdl;   // launch with programmatic stream serialization (the kernel calls griddepcontrol.wait itself)
    int pair;      // launched as conv_pair_kernel (two-CTA clusters): tm_b boxes hold BLOCK_N / 2 rows
    int num_sms;
    int balanced_grid;  // conv_pair_launch: smallest grid that needs the same number of rounds (plans for a share of the chip)
    double flops;  // algorithmic: 2*M*Cout*Cin*R*S
};

int conv_tc_pick_block_n(int c_out, int m_tiles, int num_sms);
// 128B-swizzled 2-D map over a dense row-major fp16 matrix [rows][cols] (out-of-range box elements read as zero)
bool conv_tc_encode_2d(CUtensorMap* tm, const __half* base, long cols, long rows, int box_cols, int box_rows, std::string* err);
bool conv_tc_prepare(ConvTcLaunch* L, const ConvTcProblem& q, int num_sms, std::string* err);
cudaError_t conv_tc_launch(const ConvTcLaunch& L, cudaStream_t stream);
void conv_tc_dump_timeline(const ConvTcLaunch& L);  // perf experiments only
void conv_tc_chain_reset();                          // perf experiments only: whole-encode launch chain timeline (conv_pair.cu stamps)
void conv_tc_chain_dump();
// two-CTA variant (conv_pair.cu)
bool conv_pair_supported(const ConvTcLaunch& L);
cudaError_t conv_pair_set_attr(int block_n);
cudaError_t conv_pair_launch(const ConvTcLaunch& L, int num_sms, cudaStream_t stream);
// benchmark only: TMA load rate of [128 x 64] fp16 boxes; mode 0 = 2-D tiled over [N*H*W, C], 1 = im2col (3x3, pad 1)
int tma_probe3(int mode, int c, long rows_total, int slabs, int stages, int iters, int grid, const __half* x, cudaStream_t stream, float* ms,
               std::string* err);
int tma_probe2(int mode, int c, long rows_total, int box_c, int box_r, int csz, int stages, int iters, int grid, const __half* x, cudaStream_t stream,
               float* ms, std::string* err);
int tma_probe(int mode, int c, int w, int h, int n, int stages, int iters, int grid, int distinct, const __half* x, cudaStream_t stream, float* ms,
              std::string* err);

}  // namespace k
}  // namespace smelter
