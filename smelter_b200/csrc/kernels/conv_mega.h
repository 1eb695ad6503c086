// Host interface of the persistent multi-layer convolution kernel (see conv_mega.cu).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "conv_igemm.h"

namespace smelter {
namespace k {

constexpr int kMegaMaxLayers = 56;  // 56 x (128 B of scalars + 3 tensor maps) stays below the 32 KiB kernel-parameter limit

struct alignas(16) MegaLayer {  // 128 bytes
    int M, num_m_tiles, num_n_tiles, block_n;
    int total_kb, kblocks_per_tap, taps_w, taps_h;  // total_kb = main_kb + residual k-blocks
    int mode, PQ, Q, stride_h;
    int stride_w, dil_h, dil_w, corner_h;
    int corner_w, in_h, in_w, out_pitch;
    int act;
    float clip_lo, clip_hi;
    int dep;        // layer of the run that produces this layer's input, -1 = produced before the launch
    int res_dep;    // same for the residual operand (-1 = none or external)
    int item_base;  // first global item (output tile) index of this layer
    int item_end;
    int flag_base;  // first per-m-tile completion counter of this layer
    const float* bias;
    int main_kb;    // k-blocks of the convolution proper
    int res_map;    // index into MegaParams::tm_res when a residual tensor is added, else -1
};
static_assert(sizeof(MegaLayer) == 128, "MegaLayer layout");

constexpr int kMegaMaxResidual = 24;

struct MegaParams {
    int num_layers, total_items;
    int debug;  // perf experiments (env SMELTER_MEGA_DEBUG): 1 = the scout does not wait (results racy), 4 = no release fence before signals
    int reserved;
    unsigned int* flags;       // per (layer, m-tile): epilogue warps that have stored their part (complete at 4 * num_n_tiles)
    unsigned long long* timeline;  // instrumented builds only: [layer][2] first-load / last-accumulator times
    MegaLayer L[kMegaMaxLayers];
    CUtensorMap tm_a[kMegaMaxLayers], tm_b[kMegaMaxLayers], tm_out[kMegaMaxLayers];
    CUtensorMap tm_res[kMegaMaxResidual];  // residual tensors as [32 row x 64 column] boxes (the geometry of an output map)
    CUtensorMap tm_ident[2];               // 64 x 64 identity, boxes of 64 / 128 rows (rows outside the tensor read as zero)
};
static_assert(sizeof(MegaParams) <= 32764, "kernel parameter space");

struct MegaLaunch {
    std::unique_ptr<MegaParams> params;
    size_t sync_words = 0;  // unsigned ints at params->layer_done zeroed before every launch
    int grid = 0;
    double flops = 0;
};

int conv_mega_pick_block_n(int c_out, int m_tiles, int num_kb, int num_sms);
// unsigned ints of device memory the run needs for its completion counters
size_t conv_mega_sync_words(const std::vector<ConvTcProblem>& layers);
// layers[i].x / .y / .residual / .w_packed / .bias are final device pointers; dep[i] / res_dep[i] index into `layers` or are -1
// `identity` = device pointer to a dense 64 x 64 fp16 identity matrix
bool conv_mega_prepare(MegaLaunch* out, const std::vector<ConvTcProblem>& layers, const std::vector<int>& dep, const std::vector<int>& res_dep,
                       int num_sms, unsigned int* sync_words, const __half* identity, std::string* err);

cudaError_t conv_mega_launch(const MegaLaunch& L, cudaStream_t stream);

}  // namespace k
}  // namespace smelter
