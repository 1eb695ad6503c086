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
    int total_kb, kblocks_per_tap, taps_w, taps_h;
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
    const __half* residual;  // NHWC [M, out_pitch] or null
};
static_assert(sizeof(MegaLayer) == 128, "MegaLayer layout");

struct MegaParams {
    int num_layers, total_items;
    int debug;  // perf experiments (env SMELTER_MEGA_DEBUG): 1 = skip dependency waits (results racy), 2 = skip residual loads, 4 = skip signal fences
    int reserved;
    unsigned int* layer_done;  // [kMegaMaxLayers] m-tiles completed per layer
    unsigned int* flags;       // per (layer, m-tile): epilogue warps that have stored their part (complete at 4 * num_n_tiles)
    MegaLayer L[kMegaMaxLayers];
    CUtensorMap tm_a[kMegaMaxLayers], tm_b[kMegaMaxLayers], tm_out[kMegaMaxLayers];
};
static_assert(sizeof(MegaParams) <= 32764, "kernel parameter space");

struct MegaLaunch {
    std::unique_ptr<MegaParams> params;
    size_t sync_words = 0;  // unsigned ints at params->layer_done zeroed before every launch
    int grid = 0;
    double flops = 0;
};

int conv_mega_pick_block_n(int c_out, int m_tiles, int num_sms);
// unsigned ints of device memory the run needs for its completion counters
size_t conv_mega_sync_words(const std::vector<ConvTcProblem>& layers);
// layers[i].x / .y / .residual / .w_packed / .bias are final device pointers; dep[i] / res_dep[i] index into `layers` or are -1
bool conv_mega_prepare(MegaLaunch* out, const std::vector<ConvTcProblem>& layers, const std::vector<int>& dep, const std::vector<int>& res_dep,
                       int num_sms, unsigned int* sync_words, std::string* err);
cudaError_t conv_mega_launch(const MegaLaunch& L, cudaStream_t stream);

}  // namespace k
}  // namespace smelter
