// Persistent multi-layer ("mega") variant of the tcgen05 implicit-GEMM convolution.
//
// A run of consecutive Conv nodes of the graph walk (reference Sources/Smelter/ONNXGraph.swift:169-193 emits one MPS node per
// ONNX node; MPSNNGraph then schedules them, closed source) is executed by ONE kernel launch: 148 persistent CTAs walk the
// concatenated list of output tiles of all layers of the run (global round-robin, so a layer whose tile count is not a
// multiple of the SM count does not leave SMs idle), and a tile starts as soon as the input tiles it reads are complete:
//
//   * every finished output tile bumps a per-(layer, m-tile) counter in global memory (release), after its TMA stores have
//     completed; the activation (A) producer of a dependent tile polls the counters of the m-tiles its im2col window covers
//     (acquire) before it issues the first load; the epilogue does the same for the residual operand.
//   * no grid-wide barrier and no kernel boundary between layers: the pipeline-fill / drain / launch cost (~4 us per layer on
//     B200, more than the tensor time of most ResNet-50 layers at batch 32) is paid once per run; weight (B) tiles of the next
//     layer are prefetched while the current layer is still in its last tiles because the operand ring is shared by all layers.
//   * deadlock freedom: all CTAs are co-resident (grid <= #SMs, 1 CTA/SM), every CTA visits its items in increasing global
//     order and an item only depends on items of earlier layers, i.e. on smaller global indices.
//   * every layer output of a run has its own buffer (engine.cc: no arena reuse inside a run), so the only hazard is RAW.
//
// Tile geometry is uniform so the ring never drains between layers: 128 x {64,128} x 64, six 32 KiB stages, fp32 accumulators
// double buffered in TMEM (2 x 128 columns), warp roles as in conv_igemm.cu.  The residual is read with plain global loads
// (L2 only: the data was written by other SMs during this launch) instead of TMA, which keeps the epilogue at one 4 KiB
// staging buffer per warp.
#include "conv_mega.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ptx.cuh"

namespace smelter {
namespace k {

using namespace ptx;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 512;
constexpr int kNumAProducers = 4;
constexpr int kNumBProducers = 3;
constexpr int kBProducerWarp0 = 13;
constexpr int kEpilogueWarp0 = 4;
constexpr int kEpilogueWarps = 8;
constexpr int kMmaWarp = 12;
constexpr int kStages = 6;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;       // 16 KiB
constexpr uint32_t kBSlotBytes = 128 * kBlockK * 2;       // 16 KiB (BLOCK_N <= 128)
constexpr uint32_t kStageBytes = kABytes + kBSlotBytes;   // 32 KiB
constexpr int kChunkN = 64;
constexpr uint32_t kEpiBufBytes = 32 * kChunkN * 2;       // 4 KiB per epilogue warp
constexpr uint32_t kBiasSlotBytes = kChunkN * 4;
constexpr uint32_t kBarrierBytes = 512;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + kEpilogueWarps * (kEpiBufBytes + kBiasSlotBytes) + kBarrierBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr uint32_t kTmemCols = 256;
constexpr long long kSpinLimitCycles = 4000000000LL;  // ~2 s: a dependency that never arrives traps instead of hanging the GPU

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimitCycles) __trap();
    }
}
__device__ __noinline__ float sigmoid_slow(float v) { return 1.f / (1.f + __expf(-v)); }

// Blocks until every m-tile in [lo, hi] of layer `d` is complete.  `done_mask` memoises layers known to be complete.
__device__ __forceinline__ void wait_tiles(const MegaParams& P, int d, int lo, int hi, unsigned long long& done_mask) {
    if ((done_mask >> d) & 1ull) return;
    if (P.debug & 1) return;
    const MegaLayer& D = P.L[d];
    if (ld_relaxed(P.layer_done + d) >= unsigned(D.num_m_tiles)) {
        done_mask |= 1ull << d;
    } else {
        const unsigned target = unsigned(D.num_n_tiles) * 4u;
        const unsigned* f = P.flags + D.flag_base;
        const long long t0 = clock64();
        for (int j = lo; j <= hi; ++j) {
            while (ld_relaxed(f + j) < target) {
                if (clock64() - t0 > kSpinLimitCycles) __trap();
            }
        }
    }
    fence_acq_rel_gpu();
    fence_proxy_async_all();  // the data is read through the async proxy (TMA) next
}

__global__ void __launch_bounds__(kThreads, 1) conv_mega_kernel(const __grid_constant__ MegaParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) __trap();
    const uint32_t epi_base = smem_base + kStages * kStageBytes;
    const uint32_t bias_base = epi_base + kEpilogueWarps * kEpiBufBytes;
    const uint32_t bar_base = bias_base + kEpilogueWarps * kBiasSlotBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int G = int(gridDim.x);

    if (warp == 1) {
        if (lane < kStages) {
            mbar_init(full_bar(lane), 2);
            mbar_init(empty_bar(lane), 1);
        } else if (lane < kStages + 2) {
            mbar_init(tmem_full_bar(lane - kStages), 1);
            mbar_init(tmem_empty_bar(lane - kStages), 128);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // The run reads tensors produced by earlier kernels of the stream (its external inputs) and overwrites buffers they may
    // still be reading: everything below waits for them.  (Weights are older still; the wait is cheap and happens once per run.)
    grid_dep_launch_dependents();
    grid_dep_wait();

    // item cursor: g = global item index, `layer` = the layer it belongs to (item_base <= g < item_end)
    auto seek = [&](int g, int& layer) {
        while (layer < P.num_layers && g >= P.L[layer].item_end) ++layer;
    };

    if (warp < kNumAProducers || warp >= kBProducerWarp0) {
        // ================= TMA producers =================
        const bool is_a = warp < kNumAProducers;
        const int me = is_a ? warp : warp - kBProducerWarp0;
        const int n_prod = is_a ? kNumAProducers : kNumBProducers;
        if (elect_one()) {
            unsigned long long done_mask = 0;
            uint32_t stage = uint32_t(me), phase = 0;
            uint32_t full_addr = bar_base + 8u * stage;
            uint32_t dst = smem_base + stage * kStageBytes + (is_a ? 0u : kABytes);
            int kb = me;
            int layer = 0;
            int g = int(blockIdx.x);
            seek(g, layer);
            int cur_layer = -1;
            while (layer < P.num_layers) {
                const MegaLayer& L = P.L[layer];
                if (layer != cur_layer) {
                    cur_layer = layer;
                    prefetch_tensormap(is_a ? &P.tm_a[layer] : &P.tm_b[layer]);
                }
                const int num_kb = L.total_kb;
                if (kb < num_kb) {
                    const int nn = L.num_n_tiles;
                    const int tile = g - L.item_base;
                    const int m_tile = tile / nn;
                    const int n_tile = tile - m_tile * nn;
                    const int m0 = m_tile * kBlockM;
                    const int kpt = L.kblocks_per_tap, taps_w = L.taps_w;
                    int cblk = 0, tap = 0, fs = 0, fr = 0;
                    if (kb < 8) {
                        cblk = kb;
                        while (cblk >= kpt) { cblk -= kpt; ++tap; ++fs; }
                        while (fs >= taps_w) { fs -= taps_w; ++fr; }
                    } else {
                        tap = kb / kpt; cblk = kb - tap * kpt;
                        fr = tap / taps_w; fs = tap - fr * taps_w;
                    }
                    if (is_a) {
                        int img = 0, base_h = 0, base_w = 0;
                        if (L.mode != CONV_MODE_TILED) {
                            img = m0 / L.PQ;
                            const int rem = m0 - img * L.PQ;
                            const int op = rem / L.Q;
                            const int oq = rem - op * L.Q;
                            base_h = L.corner_h + op * L.stride_h;
                            base_w = L.corner_w + oq * L.stride_w;
                        }
                        if (L.dep >= 0) {
                            int lo = m_tile, hi = m_tile;
                            if (L.mode != CONV_MODE_TILED) {
                                // input pixels the tile's receptive field can touch: whole rows from the first output pixel's top row
                                // to the last output pixel's bottom row (a superset across image borders)
                                const int m_last = min(m0 + kBlockM, L.M) - 1;
                                const int img1 = m_last / L.PQ;
                                const int op1 = (m_last - img1 * L.PQ) / L.Q;
                                const int ih_lo = max(base_h, 0);
                                const int ih_hi = min(L.corner_h + op1 * L.stride_h + (L.taps_h - 1) * L.dil_h, L.in_h - 1);
                                lo = ((img * L.in_h + ih_lo) * L.in_w) / kBlockM;
                                hi = ((img1 * L.in_h + ih_hi) * L.in_w + L.in_w - 1) / kBlockM;
                            }
                            wait_tiles(P, L.dep, lo, hi, done_mask);
                        }
                        const CUtensorMap* tm = &P.tm_a[layer];
                        if (L.mode == CONV_MODE_TILED) {
#pragma unroll 1
                            for (; kb < num_kb; kb += kNumAProducers) {
                                mbar_wait_bounded(full_addr + 8u * kStages, phase ^ 1u);
                                mbar_expect_tx(full_addr, kABytes);
                                tma_load_2d(tm, full_addr, dst, kb * kBlockK, m0);
                                stage += kNumAProducers; full_addr += 8u * kNumAProducers; dst += kNumAProducers * kStageBytes;
                                if (stage >= uint32_t(kStages)) { stage -= kStages; phase ^= 1u; full_addr -= 8u * kStages; dst -= kStages * kStageBytes; }
                            }
                        } else {
                            const int dil_w = L.dil_w, dil_h = L.dil_h;
#pragma unroll 1
                            for (; kb < num_kb; kb += kNumAProducers) {
                                mbar_wait_bounded(full_addr + 8u * kStages, phase ^ 1u);
                                mbar_expect_tx(full_addr, kABytes);
                                tma_load_im2col_4d(tm, full_addr, dst, cblk * kBlockK, base_w, base_h, img, uint16_t(fs * dil_w), uint16_t(fr * dil_h));
                                stage += kNumAProducers; full_addr += 8u * kNumAProducers; dst += kNumAProducers * kStageBytes;
                                if (stage >= uint32_t(kStages)) { stage -= kStages; phase ^= 1u; full_addr -= 8u * kStages; dst -= kStages * kStageBytes; }
                                cblk += kNumAProducers;
                                while (cblk >= kpt) { cblk -= kpt; ++fs; }
                                while (fs >= taps_w) { fs -= taps_w; ++fr; }
                            }
                        }
                    } else {
                        const CUtensorMap* tm = &P.tm_b[layer];
                        const int n0 = n_tile * L.block_n;
                        const uint32_t b_bytes = uint32_t(L.block_n) * kBlockK * 2;
#pragma unroll 1
                        for (; kb < num_kb; kb += kNumBProducers) {
                            mbar_wait_bounded(full_addr + 8u * kStages, phase ^ 1u);
                            mbar_expect_tx(full_addr, b_bytes);
                            tma_load_3d(tm, full_addr, dst, cblk * kBlockK, tap, n0);
                            stage += kNumBProducers; full_addr += 8u * kNumBProducers; dst += kNumBProducers * kStageBytes;
                            if (stage >= uint32_t(kStages)) { stage -= kStages; phase ^= 1u; full_addr -= 8u * kStages; dst -= kStages * kStageBytes; }
                            cblk += kNumBProducers;
                            while (cblk >= kpt) { cblk -= kpt; ++tap; }
                        }
                    }
                }
                kb -= num_kb;
                g += G;
                seek(g, layer);
            }
            (void)n_prod;
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        constexpr uint64_t desc_hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
        constexpr uint32_t desc_lbo = 1u << 16;
        const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | desc_lbo;
        constexpr uint32_t kStage16 = kStageBytes >> 4;
        constexpr uint32_t kB16 = kABytes >> 4;
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            uint32_t a_lo = a_lo0;
            uint32_t full_addr = bar_base;
            int layer = 0;
            int g = int(blockIdx.x);
            seek(g, layer);
            uint32_t t = 0;  // items this CTA has started
            while (layer < P.num_layers) {
                const MegaLayer& L = P.L[layer];
                const uint32_t idesc = make_idesc_f16(kBlockM, uint32_t(L.block_n));
                const int num_kb = L.total_kb;
                const uint32_t acc = t & 1u;
                mbar_wait_bounded(tmem_empty_bar(int(acc)), ((t >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * 128u;
                auto kblock = [&](uint32_t first_accumulate) {
                    mbar_wait_bounded(full_addr, phase);
                    tc_fence_after();
                    const uint64_t a_desc = desc_hi | uint64_t(a_lo);
                    const uint64_t b_desc = a_desc + kB16;
                    umma_f16(tmem_d, a_desc, b_desc, idesc, first_accumulate);
                    umma_f16(tmem_d, a_desc + 2, b_desc + 2, idesc, 1u);
                    umma_f16(tmem_d, a_desc + 4, b_desc + 4, idesc, 1u);
                    umma_f16(tmem_d, a_desc + 6, b_desc + 6, idesc, 1u);
                    umma_commit(full_addr + 8u * kStages);
                    a_lo += kStage16;
                    full_addr += 8u;
                    if (++stage == uint32_t(kStages)) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_addr = bar_base; }
                };
                kblock(0u);
                int kb = 1;
#pragma unroll 1
                for (; kb + 1 < num_kb; kb += 2) {
                    kblock(1u);
                    kblock(1u);
                }
                if (kb < num_kb) kblock(1u);
                umma_commit(tmem_full_bar(int(acc)));
                ++t;
                g += G;
                seek(g, layer);
            }
        }
    } else {
        // ================= epilogue (warps 4..11) =================
        const int ewarp = warp - kEpilogueWarp0;
        const int group = ewarp >> 2;
        const int ew = ewarp & 3;
        const uint32_t buf = epi_base + uint32_t(ewarp) * kEpiBufBytes;
        const uint32_t bias_slot = bias_base + uint32_t(ewarp) * kBiasSlotBytes;
        const uint32_t rowbuf = buf + uint32_t(lane) * 128u;
        const uint32_t sw = uint32_t(lane & 7);
        unsigned long long done_mask = 0;
        // Deferred completion signal of this warp's previous tile (warp-uniform state; lane 0, which owns the bulk store groups,
        // does the work).  The counter bump's return value is consumed one tile later, so its latency is off the critical path.
        bool pending = false;
        int pend_layer = 0, pend_mtile = 0;
        bool have_old = false;
        unsigned old_val = 0, old_target = 0;
        int old_layer = 0;
        auto signal = [&](int wait_keep) {  // the previous tile's stores are complete -> publish
            if (lane == 0) {
                if (wait_keep == 0) tma_store_wait<0>(); else tma_store_wait<1>();
                if (have_old && old_val + 1u == old_target) atomicAdd(P.layer_done + old_layer, 1u);
                if (!(P.debug & 4)) {
                    fence_proxy_async_all();
                    __threadfence();
                }
                const MegaLayer& S = P.L[pend_layer];
                old_val = atomicAdd(P.flags + S.flag_base + pend_mtile, 1u);
                old_target = unsigned(S.num_n_tiles) * 4u;
                old_layer = pend_layer;
                have_old = true;
            }
            pending = false;
        };

        int layer = 0;
        int g = int(blockIdx.x) + group * G;  // this group drains the CTA's items group, group + 2, ...
        seek(g, layer);
        uint32_t acc_phase = 0;
        while (layer < P.num_layers) {
            const MegaLayer& L = P.L[layer];
            const int nn = L.num_n_tiles;
            const int tile = g - L.item_base;
            const int m_tile = tile / nn;
            const int n_tile = tile - m_tile * nn;
            const int m_row0 = m_tile * kBlockM + ew * 32;
            const int chunks = L.block_n >> 6;
            const bool is_sigmoid = L.act == ACT_SIGMOID;
            const __half2 lo2 = __float2half2_rn(L.act == ACT_RELU ? 0.f : (L.act == ACT_CLIP ? L.clip_lo : -INFINITY));
            const __half2 hi2 = __float2half2_rn(L.act == ACT_CLIP ? L.clip_hi : INFINITY);
            const __half* res = L.residual;
            const int row = m_row0 + lane;
            const bool row_ok = row < L.M;
            // the previous tile's signal must never wait behind something that (transitively) needs it
            if (res && L.res_dep >= 0) {
                if (pending && !((done_mask >> L.res_dep) & 1ull)) signal(0);
                wait_tiles(P, L.res_dep, m_tile, m_tile, done_mask);
            }
            if (pending) {
                int ready = 1;
                if (lane == 0) ready = mbar_try_wait(tmem_full_bar(group), acc_phase) ? 1 : 0;
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (!ready) signal(0);
            }
            mbar_wait_bounded(tmem_full_bar(group), acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(group) * 128u;
            const __half* res_row = res ? res + size_t(row_ok ? row : 0) * L.out_pitch : nullptr;
            auto load_res = [&](uint4 (&r)[4], int col) {  // 32 columns of this lane's residual row, straight from L2
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    r[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (row_ok && col + q * 8 < L.out_pitch && !(P.debug & 2)) r[q] = ld_cg_v4(res_row + col + q * 8);
                }
            };
            // one 32-column half of a chunk: accumulators + bias (+ residual) -> activation -> fp16 -> staging row
            auto half = [&](uint32_t tcol, const uint4 (&r)[4], int h, bool release) {
                uint32_t v[32];
                tmem_ld_32(taddr + tcol, v);
                tmem_ld_wait();
                if (release) {  // the accumulator is in registers: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tmem_empty_bar(group));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t bo = bias_slot + uint32_t(h * 4 + q) * 32u;
                    const uint4 b0 = ld_shared_v4(bo), b1 = ld_shared_v4(bo + 16u);
                    float f[8];
                    f[0] = __uint_as_float(v[q * 8 + 0]) + __uint_as_float(b0.x); f[1] = __uint_as_float(v[q * 8 + 1]) + __uint_as_float(b0.y);
                    f[2] = __uint_as_float(v[q * 8 + 2]) + __uint_as_float(b0.z); f[3] = __uint_as_float(v[q * 8 + 3]) + __uint_as_float(b0.w);
                    f[4] = __uint_as_float(v[q * 8 + 4]) + __uint_as_float(b1.x); f[5] = __uint_as_float(v[q * 8 + 5]) + __uint_as_float(b1.y);
                    f[6] = __uint_as_float(v[q * 8 + 6]) + __uint_as_float(b1.z); f[7] = __uint_as_float(v[q * 8 + 7]) + __uint_as_float(b1.w);
                    if (res) {
                        const __half2* rh = reinterpret_cast<const __half2*>(&r[q]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 r2 = __half22float2(rh[i]);
                            f[2 * i] += r2.x;
                            f[2 * i + 1] += r2.y;
                        }
                    }
                    if (is_sigmoid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = sigmoid_slow(f[i]);
                    }
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) oh[i] = __hmin2(__hmax2(__floats2half2_rn(f[2 * i], f[2 * i + 1]), lo2), hi2);
                    st_shared_v4(rowbuf + ((uint32_t(h * 4 + q) ^ sw) << 4), o);
                }
            };
#pragma unroll 1
            for (int c = 0; c < chunks; ++c) {
                const int col0 = n_tile * L.block_n + c * kChunkN;
                // bias of this chunk -> the warp's smem slot (two columns per lane)
                const float2 bv = __ldg(reinterpret_cast<const float2*>(L.bias + col0) + lane);
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_slot + uint32_t(lane) * 8u), "f"(bv.x), "f"(bv.y) : "memory");
                uint4 r0[4], r1[4];
                if (res) load_res(r0, col0);
                if (lane == 0) tma_store_wait_read<0>();  // the previous store has read the staging buffer
                __syncwarp();                             // ... and the bias slot is published
                if (res) load_res(r1, col0 + 32);
                half(uint32_t(c * kChunkN), r0, 0, false);
                half(uint32_t(c * kChunkN + 32), r1, 1, c == chunks - 1);
                fence_proxy_async_smem();
                __syncwarp();  // also: every lane is done with the bias slot before the next chunk publishes its own
                if (lane == 0) {
                    tma_store_2d(&P.tm_out[layer], buf, col0, m_row0);
                    tma_store_commit();
                }
                if (c == 0 && pending) signal(1);  // every store group but the one just committed is complete
            }
            pending = true; pend_layer = layer; pend_mtile = m_tile;
            g += 2 * G;
            seek(g, layer);
        }
        if (pending) signal(0);
        if (lane == 0 && have_old && old_val + 1u == old_target) atomicAdd(P.layer_done + old_layer, 1u);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

int conv_mega_pick_block_n(int c_out, int m_tiles, int num_sms) {
    if (c_out <= 64) return 64;
    const long tiles128 = long(m_tiles) * ((c_out + 127) / 128);
    return tiles128 >= num_sms ? 128 : 64;
}

bool conv_mega_prepare(MegaLaunch* out, const std::vector<ConvTcProblem>& layers, const std::vector<int>& dep, const std::vector<int>& res_dep,
                       int num_sms, unsigned int* sync_words, std::string* err) {
    const int n = int(layers.size());
    if (n < 1 || n > kMegaMaxLayers) { if (err) *err = "conv_mega: layer count out of range"; return false; }
    out->params = std::make_unique<MegaParams>();
    MegaParams& P = *out->params;
    memset(&P, 0, sizeof(P));
    P.num_layers = n;
    { const char* dbg = getenv("SMELTER_MEGA_DEBUG"); P.debug = dbg ? atoi(dbg) : 0; }
    int item = 0, flag = 0;
    double flops = 0;
    for (int i = 0; i < n; ++i) {
        ConvTcProblem q = layers[size_t(i)];
        const int R = q.k_h, S = q.k_w;
        const int Pp = (q.h + q.pad_t + q.pad_b - q.dil_h * (R - 1) - 1) / q.stride_h + 1;
        const int Qq = (q.w + q.pad_l + q.pad_r - q.dil_w * (S - 1) - 1) / q.stride_w + 1;
        const long M = long(q.n) * std::max(Pp, 0) * std::max(Qq, 0);
        q.block_n = conv_mega_pick_block_n(q.c_out, int((M + kBlockM - 1) / kBlockM), num_sms);
        q.splits = 1;
        const __half* residual = q.residual;
        q.residual = nullptr;  // the mega kernel reads the residual itself
        ConvTcLaunch one;
        if (!conv_tc_prepare(&one, q, num_sms, err)) return false;
        MegaLayer& L = P.L[i];
        L.M = one.p.M; L.num_m_tiles = one.p.num_m_tiles; L.num_n_tiles = one.p.num_n_tiles; L.block_n = one.block_n;
        L.total_kb = one.p.num_taps * one.p.kblocks_per_tap; L.kblocks_per_tap = one.p.kblocks_per_tap; L.taps_w = one.p.taps_w;
        L.taps_h = one.p.num_taps / one.p.taps_w;
        L.mode = one.p.mode; L.PQ = one.p.PQ; L.Q = one.p.Q; L.stride_h = one.p.stride_h; L.stride_w = one.p.stride_w;
        L.dil_h = one.p.dil_h; L.dil_w = one.p.dil_w; L.corner_h = one.p.corner_h; L.corner_w = one.p.corner_w;
        L.in_h = q.h; L.in_w = q.w; L.out_pitch = one.p.out_pitch; L.act = one.p.act; L.clip_lo = one.p.clip_lo; L.clip_hi = one.p.clip_hi;
        L.dep = dep[size_t(i)]; L.res_dep = residual ? res_dep[size_t(i)] : -1;
        L.item_base = item; item += L.num_m_tiles * L.num_n_tiles; L.item_end = item;
        L.flag_base = flag; flag += L.num_m_tiles;
        L.bias = q.bias; L.residual = residual;
        if (L.mode == CONV_MODE_PACKED_ROW && L.dep >= 0) { if (err) *err = "conv_mega: packed-row layers must read an external tensor"; return false; }
        if (L.dep >= i || L.res_dep >= i) { if (err) *err = "conv_mega: dependencies must point to earlier layers"; return false; }
        P.tm_a[i] = one.tm_a; P.tm_b[i] = one.tm_b; P.tm_out[i] = one.tm_out;
        flops += one.flops;
    }
    P.total_items = item;
    P.layer_done = sync_words;
    P.flags = sync_words + kMegaMaxLayers;
    out->sync_words = size_t(kMegaMaxLayers) + size_t(flag);
    out->grid = std::min(item, num_sms);
    out->flops = flops;
    cudaError_t e = cudaFuncSetAttribute(conv_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(mega): ") + cudaGetErrorString(e); return false; }
    return true;
}

size_t conv_mega_sync_words(const std::vector<ConvTcProblem>& layers) {
    size_t words = kMegaMaxLayers;
    for (const ConvTcProblem& q : layers) {
        const int Pp = (q.h + q.pad_t + q.pad_b - q.dil_h * (q.k_h - 1) - 1) / q.stride_h + 1;
        const int Qq = (q.w + q.pad_l + q.pad_r - q.dil_w * (q.k_w - 1) - 1) / q.stride_w + 1;
        const long M = long(q.n) * std::max(Pp, 0) * std::max(Qq, 0);
        words += size_t((M + kBlockM - 1) / kBlockM);
    }
    return words;
}

cudaError_t conv_mega_launch(const MegaLaunch& L, cudaStream_t stream) {
    // completion counters start from zero on every launch (one small memset node in the captured graph)
    cudaError_t e = cudaMemsetAsync(L.params->layer_done, 0, L.sync_words * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(L.grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 0;  // follows a memset node: plain stream order
    return cudaLaunchKernelEx(&cfg, conv_mega_kernel, *L.params);
}

}  // namespace k
}  // namespace smelter
