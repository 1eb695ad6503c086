// Persistent multi-layer ("mega") variant of the tcgen05 implicit-GEMM convolution.
//
// A run of consecutive Conv nodes of the graph walk (reference Sources/Smelter/ONNXGraph.swift:169-193 emits one MPS node per
// ONNX node; MPSNNGraph then schedules them, closed source) is executed by ONE kernel launch: 148 persistent CTAs walk the
// concatenated list of output tiles of all layers of the run (global round-robin, so a layer whose tile count is not a
// multiple of the SM count does not leave SMs idle), and a tile starts as soon as the input tiles it reads are complete:
//
//   * every finished output tile bumps a per-(layer, m-tile) counter in global memory (release) after its TMA stores have
//     completed.  A scout warp per CTA walks the CTA's item list ahead of the producers, polls the counters of the m-tiles an
//     item's im2col window (and residual) covers, and publishes "items 0..t are ready" in shared memory; the activation
//     producers only look at that word, so the L2 round trips of the polling are off the load path.
//   * no grid-wide barrier and no kernel boundary between layers: the pipeline-fill / drain / launch cost (~4 us per layer on
//     B200, more than the tensor time of most ResNet-50 layers at batch 32) is paid once per run; weight (B) tiles of the next
//     layer are prefetched while the current layer is still in its last tiles because the operand ring is shared by all layers.
//   * the residual Add is folded into the k-loop: D += Res * I, i.e. one extra k-block per 64 output columns whose A operand is
//     the residual tile (plain 2-D TMA) and whose B operand is a 64 x 64 identity (rows outside the identity tensor are
//     zero-filled by TMA, which shifts the diagonal to the right column block).  fp16 x 1.0 is exact, the add happens in the
//     fp32 accumulator, and the epilogue is the same for every layer.
//   * completion signals are batched: an epilogue warp publishes finished tiles when it is about to wait for an accumulator
//     anyway, or when four are pending — one gpu-scope fence per batch instead of per tile (measured ~1 us each).
//   * deadlock freedom: all CTAs are co-resident (grid <= #SMs, 1 CTA/SM), every CTA visits its items in increasing global
//     order, an item only depends on items of earlier layers (smaller global indices), and a warp always publishes what it has
//     finished before it blocks.
//   * every layer output of a run has its own buffer (engine.cc: no arena reuse inside a run), so the only hazard is RAW.
//
// Tile geometry is uniform so the ring never drains between layers: 128 x {64,128} x 64, six 32 KiB stages, fp32 accumulators
// double buffered in TMEM (2 x 128 columns).  Warps: 0-3 activation producers, 4-11 epilogue (two groups), 12 MMA issuer,
// 13-14 weight producers, 15 scout.
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
constexpr int kNumBProducers = 2;
constexpr int kBProducerWarp0 = 13;
constexpr int kScoutWarp = 15;
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
constexpr int kMaxPending = 4;                         // finished tiles an epilogue warp may hold back before it must publish
constexpr long long kSpinLimitCycles = 4000000000LL;   // ~2 s: a dependency that never arrives traps instead of hanging the GPU
// barrier region (bytes from bar_base)
constexpr uint32_t kOffFull = 0, kOffEmpty = 8 * kStages, kOffTmemFull = 16 * kStages, kOffTmemEmpty = 16 * kStages + 16;
constexpr uint32_t kOffTmemSlot = 16 * kStages + 32, kOffReady = kOffTmemSlot + 8, kOffPending = 192;
static_assert(kOffReady + 4 <= kOffPending && kOffPending + kEpilogueWarps * kMaxPending * 8 <= kBarrierBytes, "barrier region layout");

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add(unsigned* p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_shared_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_volatile(uint32_t addr, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimitCycles) __trap();
    }
}
#ifndef SMELTER_CONV_INSTRUMENT
#define SMELTER_CONV_INSTRUMENT 0
#endif
constexpr bool kInstr = SMELTER_CONV_INSTRUMENT != 0;
// instrumented builds: per layer, the earliest first-load time and the latest accumulator-ready time (ns, %globaltimer)
__device__ __forceinline__ void stamp_layer(const MegaParams& P, int layer, int which) {
    if (kInstr && P.timeline) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (which == 0) atomicMin(P.timeline + 2 * layer, t);
        else atomicMax(P.timeline + 2 * layer + 1, t);
    }
}
__device__ __noinline__ float sigmoid_slow(float v) { return 1.f / (1.f + __expf(-v)); }

__global__ void __launch_bounds__(kThreads, 1) conv_mega_kernel(const __grid_constant__ MegaParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) __trap();
    const uint32_t epi_base = smem_base + kStages * kStageBytes;
    const uint32_t bias_base = epi_base + kEpilogueWarps * kEpiBufBytes;
    const uint32_t bar_base = bias_base + kEpilogueWarps * kBiasSlotBytes;
    auto full_bar = [&](int s) { return bar_base + kOffFull + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + kOffEmpty + 8u * s; };
    auto tmem_full_bar = [&](int a) { return bar_base + kOffTmemFull + 8u * a; };
    auto tmem_empty_bar = [&](int a) { return bar_base + kOffTmemEmpty + 8u * a; };
    const uint32_t tmem_slot = bar_base + kOffTmemSlot;
    const uint32_t ready_addr = bar_base + kOffReady;  // items of this CTA (in visiting order) whose inputs are complete
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
        } else if (lane == 31) {
            st_shared_volatile(ready_addr, 0u);
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

    // item cursor: g = global item index, `layer` = the layer it belongs to (item_base <= g < item_end)
    auto seek = [&](int g, int& layer) {
        while (layer < P.num_layers && g >= P.L[layer].item_end) ++layer;
    };

    if (warp < kNumAProducers) {
        // ================= activation (A) producers =================
        // Producer i issues the loads of k-blocks i, i + 4, ... of the CTA's k-block stream.  Positions advance by adds and
        // compares only (see conv_igemm.cu produce()); the first load of an item waits for the scout's ready word.
        if (elect_one()) {
            uint32_t stage = uint32_t(warp), phase = 0;
            uint32_t full_addr = bar_base + kOffFull + 8u * stage;
            uint32_t dst = smem_base + stage * kStageBytes;
            int kb = warp;
            int layer = 0;
            int g = int(blockIdx.x);
            seek(g, layer);
            uint32_t t = 0;      // ordinal of the current item in this CTA's visiting order
            uint32_t ready = 0;  // last value read from the scout's word
            int cur_layer = -1;
            while (layer < P.num_layers) {
                const MegaLayer& L = P.L[layer];
                if (layer != cur_layer) {
                    cur_layer = layer;
                    prefetch_tensormap(&P.tm_a[layer]);
                    if (L.res_map >= 0) prefetch_tensormap(&P.tm_res[L.res_map]);
                }
                const int num_kb = L.total_kb;
                if (kb < num_kb) {
                    if (ready <= t) {
                        const long long t0 = clock64();
                        while ((ready = ld_shared_volatile(ready_addr)) <= t) {
                            if (clock64() - t0 > kSpinLimitCycles) __trap();
                        }
                        fence_proxy_async_all();  // the scout's acquire is ordered before the TMA reads issued below
                    }
                    const int nn = L.num_n_tiles;
                    const int tile = g - L.item_base;
                    const int m_tile = tile / nn;
                    const int n_tile = tile - m_tile * nn;
                    const int m0 = m_tile * kBlockM;
                    const int main_kb = L.main_kb;
                    const int kpt = L.kblocks_per_tap, taps_w = L.taps_w;
                    const bool tiled = L.mode == CONV_MODE_TILED;
                    int cblk = 0, fs = 0, fr = 0;
                    int img = 0, base_h = 0, base_w = 0;
                    if (!tiled) {
                        if (kb < 8) {
                            cblk = kb;
                            while (cblk >= kpt) { cblk -= kpt; ++fs; }
                            while (fs >= taps_w) { fs -= taps_w; ++fr; }
                        } else {
                            const int tap = kb / kpt;
                            cblk = kb - tap * kpt;
                            fr = tap / taps_w; fs = tap - fr * taps_w;
                        }
                        img = m0 / L.PQ;
                        const int rem = m0 - img * L.PQ;
                        const int op = rem / L.Q;
                        const int oq = rem - op * L.Q;
                        base_h = L.corner_h + op * L.stride_h;
                        base_w = L.corner_w + oq * L.stride_w;
                    }
                    const CUtensorMap* tm = &P.tm_a[layer];
                    const int dil_w = L.dil_w, dil_h = L.dil_h;
                    if (warp == 0) stamp_layer(P, layer, 0);
#pragma unroll 1
                    for (; kb < num_kb; kb += kNumAProducers) {
                        mbar_wait_bounded(full_addr + (kOffEmpty - kOffFull), phase ^ 1u);
                        if (P.debug & 32) { mbar_arrive(full_addr); } else {
                        mbar_expect_tx(full_addr, kABytes);
                        if (kb < main_kb) {
                            if (tiled) tma_load_2d(tm, full_addr, dst, kb * kBlockK, m0);
                            else tma_load_im2col_4d(tm, full_addr, dst, cblk * kBlockK, base_w, base_h, img, uint16_t(fs * dil_w), uint16_t(fr * dil_h));
                        } else {
                            // residual k-block: 64 columns of the residual tile as the A operand (four 32-row boxes)
                            const CUtensorMap* tr = &P.tm_res[L.res_map];
                            const int c0 = n_tile * L.block_n + (kb - main_kb) * kBlockK;
#pragma unroll
                            for (int i = 0; i < 4; ++i) tma_load_2d(tr, full_addr, dst + uint32_t(i) * 4096u, c0, m0 + 32 * i);
                        }
                        }
                        stage += kNumAProducers; full_addr += 8u * kNumAProducers; dst += kNumAProducers * kStageBytes;
                        if (stage >= uint32_t(kStages)) { stage -= kStages; phase ^= 1u; full_addr -= 8u * kStages; dst -= kStages * kStageBytes; }
                        if (!tiled) {
                            cblk += kNumAProducers;
                            while (cblk >= kpt) { cblk -= kpt; ++fs; }
                            while (fs >= taps_w) { fs -= taps_w; ++fr; }
                        }
                    }
                }
                kb -= num_kb;
                g += G;
                ++t;
                seek(g, layer);
            }
        }
    } else if (warp >= kBProducerWarp0 && warp < kBProducerWarp0 + kNumBProducers) {
        // ================= weight (B) producers =================
        // No dependency on anything computed in this launch: they run ahead as far as the ring allows, also across layers.
        if (elect_one()) {
            const int me = warp - kBProducerWarp0;
            uint32_t stage = uint32_t(me), phase = 0;
            uint32_t full_addr = bar_base + kOffFull + 8u * stage;
            uint32_t dst = smem_base + stage * kStageBytes + kABytes;
            int kb = me;
            int layer = 0;
            int g = int(blockIdx.x);
            seek(g, layer);
            int cur_layer = -1;
            while (layer < P.num_layers) {
                const MegaLayer& L = P.L[layer];
                if (layer != cur_layer) {
                    cur_layer = layer;
                    prefetch_tensormap(&P.tm_b[layer]);
                }
                const int num_kb = L.total_kb;
                if (kb < num_kb) {
                    const int nn = L.num_n_tiles;
                    const int tile = g - L.item_base;
                    const int n_tile = tile % nn;
                    const int n0 = n_tile * L.block_n;
                    const int main_kb = L.main_kb;
                    const int kpt = L.kblocks_per_tap;
                    const uint32_t b_bytes = uint32_t(L.block_n) * kBlockK * 2;
                    const CUtensorMap* tm = &P.tm_b[layer];
                    const CUtensorMap* ti = &P.tm_ident[L.block_n == 128 ? 1 : 0];
                    int cblk = 0, tap = 0;
                    if (kb < 8) {
                        cblk = kb;
                        while (cblk >= kpt) { cblk -= kpt; ++tap; }
                    } else {
                        tap = kb / kpt;
                        cblk = kb - tap * kpt;
                    }
#pragma unroll 1
                    for (; kb < num_kb; kb += kNumBProducers) {
                        mbar_wait_bounded(full_addr + (kOffEmpty - kOffFull), phase ^ 1u);
                        if (P.debug & 16) { mbar_arrive(full_addr); } else {
                        mbar_expect_tx(full_addr, b_bytes);
                        if (kb < main_kb) tma_load_3d(tm, full_addr, dst, cblk * kBlockK, tap, n0);
                        else tma_load_2d(ti, full_addr, dst, 0, -kBlockK * (kb - main_kb));  // identity shifted to column block kb - main_kb
                        }
                        stage += kNumBProducers; full_addr += 8u * kNumBProducers; dst += kNumBProducers * kStageBytes;
                        if (stage >= uint32_t(kStages)) { stage -= kStages; phase ^= 1u; full_addr -= 8u * kStages; dst -= kStages * kStageBytes; }
                        cblk += kNumBProducers;
                        while (cblk >= kpt) { cblk -= kpt; ++tap; }
                    }
                }
                kb -= num_kb;
                g += G;
                seek(g, layer);
            }
        }
    } else if (warp == kScoutWarp) {
        // ================= scout =================
        // Walks this CTA's items in visiting order; for each, waits until the m-tiles its input window (and residual tile) needs
        // are complete — the 32 lanes poll up to 32 counters at a time — then publishes the item's ordinal.
        auto wait_range = [&](int d, int lo, int hi) {
            const MegaLayer& D = P.L[d];
            const unsigned target = unsigned(D.num_n_tiles) * 4u;
            const unsigned* f = P.flags + D.flag_base;
            for (int base = lo; base <= hi; base += 32) {
                const int j = base + lane;
                const long long t0 = clock64();
                bool ok = j > hi;
                while (true) {
                    if (!ok) ok = ld_acquire(f + j) >= target;
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (clock64() - t0 > kSpinLimitCycles) __trap();
                }
            }
        };
        int layer = 0;
        int g = int(blockIdx.x);
        seek(g, layer);
        uint32_t t = 0;
        while (layer < P.num_layers) {
            const MegaLayer& L = P.L[layer];
            if (!(P.debug & 1) && (L.dep >= 0 || (L.res_map >= 0 && L.res_dep >= 0))) {
                const int tile = g - L.item_base;
                const int m_tile = tile / L.num_n_tiles;
                if (L.dep >= 0) {
                    int lo = m_tile, hi = m_tile;
                    if (L.mode != CONV_MODE_TILED) {
                        // input pixels the tile's receptive field can touch: whole rows from the first output pixel's top row to the
                        // last output pixel's bottom row (a superset when the tile crosses image borders)
                        const int m0 = m_tile * kBlockM;
                        const int m_last = min(m0 + kBlockM, L.M) - 1;
                        const int img0 = m0 / L.PQ, img1 = m_last / L.PQ;
                        const int op0 = (m0 - img0 * L.PQ) / L.Q, op1 = (m_last - img1 * L.PQ) / L.Q;
                        const int ih_lo = max(L.corner_h + op0 * L.stride_h, 0);
                        const int ih_hi = min(L.corner_h + op1 * L.stride_h + (L.taps_h - 1) * L.dil_h, L.in_h - 1);
                        lo = ((img0 * L.in_h + ih_lo) * L.in_w) / kBlockM;
                        hi = ((img1 * L.in_h + ih_hi) * L.in_w + L.in_w - 1) / kBlockM;
                    }
                    wait_range(L.dep, lo, hi);
                }
                if (L.res_map >= 0 && L.res_dep >= 0) wait_range(L.res_dep, m_tile, m_tile);
            }
            ++t;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                st_shared_volatile(ready_addr, t);
            }
            g += G;
            seek(g, layer);
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
            uint32_t full_addr = bar_base + kOffFull;
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
                    umma_commit(full_addr + (kOffEmpty - kOffFull));
                    a_lo += kStage16;
                    full_addr += 8u;
                    if (++stage == uint32_t(kStages)) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_addr = bar_base + kOffFull; }
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
    } else if (warp >= kEpilogueWarp0 && warp < kEpilogueWarp0 + kEpilogueWarps) {
        // ================= epilogue (warps 4..11) =================
        const int ewarp = warp - kEpilogueWarp0;
        const int group = ewarp >> 2;
        const int ew = ewarp & 3;
        const uint32_t buf = epi_base + uint32_t(ewarp) * kEpiBufBytes;
        const uint32_t bias_slot = bias_base + uint32_t(ewarp) * kBiasSlotBytes;
        const uint32_t rowbuf = buf + uint32_t(lane) * 128u;
        const uint32_t sw = uint32_t(lane & 7);
        const uint32_t pend_addr = bar_base + kOffPending + uint32_t(ewarp) * (kMaxPending * 8u);  // lane 0's list of finished tiles
        int n_pend = 0;       // warp-uniform
        int last_chunks = 1;  // store groups of the newest pending tile
        // Publish finished tiles.  keep_last: the newest tile's stores were committed a moment ago — leave it pending rather than
        // wait for them.  Lane 0 owns the bulk store groups, so it does the work.
        auto publish = [&](bool keep_last) {
            const int n = keep_last ? n_pend - 1 : n_pend;
            if (lane == 0 && n > 0) {
                if (!keep_last) tma_store_wait<0>();
                else if (last_chunks == 1) tma_store_wait<1>();
                else tma_store_wait<2>();
                if (!(P.debug & 12)) __threadfence();  // release: the tiles' data (complete in L2) before the counters
                for (int i = 0; i < n; ++i) {
                    const uint32_t e_layer = ld_shared_volatile(pend_addr + uint32_t(i) * 8u);
                    const uint32_t e_mtile = ld_shared_volatile(pend_addr + uint32_t(i) * 8u + 4u);
                    if (P.debug & 8) red_add_release(P.flags + P.L[e_layer].flag_base + e_mtile, 1u);
                    else red_add(P.flags + P.L[e_layer].flag_base + e_mtile, 1u);
                }
                if (keep_last) {
                    st_shared_volatile(pend_addr, ld_shared_volatile(pend_addr + uint32_t(n) * 8u));
                    st_shared_volatile(pend_addr + 4u, ld_shared_volatile(pend_addr + uint32_t(n) * 8u + 4u));
                }
            }
            if (n > 0) n_pend = keep_last ? 1 : 0;
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
            // never block while holding finished tiles back: the accumulator we wait for may (transitively) need them
            if (n_pend > 0) {
                int ready = 1;
                if (lane == 0) ready = mbar_try_wait(tmem_full_bar(group), acc_phase) ? 1 : 0;
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (!ready) publish(false);
            }
            mbar_wait_bounded(tmem_full_bar(group), acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            if (ew == 0 && lane == 0) stamp_layer(P, layer, 1);
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(group) * 128u;
            // one 32-column half of a chunk: accumulators + bias -> activation -> fp16 -> staging row
            auto half = [&](uint32_t tcol, int h, bool release) {
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
                if (lane == 0) tma_store_wait_read<0>();  // the previous store has read the staging buffer
                __syncwarp();                             // ... and the bias slot is published
                half(uint32_t(c * kChunkN), 0, false);
                half(uint32_t(c * kChunkN + 32), 1, c == chunks - 1);
                fence_proxy_async_smem();
                __syncwarp();  // also: every lane is done with the bias slot before the next chunk publishes its own
                if (lane == 0) {
                    tma_store_2d(&P.tm_out[layer], buf, col0, m_row0);
                    tma_store_commit();
                }
            }
            if (lane == 0) {
                st_shared_volatile(pend_addr + uint32_t(n_pend) * 8u, uint32_t(layer));
                st_shared_volatile(pend_addr + uint32_t(n_pend) * 8u + 4u, uint32_t(m_tile));
            }
            ++n_pend;
            last_chunks = chunks;
            if (n_pend == kMaxPending) publish(true);
            g += 2 * G;
            seek(g, layer);
        }
        publish(false);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

int conv_mega_pick_block_n(int c_out, int m_tiles, int num_kb, int num_sms) {
    // Span of a layer ~ waves x (time of one tile) once the first operands are in: the k-loop costs ~0.08 / 0.13 us per k-block
    // for N = 64 / 128 (measured, B200), the epilogue ~0.35 us per 64 columns, whichever is longer.  In the small late layers a
    // partial second wave doubles the span (196 tiles of N=64 on 148 SMs are slower than 98 tiles of N=128); in the large
    // early layers the two are close and N=128 reads the activations half as often.
    if (c_out <= 64) return 64;
    auto span = [&](int bn, double t_kb, double t_epi) {
        const long tiles = long(m_tiles) * ((c_out + bn - 1) / bn);
        const long waves = (tiles + num_sms - 1) / num_sms;
        return double(waves) * std::max(num_kb * t_kb, t_epi);
    };
    const double s64 = span(64, 0.08, 0.35) * 1.1, s128 = span(128, 0.13, 0.7);
    return s128 <= s64 ? 128 : 64;
}

bool conv_mega_prepare(MegaLaunch* out, const std::vector<ConvTcProblem>& layers, const std::vector<int>& dep, const std::vector<int>& res_dep,
                       int num_sms, unsigned int* sync_words, const __half* identity, std::string* err) {
    const int n = int(layers.size());
    if (n < 1 || n > kMegaMaxLayers) { if (err) *err = "conv_mega: layer count out of range"; return false; }
    out->params = std::make_unique<MegaParams>();
    MegaParams& P = *out->params;
    memset(&P, 0, sizeof(P));
    P.num_layers = n;
    { const char* dbg = getenv("SMELTER_MEGA_DEBUG"); P.debug = dbg ? atoi(dbg) : 0; }
    if (!conv_tc_encode_2d(&P.tm_ident[0], identity, 64, 64, 64, 64, err)) return false;
    if (!conv_tc_encode_2d(&P.tm_ident[1], identity, 64, 64, 64, 128, err)) return false;
    int item = 0, flag = 0, n_res = 0;
    double flops = 0;
    for (int i = 0; i < n; ++i) {
        ConvTcProblem q = layers[size_t(i)];
        const int R = q.k_h, S = q.k_w;
        const int Pp = (q.h + q.pad_t + q.pad_b - q.dil_h * (R - 1) - 1) / q.stride_h + 1;
        const int Qq = (q.w + q.pad_l + q.pad_r - q.dil_w * (S - 1) - 1) / q.stride_w + 1;
        const long M = long(q.n) * std::max(Pp, 0) * std::max(Qq, 0);
        const int kc = q.mode == CONV_MODE_PACKED_ROW ? S * q.c_in_pitch : q.c_in_pitch;
        const int taps = q.mode == CONV_MODE_TILED ? 1 : (q.mode == CONV_MODE_PACKED_ROW ? R : R * S);
        q.block_n = conv_mega_pick_block_n(q.c_out, int((M + kBlockM - 1) / kBlockM), taps * ((kc + kBlockK - 1) / kBlockK), num_sms);
        q.splits = 1;
        q.pair = -1;  // the persistent kernel loads whole weight tiles
        ConvTcLaunch one;
        if (!conv_tc_prepare(&one, q, num_sms, err)) return false;
        MegaLayer& L = P.L[i];
        L.M = one.p.M; L.num_m_tiles = one.p.num_m_tiles; L.num_n_tiles = one.p.num_n_tiles; L.block_n = one.block_n;
        L.main_kb = one.p.num_taps * one.p.kblocks_per_tap;
        L.kblocks_per_tap = one.p.kblocks_per_tap; L.taps_w = one.p.taps_w;
        L.taps_h = one.p.num_taps / one.p.taps_w;
        L.mode = one.p.mode; L.PQ = one.p.PQ; L.Q = one.p.Q; L.stride_h = one.p.stride_h; L.stride_w = one.p.stride_w;
        L.dil_h = one.p.dil_h; L.dil_w = one.p.dil_w; L.corner_h = one.p.corner_h; L.corner_w = one.p.corner_w;
        L.in_h = q.h; L.in_w = q.w; L.out_pitch = one.p.out_pitch; L.act = one.p.act; L.clip_lo = one.p.clip_lo; L.clip_hi = one.p.clip_hi;
        L.dep = dep[size_t(i)];
        L.res_dep = -1;
        L.res_map = -1;
        L.total_kb = L.main_kb;
        if (q.residual) {
            if (n_res == kMegaMaxResidual) { if (err) *err = "conv_mega: too many residual layers in one run"; return false; }
            L.res_dep = res_dep[size_t(i)];
            L.res_map = n_res;
            P.tm_res[n_res++] = one.tm_res;  // [32 row x 64 column] boxes over the residual tensor
            L.total_kb += one.block_n / kBlockK;
        }
        L.item_base = item; item += L.num_m_tiles * L.num_n_tiles; L.item_end = item;
        L.flag_base = flag; flag += L.num_m_tiles;
        L.bias = q.bias;
        if (L.mode == CONV_MODE_PACKED_ROW && L.dep >= 0) { if (err) *err = "conv_mega: packed-row layers must read an external tensor"; return false; }
        if (L.dep >= i || L.res_dep >= i) { if (err) *err = "conv_mega: dependencies must point to earlier layers"; return false; }
        P.tm_a[i] = one.tm_a; P.tm_b[i] = one.tm_b; P.tm_out[i] = one.tm_out;
        flops += one.flops;
    }
    P.total_items = item;
    P.flags = sync_words;
    out->sync_words = size_t(flag);
    out->grid = std::min(item, num_sms);
    out->flops = flops;
    cudaError_t e = cudaFuncSetAttribute(conv_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(mega): ") + cudaGetErrorString(e); return false; }
    return true;
}

size_t conv_mega_sync_words(const std::vector<ConvTcProblem>& layers) {
    size_t words = 0;
    for (const ConvTcProblem& q : layers) {
        const int Pp = (q.h + q.pad_t + q.pad_b - q.dil_h * (q.k_h - 1) - 1) / q.stride_h + 1;
        const int Qq = (q.w + q.pad_l + q.pad_r - q.dil_w * (q.k_w - 1) - 1) / q.stride_w + 1;
        const long M = long(q.n) * std::max(Pp, 0) * std::max(Qq, 0);
        words += size_t((M + kBlockM - 1) / kBlockM);
    }
    return std::max<size_t>(words, 1);
}

cudaError_t conv_mega_launch(const MegaLaunch& L, cudaStream_t stream) {
    // completion counters start from zero on every launch (one small memset node in the captured graph)
    cudaError_t e = cudaMemsetAsync(L.params->flags, 0, L.sync_words * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    if (kInstr && getenv("SMELTER_MEGA_TIMELINE")) {  // perf experiments, outside graph capture only
        static unsigned long long* buf = nullptr;
        if (!buf) cudaMalloc(&buf, 2 * kMegaMaxLayers * sizeof(unsigned long long));
        std::vector<unsigned long long> init(2 * kMegaMaxLayers);
        for (int i = 0; i < kMegaMaxLayers; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0; }
        cudaMemcpyAsync(buf, init.data(), init.size() * 8, cudaMemcpyHostToDevice, stream);
        cudaStreamSynchronize(stream);
        L.params->timeline = buf;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned(L.grid)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
        e = cudaLaunchKernelEx(&cfg, conv_mega_kernel, *L.params);
        cudaStreamSynchronize(stream);
        std::vector<unsigned long long> h(2 * kMegaMaxLayers);
        cudaMemcpy(h.data(), buf, h.size() * 8, cudaMemcpyDeviceToHost);
        const MegaParams& P = *L.params;
        for (int i = 0; i < P.num_layers; ++i)
            fprintf(stderr, "mega layer %2d: start %8.2f us  end %8.2f us  span %7.2f  tiles %5d bn %3d kb %3d (+%d res) mode %d\n", i,
                    (h[2 * i] - h[0]) * 1e-3, (h[2 * i + 1] - h[0]) * 1e-3, (h[2 * i + 1] - h[2 * i]) * 1e-3, P.L[i].item_end - P.L[i].item_base,
                    P.L[i].block_n, P.L[i].main_kb, P.L[i].total_kb - P.L[i].main_kb, P.L[i].mode);
        return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(L.grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;  // follows a memset node: plain stream order
    return cudaLaunchKernelEx(&cfg, conv_mega_kernel, *L.params);
}

}  // namespace k
}  // namespace smelter
