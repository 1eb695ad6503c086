// Two-CTA ("pair") variant of the tcgen05 implicit-GEMM convolution: a cluster of two CTAs on one TPC computes a 256 x BLOCK_N
// output tile with `tcgen05.mma.cta_group::2`.  Each CTA loads the activation rows of its own 128-pixel half and only HALF of the
// weight tile (BLOCK_N / 2 output channels); the tensor cores of both SMs read both halves.  Per SM and k-block that is
// 16 KiB + BLOCK_N/2 x 128 B instead of 16 KiB + BLOCK_N x 128 B — the k-loop of the single-CTA kernel is bound by the bytes a
// 192 KiB operand ring can keep in flight, so fewer bytes per FLOP is k-loop speed.
//
// Protocol (leader = cluster rank 0):
//   full[s]       lives in the leader: its A and its B producer each arrive.expect_tx the bytes of BOTH CTAs' loads of that operand;
//                 the TMA loads of both CTAs are the `.cta_group::2` forms and complete the leader's barrier.
//   empty[s]      one per CTA, arrived by the leader's `tcgen05.commit.cta_group::2 ... multicast::cluster` (mask 0b11).
//   tmem_full[a]  one per CTA, same multicast commit; each CTA's epilogue drains its own 128 TMEM lanes.
//   tmem_empty[a] lives in the leader: arrivals from the epilogue threads of both CTAs (remote arrive from the peer).
//   Only the leader's MMA warp issues MMAs; the peer's takes part in the cta_group::2 TMEM allocation only.
// Projection shortcut (ConvKernelParams::side_kb): after the main k-blocks of a tile, `side_kb` more k-blocks read a second activation
// tensor (tm_a2: 1x1, own stride, same output grid) against a second weight matrix (tm_b2) into the same accumulator -- the 1x1
// "downsample" convolution of a residual block is part of its block's last GEMM instead of a launch of its own.
// Instance-norm statistics (ConvKernelParams::stats, epilogue form EPI = 2): a layer whose only reader is an InstanceNormalization
// adds the column sums and sums of squares of every 128-row tile -- read back from the fp16 staging buffer of the TMA store -- to
// per-image fp64 accumulators, so the norm behind it is one pass over the tensor (pool_norm.cu inorm_apply_kernel<true, ..>).
// Everything else (operand modes, epilogue, fusion, PDL) is conv_igemm.cu's.  Default for 64/128/256-column tiles without split-K
// (same-box A/B on ResNet-50: -1.2 % step time at batch 32, +6..8 % images/s at batch 128/256); SMELTER_NO_PAIR=1 turns it off.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv_igemm.h"
#include "pair_ptx.cuh"
#include "ptx.cuh"

#ifndef SMELTER_MMA_LOOKAHEAD
#define SMELTER_MMA_LOOKAHEAD 1
#endif

namespace smelter {
namespace k {

using namespace ptx;
using namespace pairptx;

namespace {

constexpr int kBlockM = 128;  // rows per CTA; the pair's MMA is M = 256
constexpr int kBlockK = 64;
constexpr int kThreads = 512;
constexpr int kNumProducers = 4;
constexpr int kNumBProducers = 3;
constexpr int kBProducerWarp0 = 13;
constexpr int kEpilogueWarp0 = 4;
constexpr int kEpilogueWarps = 8;
constexpr int kMmaWarp = 12;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
constexpr int kChunkN = 64;
constexpr uint32_t kEpiBufBytes = 32 * kChunkN * 2;
constexpr uint32_t kBiasSlotBytes = kChunkN * 4;
constexpr uint32_t kBarrierBytes = 512;
constexpr uint32_t kSmemLimit = 227 * 1024;

// EPI: 0 = plain epilogue, 1 = residual tensor added (two staging buffers per warp), 2 = per-tile column statistics for the instance
// norm behind the layer (ConvKernelParams::stats)
template <int BLOCK_N, int EPI>
struct Cfg {
    static constexpr bool HAS_RES = EPI == 1;
    static constexpr uint32_t kStatBytes = EPI == 2 ? 8192u : 0u;  // [item parity][group][row warp][lane] float4
    static constexpr uint32_t kBBytes = (BLOCK_N / 2) * kBlockK * 2;  // this CTA's half of the weight tile
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBufs = HAS_RES ? 2 : 1;
    static constexpr uint32_t kEpiBytes = kEpilogueWarps * (kEpiBufs * kEpiBufBytes + kBiasSlotBytes) + kStatBytes;
    static constexpr int kStagesFit = int((kSmemLimit - kEpiBytes - kBarrierBytes) / kStageBytes);
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
    static constexpr int kProducers = kStages < kNumProducers ? kStages : kNumProducers;
    static constexpr int kBProducers = kStages < kNumBProducers ? kStages : kNumBProducers;
    static constexpr int kChunks = BLOCK_N / kChunkN;
    static constexpr size_t kSmemBytes = size_t(kStages) * kStageBytes + kEpiBytes + kBarrierBytes;
};

// Epilogue timeline stamps (perf experiments only, -DSMELTER_CONV_INSTRUMENT=1 builds with SMELTER_CONV_TIMELINE set): lane 0 of the
// first epilogue warp of CTA 0 writes %globaltimer at six points of its first eight work items into p.timeline[16 + 6 * item + point].
#ifndef SMELTER_CONV_INSTRUMENT
#define SMELTER_CONV_INSTRUMENT 0
#endif
constexpr bool kInstr = SMELTER_CONV_INSTRUMENT != 0;
__device__ __forceinline__ void epi_stamp(const ConvKernelParams& p, bool who, int item, int point) {
    if (kInstr && p.timeline && who && item < 8) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.timeline[16 + 6 * item + point] = t;
    }
}

// Launch-chain stamps (same builds, SMELTER_CHAIN_TIMELINE): min / max over all CTAs of %globaltimer at point `pt` of this launch.
__device__ __forceinline__ void chain_stamp(const ConvKernelParams& p, int pt) {
    if (kInstr && p.chain) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("red.global.min.u64 [%0], %1;" ::"l"(p.chain + 2 * pt), "l"(t) : "memory");
        asm volatile("red.global.max.u64 [%0], %1;" ::"l"(p.chain + 2 * pt + 1), "l"(t) : "memory");
    }
}

enum ProducerKind : int { PROD_A_TILED = 0, PROD_A_IM2COL = 1, PROD_B = 2 };

// One elected thread per producer warp (see conv_igemm.cu produce()).  Work items are (m-pair, n-tile): cluster c visits items
// c, c + #clusters, ...; this CTA's rows are m-tile 2 * pair + rank, its weight half is rows n0 + rank * BLOCK_N / 2.
template <int KIND, int BLOCK_N, int EPI>
__device__ __forceinline__ void produce(const CUtensorMap* tm, const CUtensorMap* tm_side, const ConvKernelParams& p, uint32_t smem_base, uint32_t bar_base,
                                        int me, int n_prod, int num_items, int main_kb, uint32_t rank) {
    using C = Cfg<BLOCK_N, EPI>;
    constexpr uint32_t kTxBytes = KIND == PROD_B ? C::kBBytes : kABytes;
    const int n_clusters = int(gridDim.x) >> 1;
    const int kpt = p.kblocks_per_tap, taps_w = p.taps_w, nn = p.num_n_tiles;
    uint32_t stage = uint32_t(me), phase = 0;
    uint32_t full_addr = bar_base + 8u * stage;  // local address of full[stage]; the leader's copy is full_addr & kPeerMask
    uint32_t dst = smem_base + stage * C::kStageBytes + (KIND == PROD_B ? kABytes : 0u);
    int kb = me;
    const int side_kb = p.side_kb;
    const int num_kb = main_kb + side_kb;
    auto advance = [&]() {
        stage += uint32_t(n_prod);
        full_addr += 8u * uint32_t(n_prod);
        dst += uint32_t(n_prod) * C::kStageBytes;
        if (stage >= uint32_t(C::kStages)) {
            stage -= uint32_t(C::kStages);
            phase ^= 1u;
            full_addr -= 8u * uint32_t(C::kStages);
            dst -= uint32_t(C::kStages) * C::kStageBytes;
        }
    };
    for (int item = int(blockIdx.x) >> 1; item < num_items; item += n_clusters) {
        if (kb < num_kb) {
            const int pair = item / nn, n_tile = item - pair * nn;
            const int m0 = (2 * pair + int(rank)) * kBlockM;
            const int n0 = n_tile * BLOCK_N + int(rank) * (BLOCK_N / 2);
            int cblk = 0, tap = 0, fs = 0, fr = 0;
            int img = 0, base_h = 0, base_w = 0;
            if (KIND != PROD_A_TILED) {
                if (kb < 8) {
                    cblk = kb;
                    while (cblk >= kpt) { cblk -= kpt; ++tap; ++fs; }
                    while (fs >= taps_w) { fs -= taps_w; ++fr; }
                } else {
                    tap = kb / kpt; cblk = kb - tap * kpt;
                    fr = tap / taps_w; fs = tap - fr * taps_w;
                }
            }
            if (KIND == PROD_A_IM2COL) {
                img = m0 / p.PQ;
                const int rem = m0 - img * p.PQ;
                const int op = rem / p.Q;
                const int oq = rem - op * p.Q;
                base_h = p.corner_h + op * p.stride_h;
                base_w = p.corner_w + oq * p.stride_w;
            }
#pragma unroll 1
            for (; kb < main_kb; kb += n_prod) {
                mbar_wait_bounded(full_addr + 8u * C::kStages, phase ^ 1u);  // local empty[stage]: the pair's MMAs have consumed it
                // The leader's producer announces the bytes of BOTH CTAs' loads of this operand; the peer only issues its loads (their
                // complete_tx lands on the leader's barrier, possibly before the announcement: the count may go negative inside a phase).
                const uint32_t leader_full = full_addr & kPeerMask;
                if (rank == 0) mbar_expect_tx(full_addr, 2u * kTxBytes);
                if (KIND == PROD_A_TILED) tma2_load_2d(tm, leader_full, dst, kb * kBlockK, m0);
                else if (KIND == PROD_A_IM2COL) tma2_load_im2col_4d(tm, leader_full, dst, cblk * kBlockK, base_w, base_h, img, uint16_t(fs * p.dil_w), uint16_t(fr * p.dil_h));
                else tma2_load_3d(tm, leader_full, dst, cblk * kBlockK, tap, n0);
                advance();
                if (KIND != PROD_A_TILED) {
                    cblk += n_prod;
                    while (cblk >= kpt) { cblk -= kpt; ++tap; ++fs; }
                    while (fs >= taps_w) { fs -= taps_w; ++fr; }
                }
            }
            if (side_kb > 0 && kb < num_kb) {
                // ---- projection shortcut: k-blocks main_kb.. read the block's input (1x1, own stride) and the shortcut's weights ----
                const bool side_im2col = KIND != PROD_B && p.side_mode == CONV_MODE_IM2COL;
                if (side_im2col) {
                    img = m0 / p.PQ;
                    const int rem = m0 - img * p.PQ;
                    const int op = rem / p.Q;
                    const int oq = rem - op * p.Q;
                    base_h = op * p.side_stride_h;
                    base_w = oq * p.side_stride_w;
                }
#pragma unroll 1
                for (; kb < num_kb; kb += n_prod) {
                    mbar_wait_bounded(full_addr + 8u * C::kStages, phase ^ 1u);
                    const uint32_t leader_full = full_addr & kPeerMask;
                    if (rank == 0) mbar_expect_tx(full_addr, 2u * kTxBytes);
                    const int c0 = (kb - main_kb) * kBlockK;
                    if (KIND == PROD_B) tma2_load_3d(tm_side, leader_full, dst, c0, 0, n0);
                    else if (side_im2col) tma2_load_im2col_4d(tm_side, leader_full, dst, c0, base_w, base_h, img, uint16_t(0), uint16_t(0));
                    else tma2_load_2d(tm_side, leader_full, dst, c0, m0);
                    advance();
                }
            }
        }
        kb -= num_kb;
    }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                 const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_b2, const ConvKernelParams p) {
    using C = Cfg<BLOCK_N, EPI>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) __trap();
    const uint32_t epi_base = smem_base + C::kStages * C::kStageBytes;
    const uint32_t bias_base = epi_base + kEpilogueWarps * C::kEpiBufs * kEpiBufBytes;
    const uint32_t stat_base = bias_base + kEpilogueWarps * kBiasSlotBytes;
    const uint32_t bar_base = epi_base + C::kEpiBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
    auto res_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::kStages + 4 + w * 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 20);
    static_assert(8u * (2 * C::kStages + 21) <= kBarrierBytes, "barrier region too small");
    static_assert(C::kSmemBytes <= kSmemLimit, "shared memory budget");
    static_assert(C::kStageBytes % 1024 == 0, "stages must keep the 1024-byte alignment of the swizzled tiles");
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_clusters = int(gridDim.x) >> 1;
    const int cluster_id = int(blockIdx.x) >> 1;
    const int num_pairs = (p.num_m_tiles + 1) / 2;
    const int num_items = num_pairs * p.num_n_tiles;
    const int main_kb = p.num_taps * p.kblocks_per_tap;
    const int total_kb = main_kb + p.side_kb;
    const int my_tiles = cluster_id < num_items ? (num_items - 1 - cluster_id) / n_clusters + 1 : 0;
    constexpr bool kSplit = C::kChunks >= 2;  // both epilogue groups share every tile (see conv_igemm.cu)
    constexpr bool HAS_RES = C::HAS_RES;

    if (threadIdx.x == 0) chain_stamp(p, 0);
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_a);
        prefetch_tensormap(&tm_b);
        prefetch_tensormap(&tm_out);
        if (HAS_RES) prefetch_tensormap(&tm_res);
        if (p.side_kb) {
            prefetch_tensormap(&tm_a2);
            prefetch_tensormap(&tm_b2);
        }
    }
    if (warp == 1) {
        if (lane < C::kStages) {
            mbar_init(full_bar(lane), 2);   // the leader's A and B producer, each announcing both CTAs' bytes (only the leader's copy is used)
            mbar_init(empty_bar(lane), 1);  // the leader's multicast commit
        } else if (lane < C::kStages + 2) {
            mbar_init(tmem_full_bar(lane - C::kStages), 1);
            mbar_init(tmem_empty_bar(lane - C::kStages), 2 * (kSplit ? 256 : 128));  // epilogue threads of both CTAs (leader's copy)
        } else if (lane >= 16) {
            mbar_init(res_bar((lane - 16) >> 1, lane & 1), 1);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        tmem2_alloc(tmem_slot, C::kTmemCols);
        tmem2_relinquish();
    }
    tc_fence_before();
    cluster_sync_exec();  // both CTAs' barriers exist before anybody arrives remotely or multicasts
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (p.use_pdl) grid_dep_launch_dependents();

    if (warp < kNumProducers || warp >= kBProducerWarp0) {
        const bool is_a = warp < kNumProducers;
        const int me = is_a ? warp : warp - kBProducerWarp0;
        const int n_prod = is_a ? C::kProducers : C::kBProducers;
        if (elect_one() && me < n_prod) {
            if (is_a) {
                if (p.use_pdl) grid_dep_wait();
                if (me == 0) chain_stamp(p, 1);
                if (p.mode == CONV_MODE_TILED) produce<PROD_A_TILED, BLOCK_N, EPI>(&tm_a, &tm_a2, p, smem_base, bar_base, me, n_prod, num_items, main_kb, rank);
                else produce<PROD_A_IM2COL, BLOCK_N, EPI>(&tm_a, &tm_a2, p, smem_base, bar_base, me, n_prod, num_items, main_kb, rank);
            } else {
                produce<PROD_B, BLOCK_N, EPI>(&tm_b, &tm_b2, p, smem_base, bar_base, me, n_prod, num_items, main_kb, rank);
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (leader CTA only) =================
        constexpr uint32_t idesc = make_idesc_f16(2 * kBlockM, BLOCK_N);
        constexpr uint64_t desc_hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
        constexpr uint32_t desc_lbo = 1u << 16;
        const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | desc_lbo;
        constexpr uint32_t kStage16 = C::kStageBytes >> 4;
        constexpr uint32_t kB16 = kABytes >> 4;
        if (leader && elect_one()) {
            uint32_t stage = 0, phase = 0;
            uint32_t a_lo = a_lo0;
            uint32_t full_addr = bar_base;
#if SMELTER_MMA_LOOKAHEAD
            // The barrier of the NEXT stage is probed (non-blocking) before this stage's MMAs are issued and the answer is read after
            // them: the probe's latency hides behind the issue slots instead of sitting between two k-blocks' MMAs (measured: the
            // wait -> 4 x MMA -> commit chain of one thread, not the tensor pipe, paced the k-loop at MMA time + ~0.08 us per k-block).
            bool ready = false;
            bool stamped = false;
            auto kblock = [&](uint32_t tmem_d, uint32_t first_accumulate) {
                if (!ready) mbar_wait_bounded(full_addr, phase);
                tc_fence_after();
                if (kInstr && first_accumulate == 0u && tmem_d == tmem_base && !stamped) { chain_stamp(p, 2); stamped = true; }
                const uint64_t a_desc = desc_hi | uint64_t(a_lo);
                const uint64_t b_desc = a_desc + kB16;
                const uint32_t cur_empty = full_addr + 8u * C::kStages;
                a_lo += kStage16;
                full_addr += 8u;
                if (++stage == uint32_t(C::kStages)) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_addr = bar_base; }
                ready = mbar_test_wait(full_addr, phase);
                umma2_f16(tmem_d, a_desc, b_desc, idesc, first_accumulate);
                umma2_f16(tmem_d, a_desc + 2, b_desc + 2, idesc, 1u);
                umma2_f16(tmem_d, a_desc + 4, b_desc + 4, idesc, 1u);
                umma2_f16(tmem_d, a_desc + 6, b_desc + 6, idesc, 1u);
                umma2_commit(cur_empty);  // empty[stage] in both CTAs
            };
#else
            auto kblock = [&](uint32_t tmem_d, uint32_t first_accumulate) {
                mbar_wait_bounded(full_addr, phase);
                tc_fence_after();
                const uint64_t a_desc = desc_hi | uint64_t(a_lo);
                const uint64_t b_desc = a_desc + kB16;
                umma2_f16(tmem_d, a_desc, b_desc, idesc, first_accumulate);
                umma2_f16(tmem_d, a_desc + 2, b_desc + 2, idesc, 1u);
                umma2_f16(tmem_d, a_desc + 4, b_desc + 4, idesc, 1u);
                umma2_f16(tmem_d, a_desc + 6, b_desc + 6, idesc, 1u);
                umma2_commit(full_addr + 8u * C::kStages);  // empty[stage] in both CTAs
                a_lo += kStage16;
                full_addr += 8u;
                if (++stage == uint32_t(C::kStages)) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_addr = bar_base; }
            };
#endif
            for (int t = 0; t < my_tiles; ++t) {
                const int acc = t & 1;
                mbar_wait_bounded(tmem_empty_bar(acc), ((uint32_t(t) >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * BLOCK_N);
                kblock(tmem_d, 0u);
                int kb = 1;
#pragma unroll 1
                for (; kb + 1 < total_kb; kb += 2) {
                    kblock(tmem_d, 1u);
                    kblock(tmem_d, 1u);
                }
                if (kb < total_kb) kblock(tmem_d, 1u);
                umma2_commit(tmem_full_bar(acc));  // accumulator complete -> the epilogues of both CTAs
            }
        }
    } else {
        // ================= epilogue (warps 4..11), each CTA drains its own 128 accumulator rows =================
        if (p.use_pdl) grid_dep_wait();
        const int ewarp = warp - kEpilogueWarp0;
        const int group = ewarp >> 2;
        const int ew = ewarp & 3;
        const uint32_t buf0 = epi_base + uint32_t(ewarp) * uint32_t(C::kEpiBufs) * kEpiBufBytes;
        const uint32_t bias_slot = bias_base + uint32_t(ewarp) * kBiasSlotBytes;
        const uint32_t row_off = uint32_t(lane) * 128u;
        const uint32_t sw = uint32_t(lane & 7);
        const uint64_t pol_drop = l2_policy_evict_first();
        const bool is_sigmoid = p.act == ACT_SIGMOID;
        const __half2 lo2 = __float2half2_rn(p.act == ACT_RELU ? 0.f : (p.act == ACT_CLIP ? p.clip_lo : -INFINITY));
        const __half2 hi2 = __float2half2_rn(p.act == ACT_CLIP ? p.clip_hi : INFINITY);
        constexpr int kCPW = kSplit ? C::kChunks / 2 : C::kChunks;
        const int group_tiles = kSplit ? my_tiles : (my_tiles > group ? (my_tiles - group + 1) / 2 : 0);
        const int n_items = group_tiles * kCPW;
        auto item_coords = [&](int item, int* m_row0, int* col0) {
            const int gt = item / kCPW;
            const int j = item - gt * kCPW;
            const int c = kSplit ? group + 2 * j : j;
            const int work = cluster_id + (kSplit ? gt : 2 * gt + group) * n_clusters;
            const int pair = work / p.num_n_tiles;
            const int n_tile = work - pair * p.num_n_tiles;
            *m_row0 = (2 * pair + int(rank)) * kBlockM + ew * 32;
            *col0 = n_tile * BLOCK_N + c * kChunkN;
        };
        auto prefetch_res = [&](int item) {
            int m_row0, col0;
            item_coords(item, &m_row0, &col0);
            const int b = item & 1;
            fence_proxy_async_smem();
            mbar_expect_tx(res_bar(ewarp, b), kEpiBufBytes);
            if (p.l2_hints & 2) tma_load_2d_hint(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, col0, m_row0, pol_drop);
            else tma_load_2d(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, col0, m_row0);
        };
        if (HAS_RES && n_items > 0 && lane == 0) prefetch_res(0);
        auto load_bias = [&](int col0) {
            float2 b = __ldg(reinterpret_cast<const float2*>(p.bias + col0) + lane);
            if (p.bias2) {  // projection shortcut: its bias joins before the activation
                const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias2 + col0) + lane);
                b.x += b2.x;
                b.y += b2.y;
            }
            return b;
        };
        float2 bias_next = make_float2(0.f, 0.f);
        if (n_items > 0) {
            int m_row0, col0;
            item_coords(0, &m_row0, &col0);
            bias_next = load_bias(col0);
        }
        uint32_t res_phase = 0;
        int item = 0;
        for (int gt = 0; gt < group_tiles; ++gt) {
            const int acc = kSplit ? (gt & 1) : group;
            const uint32_t acc_parity = uint32_t(kSplit ? (gt >> 1) : gt) & 1u;
            mbar_wait_bounded(tmem_full_bar(acc), acc_parity);
            tc_fence_after();
            const bool stamper = kInstr && blockIdx.x == 0 && ewarp == 0 && lane == 0;
            if (kInstr && gt == group_tiles - 1 && ew == 0 && lane == 0) chain_stamp(p, 3);
            epi_stamp(p, stamper, item, 0);
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * BLOCK_N);
#pragma unroll 1
            for (int j = 0; j < kCPW; ++j, ++item) {
                int m_row0, col0;
                item_coords(item, &m_row0, &col0);
                const int c = kSplit ? group + 2 * j : j;
                const int b = HAS_RES ? (item & 1) : 0;
                if (HAS_RES && lane == 0 && item + 1 < n_items) {
                    tma_store_wait_read<0>();
                    prefetch_res(item + 1);
                }
                epi_stamp(p, stamper, item, 1);
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_slot + uint32_t(lane) * 8u), "f"(bias_next.x), "f"(bias_next.y) : "memory");
                if (item + 1 < n_items) {
                    int nm, ncol0;
                    item_coords(item + 1, &nm, &ncol0);
                    bias_next = load_bias(ncol0);
                }
                uint32_t v[kChunkN];
                tmem_ld_32(taddr + uint32_t(c * kChunkN), v);
                tmem_ld_32(taddr + uint32_t(c * kChunkN + 32), v + 32);
                tmem_ld_wait();
                epi_stamp(p, stamper, item, 2);
                if (j == kCPW - 1) {  // hand the accumulator back to the leader's MMA warp (remote arrive from the peer CTA)
                    tc_fence_before();
                    mbar_arrive_cluster(tmem_empty_bar(acc) & kPeerMask);
                }
                if (HAS_RES) {
                    mbar_wait_bounded(res_bar(ewarp, b), (res_phase >> b) & 1u);
                    res_phase ^= 1u << b;
                }
                epi_stamp(p, stamper, item, 3);
                __syncwarp();
                const uint32_t buf = buf0 + uint32_t(b) * kEpiBufBytes;
                uint4 out[kChunkN / 8];
                epilogue_math<kChunkN, HAS_RES>(v, out, buf + row_off, sw, bias_slot, is_sigmoid, lo2, hi2);
                if (!HAS_RES) {
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                }
#pragma unroll
                for (int g = 0; g < kChunkN / 8; ++g) st_shared_v4(buf + row_off + ((uint32_t(g) ^ sw) << 4), out[g]);
                fence_proxy_async_smem();
                __syncwarp();
                epi_stamp(p, stamper, item, 4);
                if (lane == 0) {
                    tma_store_2d(&tm_out, buf, col0, m_row0);
                    tma_store_commit();
                }
                epi_stamp(p, stamper, item, 5);
                if (EPI == 2) {
                    // Column sums and sums of squares of this warp's 32 x 64 box for the instance norm behind the layer, read back from
                    // the staging buffer (the fp16 values the norm itself would read; lane l owns columns 2l, 2l + 1: one conflict-free
                    // 128-byte row per load).  The four row warps of the group add theirs up in shared memory (fixed order); one warp
                    // per work item sends the 128-row sums to the image's fp64 accumulators.
                    const uint32_t cg = uint32_t(lane) >> 2, cw = (uint32_t(lane) & 3u) << 2;
                    float4 st = make_float4(0.f, 0.f, 0.f, 0.f);  // sum, sum of squares of column 2l; the same of column 2l + 1
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        uint32_t hv;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv) : "r"(buf + uint32_t(r) * 128u + ((cg ^ uint32_t(r & 7)) << 4) + cw) : "memory");
                        const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&hv));
                        st.x += f2.x; st.y = fmaf(f2.x, f2.x, st.y);
                        st.z += f2.y; st.w = fmaf(f2.y, f2.y, st.w);
                    }
                    const uint32_t slot0 = stat_base + (uint32_t((item & 1) * 2 + group) << 11);  // [item parity][group][row warp][lane] float4
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot0 + (uint32_t(ew) << 9) + (uint32_t(lane) << 4)), "f"(st.x), "f"(st.y), "f"(st.z), "f"(st.w) : "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
                    const int tile_row0 = m_row0 - ew * 32;
                    const int col = col0 + 2 * lane;
                    if ((item & 3) == ew && tile_row0 < p.M && col < p.out_pitch) {
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            float4 t;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(slot0 + (uint32_t(w4) << 9) + (uint32_t(lane) << 4)) : "memory");
                            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                        }
                        // replica = tile index mod stats_reps: same-address reductions serialise in L2 (~30 ns each, measured: 512 tiles
                        // of one image on one accumulator set made a 14 us layer 31 us)
                        const int img = tile_row0 / p.stats_rows;
                        const int rep = ((tile_row0 - img * p.stats_rows) >> 7) & (p.stats_reps - 1);
                        double* dst = p.stats + (size_t(img * p.stats_reps + rep) * size_t(p.out_pitch) + size_t(col)) * 2;
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst), "d"(double(a.x)) : "memory");
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 1), "d"(double(a.y)) : "memory");
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 2), "d"(double(a.z)) : "memory");
                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + 3), "d"(double(a.w)) : "memory");
                    }
                }
            }
        }
        if (kInstr && lane == 0 && group_tiles > 0) chain_stamp(p, 4);
        if (lane == 0) tma_store_wait_read<0>();
        if (kInstr && lane == 0 && group_tiles > 0) chain_stamp(p, 6);
    }

    tc_fence_before();
    cluster_sync_exec();  // nobody exits (or frees TMEM) while the other CTA may still signal its barriers or read its operands
    if (threadIdx.x == 0) chain_stamp(p, 7);
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem2_dealloc(tmem_base, C::kTmemCols);
        if (lane == 0) chain_stamp(p, 5);
    }
}

template <int BLOCK_N>
cudaError_t set_attr_t() {
    cudaError_t e = cudaFuncSetAttribute(conv_pair_kernel<BLOCK_N, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N, 0>::kSmemBytes));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv_pair_kernel<BLOCK_N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N, 2>::kSmemBytes));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(conv_pair_kernel<BLOCK_N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N, 1>::kSmemBytes));
}

template <int BLOCK_N>
cudaError_t launch_t(const ConvTcLaunch& L, int grid, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L.p.has_residual ? Cfg<BLOCK_N, 1>::kSmemBytes : L.p.stats ? Cfg<BLOCK_N, 2>::kSmemBytes : Cfg<BLOCK_N, 0>::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = L.use_pdl ? 2 : 1;
    if (L.p.has_residual) return cudaLaunchKernelEx(&cfg, conv_pair_kernel<BLOCK_N, 1>, L.tm_a, L.tm_b, L.tm_out, L.tm_res, L.tm_a2, L.tm_b2, L.p);
    if (L.p.stats) return cudaLaunchKernelEx(&cfg, conv_pair_kernel<BLOCK_N, 2>, L.tm_a, L.tm_b, L.tm_out, L.tm_res, L.tm_a2, L.tm_b2, L.p);
    return cudaLaunchKernelEx(&cfg, conv_pair_kernel<BLOCK_N, 0>, L.tm_a, L.tm_b, L.tm_out, L.tm_res, L.tm_a2, L.tm_b2, L.p);
}

}  // namespace

bool conv_pair_supported(const ConvTcLaunch& L) { return L.splits == 1 && (L.block_n == 64 || L.block_n == 128 || L.block_n == 256); }

cudaError_t conv_pair_set_attr(int block_n) {
    switch (block_n) {
        case 64: return set_attr_t<64>();
        case 128: return set_attr_t<128>();
        case 256: return set_attr_t<256>();
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t conv_pair_launch(const ConvTcLaunch& L, int num_sms, cudaStream_t stream) {
    const int pairs = (L.p.num_m_tiles + 1) / 2;
    const long items = long(pairs) * L.p.num_n_tiles;
    // Plans for a share of the chip (smShare): as many clusters as it takes to finish in the same number of rounds, not one more --
    // 50 work items on 37 clusters need two rounds whether 37 or 25 clusters run them, and the SMs left alone are what a kernel of
    // another encode in flight starts on (+0.8 % images/s with three in flight; a whole-chip plan alone loses 0.8 % by it, because
    // clusters that finish early leave their L2 bandwidth to the rest).
    long clusters = std::min<long>(items, num_sms / 2);
    if (L.balanced_grid) {
        const long rounds = (items + clusters - 1) / clusters;
        clusters = (items + rounds - 1) / rounds;
    }
    const int grid = 2 * int(clusters);
    switch (L.block_n) {
        case 64: return launch_t<64>(L, grid, stream);
        case 128: return launch_t<128>(L, grid, stream);
        case 256: return launch_t<256>(L, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace k
}  // namespace smelter
