// Back-to-back convolution pair in one launch (opt-in, SMELTER_B2B=1): conv A (any tensor-core mode, N1 = 64 or 128 output
// channels, e.g. the 3x3 of a ResNet bottleneck) followed by a 1x1 / stride-1 conv B that reads nothing but A's output (the
// bottleneck's expanding 1x1, with its fused residual add and activation).  Both work on the same 128 output pixels, so the
// intermediate never leaves the SM:
//
//   phase 1   acc1[128 x N1] (TMEM)  = im2col(x) . W_A          operands through the TMA ring, exactly conv_pair.cu's k-loop
//   epilogue1 acc1 -> +bias_A -> activation -> fp16 -> shared memory, written in the 128B-swizzled K-major layout a tcgen05 A
//             operand has (N1/64 tiles of [128 x 64]); fence.proxy.async, then one mbarrier arrival per thread
//   phase 2   for every 128-column slice of conv B:  acc2[128 x 128] = A2(smem) . W_B[slice]   (only the weights stream through
//             the ring, into the activation slot of a stage); accumulators double-buffered in TMEM
//   epilogue2 conv_pair.cu's epilogue: +bias_B (+residual through TMA) -> activation -> swizzled staging -> TMA store
//
// Two-CTA clusters and cta_group::2 MMAs as in conv_pair.cu (each CTA: its own 128 pixels, half of every weight tile).  One
// launch and one HBM round trip of the intermediate tensor less per pair; the price is a shallower operand ring (A2 takes 16 /
// 32 KiB) and one serial epilogue1 -> first phase-2 MMA hand-over per tile.
//
// Order of events that makes the single A2 buffer and the single acc1 safe without further barriers: MMAs execute in issue
// order and a commit covers everything issued before it, so acc1_full(tile t+1) implies phase 2 of tile t has finished reading
// A2; phase 1 of tile t+1 is issued after the a2_ready(t) wait, i.e. after every epilogue thread has read acc1(t).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv_igemm.h"
#include "pair_ptx.cuh"
#include "ptx.cuh"

namespace smelter {
namespace k {

using namespace ptx;
using namespace pairptx;

namespace {

constexpr int kBlockM = 128;
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
constexpr int kBlockN2 = 128;                                  // columns of conv B per accumulator
constexpr uint32_t kW2Bytes = (kBlockN2 / 2) * kBlockK * 2;    // this CTA's half of a conv-B weight tile
constexpr uint32_t kEpiBufBytes = 32 * kChunkN * 2;
constexpr uint32_t kBiasSlotBytes = kChunkN * 4;
constexpr uint32_t kBarrierBytes = 512;
constexpr uint32_t kSmemLimit = 227 * 1024;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAcc1Col = 2 * kBlockN2;

template <int N1, bool HAS_RES>
struct Cfg {
    static constexpr uint32_t kB1Bytes = (N1 / 2) * kBlockK * 2;  // this CTA's half of the conv-A weight tile
    static constexpr uint32_t kStageBytes = kABytes + kB1Bytes;
    static constexpr int kKb2 = N1 / kBlockK;                     // k-blocks of conv B = tiles of the intermediate
    static constexpr uint32_t kA2Bytes = kKb2 * kABytes;
    static constexpr int kEpiBufs = HAS_RES ? 2 : 1;
    static constexpr uint32_t kEpiBytes = kEpilogueWarps * (kEpiBufs * kEpiBufBytes + kBiasSlotBytes);
    static constexpr int kStagesFit = int((kSmemLimit - kA2Bytes - kEpiBytes - kBarrierBytes) / kStageBytes);
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kProducers = kStages < kNumProducers ? kStages : kNumProducers;
    static constexpr int kBProducers = kStages < kNumBProducers ? kStages : kNumBProducers;
    static constexpr size_t kSmemBytes = size_t(kStages) * kStageBytes + kA2Bytes + kEpiBytes + kBarrierBytes;
};

enum ProducerKind : int { PROD_A_TILED = 0, PROD_A_IM2COL = 1, PROD_B = 2 };

// One elected thread per producer warp.  Per tile the k-block sequence is [0, kb1) for conv A and [kb1, kb1 + sub * kKb2) for
// conv B.  Conv-B k-blocks carry weights only: the B producers load them (into the stage's activation slot) and announce their
// bytes, the leader's A producers just add the second arrival the barrier expects.
template <int KIND, int N1, bool HAS_RES>
__device__ __forceinline__ void produce(const CUtensorMap* tm, const CUtensorMap* tm_w2, const ConvKernelParams& p, int subtiles, uint32_t smem_base,
                                        uint32_t bar_base, int me, int n_prod, int num_items, int kb1, uint32_t rank) {
    using C = Cfg<N1, HAS_RES>;
    const int n_clusters = int(gridDim.x) >> 1;
    const int kpt = p.kblocks_per_tap, taps_w = p.taps_w;
    const int num_kb = kb1 + subtiles * C::kKb2;
    uint32_t stage = uint32_t(me), phase = 0;
    uint32_t full_addr = bar_base + 8u * stage;
    uint32_t dst = smem_base + stage * C::kStageBytes;  // start of the stage (activation slot)
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
    int kb = me;
    for (int item = int(blockIdx.x) >> 1; item < num_items; item += n_clusters) {
        if (kb < num_kb) {
            const int m0 = (2 * item + int(rank)) * kBlockM;
            int cblk = 0, tap = 0, fs = 0, fr = 0;
            int img = 0, base_h = 0, base_w = 0;
            if (KIND != PROD_A_TILED && kb < kb1) {
                tap = kb / kpt; cblk = kb - tap * kpt;
                fr = tap / taps_w; fs = tap - fr * taps_w;
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
            for (; kb < kb1; kb += n_prod) {
                mbar_wait_bounded(full_addr + 8u * C::kStages, phase ^ 1u);  // local empty[stage]
                const uint32_t leader_full = full_addr & kPeerMask;
                if (rank == 0) mbar_expect_tx(full_addr, 2u * (KIND == PROD_B ? C::kB1Bytes : kABytes));
                if (KIND == PROD_A_TILED) tma2_load_2d(tm, leader_full, dst, kb * kBlockK, m0);
                else if (KIND == PROD_A_IM2COL) tma2_load_im2col_4d(tm, leader_full, dst, cblk * kBlockK, base_w, base_h, img, uint16_t(fs * p.dil_w), uint16_t(fr * p.dil_h));
                else tma2_load_3d(tm, leader_full, dst + kABytes, cblk * kBlockK, tap, int(rank) * (N1 / 2));
                advance();
                if (KIND != PROD_A_TILED) {
                    cblk += n_prod;
                    while (cblk >= kpt) { cblk -= kpt; ++tap; ++fs; }
                    while (fs >= taps_w) { fs -= taps_w; ++fr; }
                }
            }
#pragma unroll 1
            for (; kb < num_kb; kb += n_prod) {
                if (KIND == PROD_B) {
                    mbar_wait_bounded(full_addr + 8u * C::kStages, phase ^ 1u);
                    const uint32_t leader_full = full_addr & kPeerMask;
                    if (rank == 0) mbar_expect_tx(full_addr, 2u * kW2Bytes);
                    const int j = kb - kb1;
                    const int sub = j / C::kKb2, kk = j - sub * C::kKb2;
                    tma2_load_2d(tm_w2, leader_full, dst, kk * kBlockK, sub * kBlockN2 + int(rank) * (kBlockN2 / 2));
                } else {
                    // Both CTAs' activation producers keep waiting for the stage although only the leader's has something to do (the
                    // second arrival the barrier expects): a producer that skipped these waits could get two ring revolutions ahead
                    // of the consumer, where the one-bit phase parity of its next wait would pass too early.
                    mbar_wait_bounded(full_addr + 8u * C::kStages, phase ^ 1u);
                    if (rank == 0) mbar_arrive(full_addr);
                }
                advance();
            }
        }
        kb -= num_kb;
    }
}

template <int N1, bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 1)
conv_b2b_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b1, const __grid_constant__ CUtensorMap tm_b2,
                const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res, const ConvKernelParams p,
                const ConvB2bParams p2) {
    using C = Cfg<N1, HAS_RES>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) __trap();
    const uint32_t a2_base = smem_base + C::kStages * C::kStageBytes;
    const uint32_t epi_base = a2_base + C::kA2Bytes;
    const uint32_t bias_base = epi_base + kEpilogueWarps * C::kEpiBufs * kEpiBufBytes;
    const uint32_t bar_base = epi_base + C::kEpiBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto acc2_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
    auto acc2_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
    auto res_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::kStages + 4 + w * 2 + b); };
    const uint32_t acc1_full_bar = bar_base + 8u * (2 * C::kStages + 20);
    const uint32_t a2_ready_bar = bar_base + 8u * (2 * C::kStages + 21);
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 22);
    static_assert(8u * (2 * C::kStages + 23) <= kBarrierBytes, "barrier region too small");
    static_assert(C::kSmemBytes <= kSmemLimit, "shared memory budget");
    static_assert(C::kStageBytes % 1024 == 0 && C::kA2Bytes % 1024 == 0, "swizzled tiles need 1024-byte alignment");
    static_assert(C::kStages >= 3, "operand ring too shallow");
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_clusters = int(gridDim.x) >> 1;
    const int cluster_id = int(blockIdx.x) >> 1;
    const int num_items = (p.num_m_tiles + 1) / 2;  // one work item per pair of m-tiles: the N1 columns of conv A are one tile
    const int kb1 = p.num_taps * p.kblocks_per_tap;
    const int subtiles = p2.subtiles;
    const int my_items = cluster_id < num_items ? (num_items - 1 - cluster_id) / n_clusters + 1 : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_a);
        prefetch_tensormap(&tm_b1);
        prefetch_tensormap(&tm_b2);
        prefetch_tensormap(&tm_out);
        if (HAS_RES) prefetch_tensormap(&tm_res);
    }
    if (warp == 1) {
        if (lane < C::kStages) {
            mbar_init(full_bar(lane), 2);
            mbar_init(empty_bar(lane), 1);
        } else if (lane < C::kStages + 2) {
            mbar_init(acc2_full_bar(lane - C::kStages), 1);
            mbar_init(acc2_empty_bar(lane - C::kStages), 2 * 256);  // every epilogue thread of both CTAs (leader's copy)
        } else if (lane == C::kStages + 2) {
            mbar_init(acc1_full_bar, 1);
            mbar_init(a2_ready_bar, 2 * 256);
        } else if (lane >= 16) {
            mbar_init(res_bar((lane - 16) >> 1, lane & 1), 1);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        tmem2_alloc(tmem_slot, kTmemCols);
        tmem2_relinquish();
    }
    tc_fence_before();
    cluster_sync_all();
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
                if (p.mode == CONV_MODE_TILED) produce<PROD_A_TILED, N1, HAS_RES>(&tm_a, &tm_b2, p, subtiles, smem_base, bar_base, me, n_prod, num_items, kb1, rank);
                else produce<PROD_A_IM2COL, N1, HAS_RES>(&tm_a, &tm_b2, p, subtiles, smem_base, bar_base, me, n_prod, num_items, kb1, rank);
            } else {
                produce<PROD_B, N1, HAS_RES>(&tm_b1, &tm_b2, p, subtiles, smem_base, bar_base, me, n_prod, num_items, kb1, rank);
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer (leader CTA only) =================
        constexpr uint32_t idesc1 = make_idesc_f16(2 * kBlockM, N1);
        constexpr uint32_t idesc2 = make_idesc_f16(2 * kBlockM, kBlockN2);
        constexpr uint64_t desc_hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
        constexpr uint32_t desc_lbo = 1u << 16;
        const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | desc_lbo;
        const uint32_t a2_lo0 = ((a2_base & 0x3FFFFu) >> 4) | desc_lbo;
        constexpr uint32_t kStage16 = C::kStageBytes >> 4;
        constexpr uint32_t kB16 = kABytes >> 4;
        if (leader && elect_one()) {
            uint32_t stage = 0, phase = 0;
            uint32_t a_lo = a_lo0;
            uint32_t full_addr = bar_base;
            auto next_stage = [&]() {
                a_lo += kStage16;
                full_addr += 8u;
                if (++stage == uint32_t(C::kStages)) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_addr = bar_base; }
            };
            uint32_t sub_count = 0;  // conv-B accumulators issued so far (selects the TMEM buffer and its parity)
            for (int t = 0; t < my_items; ++t) {
                // ---- phase 1: acc1 = im2col(x) . W_A ----
                const uint32_t acc1 = tmem_base + kAcc1Col;
#pragma unroll 1
                for (int kb = 0; kb < kb1; ++kb) {
                    mbar_wait_bounded(full_addr, phase);
                    tc_fence_after();
                    const uint64_t a_desc = desc_hi | uint64_t(a_lo);
                    const uint64_t b_desc = a_desc + kB16;
                    umma2_f16(acc1, a_desc, b_desc, idesc1, kb > 0 ? 1u : 0u);
                    umma2_f16(acc1, a_desc + 2, b_desc + 2, idesc1, 1u);
                    umma2_f16(acc1, a_desc + 4, b_desc + 4, idesc1, 1u);
                    umma2_f16(acc1, a_desc + 6, b_desc + 6, idesc1, 1u);
                    umma2_commit(full_addr + 8u * C::kStages);
                    next_stage();
                }
                umma2_commit(acc1_full_bar);
                // ---- the fp16 intermediate of both CTAs is in shared memory ----
                mbar_wait_bounded(a2_ready_bar, uint32_t(t) & 1u);
                tc_fence_after();
                // ---- phase 2: one accumulator per 128 output channels of conv B ----
#pragma unroll 1
                for (int sub = 0; sub < subtiles; ++sub, ++sub_count) {
                    const uint32_t acc = sub_count & 1u;
                    mbar_wait_bounded(acc2_empty_bar(int(acc)), ((sub_count >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t acc2 = tmem_base + acc * uint32_t(kBlockN2);
#pragma unroll 1
                    for (int kk = 0; kk < C::kKb2; ++kk) {
                        mbar_wait_bounded(full_addr, phase);
                        tc_fence_after();
                        const uint64_t a_desc = desc_hi | uint64_t(a2_lo0 + uint32_t(kk) * kB16);
                        const uint64_t b_desc = desc_hi | uint64_t(a_lo);  // conv-B weights sit in the stage's activation slot
                        umma2_f16(acc2, a_desc, b_desc, idesc2, kk > 0 ? 1u : 0u);
                        umma2_f16(acc2, a_desc + 2, b_desc + 2, idesc2, 1u);
                        umma2_f16(acc2, a_desc + 4, b_desc + 4, idesc2, 1u);
                        umma2_f16(acc2, a_desc + 6, b_desc + 6, idesc2, 1u);
                        umma2_commit(full_addr + 8u * C::kStages);
                        next_stage();
                    }
                    umma2_commit(acc2_full_bar(int(acc)));
                }
            }
        }
    } else {
        // ================= epilogue warps 4..11 =================
        if (p.use_pdl) grid_dep_wait();
        const int ewarp = warp - kEpilogueWarp0;
        const int group = ewarp >> 2;
        const int ew = ewarp & 3;
        const uint32_t buf0 = epi_base + uint32_t(ewarp) * uint32_t(C::kEpiBufs) * kEpiBufBytes;
        const uint32_t bias_slot = bias_base + uint32_t(ewarp) * kBiasSlotBytes;
        const uint32_t row_off = uint32_t(lane) * 128u;
        const uint32_t sw = uint32_t(lane & 7);
        const uint64_t pol_drop = l2_policy_evict_first();
        const bool is_sigmoid = p2.act2 == ACT_SIGMOID;
        const __half2 lo2 = __float2half2_rn(p2.act2 == ACT_RELU ? 0.f : (p2.act2 == ACT_CLIP ? p2.clip2_lo : -INFINITY));
        const __half2 hi2 = __float2half2_rn(p2.act2 == ACT_CLIP ? p2.clip2_hi : INFINITY);
        const __half2 mid_lo = __float2half2_rn(p.act == ACT_RELU ? 0.f : (p.act == ACT_CLIP ? p.clip_lo : -INFINITY));
        const __half2 mid_hi = __float2half2_rn(p.act == ACT_CLIP ? p.clip_hi : INFINITY);
        const uint32_t a2_row = a2_base + uint32_t(ew * 32 + lane) * 128u;  // this thread's pixel inside every [128 x 64] tile of A2
        // conv-B output items of this group: (tile t, sub-tile s) -> the group's 64-column chunk of that accumulator
        const int n_out = my_items * subtiles;
        auto out_coords = [&](int o, int* m_row0, int* col0) {
            const int t = o / subtiles;
            const int s = o - t * subtiles;
            const int item = cluster_id + t * n_clusters;
            *m_row0 = (2 * item + int(rank)) * kBlockM + ew * 32;
            *col0 = s * kBlockN2 + group * kChunkN;
        };
        auto prefetch_res = [&](int o) {
            int m_row0, col0;
            out_coords(o, &m_row0, &col0);
            const int b = o & 1;
            fence_proxy_async_smem();
            mbar_expect_tx(res_bar(ewarp, b), kEpiBufBytes);
            if (p2.l2_hints & 2) tma_load_2d_hint(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, col0, m_row0, pol_drop);
            else tma_load_2d(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, col0, m_row0);
        };
        if (HAS_RES && n_out > 0 && lane == 0) prefetch_res(0);
        float2 bias_next = make_float2(0.f, 0.f);
        if (n_out > 0) bias_next = __ldg(reinterpret_cast<const float2*>(p2.bias2 + group * kChunkN) + lane);
        uint32_t res_phase = 0;
        int o = 0;
        for (int t = 0; t < my_items; ++t) {
            // ---- epilogue 1: acc1 -> bias, activation, fp16 -> A2 (K-major, 128B swizzle); 32-column pieces alternate between the groups ----
            mbar_wait_bounded(acc1_full_bar, uint32_t(t) & 1u);
            tc_fence_after();
            const uint32_t t1 = tmem_base + (uint32_t(ew * 32) << 16) + kAcc1Col;
#pragma unroll 1
            for (int piece = group; piece < N1 / 32; piece += 2) {
                uint32_t v[32];
                tmem_ld_32(t1 + uint32_t(piece * 32), v);
                tmem_ld_wait();
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + piece * 32);
                const uint32_t tile = a2_row + uint32_t(piece >> 1) * kABytes;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
                    uint4 q;
                    __half2* qh = reinterpret_cast<__half2*>(&q);
                    qh[0] = __floats2half2_rn(__uint_as_float(v[g * 8 + 0]) + ba.x, __uint_as_float(v[g * 8 + 1]) + ba.y);
                    qh[1] = __floats2half2_rn(__uint_as_float(v[g * 8 + 2]) + ba.z, __uint_as_float(v[g * 8 + 3]) + ba.w);
                    qh[2] = __floats2half2_rn(__uint_as_float(v[g * 8 + 4]) + bb.x, __uint_as_float(v[g * 8 + 5]) + bb.y);
                    qh[3] = __floats2half2_rn(__uint_as_float(v[g * 8 + 6]) + bb.z, __uint_as_float(v[g * 8 + 7]) + bb.w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) qh[i] = __hmin2(__hmax2(qh[i], mid_lo), mid_hi);
                    const uint32_t g16 = uint32_t((piece & 1) * 4 + g);  // 16-byte group inside the 128-byte row of the tile
                    st_shared_v4(tile + ((g16 ^ sw) << 4), q);
                }
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
            tc_fence_before();
            mbar_arrive_cluster(a2_ready_bar & kPeerMask);

            // ---- epilogue 2: conv_pair.cu's, one 64-column chunk per group and accumulator ----
#pragma unroll 1
            for (int s = 0; s < subtiles; ++s, ++o) {
                const int acc = o & 1;
                mbar_wait_bounded(acc2_full_bar(acc), uint32_t(o >> 1) & 1u);
                tc_fence_after();
                int m_row0, col0;
                out_coords(o, &m_row0, &col0);
                const int b = HAS_RES ? (o & 1) : 0;
                if (HAS_RES && lane == 0 && o + 1 < n_out) {
                    tma_store_wait_read<0>();
                    prefetch_res(o + 1);
                }
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_slot + uint32_t(lane) * 8u), "f"(bias_next.x), "f"(bias_next.y) : "memory");
                if (o + 1 < n_out) {
                    int nm, ncol0;
                    out_coords(o + 1, &nm, &ncol0);
                    bias_next = __ldg(reinterpret_cast<const float2*>(p2.bias2 + ncol0) + lane);
                }
                uint32_t v[kChunkN];
                const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * kBlockN2 + group * kChunkN);
                tmem_ld_32(taddr, v);
                tmem_ld_32(taddr + 32u, v + 32);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive_cluster(acc2_empty_bar(acc) & kPeerMask);
                if (HAS_RES) {
                    mbar_wait_bounded(res_bar(ewarp, b), (res_phase >> b) & 1u);
                    res_phase ^= 1u << b;
                }
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
                if (lane == 0) {
                    tma_store_2d(&tm_out, buf, col0, m_row0);
                    tma_store_commit();
                }
            }
        }
        if (lane == 0) tma_store_wait_read<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem2_dealloc(tmem_base, kTmemCols);
    }
}

template <int N1>
cudaError_t set_attr_t() {
    cudaError_t e = cudaFuncSetAttribute(conv_b2b_kernel<N1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<N1, false>::kSmemBytes));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(conv_b2b_kernel<N1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<N1, true>::kSmemBytes));
}

template <int N1>
cudaError_t launch_t(const ConvB2bLaunch& L, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(L.grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L.p2.has_residual ? Cfg<N1, true>::kSmemBytes : Cfg<N1, false>::kSmemBytes;
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
    if (L.p2.has_residual) return cudaLaunchKernelEx(&cfg, conv_b2b_kernel<N1, true>, L.tm_a, L.tm_b1, L.tm_b2, L.tm_out, L.tm_res, L.p, L.p2);
    return cudaLaunchKernelEx(&cfg, conv_b2b_kernel<N1, false>, L.tm_a, L.tm_b1, L.tm_b2, L.tm_out, L.tm_res, L.p, L.p2);
}

}  // namespace

bool conv_b2b_supported(const ConvB2bProblem& q, int num_sms) {
    const ConvTcProblem& a = q.first;
    if (num_sms < 2 || (a.c_out != 64 && a.c_out != 128) || a.c_out_pitch != a.c_out) return false;
    if (a.mode != CONV_MODE_TILED && a.mode != CONV_MODE_IM2COL) return false;
    if (a.act == ACT_SIGMOID || a.residual || a.side_x) return false;
    if (q.c_out2 < 64 || q.c_out2_pitch % 8) return false;
    return true;
}

bool conv_b2b_prepare(ConvB2bLaunch* L, const ConvB2bProblem& q, int num_sms, std::string* err) {
    if (!conv_b2b_supported(q, num_sms)) { if (err) *err = "conv_b2b: unsupported pair of convolutions"; return false; }
    // geometry, the activation map and the (half-tile) weight map of conv A are conv_tc_prepare's two-CTA form
    ConvTcProblem a = q.first;
    a.block_n = a.c_out;
    a.pair = 1;
    a.splits = 1;
    a.residual = nullptr;
    a.y = q.y;  // only so that the (unused) output map has a base address
    ConvTcLaunch tmp;
    if (!conv_tc_prepare(&tmp, a, num_sms, err)) return false;
    if (!tmp.pair || tmp.block_n != a.c_out) { if (err) *err = "conv_b2b: conv A did not get the two-CTA plan"; return false; }
    memset(L, 0, sizeof(*L));
    L->tm_a = tmp.tm_a;
    L->tm_b1 = tmp.tm_b;
    L->p = tmp.p;
    L->p.has_residual = 0;
    L->n1 = a.c_out;
    L->use_pdl = tmp.use_pdl;
    const long M = tmp.p.M;
    if (!conv_tc_encode_2d(&L->tm_b2, q.w2_packed, a.c_out, q.c_out2, kBlockK, kBlockN2 / 2, err)) return false;
    if (!conv_tc_encode_2d(&L->tm_out, q.y, q.c_out2_pitch, M, kChunkN, 32, err)) return false;
    if (!conv_tc_encode_2d(&L->tm_res, q.residual ? q.residual : q.y, q.c_out2_pitch, M, kChunkN, 32, err)) return false;
    L->p2.bias2 = q.bias2;
    L->p2.act2 = q.act2;
    L->p2.clip2_lo = q.clip2_lo;
    L->p2.clip2_hi = q.clip2_hi;
    L->p2.subtiles = (q.c_out2 + kBlockN2 - 1) / kBlockN2;
    L->p2.has_residual = q.residual ? 1 : 0;
    L->p2.l2_hints = q.l2_hints;
    const int items = (tmp.p.num_m_tiles + 1) / 2;
    L->grid = 2 * std::min(items, num_sms / 2);
    L->flops = tmp.flops + 2.0 * double(M) * q.c_out2 * a.c_out;
    cudaError_t e = a.c_out == 64 ? set_attr_t<64>() : set_attr_t<128>();
    if (e != cudaSuccess) { if (err) *err = std::string("conv_b2b: cudaFuncSetAttribute: ") + cudaGetErrorString(e); return false; }
    return true;
}

cudaError_t conv_b2b_launch(const ConvB2bLaunch& L, cudaStream_t stream) {
    return L.n1 == 64 ? launch_t<64>(L, stream) : launch_t<128>(L, stream);
}

}  // namespace k
}  // namespace smelter
