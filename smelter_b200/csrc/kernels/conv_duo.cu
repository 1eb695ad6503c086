// Half-footprint two-CTA convolution: the tcgen05 implicit GEMM of conv_pair.cu (256 x BLOCK_N tiles, `tcgen05.mma.cta_group::2`, each
// CTA loads its own activation rows and half of the weight tile) re-cut so that TWO CTAs fit on one SM:
//   256 threads (8 warps), <= 113 KiB of shared memory, <= 256 TMEM columns per CTA.
// Why: a convolution launch spends 4-5 us outside its k-loop (launch gap, first TMA round trip, last tile's epilogue, teardown) and
// ResNet-50 at batch 32 is 50 such launches whose k-loops are 3-7 us.  With one 227 KiB CTA per SM nothing can overlap those phases:
// the next kernel's CTAs cannot become resident before this kernel's exit, and encodes of different batches on different streams
// serialise.  With half-size CTAs (a) a layer with more than 74 work items keeps two tiles in flight per SM pair, one CTA's epilogue
// behind the other's k-loop, and (b) kernels of independent encodes (another stream) fill each other's bubbles.
//
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp id), 4 MMA issuer (leader CTA) + TMEM allocator, 5-6 TMA producers of the
// activation operand, 7 TMA producer of the weight operand.  Barrier protocol as conv_pair.cu (leader = cluster rank 0).
//
// The residual add runs on the tensor core: after the main (and projection-shortcut) k-blocks, BLOCK_N / 64 more k-blocks take the
// residual tile [256 x 64] as the A operand against 64 columns of an identity matrix as B, so D += R lands in the fp32 accumulator
// exactly.  The epilogue is the same straight-line bias / activation / fp16 / TMA-store block for every layer, there is no
// residual staging buffer (the ring is one stage deeper for it) and no second template variant.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

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
constexpr int kThreads = 256;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kAProdWarp0 = 5;
constexpr int kAProducers = 2;
constexpr int kBProdWarp = 7;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
constexpr int kChunkN = 64;
constexpr uint32_t kEpiBufBytes = 32 * kChunkN * 2;
constexpr uint32_t kBiasSlotBytes = 32 * 4;   // half a chunk of fp32 bias values per warp, refilled between the halves
constexpr uint32_t kBarrierBytes = 256;
// two CTAs per SM: 2 x (dynamic + 1 KiB reserved) <= 228 KiB
constexpr uint32_t kSmemLimit = 113 * 1024;
constexpr int kIdentN = 256;

__device__ __half g_identity[kIdentN * kIdentN];

__global__ void identity_fill_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kIdentN * kIdentN) g_identity[i] = __float2half_rn((i / kIdentN) == (i % kIdentN) ? 1.f : 0.f);
}

template <int BLOCK_N>
struct Cfg {
    static constexpr uint32_t kBBytes = (BLOCK_N / 2) * kBlockK * 2;  // this CTA's half of the weight tile
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kEpiBytes = kEpiWarps * (kEpiBufBytes + kBiasSlotBytes);
    static constexpr int kStagesFit = int((kSmemLimit - kEpiBytes - kBarrierBytes) / kStageBytes);
    static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
    static constexpr int kAccBufs = BLOCK_N <= 128 ? 2 : 1;  // 256-column tiles have one accumulator: the co-resident CTA is the overlap
    static constexpr uint32_t kTmemCols = BLOCK_N <= 64 ? 128 : 256;
    static constexpr int kChunks = BLOCK_N / kChunkN;
    static constexpr size_t kSmemBytes = size_t(kStages) * kStageBytes + kEpiBytes + kBarrierBytes;
};

enum ProducerKind : int { PROD_A_TILED = 0, PROD_A_IM2COL = 1, PROD_B = 2 };

// One elected thread per producer warp.  Work items are (m-pair, n-tile): cluster c visits items c, c + #clusters, ...; this CTA's
// rows are m-tile 2 * pair + rank, its weight half is rows n0 + rank * BLOCK_N / 2.  A producer owns every n_prod-th k-block of the
// concatenated k-block sequence of its cluster's items: [main | projection shortcut | residual] per item.
template <int KIND, int BLOCK_N>
__device__ __forceinline__ void produce(const CUtensorMap* tm, const CUtensorMap* tm_side, const CUtensorMap* tm_res, const ConvKernelParams& p,
                                        uint32_t smem_base, uint32_t bar_base, int me, int n_prod, int num_items, int main_kb, uint32_t rank) {
    using C = Cfg<BLOCK_N>;
    constexpr uint32_t kTxBytes = KIND == PROD_B ? C::kBBytes : kABytes;
    const int n_clusters = int(gridDim.x) >> 1;
    const int kpt = p.kblocks_per_tap, taps_w = p.taps_w, nn = p.num_n_tiles;
    uint32_t stage = uint32_t(me), phase = 0;
    uint32_t full_addr = bar_base + 8u * stage;  // local address of full[stage]; the leader's copy is full_addr & kPeerMask
    uint32_t dst = smem_base + stage * C::kStageBytes + (KIND == PROD_B ? kABytes : 0u);
    int kb = me;
    const int side_kb = p.side_kb;
    const int side_end = main_kb + side_kb;
    const int num_kb = side_end + p.res_kb;
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
    // wait for the slot, announce the bytes of BOTH CTAs' loads of this operand on the leader's barrier (the peer only issues loads)
    auto acquire = [&]() -> uint32_t {
        mbar_wait_bounded(full_addr + 8u * C::kStages, phase ^ 1u);  // local empty[stage]: the pair's MMAs have consumed it
        if (rank == 0) mbar_expect_tx(full_addr, 2u * kTxBytes);
        return full_addr & kPeerMask;
    };
    for (int item = int(blockIdx.x) >> 1; item < num_items; item += n_clusters) {
        if (kb < num_kb) {
            const int pair = item / nn, n_tile = item - pair * nn;
            const int m0 = (2 * pair + int(rank)) * kBlockM;
            const int n0 = n_tile * BLOCK_N + int(rank) * (BLOCK_N / 2);
            int cblk = 0, tap = 0, fs = 0, fr = 0;
            int img = 0, base_h = 0, base_w = 0;
            if (KIND != PROD_A_TILED && kb < main_kb) {
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
            for (; kb < main_kb; kb += n_prod) {
                const uint32_t leader_full = acquire();
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
            if (side_kb > 0 && kb < side_end) {
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
                for (; kb < side_end; kb += n_prod) {
                    const uint32_t leader_full = acquire();
                    const int c0 = (kb - main_kb) * kBlockK;
                    if (KIND == PROD_B) tma2_load_3d(tm_side, leader_full, dst, c0, 0, n0);
                    else if (side_im2col) tma2_load_im2col_4d(tm_side, leader_full, dst, c0, base_w, base_h, img, uint16_t(0), uint16_t(0));
                    else tma2_load_2d(tm_side, leader_full, dst, c0, m0);
                    advance();
                }
            }
            // ---- residual: k-block j adds columns 64 j .. 64 j + 63 of the tile: A = residual[m0.., n_tile * BLOCK_N + 64 j ..],
            //      B = rows (this CTA's half of the tile's columns) x columns 64 j .. of the identity ----
#pragma unroll 1
            for (; kb < num_kb; kb += n_prod) {
                const uint32_t leader_full = acquire();
                const int c0 = (kb - side_end) * kBlockK;
                if (KIND == PROD_B) tma2_load_2d(tm_res, leader_full, dst, c0, int(rank) * (BLOCK_N / 2));
                else tma2_load_2d(tm_res, leader_full, dst, n_tile * BLOCK_N + c0, m0);
                advance();
            }
        }
        kb -= num_kb;
    }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 2)
conv_duo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_ident, const __grid_constant__ CUtensorMap tm_a2,
                const __grid_constant__ CUtensorMap tm_b2, const ConvKernelParams p) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) __trap();
    const uint32_t epi_base = smem_base + C::kStages * C::kStageBytes;
    const uint32_t bias_base = epi_base + kEpiWarps * kEpiBufBytes;
    const uint32_t bar_base = epi_base + C::kEpiBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);
    static_assert(8u * (2 * C::kStages + 5) <= kBarrierBytes, "barrier region too small");
    static_assert(C::kSmemBytes <= kSmemLimit, "shared memory budget (two CTAs per SM)");
    static_assert(C::kStages >= kAProducers, "ring shallower than the producer count");
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
    const int total_kb = main_kb + p.side_kb + p.res_kb;
    const int my_tiles = cluster_id < num_items ? (num_items - 1 - cluster_id) / n_clusters + 1 : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_a);
        prefetch_tensormap(&tm_b);
        prefetch_tensormap(&tm_out);
        if (p.res_kb) {
            prefetch_tensormap(&tm_res);
            prefetch_tensormap(&tm_ident);
        }
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
            mbar_init(tmem_empty_bar(lane - C::kStages), 2 * kEpiWarps);  // one arrival per epilogue warp of both CTAs (leader's copy)
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        tmem2_alloc(tmem_slot, C::kTmemCols);
        tmem2_relinquish();
    }
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers exist before anybody arrives remotely or multicasts
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Programmatic dependent launch: the next kernel of this stream is released when every CTA has reached the epilogue of its LAST
    // tile (below), not at kernel start: a dependent released early only sits in griddepcontrol.wait on the SM's second CTA slot,
    // which is exactly the slot a kernel of another stream (an independent encode) could be computing in.
    if (p.use_pdl && my_tiles == 0) grid_dep_launch_dependents();

    if (warp >= kAProdWarp0) {
        const bool is_a = warp < kBProdWarp;
        if (elect_one()) {
            if (is_a) {
                // everything the previous grid wrote is read through these loads (activations, shortcut input, residual)
                if (p.use_pdl) grid_dep_wait();
                const int me = warp - kAProdWarp0;
                if (p.mode == CONV_MODE_TILED) produce<PROD_A_TILED, BLOCK_N>(&tm_a, &tm_a2, &tm_res, p, smem_base, bar_base, me, kAProducers, num_items, main_kb, rank);
                else produce<PROD_A_IM2COL, BLOCK_N>(&tm_a, &tm_a2, &tm_res, p, smem_base, bar_base, me, kAProducers, num_items, main_kb, rank);
            } else {
                produce<PROD_B, BLOCK_N>(&tm_b, &tm_b2, &tm_ident, p, smem_base, bar_base, 0, 1, num_items, main_kb, rank);
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
            auto kblock = [&](uint32_t tmem_d, uint32_t first_accumulate) {
                if (!ready) mbar_wait_bounded(full_addr, phase);
                tc_fence_after();
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
                const int acc = C::kAccBufs == 2 ? (t & 1) : 0;
                const uint32_t use = uint32_t(C::kAccBufs == 2 ? (t >> 1) : t);
                mbar_wait_bounded(tmem_empty_bar(acc), (use & 1u) ^ 1u);
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
        // ================= epilogue (warps 0..3), each CTA drains its own 128 accumulator rows =================
        const int ew = warp;
        const uint32_t buf = epi_base + uint32_t(ew) * kEpiBufBytes;
        const uint32_t bias_slot = bias_base + uint32_t(ew) * kBiasSlotBytes;
        const uint32_t row_off = uint32_t(lane) * 128u;
        const uint32_t sw = uint32_t(lane & 7);
        const bool is_sigmoid = p.act == ACT_SIGMOID;
        const __half2 lo2 = __float2half2_rn(p.act == ACT_RELU ? 0.f : (p.act == ACT_CLIP ? p.clip_lo : -INFINITY));
        const __half2 hi2 = __float2half2_rn(p.act == ACT_CLIP ? p.clip_hi : INFINITY);
        auto load_bias = [&](int col) {  // this lane's two bias values of a chunk: columns col + lane and col + 32 + lane
            float2 b = make_float2(__ldg(p.bias + col + lane), __ldg(p.bias + col + 32 + lane));
            if (p.bias2) {  // projection shortcut: its bias joins before the activation
                b.x += __ldg(p.bias2 + col + lane);
                b.y += __ldg(p.bias2 + col + 32 + lane);
            }
            return b;
        };
        float2 bias_next = make_float2(0.f, 0.f);
        if (my_tiles > 0) bias_next = load_bias((cluster_id % p.num_n_tiles) * BLOCK_N);
        for (int t = 0; t < my_tiles; ++t) {
            const int acc = C::kAccBufs == 2 ? (t & 1) : 0;
            const uint32_t use = uint32_t(C::kAccBufs == 2 ? (t >> 1) : t);
            const int work = cluster_id + t * n_clusters;
            const int pair = work / p.num_n_tiles;
            const int n_tile = work - pair * p.num_n_tiles;
            const int m_row0 = (2 * pair + int(rank)) * kBlockM + ew * 32;
            mbar_wait_bounded(tmem_full_bar(acc), use & 1u);
            tc_fence_after();
            if (p.use_pdl && t + 1 == my_tiles && ew == 0 && lane == 0) grid_dep_launch_dependents();
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * BLOCK_N);
#pragma unroll 1
            for (int c = 0; c < C::kChunks; ++c) {
                const int col0 = n_tile * BLOCK_N + c * kChunkN;
                const float2 bias_cur = bias_next;
                {  // bias of the next chunk (this tile's, or the first chunk of the next tile)
                    int ncol = col0 + kChunkN;
                    if (c + 1 == C::kChunks) ncol = t + 1 < my_tiles ? ((work + n_clusters) % p.num_n_tiles) * BLOCK_N : 0;
                    bias_next = load_bias(ncol);
                }
                uint32_t v[kChunkN];
                tmem_ld_32(taddr + uint32_t(c * kChunkN), v);
                tmem_ld_32(taddr + uint32_t(c * kChunkN + 32), v + 32);
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_slot + uint32_t(lane) * 4u), "f"(bias_cur.x) : "memory");
                tmem_ld_wait();
                if (c == C::kChunks - 1) {  // hand the accumulator back to the leader's MMA warp (one arrival per warp; remote from the peer CTA)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(acc) & kPeerMask);
                }
                __syncwarp();
                uint4 out[kChunkN / 8];
                epilogue_math<32, false>(*reinterpret_cast<const uint32_t(*)[32]>(&v[0]), *reinterpret_cast<uint4(*)[4]>(&out[0]), 0u, sw, bias_slot, is_sigmoid,
                                         lo2, hi2);
                __syncwarp();
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_slot + uint32_t(lane) * 4u), "f"(bias_cur.y) : "memory");
                __syncwarp();
                epilogue_math<32, false>(*reinterpret_cast<const uint32_t(*)[32]>(&v[32]), *reinterpret_cast<uint4(*)[4]>(&out[4]), 0u, sw, bias_slot, is_sigmoid,
                                         lo2, hi2);
                if (lane == 0) tma_store_wait_read<0>();  // the previous chunk's store has read the staging tile
                __syncwarp();
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
    cluster_sync_all();  // nobody exits (or frees TMEM) while the other CTA may still signal its barriers or read its operands
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem2_dealloc(tmem_base, C::kTmemCols);
    }
}

template <int BLOCK_N>
cudaError_t launch_t(const ConvTcLaunch& L, int grid, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg<BLOCK_N>::kSmemBytes;
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
    return cudaLaunchKernelEx(&cfg, conv_duo_kernel<BLOCK_N>, L.tm_a, L.tm_b, L.tm_out, L.tm_res, L.tm_ident, L.tm_a2, L.tm_b2, L.p);
}

template <int BLOCK_N>
cudaError_t set_attr_t() {
    cudaError_t e = cudaFuncSetAttribute(conv_duo_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N>::kSmemBytes));
    if (e != cudaSuccess) return e;
    // the carve-out must leave room for two of these CTAs per SM
    return cudaFuncSetAttribute(conv_duo_kernel<BLOCK_N>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

}  // namespace

bool conv_duo_supported(int block_n, int splits) { return splits == 1 && (block_n == 64 || block_n == 128 || block_n == 256); }

cudaError_t conv_duo_set_attr(int block_n) {
    switch (block_n) {
        case 64: return set_attr_t<64>();
        case 128: return set_attr_t<128>();
        case 256: return set_attr_t<256>();
        default: return cudaErrorInvalidValue;
    }
}

// 256 x 256 fp16 identity in device memory (per device, filled on first use): the B operand of the residual k-blocks.
cudaError_t conv_duo_identity(const __half** out) {
    static std::mutex mu;
    static bool ready[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    void* addr = nullptr;
    e = cudaGetSymbolAddress(&addr, g_identity);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && !ready[dev]) {
        identity_fill_kernel<<<(kIdentN * kIdentN + 255) / 256, 256>>>();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return e;
        ready[dev] = true;
    }
    *out = static_cast<const __half*>(addr);
    return cudaSuccess;
}

cudaError_t conv_duo_launch(const ConvTcLaunch& L, int num_sms, cudaStream_t stream) {
    const int pairs = (L.p.num_m_tiles + 1) / 2;
    const long items = long(pairs) * L.p.num_n_tiles;
    // up to two clusters per SM pair
    const int grid = 2 * int(std::min<long>(items, 2 * (num_sms / 2)));
    switch (L.block_n) {
        case 64: return launch_t<64>(L, grid, stream);
        case 128: return launch_t<128>(L, grid, stream);
        case 256: return launch_t<256>(L, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace k
}  // namespace smelter
