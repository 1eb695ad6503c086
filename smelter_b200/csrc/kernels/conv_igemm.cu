// Implicit-GEMM convolution / fully-connected kernel for sm_100a.
//
// Replaces the arithmetic behind `MPSCNNConvolutionNode` / `MPSCNNFullyConnectedNode`
// (reference Sources/Smelter/Converters.swift:253-256, 299-302), which lives in Apple's closed MPS.
//
//   D[M = N*P*Q, Cout] = A[M, K = taps * Cin_pad] * W[Cout, K]^T   (+ bias, + residual, activation) -> fp16 NHWC
//
// * A is never materialised: a TMA *im2col* tensor map over the NHWC activation delivers, per filter tap and
//   64-channel block, a [128 pixel x 64 channel] tile straight into 128B-swizzled shared memory (zero fill for
//   padding, batch/row wrap-around handled by the TMA bounding box).  1x1/stride-1 layers use a plain 2-D tile map.
//   Small-Cin layers (network stems, Cin <= 8) use the same im2col path over a *virtual* tensor whose "channel"
//   axis is one filter row (S taps x 8 channels are contiguous in NHWC), so a 7x7x3 stem needs 7 K-blocks, not 49.
// * W is pre-packed [Cout][tap][Cin_pad] fp16 (the OHWI order ONNX2MPS.py:75 produces) and fetched with a 3-D map.
// * tcgen05.mma (kind::f16, fp32 accumulate) issued by one thread; accumulators double-buffered in TMEM so the
//   epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, one per SM.
// * Warp roles (16 warps): 0-3 TMA producers of the activation operand, 4-11 epilogue (TMEM -> regs -> bias / residual /
//   activation -> fp16 -> swizzled smem -> TMA store), 12 MMA issuer + TMEM allocator, 13-15 TMA producers of the weight operand.
#include "conv_igemm.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ptx.cuh"

namespace smelter {
namespace k {

using namespace ptx;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // fp16 elements = one 128-byte swizzle row
// 16 warps (512 threads x 128 registers = the whole register file): 0-3 TMA producers of the A (activation) operand,
// 4-11 epilogue (two groups of four), 12 MMA issuer + TMEM allocator, 13-15 TMA producers of the B (weight) operand.
// Several producers because TMA operations issued by one thread complete strictly one after another (~0.35 us each on
// B200, whatever the box size — measured with tools/tma_probe.py): a single issuing thread caps an SM at one 16 KiB
// box per 0.35 us (48 GB/s); independent issuers scale that up to the L2 limit (~125 GB/s per SM with all SMs busy).
constexpr int kThreads = 512;
constexpr int kNumProducers = 4;      // A operand (im2col boxes are the slower ones to issue)
constexpr int kNumBProducers = 3;     // B operand
constexpr int kBProducerWarp0 = 13;
constexpr int kEpilogueWarp0 = 4;
constexpr int kEpilogueWarps = 8;
constexpr int kMmaWarp = 12;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
constexpr int kChunkN = 64;                           // epilogue column chunk = one 128-byte row of fp16
constexpr uint32_t kEpiBufBytes = 32 * kChunkN * 2;   // one warp's [32 rows x 64 cols] staging tile
constexpr uint32_t kBiasSlotBytes = kChunkN * 4;     // one warp's bias values for the current chunk (fp32)
constexpr uint32_t kBarrierBytes = 512;
constexpr uint32_t kSmemLimit = 227 * 1024;

// Shared memory: [operand ring: kStages x (A tile + B tile)] [epilogue staging] [bias slots] [barriers].
// Whatever the epilogue does not need goes to the ring: the k-loop is bound by (bytes in flight) / (L2 latency), so every
// extra stage is throughput.  Without a residual each epilogue warp stages through ONE 4 KiB buffer (results are packed in
// registers and written once the previous TMA store has read the buffer); with a residual it keeps two, because the
// residual chunk of the next item is TMA-prefetched into the other buffer while this one is computed and stored.
template <int BLOCK_N, bool HAS_RES>
struct Cfg {
    static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBufs = HAS_RES ? 2 : 1;
    static constexpr uint32_t kEpiBytes = kEpilogueWarps * (kEpiBufs * kEpiBufBytes + kBiasSlotBytes);
    static constexpr int kStagesFit = int((kSmemLimit - kEpiBytes - kBarrierBytes) / kStageBytes);
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
    // Producers in use: never more than the ring has stages.  A producer only knows (from its own previous wait) that
    // every k-block up to g - kStages has been consumed; its next k-block g + P needs k-block g + P - kStages consumed and
    // the parity wait is only unambiguous if k-block g + P - 2*kStages already is, i.e. P <= kStages.
    static constexpr int kProducers = kStages < kNumProducers ? kStages : kNumProducers;
    static constexpr int kBProducers = kStages < kNumBProducers ? kStages : kNumBProducers;
    static constexpr int kChunks = BLOCK_N >= kChunkN ? BLOCK_N / kChunkN : 1;
    static constexpr int kChunkCols = BLOCK_N >= kChunkN ? kChunkN : BLOCK_N;  // columns of a chunk that carry data
    // the dynamic shared-memory window starts 1024-byte aligned (checked in the kernel), so no alignment slack is reserved
    static constexpr size_t kSmemBytes = size_t(kStages) * kStageBytes + kEpiBytes + kBarrierBytes;
};

// Instrumented builds only (-DSMELTER_CONV_INSTRUMENT=1): %globaltimer stamps of CTA 0 and the SMELTER_CONV_DEBUG ablation flags
// (16 = producers arrive without loading, 32 = no operand barriers at all, 64 = no MMAs).  Compiled out of the product kernel:
// every extra instruction in the single-thread MMA issue loop is on the critical path (measured: the loop, not the tensor pipe,
// set the k-block rate when it carried these checks).
#ifndef SMELTER_CONV_INSTRUMENT
#define SMELTER_CONV_INSTRUMENT 0
#endif
constexpr bool kInstr = SMELTER_CONV_INSTRUMENT != 0;
constexpr int kChainLaunches = 256;
unsigned long long* g_chain_buf = nullptr;
int g_chain_seq = 0;
char g_chain_desc[kChainLaunches][96];
constexpr int kMmaLoopVariant = 1;  // MMA issue loop: 0 = one elected lane, 1 = same unrolled by two, 2 = whole warp + elected issue

__device__ __forceinline__ void stamp(const ConvKernelParams& p, int slot) {
    if (kInstr && p.timeline && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.timeline[slot] = t;
    }
}
__device__ __forceinline__ bool dbg(const ConvKernelParams& p, int flag) { return kInstr && (p.debug_flags & flag); }

// Sigmoid is rare on this path (no BASELINE model fuses it): keep it out of line so the hot epilogue stays small.
__device__ __noinline__ float sigmoid1(float v) { return 1.f / (1.f + __expf(-v)); }

// One [32 row x kCols column] chunk of the epilogue for one warp: v = fp32 accumulators of this lane's row; +bias (fp32, from
// the warp's smem bias slot), +residual (fp32, read from the staging row where TMA put it), activation, fp16, packed into
// registers (out[g] = columns 8g..8g+7).  Software-pipelined over the 8-column groups: the bias (and residual) of group g+1
// are fetched from shared memory while group g is computed, so no LDS latency sits on the dependency chain.
template <int kCols, bool HAS_RES>
__device__ __forceinline__ void epilogue_math(const uint32_t (&v)[kCols], uint4 (&out)[kCols / 8], uint32_t rowbuf, uint32_t sw, uint32_t bias_slot,
                                              bool is_sigmoid, __half2 lo2, __half2 hi2) {
    constexpr int kGroups = kCols / 8;
    uint4 nb0 = ld_shared_v4(bias_slot), nb1 = ld_shared_v4(bias_slot + 16u);
    uint4 nrv = make_uint4(0u, 0u, 0u, 0u);
    if (HAS_RES) nrv = ld_shared_v4(rowbuf + (sw << 4));
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const uint4 bq0 = nb0, bq1 = nb1, rv = nrv;
        if (g + 1 < kGroups) {
            nb0 = ld_shared_v4(bias_slot + uint32_t(g + 1) * 32u);
            nb1 = ld_shared_v4(bias_slot + uint32_t(g + 1) * 32u + 16u);
            if (HAS_RES) nrv = ld_shared_v4(rowbuf + ((uint32_t(g + 1) ^ sw) << 4));
        }
        float f[8];
        f[0] = __uint_as_float(v[g * 8 + 0]) + __uint_as_float(bq0.x); f[1] = __uint_as_float(v[g * 8 + 1]) + __uint_as_float(bq0.y);
        f[2] = __uint_as_float(v[g * 8 + 2]) + __uint_as_float(bq0.z); f[3] = __uint_as_float(v[g * 8 + 3]) + __uint_as_float(bq0.w);
        f[4] = __uint_as_float(v[g * 8 + 4]) + __uint_as_float(bq1.x); f[5] = __uint_as_float(v[g * 8 + 5]) + __uint_as_float(bq1.y);
        f[6] = __uint_as_float(v[g * 8 + 6]) + __uint_as_float(bq1.z); f[7] = __uint_as_float(v[g * 8 + 7]) + __uint_as_float(bq1.w);
        if (HAS_RES) {
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 r2 = __half22float2(rh[i]);
                f[2 * i] += r2.x;
                f[2 * i + 1] += r2.y;
            }
        }
        if (is_sigmoid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = sigmoid1(f[i]);
        }
        __half2* oh = reinterpret_cast<__half2*>(&out[g]);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh[i] = __hmin2(__hmax2(__floats2half2_rn(f[2 * i], f[2 * i + 1]), lo2), hi2);
    }
}
// Packed results -> this lane's row of the staging tile (128B-swizzled: 16-byte piece g of row r sits at piece g ^ (r & 7)).
template <int kGroups>
__device__ __forceinline__ void epilogue_stage(const uint4 (&out)[kGroups], uint32_t rowbuf, uint32_t sw) {
#pragma unroll
    for (int g = 0; g < kGroups; ++g) st_shared_v4(rowbuf + ((uint32_t(g) ^ sw) << 4), out[g]);
}

// ---- TMA producer loop (one thread) -------------------------------------------------------------------------
// The issuing thread's instruction stream is the producer's throughput: measured on B200, a loop that re-derived the filter
// tap with an integer division and walked every k-block (own or not) spent ~0.35 us per TMA instruction, slower than the
// tensor core consumes a k-block.  So: the thread visits only its own k-blocks (stride n_prod), all positions (channel
// block, filter tap, ring stage, barrier addresses) advance by adds and compares, per-item divisions are done once per
// output tile, and the three operand kinds get their own loop bodies.
enum ProducerKind : int { PROD_A_TILED = 0, PROD_A_IM2COL = 1, PROD_B = 2 };

template <int KIND, int BLOCK_N, bool HAS_RES>
__device__ __forceinline__ void produce(const CUtensorMap* tm, const ConvKernelParams& p, uint32_t smem_base, uint32_t bar_base, int me, int n_prod,
                                        int num_tiles, int total_kb) {
    using C = Cfg<BLOCK_N, HAS_RES>;
    constexpr uint32_t kTxBytes = KIND == PROD_B ? C::kBBytes : kABytes;
    const int grid = int(gridDim.x);
    const int kpt = p.kblocks_per_tap, taps_w = p.taps_w, nn = p.num_n_tiles;
    // ring position of my next k-block
    uint32_t stage = uint32_t(me), phase = 0;
    uint32_t full_addr = bar_base + 8u * stage;                                            // empty barrier = full_addr + 8 * kStages
    uint32_t dst = smem_base + stage * C::kStageBytes + (KIND == PROD_B ? kABytes : 0u);
    int kb = me;  // my next k-block, relative to the current item's first
    // (m_tile, n_tile) of the current item advance by a constant per item when there is no k-split
    int item = int(blockIdx.x);
    int m_tile = item / nn, n_tile = item - m_tile * nn;
    const int dm = grid / nn, dn = grid - dm * nn;
    for (; item < num_tiles; item += grid) {
        int kb0 = 0, num_kb = total_kb;
        if (p.splits > 1) {
            const int tile = item / p.splits;
            const int split = item - tile * p.splits;
            kb0 = split * p.kb_per_split;
            num_kb = min(p.kb_per_split, total_kb - kb0);
            m_tile = tile / nn;
            n_tile = tile - m_tile * nn;
        }
        if (kb < num_kb) {
            const int m0 = m_tile * kBlockM;
            const int n0 = n_tile * BLOCK_N;
            // position of k-block kb0 + kb inside the filter: channel block, tap (linear), tap column / row
            int cblk = 0, tap = 0, fs = 0, fr = 0;
            int img = 0, base_h = 0, base_w = 0;
            if (KIND != PROD_A_TILED) {
                const int k_abs = kb0 + kb;
                if (k_abs < 8) {  // the usual case (kb < n_prod, no split): a few carries instead of two divisions
                    cblk = k_abs;
                    while (cblk >= kpt) { cblk -= kpt; ++tap; ++fs; }
                    while (fs >= taps_w) { fs -= taps_w; ++fr; }
                } else {
                    tap = k_abs / kpt; cblk = k_abs - tap * kpt;
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
            for (; kb < num_kb; kb += n_prod) {
                mbar_wait(full_addr + 8u * C::kStages, phase ^ 1u);
                if (kInstr && dbg(p, 16)) {
                    mbar_arrive(full_addr);
                } else {
                    mbar_expect_tx(full_addr, kTxBytes);
                    if (KIND == PROD_A_TILED) tma_load_2d(tm, full_addr, dst, (kb0 + kb) * kBlockK, m0);
                    else if (KIND == PROD_A_IM2COL) tma_load_im2col_4d(tm, full_addr, dst, cblk * kBlockK, base_w, base_h, img, uint16_t(fs * p.dil_w), uint16_t(fr * p.dil_h));
                    else tma_load_3d(tm, full_addr, dst, cblk * kBlockK, tap, n0);
                }
                // advance everything by n_prod k-blocks
                stage += uint32_t(n_prod);
                full_addr += 8u * uint32_t(n_prod);
                dst += uint32_t(n_prod) * C::kStageBytes;
                if (stage >= uint32_t(C::kStages)) {
                    stage -= uint32_t(C::kStages);
                    phase ^= 1u;
                    full_addr -= 8u * uint32_t(C::kStages);
                    dst -= uint32_t(C::kStages) * C::kStageBytes;
                }
                if (KIND != PROD_A_TILED) {
                    cblk += n_prod;
                    while (cblk >= kpt) { cblk -= kpt; ++tap; ++fs; }
                    while (fs >= taps_w) { fs -= taps_w; ++fr; }
                }
            }
        }
        kb -= num_kb;
        n_tile += dn;
        m_tile += dm;
        if (n_tile >= nn) { n_tile -= nn; ++m_tile; }
    }
}

template <int BLOCK_N, bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res, const ConvKernelParams p) {
    using C = Cfg<BLOCK_N, HAS_RES>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment; the dynamic window starts aligned (no static shared memory in this kernel).
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) __trap();
    const uint32_t epi_base = smem_base + C::kStages * C::kStageBytes;  // 1024-aligned: stage sizes are multiples of 1024
    const uint32_t bias_base = epi_base + kEpilogueWarps * C::kEpiBufs * kEpiBufBytes;
    const uint32_t bar_base = epi_base + C::kEpiBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
    auto res_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::kStages + 4 + w * 2 + b); };  // w in [0, 8)
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 20);
    const uint32_t flag_slot = bar_base + 8u * (2 * C::kStages + 21);  // two 8-byte slots: "this group reduces the tile" (split-K)
    static_assert(8u * (2 * C::kStages + 23) <= kBarrierBytes, "barrier region too small");
    static_assert(C::kSmemBytes <= kSmemLimit, "shared memory budget");
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // A work item is (output tile, k-split): with p.splits > 1 several CTAs accumulate disjoint k-block ranges of the same
    // tile and the last one to finish reduces the fp32 partials (see the split-K epilogue).  "tile" below means work item.
    const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.splits;
    const int total_kb = p.num_taps * p.kblocks_per_tap;
    const int my_tiles = int(blockIdx.x) < num_tiles ? (num_tiles - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;
    if (threadIdx.x == 0) stamp(p, 0);  // kernel entry

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_a);
        prefetch_tensormap(&tm_b);
        prefetch_tensormap(&tm_out);
        if (HAS_RES) prefetch_tensormap(&tm_res);
    }
    if (warp == 1) {  // one barrier per lane, all initialised in parallel
        if (lane < C::kStages) {
            mbar_init(full_bar(lane), 2);   // one arrive.expect_tx from the A producer, one from the B producer
            mbar_init(empty_bar(lane), 1);
        } else if (lane < C::kStages + 2) {
            mbar_init(tmem_full_bar(lane - C::kStages), 1);
            // arrivals per accumulator hand-back: one group of four warps, or both when they share every tile (see the epilogue)
            mbar_init(tmem_empty_bar(lane - C::kStages), (C::kChunks >= 2 && p.splits == 1) ? 256 : 128);
        } else if (lane >= 16) {
            mbar_init(res_bar((lane - 16) >> 1, lane & 1), 1);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(tmem_slot, C::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (threadIdx.x == 0) stamp(p, 1);  // prologue done
    // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no global data and may overlap the
    // previous kernel's tail.  From here on each role waits for the previous grid only if it touches what that grid produced:
    // the activation (A) producers and the epilogue (residual loads, output stores, split-K workspace) do; the weight (B)
    // producers and the MMA issuer do not, so the weight tiles of the first k-blocks are already in flight when the wait ends.
    if (p.use_pdl) grid_dep_launch_dependents();

    if (warp < kNumProducers || warp >= kBProducerWarp0) {
        // ================= TMA producers =================
        // Producer i of an operand issues the loads of k-blocks i, i + n, i + 2n, ... of the CTA's k-block stream (n producers
        // per operand).  Both operands of a k-block land in the same stage and complete the same full barrier.  One elected
        // lane per producer warp runs the loop; see produce() for why the loop is written the way it is.
        const bool is_a = warp < kNumProducers;
        const int me = is_a ? warp : warp - kBProducerWarp0;
        const int n_prod = is_a ? C::kProducers : C::kBProducers;
        if (elect_one() && me < n_prod && !dbg(p, 32)) {
            if (is_a) {
                if (p.use_pdl) grid_dep_wait();
                if (threadIdx.x == 0) stamp(p, 2);  // dependencies resolved
                if (p.mode == CONV_MODE_TILED) produce<PROD_A_TILED, BLOCK_N, HAS_RES>(&tm_a, p, smem_base, bar_base, me, n_prod, num_tiles, total_kb);
                else produce<PROD_A_IM2COL, BLOCK_N, HAS_RES>(&tm_a, p, smem_base, bar_base, me, n_prod, num_tiles, total_kb);
            } else {
                produce<PROD_B, BLOCK_N, HAS_RES>(&tm_b, p, smem_base, bar_base, me, n_prod, num_tiles, total_kb);
            }
        }
    } else if (warp == kMmaWarp) {
        // ================= MMA issuer =================
        // One elected lane issues the tcgen05 instructions.  This single instruction stream feeds the tensor core, so the loop
        // body is kept to the bare sequence (wait, 4 x MMA, commit): descriptor words and barrier addresses advance by adds,
        // the first k-block of a tile (which overwrites the accumulator) is peeled so the accumulate flag is an immediate.
        constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N);
        constexpr uint64_t desc_hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);  // SBO, version, SWIZZLE_128B
        constexpr uint32_t desc_lbo = 1u << 16;
        const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | desc_lbo;
        constexpr uint32_t kStage16 = C::kStageBytes >> 4;
        constexpr uint32_t kB16 = kABytes >> 4;
        uint32_t stage = 0, phase = 0;
        uint32_t a_lo = a_lo0;
        uint32_t full_addr = bar_base;  // full_bar(stage); empty_bar(stage) = full_addr + 8 * kStages
        auto issue = [&](uint32_t tmem_d, uint32_t first_accumulate) {
            const uint64_t a_desc = desc_hi | uint64_t(a_lo);
            const uint64_t b_desc = a_desc + kB16;
            if (!dbg(p, 64)) {
                umma_f16(tmem_d, a_desc, b_desc, idesc, first_accumulate);
                umma_f16(tmem_d, a_desc + 2, b_desc + 2, idesc, 1u);   // +32 bytes along K inside the swizzle atom
                umma_f16(tmem_d, a_desc + 4, b_desc + 4, idesc, 1u);
                umma_f16(tmem_d, a_desc + 6, b_desc + 6, idesc, 1u);
            }
            if (!dbg(p, 32)) umma_commit(full_addr + 8u * C::kStages);  // frees the smem slot once these MMAs retire
            else if (dbg(p, 1024)) umma_commit(res_bar(7, 1));          // experiment: cost of the commit alone (nobody waits on it)
        };
        auto advance = [&]() {
            a_lo += kStage16;
            full_addr += 8u;
            if (++stage == uint32_t(C::kStages)) { stage = 0; phase ^= 1u; a_lo = a_lo0; full_addr = bar_base; }
        };
        auto tile_kb = [&](int t) {
            int num_kb = total_kb;
            if (p.splits > 1) {
                const int split = (int(blockIdx.x) + t * int(gridDim.x)) % p.splits;
                num_kb = min(p.kb_per_split, total_kb - split * p.kb_per_split);
            }
            return num_kb;
        };
        const int variant = kInstr ? ((p.debug_flags >> 8) & 3) : kMmaLoopVariant;
        if (variant == 2) {
            // whole warp walks the loop (uniform registers, no R2UR); one elected lane issues
            for (int t = 0; t < my_tiles; ++t) {
                const int acc = t & 1;
                const int num_kb = tile_kb(t);
                mbar_wait(tmem_empty_bar(acc), ((uint32_t(t) >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * BLOCK_N);
#pragma unroll 1
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (!dbg(p, 32)) mbar_wait(full_addr, phase);
                    tc_fence_after();
                    if (elect_one()) issue(tmem_d, kb > 0 ? 1u : 0u);
                    advance();
                }
                if (elect_one()) umma_commit(tmem_full_bar(acc));
            }
        } else if (elect_one()) {
            // Only the elected lane runs the loop.
            auto kblock = [&](uint32_t tmem_d, uint32_t first_accumulate) {
                if (!dbg(p, 32)) mbar_wait(full_addr, phase);
                tc_fence_after();
                issue(tmem_d, first_accumulate);
                advance();
            };
            for (int t = 0; t < my_tiles; ++t) {
                const int acc = t & 1;                       // accumulator buffer == epilogue group
                const int num_kb = tile_kb(t);
                mbar_wait(tmem_empty_bar(acc), ((uint32_t(t) >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * BLOCK_N);
                kblock(tmem_d, 0u);
                if (t == 0) stamp(p, 3);  // first operands landed
                int kb = 1;
                if (variant == 1) {
#pragma unroll 1
                    for (; kb + 1 < num_kb; kb += 2) {
                        kblock(tmem_d, 1u);
                        kblock(tmem_d, 1u);
                    }
                }
#pragma unroll 1
                for (; kb < num_kb; ++kb) kblock(tmem_d, 1u);
                umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue group `acc`
                if (t == 0) stamp(p, 4);  // all MMAs of the first tile issued
            }
        }
    } else {
        // ================= epilogue (warps 4..11) =================
        // Two groups of four warps; group g drains accumulator buffer g, i.e. this CTA's tiles g, g+2, g+4, ... so the
        // two groups interleave on the four SM sub-partitions and hide each other's latencies.  Within a group each warp
        // owns 32 accumulator rows (its TMEM lane quarter) and streams them out in 64-column chunks:
        //   TMEM -> registers -> +bias (+residual) in fp32 -> fp16 -> activation clamp (packed half2, still in registers) ->
        //   128B-swizzled smem staging -> TMA store.
        // Without a residual the warp has one staging buffer: the packed results wait in registers until the previous chunk's
        // TMA store has read it.  With a residual there are two: the residual chunk of the next item is TMA-prefetched into
        // the other buffer and the result overwrites it in place (every thread reads and writes only its own 16-byte pieces).
        // All global traffic is whole 128-byte lines issued by the TMA unit.
        if (p.use_pdl) grid_dep_wait();
        const int ewarp = warp - kEpilogueWarp0;  // 0..7
        const int group = ewarp >> 2;
        const int ew = ewarp & 3;                 // == warp % 4: the TMEM lane quarter this warp may read
        const uint32_t buf0 = epi_base + uint32_t(ewarp) * uint32_t(C::kEpiBufs) * kEpiBufBytes;
        const uint32_t bias_slot = bias_base + uint32_t(ewarp) * kBiasSlotBytes;
        const uint32_t row_off = uint32_t(lane) * 128u;
        const uint32_t sw = uint32_t(lane & 7);
        const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
        const bool is_sigmoid = p.act == ACT_SIGMOID;
        const __half2 lo2 = __float2half2_rn(p.act == ACT_RELU ? 0.f : (p.act == ACT_CLIP ? p.clip_lo : -INFINITY));
        const __half2 hi2 = __float2half2_rn(p.act == ACT_CLIP ? p.clip_hi : INFINITY);
        // One chunk from accumulators in registers to a committed TMA store.  `b` selects the staging buffer (always 0 without
        // a residual); with a residual the caller has already waited for the chunk's residual to land in that buffer.
        auto finish_chunk = [&](const uint32_t (&v)[C::kChunkCols], int b, int col0, int m_row0) {
            const uint32_t buf = buf0 + uint32_t(b) * kEpiBufBytes;
            uint4 out[C::kChunkCols / 8];
            epilogue_math<C::kChunkCols, HAS_RES>(v, out, buf + row_off, sw, bias_slot, is_sigmoid, lo2, hi2);
            if (!HAS_RES) {
                if (lane == 0) tma_store_wait_read<0>();  // the previous chunk's store has read the (only) buffer
                __syncwarp();
            }
            epilogue_stage<C::kChunkCols / 8>(out, buf + row_off, sw);
            fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
            __syncwarp();              // also: every lane is done with the bias slot before the next chunk's bias is published
            if (lane == 0) {
                // rows >= M and columns >= out_pitch are clipped by the map
                if (p.l2_hints & 1) tma_store_2d_hint(&tm_out, buf, col0, m_row0, pol_keep);
                else tma_store_2d(&tm_out, buf, col0, m_row0);
                tma_store_commit();
            }
        };
        if (p.splits == 1) {
        // Which (tile, chunk) pairs this warp drains.  Tiles of one 64-column chunk: the two groups alternate whole tiles
        // (group g = accumulator buffer g).  Wider tiles: BOTH groups work on every tile, group g taking chunks g, g+2, ... —
        // same throughput, but a tile's epilogue takes half as long, which is what the last tile of a CTA (and every tile of a
        // layer with at most one tile per CTA) exposes.
        constexpr bool kSplit = C::kChunks >= 2;
        constexpr int kCPW = kSplit ? C::kChunks / 2 : C::kChunks;  // chunks per warp per tile
        const int group_tiles = kSplit ? my_tiles : (my_tiles > group ? (my_tiles - group + 1) / 2 : 0);
        const int n_items = group_tiles * kCPW;  // (tile, chunk) stream of this warp
        auto item_coords = [&](int item, int* m_row0, int* col0) {
            const int gt = item / kCPW;           // kCPW is a power of two: shifts
            const int j = item - gt * kCPW;
            const int c = kSplit ? group + 2 * j : j;
            const int tile = int(blockIdx.x) + (kSplit ? gt : 2 * gt + group) * int(gridDim.x);
            const int m_tile = tile / p.num_n_tiles;
            const int n_tile = tile - m_tile * p.num_n_tiles;
            *m_row0 = m_tile * kBlockM + ew * 32;
            *col0 = n_tile * BLOCK_N + c * kChunkN;
        };
        auto prefetch_res = [&](int item) {  // lane 0 only; the target buffer must no longer be read by a TMA store
            int m_row0, col0;
            item_coords(item, &m_row0, &col0);
            const int b = item & 1;
            fence_proxy_async_smem();
            mbar_expect_tx(res_bar(ewarp, b), kEpiBufBytes);
            if (p.l2_hints & 2) tma_load_2d_hint(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, col0, m_row0, pol_drop);
            else tma_load_2d(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, col0, m_row0);
        };
        if (HAS_RES && n_items > 0 && lane == 0) prefetch_res(0);
        // bias of the next chunk, two columns per lane, fetched one chunk ahead (weights: no dependency on the previous kernel)
        float2 bias_next = make_float2(0.f, 0.f);
        if (n_items > 0) {
            int m_row0, col0;
            item_coords(0, &m_row0, &col0);
            bias_next = __ldg(reinterpret_cast<const float2*>(p.bias + col0) + lane);
        }
        uint32_t res_phase = 0;  // bit b = parity of res_bar(ewarp, b)
        int item = 0;
        for (int gt = 0; gt < group_tiles; ++gt) {
            const int acc = kSplit ? (gt & 1) : group;                             // accumulator buffer of this tile
            const uint32_t acc_parity = uint32_t(kSplit ? (gt >> 1) : gt) & 1u;    // how often this warp has used that buffer's barrier
            mbar_wait(tmem_full_bar(acc), acc_parity);  // (one polling lane + nanosleep back-off measured slower than all lanes waiting)
            if (gt == 0 && ewarp == 0 && lane == 0) stamp(p, 5);  // first accumulator ready
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * BLOCK_N);
#pragma unroll 1
            for (int j = 0; j < kCPW; ++j, ++item) {
                int m_row0, col0;
                item_coords(item, &m_row0, &col0);
                const int c = kSplit ? group + 2 * j : j;
                const int b = HAS_RES ? (item & 1) : 0;
                if (HAS_RES && lane == 0 && item + 1 < n_items) {
                    tma_store_wait_read<0>();  // the store of item-1 has finished reading buffer b^1
                    prefetch_res(item + 1);
                }
                // publish this chunk's bias to the warp through smem, then start fetching the next chunk's
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_slot + uint32_t(lane) * 8u), "f"(bias_next.x), "f"(bias_next.y) : "memory");
                if (item + 1 < n_items) {
                    int nm, ncol0;
                    item_coords(item + 1, &nm, &ncol0);
                    bias_next = __ldg(reinterpret_cast<const float2*>(p.bias + ncol0) + lane);
                }
                uint32_t v[C::kChunkCols];
                tmem_ld_32(taddr + uint32_t(c * kChunkN), v);
                if (C::kChunkCols > 32) tmem_ld_32(taddr + uint32_t(c * kChunkN + 32), v + (C::kChunkCols > 32 ? 32 : 0));
                tmem_ld_wait();
                if (j == kCPW - 1) {  // this warp's part of the accumulator is in registers: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tmem_empty_bar(acc));
                }
                if (HAS_RES) {
                    mbar_wait(res_bar(ewarp, b), (res_phase >> b) & 1u);
                    res_phase ^= 1u << b;
                }
                __syncwarp();  // the bias slot is published (and, with a residual, lane 0's wait_group.read covers the warp)
                finish_chunk(v, b, col0, m_row0);
            }
        }
        } else {
            // ---------------- split-K epilogue ----------------
            // Every CTA that worked on a k-range of the tile dumps its raw fp32 accumulator to its slab of the workspace,
            // then bumps the tile's counter; the CTA that arrives last sums all slabs and runs the normal epilogue
            // (bias, residual, activation, TMA store) and resets the counter for the next launch.
            const int group_items = my_tiles > group ? (my_tiles - group + 1) / 2 : 0;
            volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem_raw + (flag_slot + 8u * uint32_t(group) - smem_base));
            const int row = ew * 32 + lane;
            uint32_t acc_phase = 0, res_phase = 0;
            int n_stores = 0;  // TMA stores this warp has committed (selects the staging buffer)
            for (int gi = 0; gi < group_items; ++gi) {
                const int item = int(blockIdx.x) + (2 * gi + group) * int(gridDim.x);
                const int tile = item / p.splits;
                const int m_tile = tile / p.num_n_tiles;
                const int n_tile = tile - m_tile * p.num_n_tiles;
                mbar_wait(tmem_full_bar(group), acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(group * BLOCK_N);
                // workspace layout per work item: [16-byte piece = 4 columns][row], so the 32 lanes of a warp (32 consecutive
                // rows) write / read 512 contiguous bytes per instruction although every lane owns a whole row
                uint4* slab = reinterpret_cast<uint4*>(p.ws) + size_t(item) * (kBlockM * BLOCK_N / 4) + row;
#pragma unroll 1
                for (int c = 0; c < C::kChunks; ++c) {
                    uint32_t v[C::kChunkCols];
                    tmem_ld_32(taddr + uint32_t(c * kChunkN), v);
                    if (C::kChunkCols > 32) tmem_ld_32(taddr + uint32_t(c * kChunkN + 32), v + (C::kChunkCols > 32 ? 32 : 0));
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < C::kChunkCols / 4; ++j)
                        slab[size_t(c * (kChunkN / 4) + j) * kBlockM] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                tc_fence_before();
                mbar_arrive(tmem_empty_bar(group));
                __threadfence();  // partials visible device-wide before the counter moves
                asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
                if (ew == 0 && lane == 0) {
                    const unsigned old = atomicAdd(p.counters + tile, 1u);
                    const bool last = old == unsigned(p.splits - 1);
                    if (last) p.counters[tile] = 0u;
                    *flag = last ? 1u : 0u;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
                if (*flag == 0u) continue;
                __threadfence();
                const int m_row0 = m_tile * kBlockM + ew * 32;
                int col0 = n_tile * BLOCK_N;
                float2 bias_next = __ldg(reinterpret_cast<const float2*>(p.bias + col0) + lane);
                auto load_res = [&](int cc, int b) {  // lane 0 only
                    fence_proxy_async_smem();
                    mbar_expect_tx(res_bar(ewarp, b), kEpiBufBytes);
                    tma_load_2d(&tm_res, res_bar(ewarp, b), buf0 + uint32_t(b) * kEpiBufBytes, n_tile * BLOCK_N + cc * kChunkN, m_row0);
                };
                if (HAS_RES && lane == 0) {
                    tma_store_wait_read<0>();
                    load_res(0, n_stores & 1);
                }
#pragma unroll 1
                for (int c = 0; c < C::kChunks; ++c, ++n_stores, col0 += kChunkN) {
                    const int b = HAS_RES ? (n_stores & 1) : 0;
                    if (HAS_RES && lane == 0 && c + 1 < C::kChunks) {
                        tma_store_wait_read<0>();
                        load_res(c + 1, b ^ 1);
                    }
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bias_slot + uint32_t(lane) * 8u), "f"(bias_next.x), "f"(bias_next.y) : "memory");
                    if (c + 1 < C::kChunks) bias_next = __ldg(reinterpret_cast<const float2*>(p.bias + col0 + kChunkN) + lane);
                    float acc[C::kChunkCols];
#pragma unroll
                    for (int i = 0; i < C::kChunkCols; ++i) acc[i] = 0.f;
                    for (int sp = 0; sp < p.splits; ++sp) {
                        const float4* q = reinterpret_cast<const float4*>(p.ws) + (size_t(tile) * p.splits + sp) * (kBlockM * BLOCK_N / 4) +
                                          size_t(c * (kChunkN / 4)) * kBlockM + row;
#pragma unroll
                        for (int j = 0; j < C::kChunkCols / 4; ++j) {
                            const float4 t4 = __ldcg(q + size_t(j) * kBlockM);
                            acc[4 * j] += t4.x; acc[4 * j + 1] += t4.y; acc[4 * j + 2] += t4.z; acc[4 * j + 3] += t4.w;
                        }
                    }
                    uint32_t v[C::kChunkCols];
#pragma unroll
                    for (int i = 0; i < C::kChunkCols; ++i) v[i] = __float_as_uint(acc[i]);
                    if (HAS_RES) {
                        mbar_wait(res_bar(ewarp, b), (res_phase >> b) & 1u);
                        res_phase ^= 1u << b;
                    }
                    __syncwarp();
                    finish_chunk(v, b, col0, m_row0);
                }
            }
        }
        if (ewarp == 0 && lane == 0) stamp(p, 6);  // last store issued
        if (lane == 0) tma_store_wait_read<0>();  // smem must stay valid until read; global visibility comes with grid completion
        if (ewarp == 0 && lane == 0) stamp(p, 7);  // stores read
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
        if (lane == 0) stamp(p, 8);  // exit
    }
}

// ---- host side -----------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled g_encode_tiled = nullptr;
PFN_encodeIm2col g_encode_im2col = nullptr;
int g_driver_version = 0;

bool load_driver_entry_points(std::string* err) {
    if (g_encode_tiled && g_encode_im2col) return true;
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
        if (err) *err = "cuTensorMapEncodeTiled not available from the CUDA driver";
        return false;
    }
    g_encode_tiled = reinterpret_cast<PFN_encodeTiled>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
        if (err) *err = "cuTensorMapEncodeIm2col not available from the CUDA driver";
        return false;
    }
    g_encode_im2col = reinterpret_cast<PFN_encodeIm2col>(fn);
    cudaDriverGetVersion(&g_driver_version);
    return true;
}

template <int BLOCK_N>
cudaError_t set_attr_t() {
    // per device; cheap, and done at prepare time so that launches are legal inside stream capture
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N, false>::kSmemBytes));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N, true>::kSmemBytes));
}
cudaError_t set_attr(int block_n) {
    switch (block_n) {
        case 32: return set_attr_t<32>();
        case 64: return set_attr_t<64>();
        case 128: return set_attr_t<128>();
        case 256: return set_attr_t<256>();
        default: return cudaErrorInvalidValue;
    }
}

template <int BLOCK_N>
cudaError_t launch_t(const ConvTcLaunch& L, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(L.grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = L.p.has_residual ? Cfg<BLOCK_N, true>::kSmemBytes : Cfg<BLOCK_N, false>::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = L.use_pdl ? 1 : 0;
    if (L.p.has_residual) return cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BLOCK_N, true>, L.tm_a, L.tm_b, L.tm_out, L.tm_res, L.p);
    return cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BLOCK_N, false>, L.tm_a, L.tm_b, L.tm_out, L.tm_res, L.p);
}

}  // namespace

bool conv_tc_encode_2d(CUtensorMap* tm, const __half* base, long cols, long rows, int box_cols, int box_rows, std::string* err) {
    if (!load_driver_entry_points(err)) return false;
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(cols) * 2};
    cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(2d) failed: " + std::to_string(int(r));
        return false;
    }
    return true;
}

int conv_tc_pick_block_n(int c_out, int m_tiles, int num_sms) {
    // Smallest tile that covers Cout for narrow layers; for wide layers pick the width that leaves the fewest
    // idle SMs in the last wave (tiles are persistent-scheduled round-robin over num_sms CTAs).
    if (c_out <= 32) return 32;
    if (c_out <= 64) return 64;
    int best = 128;
    double best_cost = 1e30;
    const int cands[3] = {64, 128, 256};
    for (int bn : cands) {
        const int n_tiles = (c_out + bn - 1) / bn;
        const long tiles = long(n_tiles) * m_tiles;
        const long waves = (tiles + num_sms - 1) / num_sms;
        // cost ~ waves * per-tile time; per-tile time ~ bn columns of MMA (+ a fixed overhead, which
        // penalises very narrow tiles) and the padded columns of the last n-tile are wasted work.
        const double cost = double(waves) * (bn + 24.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

ConvTcPlanInfo conv_tc_plan(const ConvTcProblem& q, int num_sms) {
    const int R = q.k_h, S = q.k_w;
    const int P = (q.h + q.pad_t + q.pad_b - q.dil_h * (R - 1) - 1) / q.stride_h + 1;
    const int Q = (q.w + q.pad_l + q.pad_r - q.dil_w * (S - 1) - 1) / q.stride_w + 1;
    const long M = long(q.n) * std::max(P, 0) * std::max(Q, 0);
    const int m_tiles = int((M + kBlockM - 1) / kBlockM);
    const int kc = q.mode == CONV_MODE_PACKED_ROW ? S * q.c_in_pitch : q.c_in_pitch;
    const int taps = q.mode == CONV_MODE_TILED ? 1 : (q.mode == CONV_MODE_PACKED_ROW ? R : R * S);
    const int side_kb = q.side_x ? (q.side_c_in_pitch + kBlockK - 1) / kBlockK : 0;
    const int num_kb = taps * ((kc + kBlockK - 1) / kBlockK) + side_kb;
    ConvTcPlanInfo best{};
    best.block_n = q.block_n ? q.block_n : conv_tc_pick_block_n(q.c_out, m_tiles, num_sms);
    best.splits = 1;
    const bool pair_mode = q.pair > 0 || (q.pair == 0 && !getenv("SMELTER_NO_PAIR"));
    if (pair_mode && !q.block_n && q.c_out > 64) {
        // Two-CTA clusters (conv_pair.cu): work items are (m-tile pair, n-tile) on num_sms / 2 clusters.  Per k-block a pair needs
        // ~0.125 / 0.15 / 0.26 us for N = 64 / 128 / 256 (N = 256 is bound by its MMAs, the others by operand delivery).  A [32 x 64]
        // epilogue item costs a warp ~1 us; 64-column tiles alternate between the two epilogue groups (one tile per 0.5 us), wider ones
        // are shared by both (1 us per 128 columns), and the last tile's epilogue is not hidden behind anything.  Checked against
        // every layer of tools/conv_layers.py under SMELTER_FORCE_BN=64/128/256 (profiles/r2_force_bn_layers.txt): same ordering.
        // That is the model for a plan that has the chip to itself (0.5248 -> 0.5185 ms per ResNet-50 step).  Plans for a share of the
        // chip run next to other encodes' kernels and compete for L2 bandwidth; there the narrower tiles it prefers (more operand
        // bytes per FLOP) lose 1.6 % images/s with three in flight, so share plans keep the first model: epilogue 0.35 us per 64
        // columns, no tail term.
        const bool alone = num_sms >= 120;
        const long pairs = (m_tiles + 1) / 2;
        double best_span = 1e30;
        const int cands[3] = {64, 128, 256};
        const double t_kb[3] = {0.125, 0.15, 0.26};
        const double epi_rate[3] = {0.5, 1.0, 2.0}, epi_tail[3] = {1.0, 1.0, 2.0};
        for (int i = 0; i < 3; ++i) {
            const int bn = cands[i];
            if (bn / 2 >= q.c_out) continue;
            const long items = pairs * ((q.c_out + bn - 1) / bn);
            const long waves = (items + num_sms / 2 - 1) / (num_sms / 2);
            const double span = alone ? double(waves) * std::max(num_kb * t_kb[i], epi_rate[i]) + epi_tail[i]
                                      : double(waves) * std::max(num_kb * t_kb[i], 0.35 * (bn / 64));
            if (span < best_span * 0.97) { best_span = span; best.block_n = bn; }
        }
    }
    if (const char* force = getenv("SMELTER_FORCE_BN")) {  // tuning experiments only
        const int v = atoi(force);
        if (v == 32 || v == 64 || v == 128 || v == 256) best.block_n = v;
    }
    const bool forced = q.block_n != 0 || q.splits == 1 || q.side_x || getenv("SMELTER_FORCE_BN") || getenv("SMELTER_NO_SPLITK");
    if (q.splits > 1) {
        best.splits = q.splits;
    } else if (!forced && q.c_out > 32) {
        // Single-wave layers (fewer tiles than SMs) are bound by the serial k-loop of one CTA: wider tiles make each k-block
        // do more work per MMA-issue slot, and splitting k across otherwise idle SMs shortens the chain.  Costs in us from
        // measurements on B200: k-block of N=64/128/256: 0.22/0.25/0.33; epilogue chunk 0.35; the split-K reduction costs a
        // device-wide fence + counter (~2.5) and one L2 round trip (~0.8) per split per 64-column chunk in the reducing CTA
        // (measured: splitting the 36..72 k-block layers of ResNet-50 at batch 32 lost time, so the model must be this honest).
        auto cost = [&](int bn, int splits) {
            const double t_kb = bn <= 64 ? 0.22 : (bn <= 128 ? 0.25 : 0.33);
            const int kbs = (num_kb + splits - 1) / splits;
            return kbs * t_kb + (bn / 64) * 0.35 + (splits > 1 ? 2.5 + (bn / 64) * splits * 0.8 : 0.0);
        };
        const int base_tiles = m_tiles * ((q.c_out + best.block_n - 1) / best.block_n);
        if (base_tiles <= num_sms) {
            double best_cost = cost(best.block_n, 1);
            const int cands[3] = {64, 128, 256};
            for (int bn : cands) {
                if (bn / 2 >= q.c_out) continue;  // more than half of the tile would be padding
                const int tiles = m_tiles * ((q.c_out + bn - 1) / bn);
                for (int splits = 1; splits <= 8; ++splits) {
                    if (tiles * splits > num_sms) break;
                    const int kbs = (num_kb + splits - 1) / splits;
                    if (splits > 1 && (kbs < 4 || (splits - 1) * kbs >= num_kb)) continue;
                    const double c = cost(bn, splits);
                    if (c < best_cost * 0.85) {  // only move for a clear win
                        best_cost = c / 0.85;
                        best.block_n = bn;
                        best.splits = splits;
                    }
                }
            }
        }
    }
    const int tiles = m_tiles * ((q.c_out + best.block_n - 1) / best.block_n);
    best.ws_bytes = best.splits > 1 ? size_t(tiles) * best.splits * kBlockM * best.block_n * sizeof(float) : 0;
    best.counter_bytes = best.splits > 1 ? size_t(tiles) * sizeof(unsigned int) : 0;
    return best;
}

bool conv_tc_side_supported(const ConvTcProblem& q, int num_sms) {
    if (!q.side_x || q.pair < 0 || (q.pair == 0 && getenv("SMELTER_NO_PAIR")) || num_sms < 2 || q.splits > 1) return false;
    if (q.side_c_in_pitch % 8 || q.side_stride_h < 1 || q.side_stride_w < 1 || q.side_stride_h > 8 || q.side_stride_w > 8) return false;
    const int P = (q.h + q.pad_t + q.pad_b - q.dil_h * (q.k_h - 1) - 1) / q.stride_h + 1;
    const int Q = (q.w + q.pad_l + q.pad_r - q.dil_w * (q.k_w - 1) - 1) / q.stride_w + 1;
    if ((q.side_h - 1) / q.side_stride_h + 1 != P || (q.side_w - 1) / q.side_stride_w + 1 != Q) return false;
    auto pairable = [&](const ConvTcProblem& pq) {
        const ConvTcPlanInfo info = conv_tc_plan(pq, num_sms);
        return info.splits == 1 && (info.block_n == 64 || info.block_n == 128 || info.block_n == 256);
    };
    ConvTcProblem alone = q;
    alone.side_x = nullptr;
    return pairable(alone) && pairable(q);
}

bool conv_tc_stats_supported(const ConvTcProblem& q, int num_sms) {
    if (q.pair < 0 || (q.pair == 0 && getenv("SMELTER_NO_PAIR")) || num_sms < 2 || q.splits > 1 || q.residual || q.side_x || q.act != ACT_NONE) return false;
    if (q.n > 65535) return false;  // the one-pass norm's grid.y is the image
    const int P = (q.h + q.pad_t + q.pad_b - q.dil_h * (q.k_h - 1) - 1) / q.stride_h + 1;
    const int Q = (q.w + q.pad_l + q.pad_r - q.dil_w * (q.k_w - 1) - 1) / q.stride_w + 1;
    if (P <= 0 || Q <= 0 || (long(P) * Q) % 128) return false;
    const ConvTcPlanInfo info = conv_tc_plan(q, num_sms);
    return info.splits == 1 && (info.block_n == 64 || info.block_n == 128 || info.block_n == 256);
}

int conv_tc_stats_reps(const ConvTcProblem& q) {
    const int P = (q.h + q.pad_t + q.pad_b - q.dil_h * (q.k_h - 1) - 1) / q.stride_h + 1;
    const int Q = (q.w + q.pad_l + q.pad_r - q.dil_w * (q.k_w - 1) - 1) / q.stride_w + 1;
    const long tiles = long(P) * Q / kBlockM;
    int reps = 1;
    while (reps < 16 && reps * 2 * 16 <= tiles && long(q.c_out_pitch) * reps * 2 <= 1024) reps <<= 1;
    return reps;
}

namespace {

// A-operand map of a TILED / IM2COL / PACKED_ROW problem: [128 pixels x 64 channels] boxes, 128-byte swizzle.
bool encode_a_map(CUtensorMap* tm, int mode, const __half* x, int n, int h, int w, int c_in_pitch, int kc, long M, int R, int S, int Q,
                  int stride_h, int stride_w, int dil_h, int dil_w, int pad_t, int pad_l, int pad_b, int pad_r, std::string* err) {
    if (mode == CONV_MODE_TILED) {
        cuuint64_t dims[2] = {cuuint64_t(kc), cuuint64_t(M)};
        cuuint64_t strides[1] = {cuuint64_t(kc) * 2};
        cuuint32_t box[2] = {kBlockK, kBlockM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(x), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            if (err) *err = "cuTensorMapEncodeTiled(activations) failed: " + std::to_string(int(r));
            return false;
        }
        return true;
    }
    cuuint64_t dims[4];
    cuuint64_t strides[3];
    int lower[2], upper[2];       // {W, H}
    cuuint32_t estr[4];
    if (mode == CONV_MODE_IM2COL) {
        dims[0] = cuuint64_t(kc); dims[1] = cuuint64_t(w); dims[2] = cuuint64_t(h); dims[3] = cuuint64_t(n);
        strides[0] = cuuint64_t(kc) * 2;
        strides[1] = strides[0] * w;
        strides[2] = strides[1] * h;
        lower[0] = -pad_l; lower[1] = -pad_t;
        upper[0] = pad_r - (S - 1) * dil_w;
        upper[1] = pad_b - (R - 1) * dil_h;
        estr[0] = 1; estr[1] = cuuint32_t(stride_w); estr[2] = cuuint32_t(stride_h); estr[3] = 1;
    } else {  // packed row: virtual tensor {S*8, Q, H, N}, overlapping W stride
        // When the last window of a row can read a full 64-element (128-byte) run without leaving the image row, expose
        // 64 "channels": the extra elements are the next pixel's data and meet zero weights (the B box zero-fills past
        // S*8), and the TMA no longer has to zero-fill the tail of every 112-byte row.
        const long last_window_end = long(Q - 1) * stride_w * c_in_pitch + kBlockK;
        const int kc_a = (kc < kBlockK && last_window_end <= long(w) * c_in_pitch) ? kBlockK : kc;
        dims[0] = cuuint64_t(kc_a); dims[1] = cuuint64_t(Q); dims[2] = cuuint64_t(h); dims[3] = cuuint64_t(n);
        strides[0] = cuuint64_t(stride_w) * c_in_pitch * 2;
        strides[1] = cuuint64_t(w) * c_in_pitch * 2;
        strides[2] = strides[1] * h;
        lower[0] = 0; lower[1] = 0;
        upper[0] = 0; upper[1] = -(R - 1) * dil_h;
        estr[0] = 1; estr[1] = 1; estr[2] = cuuint32_t(stride_h); estr[3] = 1;
    }
    if (lower[0] < -128 || lower[1] < -128 || upper[0] < -128 || upper[1] < -128 || upper[0] > 127 || upper[1] > 127 ||
        estr[1] > 8 || estr[2] > 8) {
        if (err) *err = "conv: padding/stride outside the im2col TMA range";
        return false;
    }
    CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(x), dims, strides, lower, upper,
                                 kBlockK, kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeIm2col failed: " + std::to_string(int(r));
        return false;
    }
    // Driver quirk (drivers <= CUDA 13.1): im2col descriptors of tensors smaller than 128 KiB get a flag bit
    // that must be cleared, otherwise loads fault.  Same workaround NVIDIA's CUTLASS applies.
    if (g_driver_version <= 13010) {
        const size_t bytes = size_t(strides[2]) * n;
        if (bytes < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
    }
    return true;
}

// Weight map [Cout][taps][kc] with [box_rows x 1 x 64] boxes.
bool encode_b_map(CUtensorMap* tm, const __half* w_packed, int kc, int taps, int c_out, int box_rows, std::string* err) {
    cuuint64_t dims[3] = {cuuint64_t(kc), cuuint64_t(taps), cuuint64_t(c_out)};
    cuuint64_t strides[2] = {cuuint64_t(kc) * 2, cuuint64_t(kc) * 2 * taps};
    cuuint32_t box[3] = {kBlockK, 1, cuuint32_t(box_rows)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(w_packed), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled(weights) failed: " + std::to_string(int(r));
        return false;
    }
    return true;
}

}  // namespace

bool conv_tc_prepare(ConvTcLaunch* L, const ConvTcProblem& q, int num_sms, std::string* err) {
    if (!load_driver_entry_points(err)) return false;
    memset(L, 0, sizeof(*L));
    const int R = q.k_h, S = q.k_w;
    ConvKernelParams& p = L->p;

    // Geometry of the (possibly virtual) activation tensor the A-operand map walks.
    int mode = q.mode;
    const int P = (q.h + q.pad_t + q.pad_b - q.dil_h * (R - 1) - 1) / q.stride_h + 1;
    const int Q = (q.w + q.pad_l + q.pad_r - q.dil_w * (S - 1) - 1) / q.stride_w + 1;
    if (P <= 0 || Q <= 0) { if (err) *err = "conv: empty output"; return false; }
    const long M = long(q.n) * P * Q;
    if (M >= (1L << 31)) { if (err) *err = "conv: M too large"; return false; }

    int kc;          // extent of the A "channel" axis (elements)
    int taps, taps_w;
    if (mode == CONV_MODE_TILED) {
        if (R != 1 || S != 1 || q.stride_h != 1 || q.stride_w != 1 || q.pad_t || q.pad_l || q.pad_b || q.pad_r) {
            if (err) *err = "conv: tiled mode needs a 1x1/stride-1/pad-0 problem";
            return false;
        }
        kc = q.c_in_pitch; taps = 1; taps_w = 1;
    } else if (mode == CONV_MODE_IM2COL) {
        kc = q.c_in_pitch; taps = R * S; taps_w = S;
    } else if (mode == CONV_MODE_PACKED_ROW) {
        if (q.c_in_pitch % 8 || q.c_in_pitch > 64 || q.dil_w != 1 || q.pad_t || q.pad_l || q.pad_b || q.pad_r || S * q.c_in_pitch > 1024) {
            if (err) *err = "conv: packed-row mode needs a Cin pitch of 8..64, dilation_w 1 and materialised padding";
            return false;
        }
        kc = S * q.c_in_pitch; taps = R; taps_w = 1;
    } else {
        if (err) *err = "conv: unknown mode";
        return false;
    }
    if (kc % 8) { if (err) *err = "conv: channel pitch must be a multiple of 8"; return false; }

    const ConvTcPlanInfo plan = conv_tc_plan(q, num_sms);
    const int block_n = plan.block_n;
    if (plan.splits > 1 && (!q.split_ws || !q.split_counters)) {
        if (err) *err = "conv: split-K chosen but no workspace supplied";
        return false;
    }
    L->block_n = block_n;
    L->use_pdl = getenv("SMELTER_NO_PDL") ? 0 : 1;
    p.use_pdl = L->use_pdl;
    p.l2_hints = q.l2_hints;
    if (const char* h = getenv("SMELTER_L2_HINTS")) p.l2_hints = atoi(h);  // experiments: force for every layer
    { const char* dbg = getenv("SMELTER_CONV_DEBUG"); p.debug_flags = dbg ? atoi(dbg) : 0; }
    p.timeline = nullptr;
    if (getenv("SMELTER_CONV_TIMELINE")) {
        static unsigned long long* buf = nullptr;
        if (!buf && cudaMalloc(&buf, 64 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(buf, 0, 64 * sizeof(unsigned long long));
        p.timeline = buf;
    }
    p.chain = nullptr;
    if (getenv("SMELTER_CHAIN_TIMELINE")) {
        if (!g_chain_buf && cudaMalloc(&g_chain_buf, kChainLaunches * 16 * sizeof(unsigned long long)) == cudaSuccess) conv_tc_chain_reset();
        if (g_chain_buf && g_chain_seq < kChainLaunches) {
            snprintf(g_chain_desc[g_chain_seq], sizeof g_chain_desc[0], "bn%-3d M=%-6ld Cout=%-4d K=%-5d%s%s", block_n, M, q.c_out, q.c_in * R * S + (q.side_x ? q.side_c_in : 0),
                     q.residual ? " +res" : "", q.side_x ? " +side" : "");
            p.chain = g_chain_buf + 16 * g_chain_seq++;
        }
    }
    {
        cudaError_t e = set_attr(block_n);
        if (e == cudaSuccess && (q.pair > 0 || (q.pair == 0 && !getenv("SMELTER_NO_PAIR"))) && (block_n == 64 || block_n == 128 || block_n == 256)) {
            e = conv_pair_set_attr(block_n);
        }
        if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return false; }
    }
    p.M = int(M);
    p.out_pitch = q.c_out_pitch;
    p.num_m_tiles = int((M + kBlockM - 1) / kBlockM);
    p.num_n_tiles = (q.c_out + block_n - 1) / block_n;
    p.num_taps = taps;
    p.taps_w = taps_w;
    p.kblocks_per_tap = (kc + kBlockK - 1) / kBlockK;
    p.P = P; p.Q = Q; p.PQ = P * Q;
    p.stride_h = q.stride_h;
    p.stride_w = (mode == CONV_MODE_PACKED_ROW) ? 1 : q.stride_w;
    p.dil_h = q.dil_h; p.dil_w = q.dil_w;
    p.corner_h = -q.pad_t;
    p.corner_w = -q.pad_l;
    p.mode = mode;
    p.bias = q.bias; p.has_residual = q.residual ? 1 : 0;
    p.splits = plan.splits;
    p.kb_per_split = (p.num_taps * p.kblocks_per_tap + plan.splits - 1) / plan.splits;
    p.ws = plan.splits > 1 ? q.split_ws : nullptr;
    p.counters = plan.splits > 1 ? q.split_counters : nullptr;
    p.act = q.act; p.clip_lo = q.clip_lo; p.clip_hi = q.clip_hi;
    L->grid = int(std::min<long>(long(p.num_m_tiles) * p.num_n_tiles * plan.splits, num_sms));
    L->splits = plan.splits;
    L->flops = 2.0 * double(M) * q.c_out * double(q.c_in) * R * S;

    // ---- B: packed weights [Cout][taps][kc]; two-CTA clusters (conv_pair.cu): each CTA loads half of the weight tile ----
    const bool want_pair = (q.pair > 0 || (q.pair == 0 && !getenv("SMELTER_NO_PAIR"))) && plan.splits == 1 && (block_n == 64 || block_n == 128 || block_n == 256) && num_sms >= 2;
    L->pair = want_pair ? 1 : 0;
    if (q.stats) {
        if (!want_pair || q.residual || q.act != ACT_NONE || (long(P) * Q) % kBlockM) { if (err) *err = "conv: epilogue statistics need the two-CTA kernel (see conv_tc_stats_supported)"; return false; }
        p.stats = q.stats;
        p.stats_rows = P * Q;
        p.stats_reps = conv_tc_stats_reps(q);
    }
    L->num_sms = num_sms;
    L->balanced_grid = 0;
    if (!encode_b_map(&L->tm_b, q.w_packed, kc, taps, q.c_out, want_pair ? block_n / 2 : block_n, err)) return false;
    // ---- D (and the residual, same geometry): [M, out_pitch] fp16, stored / loaded as [32 rows x 64 cols] swizzled boxes ----
    for (int which = 0; which < 2; ++which) {
        const __half* base = which == 0 ? q.y : (q.residual ? q.residual : q.y);
        cuuint64_t dims[2] = {cuuint64_t(q.c_out_pitch), cuuint64_t(M)};
        cuuint64_t strides[1] = {cuuint64_t(q.c_out_pitch) * 2};
        cuuint32_t box[2] = {cuuint32_t(kChunkN), cuuint32_t(32)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode_tiled(which == 0 ? &L->tm_out : &L->tm_res, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides,
                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            if (err) *err = "cuTensorMapEncodeTiled(output) failed: " + std::to_string(int(r));
            return false;
        }
    }
    // ---- A ----
    if (!encode_a_map(&L->tm_a, mode, q.x, q.n, q.h, q.w, q.c_in_pitch, kc, M, R, S, Q, q.stride_h, q.stride_w, q.dil_h, q.dil_w, q.pad_t, q.pad_l,
                      q.pad_b, q.pad_r, err))
        return false;
    // ---- projection shortcut: 1x1 / no padding over side_x, sampled with the shortcut's stride onto the same P x Q grid ----
    if (q.side_x) {
        if (!want_pair) { if (err) *err = "conv: a projection shortcut needs the two-CTA kernel (see conv_tc_side_supported)"; return false; }
        const int sp = (q.side_h - 1) / q.side_stride_h + 1, sq = (q.side_w - 1) / q.side_stride_w + 1;
        if (sp != P || sq != Q || !q.side_w_packed) { if (err) *err = "conv: projection shortcut does not produce the output's shape"; return false; }
        const bool unit = q.side_stride_h == 1 && q.side_stride_w == 1;
        p.side_mode = unit ? CONV_MODE_TILED : CONV_MODE_IM2COL;
        p.side_kb = (q.side_c_in_pitch + kBlockK - 1) / kBlockK;
        p.side_stride_h = q.side_stride_h;
        p.side_stride_w = q.side_stride_w;
        p.bias2 = q.side_bias;
        if (!encode_a_map(&L->tm_a2, p.side_mode, q.side_x, q.n, q.side_h, q.side_w, q.side_c_in_pitch, q.side_c_in_pitch, M, 1, 1, Q, q.side_stride_h,
                          q.side_stride_w, 1, 1, 0, 0, 0, 0, err))
            return false;
        if (!encode_b_map(&L->tm_b2, q.side_w_packed, q.side_c_in_pitch, 1, q.c_out, block_n / 2, err)) return false;
        L->flops += 2.0 * double(M) * q.c_out * double(q.side_c_in);
    }
    return true;
}


// ---- TMA load-rate probe (benchmarks only): one thread per CTA keeps `stages` loads of a [128 pixel x 64 channel]
//      box in flight from an NHWC tensor and nothing consumes them; reports what the TMA unit can deliver.
namespace {
__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int mode, int stages, int iters, int tiles_total, int PQ, int Q, int taps_w, int taps,
                 int distinct) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + uint32_t(stages) * kABytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int issued = 0, done = 0;
        uint32_t phase = 0;
        int tile = distinct ? int(blockIdx.x) : 0;
        int tap = 0;
        while (done < iters) {
            while (issued < iters && issued - done < stages) {
                const int s = issued % stages;
                mbar_expect_tx(bars + 8u * s, kABytes);
                const int m0 = (tile % tiles_total) * kBlockM;
                if (mode == 0) {
                    tma_load_2d(&tm, bars + 8u * s, base + uint32_t(s) * kABytes, 0, m0);
                } else {
                    const int img = m0 / PQ, rem = m0 - img * PQ, op = rem / Q, oq = rem - op * Q;
                    tma_load_im2col_4d(&tm, bars + 8u * s, base + uint32_t(s) * kABytes, 0, oq - 1, op - 1, img, uint16_t(tap % taps_w), uint16_t(tap / taps_w));
                }
                if (++tap == taps) { tap = 0; tile += distinct ? int(gridDim.x) : 1; }
                ++issued;
            }
            const int s = done % stages;
            mbar_wait(bars + 8u * s, phase);
            if (++done % stages == 0) phase ^= 1;
        }
    }
}
}  // namespace

int tma_probe(int mode, int c, int w, int h, int n, int stages, int iters, int grid, int distinct, const __half* x, cudaStream_t stream, float* ms,
              std::string* err) {
    if (!load_driver_entry_points(err)) return 1;
    CUtensorMap tm;
    const long M = long(n) * h * w;
    if (mode == 0) {
        cuuint64_t dims[2] = {cuuint64_t(c), cuuint64_t(M)};
        cuuint64_t strides[1] = {cuuint64_t(c) * 2};
        cuuint32_t box[2] = {kBlockK, kBlockM};
        cuuint32_t estr[2] = {1, 1};
        if (g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            if (err) *err = "encode tiled failed";
            return 1;
        }
    } else {
        cuuint64_t dims[4] = {cuuint64_t(c), cuuint64_t(w), cuuint64_t(h), cuuint64_t(n)};
        cuuint64_t strides[3] = {cuuint64_t(c) * 2, cuuint64_t(c) * 2 * w, cuuint64_t(c) * 2 * w * h};
        int lower[2] = {-1, -1}, upper[2] = {-1, -1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (g_encode_im2col(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(x), dims, strides, lower, upper, kBlockK, kBlockM, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            if (err) *err = "encode im2col failed";
            return 1;
        }
    }
    const size_t smem = size_t(stages) * kABytes + 1024 + 256;
    cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int tiles_total = int((M + kBlockM - 1) / kBlockM);
    const int taps = mode == 0 ? 1 : 9;
    tma_probe_kernel<<<grid, 128, smem, stream>>>(tm, mode, stages, iters, tiles_total, h * w, w, 3, taps, distinct);  // warm-up
    cudaEventRecord(e0, stream);
    tma_probe_kernel<<<grid, 128, smem, stream>>>(tm, mode, stages, iters, tiles_total, h * w, w, 3, taps, distinct);
    cudaEventRecord(e1, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return 1;
    }
    return 0;
}


namespace {
// mode 2: tiled 2-D box {box_c elems, box_r rows} (16 KiB) without swizzle; mode 3: sw128 {64,128} box fetched by a cluster of
// `csz` CTAs, each loading 128/csz rows and multicasting them to every CTA of the cluster.
__global__ void __launch_bounds__(128, 1)
tma_probe2_kernel(const __grid_constant__ CUtensorMap tm, int mode, int stages, int iters, int tiles_total, int box_c, int box_r, int csz) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + uint32_t(stages) * kABytes;
    uint32_t rank = 0;
    if (csz > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1);
        fence_barrier_init();
    }
    if (csz > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int issued = 0, done = 0;
        uint32_t phase = 0;
        int tile = int(blockIdx.x) / csz;
        while (done < iters) {
            while (issued < iters && issued - done < stages) {
                const int s = issued % stages;
                mbar_expect_tx(bars + 8u * s, kABytes);
                const int m0 = (tile % tiles_total) * kBlockM;
                const uint32_t dst = base + uint32_t(s) * kABytes;
                if (mode == 2) {
                    tma_load_2d(&tm, bars + 8u * s, dst, 0, (tile % tiles_total) * box_r);
                } else {
                    const int rows = kBlockM / csz;
                    const uint16_t mask = uint16_t((1u << csz) - 1u);
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                        ::"r"(dst + rank * uint32_t(rows) * 128u), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bars + 8u * s), "r"(0), "r"(m0 + int(rank) * rows), "h"(mask)
                        : "memory");
                }
                tile += int(gridDim.x) / csz;
                ++issued;
            }
            const int s = done % stages;
            mbar_wait(bars + 8u * s, phase);
            if (++done % stages == 0) phase ^= 1;
        }
    }
    if (csz > 1) {  // nobody may exit while peers still multicast into its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}
}  // namespace

int tma_probe2(int mode, int c, long rows_total, int box_c, int box_r, int csz, int stages, int iters, int grid, const __half* x, cudaStream_t stream,
               float* ms, std::string* err) {
    if (!load_driver_entry_points(err)) return 1;
    CUtensorMap tm;
    // mode 2 with box_c > c: overlapping rows (row pitch c elements, row length box_c) like the packed-row stem operand
    cuuint64_t dims[2] = {cuuint64_t(mode == 2 && box_c > c ? box_c : c), cuuint64_t(rows_total)};
    cuuint64_t strides[1] = {cuuint64_t(c) * 2};
    cuuint32_t box[2] = {cuuint32_t(mode == 2 ? box_c : kBlockK), cuuint32_t(mode == 2 ? box_r : kBlockM / csz)};
    cuuint32_t estr[2] = {1, 1};
    if (g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       mode == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        if (err) *err = "encode tiled failed";
        return 1;
    }
    const size_t smem = size_t(stages) * kABytes + 1024 + 256;
    cudaFuncSetAttribute(tma_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int tiles_total = int(rows_total / (mode == 2 ? box_r : kBlockM));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(csz);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tma_probe2_kernel, tm, mode, stages, iters, tiles_total, box_c, box_r, csz);
    cudaEventRecord(e0, stream);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, tma_probe2_kernel, tm, mode, stages, iters, tiles_total, box_c, box_r, csz);
    cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return 1;
    }
    return 0;
}


namespace {
// mode 4: one issuer thread, 3-D box {64, 128, slabs} (slabs x 16 KiB per instruction) over the tensor viewed as
// {64, rows, C/64}.  mode 5: `slabs` issuer warps, each streaming its own 16 KiB boxes through its own stage ring.
__global__ void __launch_bounds__(256, 1)
tma_probe3_kernel(const __grid_constant__ CUtensorMap tm, int mode, int stages, int iters, int tiles_total, int slabs) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int issuers = mode >= 5 ? slabs : 1;
    const uint32_t box_bytes = mode == 4 ? uint32_t(slabs) * kABytes : kABytes;
    const uint32_t ring_bytes = uint32_t(stages) * box_bytes;
    const uint32_t bars = base + uint32_t(issuers) * ring_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages * issuers; ++s) mbar_init(bars + 8u * s, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int me = mode == 6 ? int(threadIdx.x) : warp;  // mode 6: the issuers are lanes of warp 0
    if ((mode == 6 ? (warp == 0 && me < issuers) : ((threadIdx.x & 31) == 0 && warp < issuers))) {
        const uint32_t my_base = base + uint32_t(me) * ring_bytes;
        const uint32_t my_bars = bars + 8u * uint32_t(me * stages);
        int issued = 0, done = 0;
        uint32_t phase = 0;
        int tile = int(blockIdx.x) * issuers + me;
        while (done < iters) {
            while (issued < iters && issued - done < stages) {
                const int s = issued % stages;
                mbar_expect_tx(my_bars + 8u * s, box_bytes);
                const int m0 = (tile % tiles_total) * kBlockM;
                if (mode == 4) tma_load_3d(&tm, my_bars + 8u * s, my_base + uint32_t(s) * box_bytes, 0, m0, 0);
                else tma_load_2d(&tm, my_bars + 8u * s, my_base + uint32_t(s) * box_bytes, 0, m0);
                tile += int(gridDim.x) * issuers;
                ++issued;
            }
            const int s = done % stages;
            mbar_wait(my_bars + 8u * s, phase);
            if (++done % stages == 0) phase ^= 1;
        }
    }
}
}  // namespace

int tma_probe3(int mode, int c, long rows_total, int slabs, int stages, int iters, int grid, const __half* x, cudaStream_t stream, float* ms,
               std::string* err) {
    if (!load_driver_entry_points(err)) return 1;
    CUtensorMap tm;
    CUresult r;
    if (mode == 4) {
        cuuint64_t dims[3] = {64, cuuint64_t(rows_total), cuuint64_t(c / 64)};
        cuuint64_t strides[2] = {cuuint64_t(c) * 2, 128};
        cuuint32_t box[3] = {64, kBlockM, cuuint32_t(slabs)};
        cuuint32_t estr[3] = {1, 1, 1};
        r = g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[2] = {cuuint64_t(c), cuuint64_t(rows_total)};
        cuuint64_t strides[1] = {cuuint64_t(c) * 2};
        cuuint32_t box[2] = {kBlockK, kBlockM};
        cuuint32_t estr[2] = {1, 1};
        r = g_encode_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        if (err) *err = "encode failed " + std::to_string(int(r));
        return 1;
    }
    const int issuers = mode >= 5 ? slabs : 1;
    const size_t smem = size_t(stages) * issuers * (mode == 4 ? slabs : 1) * kABytes + 1024 + 512;
    if (smem > kSmemLimit) {
        if (err) *err = "too much shared memory";
        return 1;
    }
    cudaFuncSetAttribute(tma_probe3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int tiles_total = int(rows_total / kBlockM);
    tma_probe3_kernel<<<grid, 256, smem, stream>>>(tm, mode, stages, iters, tiles_total, slabs);
    cudaEventRecord(e0, stream);
    tma_probe3_kernel<<<grid, 256, smem, stream>>>(tm, mode, stages, iters, tiles_total, slabs);
    cudaEventRecord(e1, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        return 1;
    }
    return 0;
}

// Whole-encode launch chain (instrumented builds, SMELTER_CHAIN_TIMELINE=1): every conv_pair launch prepared while the switch is set
// owns 16 words = {min, max} over its CTAs of: 0 entry, 1 griddepcontrol.wait returned (first activation producer), 2 first operands
// landed (leader MMA thread), 3 last accumulator ready (first epilogue warp), 4 last store issued, 5 exit.
void conv_tc_chain_reset() {
    if (!g_chain_buf) return;
    std::vector<unsigned long long> h(size_t(kChainLaunches) * 16);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (i & 1) ? 0ull : ~0ull;
    cudaMemcpy(g_chain_buf, h.data(), h.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
}
void conv_tc_chain_dump() {
    if (!g_chain_buf) return;
    std::vector<unsigned long long> h(size_t(kChainLaunches) * 16);
    cudaMemcpy(h.data(), g_chain_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int l = 0; l < g_chain_seq; ++l) if (h[size_t(l) * 16] < t0) t0 = h[size_t(l) * 16];
    fprintf(stderr, "chain timeline (us since the first entry): launch | entry min..max | go min..max | first_operands min..max | last_acc min..max | last_store min..max | dealloc'd min..max | stores_read min..max | teardown_sync min..max\n");
    for (int l = 0; l < g_chain_seq; ++l) {
        const unsigned long long* r = &h[size_t(l) * 16];
        if (r[0] == ~0ull) continue;
        fprintf(stderr, "%3d %-44s", l, g_chain_desc[l]);
        for (int pt = 0; pt < 8; ++pt) {
            if (r[2 * pt] == ~0ull) fprintf(stderr, " |      -      -");
            else fprintf(stderr, " | %7.2f %7.2f", double((long long)(r[2 * pt] - t0)) * 1e-3, double((long long)(r[2 * pt + 1] - t0)) * 1e-3);
        }
        fprintf(stderr, "\n");
    }
}

void conv_tc_dump_timeline(const ConvTcLaunch& L) {
    if (!L.p.timeline) return;
    unsigned long long h[64];
    cudaMemcpy(h, L.p.timeline, sizeof h, cudaMemcpyDeviceToHost);
    if (L.pair) {  // conv_pair.cu: six stamps per work item of the first epilogue warp (see epi_stamp)
        fprintf(stderr, "pair epilogue timeline (ns since the first stamp; acc_ready, store_read+res_prefetch, tmem_loaded, residual_landed, staged, store_issued):\n");
        for (int item = 0; item < 8; ++item) {
            if (!h[16 + 6 * item]) break;
            fprintf(stderr, "  item %d:", item);
            for (int pt = 0; pt < 6; ++pt) fprintf(stderr, " %lld", (long long)(h[16 + 6 * item + pt] - h[16]));
            fprintf(stderr, "\n");
        }
        cudaMemset(L.p.timeline, 0, sizeof h);
        return;
    }
    const char* names[9] = {"entry", "prologue_done", "deps_resolved", "first_operands", "mma_issued", "acc_ready", "last_store_issued", "stores_read", "exit"};
    fprintf(stderr, "timeline(ns since entry):");
    for (int i = 0; i < 9; ++i) fprintf(stderr, " %s=%lld", names[i], (long long)(h[i] - h[0]));
    fprintf(stderr, "\n");
}

cudaError_t conv_tc_launch(const ConvTcLaunch& L, cudaStream_t stream) {
    if (L.pair) return conv_pair_launch(L, L.num_sms, stream);
    switch (L.block_n) {
        case 32: return launch_t<32>(L, stream);
        case 64: return launch_t<64>(L, stream);
        case 128: return launch_t<128>(L, stream);
        case 256: return launch_t<256>(L, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace k
}  // namespace smelter
