// cta_group::2 / cluster forms of the PTX wrappers and the epilogue arithmetic shared by the two-CTA convolution kernels
// (conv_pair.cu).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace smelter {
namespace k {
namespace pairptx {

using namespace ptx;

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;            // clears the CTA-rank bit of a shared::cluster address -> the leader's copy
constexpr long long kSpinLimitCycles = 4000000000LL;  // a broken hand-shake traps instead of hanging the GPU

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only rendezvous of the two CTAs: `arrive.release` is a cluster-scope memory fence that waits for the CTA's outstanding
// global writes (measured with the launch-chain stamps: the teardown barrier of conv_pair_kernel returned ~1 us after the last
// epilogue warp reached it, once per launch).  The hand-shakes around it carry their own ordering: mbarrier inits are published by
// fence.mbarrier_init.release.cluster, CTA-local shared memory by the bar.sync in front, tcgen05 work by its fences and commits.
#ifndef SMELTER_CLUSTER_RELAXED
#define SMELTER_CLUSTER_RELAXED 1
#endif
__device__ __forceinline__ void cluster_sync_exec() {
#if SMELTER_CLUSTER_RELAXED
    __syncthreads();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
    cluster_sync_all();
#endif
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kSpinLimitCycles) __trap();
    }
}
__device__ __forceinline__ void tma2_load_2d(const void* desc, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma2_load_3d(const void* desc, uint32_t leader_bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const void* desc, uint32_t leader_bar, uint32_t dst, int c, int w, int h, int n, uint16_t off_w,
                                                    uint16_t off_h) {
    asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
                 : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {  // arrives on `bar` in BOTH CTAs when the MMAs issued so far have retired
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(uint16_t(3)) : "memory");
}

static __device__ __noinline__ float sigmoid1(float v) { return 1.f / (1.f + __expf(-v)); }

// CLAMP: 0 = no activation clamp, 1 = lower bound only (ReLU), 2 = both bounds (Clip)
template <int kCols, bool HAS_RES, bool SIGMOID, int CLAMP>
__device__ __forceinline__ void epilogue_math_t(const uint32_t (&v)[kCols], uint4 (&out)[kCols / 8], uint32_t rowbuf, uint32_t sw, uint32_t bias_slot,
                                                __half2 lo2, __half2 hi2) {
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
        if (SIGMOID) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = sigmoid1(f[i]);
        }
        __half2* oh = reinterpret_cast<__half2*>(&out[g]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            if (CLAMP >= 1) h = __hmax2(h, lo2);
            if (CLAMP >= 2) h = __hmin2(h, hi2);
            oh[i] = h;
        }
    }
}

// The sigmoid test is made once per chunk, not once per 8 columns: a branch around the out-of-line calls inside the unrolled loop
// cuts it into eight scheduling regions, each a serial LDS -> FADD -> F2FP -> HMNMX2 chain, and the epilogue's arithmetic
// (measured with the timeline stamps: 0.55-0.7 us of a 1.2 us work item) cannot overlap across column groups.
template <int kCols, bool HAS_RES>
__device__ __forceinline__ void epilogue_math(const uint32_t (&v)[kCols], uint4 (&out)[kCols / 8], uint32_t rowbuf, uint32_t sw, uint32_t bias_slot,
                                              bool is_sigmoid, __half2 lo2, __half2 hi2) {
    // lo2 / hi2 are -inf / +inf where the activation has no bound; the bounded forms are picked by comparing against those
    const bool has_lo = __hgt(__low2half(lo2), __float2half_rn(-INFINITY)), has_hi = __hlt(__low2half(hi2), __float2half_rn(INFINITY));
    if (is_sigmoid) epilogue_math_t<kCols, HAS_RES, true, 0>(v, out, rowbuf, sw, bias_slot, lo2, hi2);
    else if (has_hi) epilogue_math_t<kCols, HAS_RES, false, 2>(v, out, rowbuf, sw, bias_slot, lo2, hi2);
    else if (has_lo) epilogue_math_t<kCols, HAS_RES, false, 1>(v, out, rowbuf, sw, bias_slot, lo2, hi2);
    else epilogue_math_t<kCols, HAS_RES, false, 0>(v, out, rowbuf, sw, bias_slot, lo2, hi2);
}

}  // namespace pairptx
}  // namespace k
}  // namespace smelter
