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
// * Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (TMEM -> regs -> bias /
//   residual / activation -> fp16 -> global).
#include "conv_igemm.h"

#include <cstdio>
#include <cstring>

#include "ptx.cuh"

namespace smelter {
namespace k {

using namespace ptx;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // fp16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 256;
constexpr int kEpilogueWarp0 = 4;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;

template <int BLOCK_N>
struct Cfg {
    static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    // leave 2 KB for barriers + alignment slack out of 227 KB
    static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
    static constexpr uint32_t kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
    static constexpr size_t kSmemBytes = size_t(kStages) * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float v, int act, float lo, float hi) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_CLIP) return fminf(fmaxf(v, lo), hi);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const ConvKernelParams p) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment.
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = p.num_taps * p.kblocks_per_tap;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_a);
        prefetch_tensormap(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, C::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.num_n_tiles;
                const int n_tile = tile - m_tile * p.num_n_tiles;
                const int m0 = m_tile * kBlockM;
                const int n0 = n_tile * BLOCK_N;
                int img = 0, base_h = 0, base_w = 0;
                if (p.mode != CONV_MODE_TILED) {
                    img = m0 / p.PQ;
                    const int rem = m0 - img * p.PQ;
                    const int op = rem / p.Q;
                    const int oq = rem - op * p.Q;
                    base_h = p.corner_h + op * p.stride_h;
                    base_w = p.corner_w + oq * p.stride_w;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.kblocks_per_tap;
                    const int cblk = kb - tap * p.kblocks_per_tap;
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t a_dst = smem_base + stage * C::kStageBytes;
                    const uint32_t b_dst = a_dst + kABytes;
                    mbar_expect_tx(full_bar(stage), C::kStageBytes);
                    if (p.mode == CONV_MODE_TILED) {
                        tma_load_2d(&tm_a, full_bar(stage), a_dst, cblk * kBlockK, m0);
                    } else {
                        const int r = tap / p.taps_w;
                        const int s = tap - r * p.taps_w;
                        tma_load_im2col_4d(&tm_a, full_bar(stage), a_dst, cblk * kBlockK, base_w, base_h, img,
                                           uint16_t(s * p.dil_w), uint16_t(r * p.dil_h));
                    }
                    tma_load_3d(&tm_b, full_bar(stage), b_dst, cblk * kBlockK, tap, n0);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * C::kStageBytes;
                    const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
                    const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance K inside the swizzle atom: +32 bytes -> +2 in the (>>4) address field
                        umma_f16(tmem_d, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ================= epilogue =================
        const int ew = warp - kEpilogueWarp0;  // == warp % 4: the TMEM lane quarter this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.num_n_tiles;
            const int n_tile = tile - m_tile * p.num_n_tiles;
            const int m = m_tile * kBlockM + ew * 32 + lane;
            const int n0 = n_tile * BLOCK_N;
            mbar_wait(tmem_full_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * BLOCK_N);
            const bool row_ok = m < p.M;
            __half* out_row = p.out + size_t(row_ok ? m : 0) * p.out_pitch;
            const __half* res_row = p.residual ? p.residual + size_t(row_ok ? m : 0) * p.out_pitch : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                uint32_t v[16];
                tmem_ld_16(taddr + uint32_t(c0), v);
                tmem_ld_wait();
                const int col = n0 + c0;
                if (row_ok && col < p.out_pitch) {
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b = __ldg(b4 + i);
                        f[4 * i + 0] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
                    }
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        if (col + 8 * h8 < p.out_pitch) {
                            if (res_row) {
                                const uint4 rv = *reinterpret_cast<const uint4*>(res_row + col + 8 * h8);
                                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 r2 = __half22float2(rh[i]);
                                    f[8 * h8 + 2 * i] += r2.x;
                                    f[8 * h8 + 2 * i + 1] += r2.y;
                                }
                            }
                            uint4 ov;
                            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float a = apply_act(f[8 * h8 + 2 * i], p.act, p.clip_lo, p.clip_hi);
                                const float b = apply_act(f[8 * h8 + 2 * i + 1], p.act, p.clip_lo, p.clip_hi);
                                oh[i] = __floats2half2_rn(a, b);
                            }
                            *reinterpret_cast<uint4*>(out_row + col + 8 * h8) = ov;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tmem_empty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
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
    return cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BLOCK_N>::kSmemBytes));
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
    conv_igemm_kernel<BLOCK_N><<<L.grid, kThreads, Cfg<BLOCK_N>::kSmemBytes, stream>>>(L.tm_a, L.tm_b, L.p);
    return cudaGetLastError();
}

}  // namespace

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
        if (q.c_in_pitch != 8 || q.dil_w != 1 || q.pad_t || q.pad_l || q.pad_b || q.pad_r) {
            if (err) *err = "conv: packed-row mode needs Cin pitch 8, dilation_w 1 and materialised padding";
            return false;
        }
        kc = S * 8; taps = R; taps_w = 1;
    } else {
        if (err) *err = "conv: unknown mode";
        return false;
    }
    if (kc % 8) { if (err) *err = "conv: channel pitch must be a multiple of 8"; return false; }

    const int block_n = q.block_n ? q.block_n : conv_tc_pick_block_n(q.c_out, int((M + kBlockM - 1) / kBlockM), num_sms);
    L->block_n = block_n;
    {
        cudaError_t e = set_attr(block_n);
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
    p.bias = q.bias; p.residual = q.residual; p.out = q.y;
    p.act = q.act; p.clip_lo = q.clip_lo; p.clip_hi = q.clip_hi;
    L->grid = int(std::min<long>(long(p.num_m_tiles) * p.num_n_tiles, num_sms));
    L->flops = 2.0 * double(M) * q.c_out * double(q.c_in) * R * S;

    // ---- B: packed weights [Cout][taps][kc] ----
    {
        cuuint64_t dims[3] = {cuuint64_t(kc), cuuint64_t(taps), cuuint64_t(q.c_out)};
        cuuint64_t strides[2] = {cuuint64_t(kc) * 2, cuuint64_t(kc) * 2 * taps};
        cuuint32_t box[3] = {kBlockK, 1, cuuint32_t(block_n)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = g_encode_tiled(&L->tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(q.w_packed), dims, strides,
                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            if (err) *err = "cuTensorMapEncodeTiled(weights) failed: " + std::to_string(int(r));
            return false;
        }
    }
    // ---- A ----
    if (mode == CONV_MODE_TILED) {
        cuuint64_t dims[2] = {cuuint64_t(kc), cuuint64_t(M)};
        cuuint64_t strides[1] = {cuuint64_t(kc) * 2};
        cuuint32_t box[2] = {kBlockK, kBlockM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode_tiled(&L->tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(q.x), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            if (err) *err = "cuTensorMapEncodeTiled(activations) failed: " + std::to_string(int(r));
            return false;
        }
    } else {
        cuuint64_t dims[4];
        cuuint64_t strides[3];
        int lower[2], upper[2];       // {W, H}
        cuuint32_t estr[4];
        if (mode == CONV_MODE_IM2COL) {
            dims[0] = cuuint64_t(kc); dims[1] = cuuint64_t(q.w); dims[2] = cuuint64_t(q.h); dims[3] = cuuint64_t(q.n);
            strides[0] = cuuint64_t(kc) * 2;
            strides[1] = strides[0] * q.w;
            strides[2] = strides[1] * q.h;
            lower[0] = -q.pad_l; lower[1] = -q.pad_t;
            upper[0] = q.pad_r - (S - 1) * q.dil_w;
            upper[1] = q.pad_b - (R - 1) * q.dil_h;
            estr[0] = 1; estr[1] = cuuint32_t(q.stride_w); estr[2] = cuuint32_t(q.stride_h); estr[3] = 1;
        } else {  // packed row: virtual tensor {S*8, Q, H, N}, overlapping W stride
            dims[0] = cuuint64_t(kc); dims[1] = cuuint64_t(Q); dims[2] = cuuint64_t(q.h); dims[3] = cuuint64_t(q.n);
            strides[0] = cuuint64_t(q.stride_w) * 16;
            strides[1] = cuuint64_t(q.w) * 16;
            strides[2] = strides[1] * q.h;
            lower[0] = 0; lower[1] = 0;
            upper[0] = 0; upper[1] = -(R - 1) * q.dil_h;
            estr[0] = 1; estr[1] = 1; estr[2] = cuuint32_t(q.stride_h); estr[3] = 1;
        }
        if (lower[0] < -128 || lower[1] < -128 || upper[0] < -128 || upper[1] < -128 || upper[0] > 127 || upper[1] > 127 ||
            estr[1] > 8 || estr[2] > 8) {
            if (err) *err = "conv: padding/stride outside the im2col TMA range";
            return false;
        }
        CUresult r = g_encode_im2col(&L->tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(q.x), dims, strides, lower, upper,
                                     kBlockK, kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            if (err) *err = "cuTensorMapEncodeIm2col failed: " + std::to_string(int(r));
            return false;
        }
        // Driver quirk (drivers <= CUDA 13.1): im2col descriptors of tensors smaller than 128 KiB get a flag bit
        // that must be cleared, otherwise loads fault.  Same workaround NVIDIA's CUTLASS applies.
        if (g_driver_version <= 13010) {
            const size_t bytes = size_t(strides[2]) * q.n;
            if (bytes < 131072) reinterpret_cast<uint64_t*>(&L->tm_a)[1] &= ~(1ull << 21);
        }
    }
    return true;
}

cudaError_t conv_tc_launch(const ConvTcLaunch& L, cudaStream_t stream) {
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
