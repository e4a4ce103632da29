// stem_tc.cu — the network stem on tcgen05 tensor cores (BF16 path):
//   conv 7x7/2 pad 3 (3 -> 64) + folded BN + ReLU + max-pool 3x3/2 pad 1, fused in ONE kernel,
// replacing conv2dForwardKernel + batchNorm2dForwardKernel + reluForwardKernel + maxPool2dKernel as
// chained at /root/reference/cuda/inference/main.cu:176-192 (on CUDA cores the stem is
// FP32-compute-bound, 60 GFLOP per 256-batch; on tensor cores it is a ~0.1 ms memory pass).
//
// Trick: no im2col is ever built. A pre-pass (stem_pack_kernel) rewrites the FP32 NCHW image as
// zero-padded NHWC4 BF16 ("RGB0" pixels of 8 bytes, 232 pixels per row). Two adjacent pixels form a
// 16-byte "super-pixel"; because the conv stride is 2, the window of output column ow starts at
// super-pixel ow and spans 4 of them (7 taps + 1 zero-weight tap). So for a fixed filter row kh the
// A operand A[ow][j*8+e] = row[(ow+j)*8 + e] is a Hankel matrix that a NO-SWIZZLE K-major UMMA
// descriptor reads directly from the contiguous row in shared memory: 16 bytes between consecutive M
// rows (SBO = 128 per 8 rows) and 16 bytes between consecutive K core matrices (LBO = 16) — the
// descriptor simply overlaps. One output row (128 lanes, 112 valid) x 64 channels accumulates over
// 7 kh x 2 MMAs (K = 16 each) into a 64-column TMEM slot.
//
// Work unit = (image, 3 pooled rows) = 7 conv rows (one is shared with the neighbour unit, 7/6
// redundancy) = 19 padded input rows (35 KB, ONE bulk copy) -> 7 TMEM slots (448 columns).
// Warp roles: warp 0 producer (bulk copies), warp 1 MMA issuer, warp 2 TMEM alloc, warps 4..7
// epilogue: bias + ReLU, vertical max over 3 slots in registers, horizontal max through shared
// memory, coalesced NHWC BF16 stores of the pooled row.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "sm100_ptx.cuh"

namespace rnb {

namespace {

constexpr int IMG = 224;
constexpr int PAD_W = 232;            // 3 + 224 + 5 pixels
constexpr int PAD_H = 235;            // 5 + 224 + 6 rows (padded row pr = ih + 5)
constexpr int ROW_BYTES = PAD_W * 8;  // 1856
constexpr int CONV = 112, POOL = 56;
constexpr int ROWS_PER_UNIT = 7;      // conv rows (TMEM slots)
constexpr int POOLED_PER_UNIT = 3;
constexpr int UNITS_PER_IMG = (POOL + POOLED_PER_UNIT - 1) / POOLED_PER_UNIT;  // 19
constexpr int IN_ROWS = 2 * ROWS_PER_UNIT + 5;                                  // 19
constexpr int IN_BYTES = IN_ROWS * ROW_BYTES;                                   // 35264
constexpr int IN_SLOT_BYTES = ((IN_BYTES + 512 + 1023) / 1024) * 1024;          // + read-past slack
constexpr int W_BYTES = 28 * 1024;    // [kh*4 + j][64 oc][8 e] bf16
constexpr int VBUF_BYTES = 112 * 128; // [ow][64 ch] bf16, 16-byte chunks XOR-swizzled by (ow & 7)
constexpr int NBAR = 4 + 2 * ROWS_PER_UNIT;
constexpr int STEM_SMEM = 1024 + 2 * IN_SLOT_BYTES + W_BYTES + 2 * VBUF_BYTES + NBAR * 8 + 16;
constexpr int EPI_THREADS = 256;                 // 8 epilogue warps
constexpr int STEM_TC_THREADS = 128 + EPI_THREADS;

// x [B,3,224,224] fp32 -> xp [B][235][232][4] bf16, zero border and zero 4th channel.
// One thread per group of 4 padded pixels (PAD_W = 232 = 58 groups): interior groups read one float4
// from each colour plane (image column 4g-4 .. 4g-1 for group g >= 1; 224 = 56 groups) and write 32 B.
__global__ void stem_pack_kernel(const float* __restrict__ x, uint4* __restrict__ xp, int B) {
    constexpr int GROUPS = PAD_W / 4;  // 58
    const int64_t total = 1LL * B * PAD_H * GROUPS;
    ptx::griddep_launch_dependents();
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int gidx = static_cast<int>(i % GROUPS);
        int64_t t = i / GROUPS;
        const int pr = static_cast<int>(t % PAD_H);
        const int b = static_cast<int>(t / PAD_H);
        const int ih = pr - 5;
        float r[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f}, bl[4] = {0.f, 0.f, 0.f, 0.f};
        if (ih >= 0 && ih < IMG) {
            // padded pixels 4g..4g+3 are image columns 4g-3..4g: not 16-byte aligned, so read the two
            // aligned float4s that cover them (columns 4g-4..4g-1 and 4g..4g+3) and pick
            const float* row = x + (1LL * b * 3 * IMG + ih) * IMG;
            const int c0 = 4 * gidx - 4;
            float lo[3][4], hi[3][4];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* pl = row + 1LL * c * IMG * IMG;
                const float4 a = (c0 >= 0 && c0 < IMG) ? __ldg(reinterpret_cast<const float4*>(pl + c0))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 d = (c0 + 4 < IMG) ? __ldg(reinterpret_cast<const float4*>(pl + c0 + 4))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                lo[c][0] = a.x; lo[c][1] = a.y; lo[c][2] = a.z; lo[c][3] = a.w;
                hi[c][0] = d.x; hi[c][1] = d.y; hi[c][2] = d.z; hi[c][3] = d.w;
            }
            // image column of padded pixel 4g+j is 4g+j-3 = c0 + 1 + j
            r[0] = lo[0][1]; r[1] = lo[0][2]; r[2] = lo[0][3]; r[3] = hi[0][0];
            g[0] = lo[1][1]; g[1] = lo[1][2]; g[2] = lo[1][3]; g[3] = hi[1][0];
            bl[0] = lo[2][1]; bl[1] = lo[2][2]; bl[2] = lo[2][3]; bl[3] = hi[2][0];
        }
        uint4 o0, o1;
        auto px = [&](int j, uint32_t& a, uint32_t& c) {
            const __nv_bfloat162 rg = __floats2bfloat162_rn(r[j], g[j]);
            const __nv_bfloat162 b0 = __floats2bfloat162_rn(bl[j], 0.f);
            a = *reinterpret_cast<const uint32_t*>(&rg);
            c = *reinterpret_cast<const uint32_t*>(&b0);
        };
        px(0, o0.x, o0.y); px(1, o0.z, o0.w); px(2, o1.x, o1.y); px(3, o1.z, o1.w);
        uint4* dst = xp + 2 * i;
        dst[0] = o0;
        dst[1] = o1;
    }
}

// Same layout pre-pass from the DECODED image: x [B][224][224][3] uint8 HWC (what a JPEG decoder + resize/crop
// yields, convert_imgs_to_bin.py:12) -> xp, with the /255 + mean/std normalisation of convert_imgs_to_bin.py:18
// (torchvision ToTensor + Normalize) evaluated in FP32 exactly as torch does: (float(u8) / 255 - mean) / std.
// 4x fewer input bytes than the FP32 NCHW tensor (the host->device copy is what bounds end-to-end throughput).
struct StemNorm {
    float mean[3], std[3];
};
__global__ void stem_pack_u8_kernel(const uint8_t* __restrict__ x, uint4* __restrict__ xp, int B, StemNorm nm) {
    constexpr int GROUPS = PAD_W / 4;  // 58
    const int64_t total = 1LL * B * PAD_H * GROUPS;
    ptx::griddep_launch_dependents();
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int gidx = static_cast<int>(i % GROUPS);
        int64_t t = i / GROUPS;
        const int pr = static_cast<int>(t % PAD_H);
        const int b = static_cast<int>(t / PAD_H);
        const int ih = pr - 5;
        float v[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j][0] = v[j][1] = v[j][2] = 0.f;
        if (ih >= 0 && ih < IMG) {
            const uint8_t* row = x + (1LL * b * IMG + ih) * IMG * 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int iw = 4 * gidx - 3 + j;  // image column of padded pixel 4g + j
                if (iw >= 0 && iw < IMG) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float f = __fdiv_rn(static_cast<float>(__ldg(row + iw * 3 + c)), 255.f);
                        v[j][c] = __fdiv_rn(__fsub_rn(f, nm.mean[c]), nm.std[c]);
                    }
                }
            }
        }
        uint4 o0, o1;
        auto px = [&](int j, uint32_t& a, uint32_t& c) {
            const __nv_bfloat162 rg = __floats2bfloat162_rn(v[j][0], v[j][1]);
            const __nv_bfloat162 b0 = __floats2bfloat162_rn(v[j][2], 0.f);
            a = *reinterpret_cast<const uint32_t*>(&rg);
            c = *reinterpret_cast<const uint32_t*>(&b0);
        };
        px(0, o0.x, o0.y); px(1, o0.z, o0.w); px(2, o1.x, o1.y); px(3, o1.z, o1.w);
        uint4* dst = xp + 2 * i;
        dst[0] = o0;
        dst[1] = o1;
    }
}

// x [B][H][W][3] uint8 HWC -> out [B][3][H][W] fp32 NCHW, normalised as above (generic path: TF32 stem,
// CUDA-core stem). One thread per pixel.
__global__ void u8_hwc_to_f32_nchw_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, int64_t pixels,
                                          int HW, StemNorm nm) {
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < pixels; i += 1LL * gridDim.x * blockDim.x) {
        const int64_t b = i / HW, p = i - b * HW;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f = __fdiv_rn(static_cast<float>(__ldg(x + i * 3 + c)), 255.f);
            out[(b * 3 + c) * HW + p] = __fdiv_rn(__fsub_rn(f, nm.mean[c]), nm.std[c]);
        }
    }
}

// w [64][3][7][7] fp32 + BN -> wk [kh*4+j][oc][e] bf16 with e = (kw - 2j)*4 + c, kw in {2j, 2j+1};
// kw = 7 and c = 3 are zero. bias[oc] = folded BN shift.
__global__ void stem_pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                         const float* __restrict__ bn_b, const float* __restrict__ bn_m,
                                         const float* __restrict__ bn_v, __nv_bfloat16* __restrict__ wk,
                                         float* __restrict__ bias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 28 * 64 * 8) return;
    const int e = i & 7, oc = (i >> 3) & 63, chunk = i >> 9;
    const int kh = chunk >> 2, j = chunk & 3;
    const int kw = 2 * j + (e >> 2), c = e & 3;
    double scale = 1.0, shift = 0.0;
    if (bn_w) {
        scale = static_cast<double>(bn_w[oc]) / sqrt(static_cast<double>(bn_v[oc]) + 1e-5);
        shift = static_cast<double>(bn_b[oc]) - static_cast<double>(bn_m[oc]) * scale;
    }
    float v = 0.f;
    if (kw < 7 && c < 3) v = static_cast<float>(static_cast<double>(w[((oc * 3 + c) * 7 + kh) * 7 + kw]) * scale);
    wk[i] = __float2bfloat16_rn(v);
    if (chunk == 0 && e == 0) bias[oc] = static_cast<float>(shift);
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                              uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            ptx::smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(STEM_TC_THREADS, 1)
stem_tc_kernel(const uint8_t* __restrict__ xp, const uint8_t* __restrict__ wk,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B) {
    using namespace ptx;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* in_slot = smem;                              // 2 x IN_SLOT_BYTES
    uint8_t* wsm = smem + 2 * IN_SLOT_BYTES;              // W_BYTES
    uint8_t* vbuf = wsm + W_BYTES;                        // 2 x VBUF_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(vbuf + 2 * VBUF_BYTES);
    uint64_t* in_full = bars;        // [2]
    uint64_t* in_empty = bars + 2;   // [2]
    uint64_t* slot_full = bars + 4;  // [7]
    uint64_t* slot_empty = bars + 4 + ROWS_PER_UNIT;  // [7]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5;
    const int num_units = B * UNITS_PER_IMG;
    // Each CTA takes a CONTIGUOUS range of units, so that most units follow the previous unit of the same image: the conv
    // row the two share (slot 0 of the new unit = slot 6 of the old one) is then neither recomputed nor re-read — the
    // epilogue already holds it in registers (`carry`). One seventh of the MMAs of this tensor-bound kernel disappears.
    const int upc = num_units / static_cast<int>(gridDim.x), urem = num_units % static_cast<int>(gridDim.x);
    const int u_begin = static_cast<int>(blockIdx.x) * upc + min(static_cast<int>(blockIdx.x), urem);
    const int u_end = u_begin + upc + (static_cast<int>(blockIdx.x) < urem ? 1 : 0);
    auto continues = [&](int u) { return u > u_begin && (u % UNITS_PER_IMG) != 0; };

    if (threadIdx.x == 32) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&in_full[i], 1);
            mbar_init(&in_empty[i], 1);
        }
        for (int i = 0; i < ROWS_PER_UNIT; ++i) {
            mbar_init(&slot_full[i], 1);
            mbar_init(&slot_empty[i], EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    // weights -> smem (28 KB, once per CTA), zero the read-past slack of both input slots
    for (int i = threadIdx.x; i < W_BYTES / 16; i += STEM_TC_THREADS)
        reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(wk) + i);
    for (int i = threadIdx.x; i < 2 * (IN_SLOT_BYTES - IN_BYTES) / 16; i += STEM_TC_THREADS) {
        const int s = i / ((IN_SLOT_BYTES - IN_BYTES) / 16), o = i % ((IN_SLOT_BYTES - IN_BYTES) / 16);
        reinterpret_cast<uint4*>(in_slot + s * IN_SLOT_BYTES + IN_BYTES)[o] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();  // generic-proxy writes above are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // PDL: the prologue above (barriers, TMEM, weights — constants) overlaps the tail of the layout pre-pass, and
    // the first conv kernel's prologue overlaps this kernel's tail
    griddep_launch_dependents();
    griddep_wait();

    if (warp == 0) {
        // ===================================================== producer: one bulk copy per unit
        int it = 0;
        for (int u = u_begin; u < u_end; ++u, ++it) {
            const int s = it & 1;
            mbar_wait(&in_empty[s], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                const int b = u / UNITS_PER_IMG, v = u - b * UNITS_PER_IMG;
                const uint8_t* src = xp + (1LL * b * PAD_H + 12 * v) * ROW_BYTES;  // padded rows 12v .. 12v+18
                mbar_expect_tx(&in_full[s], IN_BYTES);
                bulk_copy_g2s(in_slot + s * IN_SLOT_BYTES, src, IN_BYTES, &in_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, 64);
        const uint64_t a_desc0 = umma_smem_desc(smem_u32(in_slot), 16, 128, UMMA_LAYOUT_NONE);
        const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), 1024, 128, UMMA_LAYOUT_NONE);
        int it = 0;
        uint32_t n0 = 0;  // uses of slot 0 so far
        for (int u = u_begin; u < u_end; ++u, ++it) {
            const int s = it & 1;
            mbar_wait(&in_full[s], (it >> 1) & 1);
            const bool cont = continues(u);
            for (int r = cont ? 1 : 0; r < ROWS_PER_UNIT; ++r) {
                // slot 0 is skipped by continuing units: its barriers count their own uses
                mbar_wait(&slot_empty[r], r == 0 ? (n0 & 1) ^ 1 : (it & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + r * 64;
#pragma unroll
                    for (int kh = 0; kh < 7; ++kh) {
                        // conv row r of the unit reads padded input rows 2r + kh of the slot
                        const uint64_t a_row = a_desc0 + static_cast<uint64_t>(
                                                             (s * IN_SLOT_BYTES + (2 * r + kh) * ROW_BYTES) >> 4);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint64_t ad = a_row + static_cast<uint64_t>(i * 2);            // +32 B: j += 2
                            const uint64_t bd = b_desc0 + static_cast<uint64_t>(((kh * 4 + 2 * i) * 1024) >> 4);
                            mma_f16_ss(d_tmem, ad, bd, idesc, (kh | i) != 0);
                        }
                    }
                    tc_commit(&slot_full[r]);
                    if (r == ROWS_PER_UNIT - 1) tc_commit(&in_empty[s]);  // input slot fully consumed
                }
                __syncwarp();
            }
            if (!cont) ++n0;
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue
        // 8 warps: warp quarter q = warp & 3 owns TMEM lanes 32q..32q+31 (conv columns ow), and the
        // two warps of a quarter split the 64 channels (half 0 / half 1).
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int et = q * 32 + (threadIdx.x & 31);   // conv output column ow == TMEM lane
        const int etid = threadIdx.x - 128;           // 0..255 within the epilogue group
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 32;
        float bias_r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) bias_r[i] = __ldg(bias + half * 32 + i);
        int it = 0;
        int vb = 0;  // vbuf ping-pong
        uint32_t n0 = 0;  // uses of slot 0 so far
        float carry[32];  // the last conv row of the previous pooled row (and of the previous unit)
#pragma unroll
        for (int i = 0; i < 32; ++i) carry[i] = -INFINITY;
        for (int u = u_begin; u < u_end; ++u, ++it) {
            const int b = u / UNITS_PER_IMG, v = u - b * UNITS_PER_IMG;
            const uint32_t par = it & 1;
            const bool cont = continues(u);
#pragma unroll
            for (int p = 0; p < POOLED_PER_UNIT; ++p) {
                const int ph = v * POOLED_PER_UNIT + p;
                // slots 2p, 2p+1, 2p+2 hold conv rows oh = 6v - 1 + slot
                if (p == 0 && !cont) mbar_wait(&slot_full[0], n0 & 1);
                mbar_wait(&slot_full[2 * p + 1], par);
                mbar_wait(&slot_full[2 * p + 2], par);
                tc_fence_after();
                // vertical max over the three conv rows of this pooled row. max_k(x_k + b) =
                // max_k(x_k) + b and ReLU output is >= 0, so a missing row (-1 or 112) is simply
                // left out; the middle row 2*ph always exists for a stored pooled row.
                // the conv row shared
                // by two consecutive pooled rows (slot 2p+2 == slot 2(p+1)) is read once and carried in
                // registers.
                float m[32];
#pragma unroll
                for (int k = (p == 0 ? 0 : 1); k < 3; ++k) {
                    if (p == 0 && k == 0 && cont) continue;  // that row is `carry`
                    const int oh = 6 * v - 1 + 2 * p + k;
                    uint32_t raw[32];
                    __syncwarp();
                    tmem_ld_32x32(lane_addr + (2 * p + k) * 64, raw);
                    tmem_ld_wait();
                    if (k == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) m[i] = oh >= 0 ? __uint_as_float(raw[i]) : -INFINITY;
                    } else {
                        const bool valid = oh < CONV;
                        if (k == 1 && (p > 0 || cont)) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) m[i] = carry[i];
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float x = valid ? __uint_as_float(raw[i]) : -INFINITY;
                            m[i] = fmaxf(m[i], x);
                            if (k == 2) carry[i] = x;
                        }
                    }
                }
                if (et < CONV) {
                    uint8_t* vrow = vbuf + vb * VBUF_BYTES + et * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) x[e] = fmaxf(m[j * 8 + e] + bias_r[j * 8 + e], 0.f);
                        uint4 o;
                        o.x = pack_bf16x2(x[0], x[1]);
                        o.y = pack_bf16x2(x[2], x[3]);
                        o.z = pack_bf16x2(x[4], x[5]);
                        o.w = pack_bf16x2(x[6], x[7]);
                        *reinterpret_cast<uint4*>(vrow + (((half * 4 + j) ^ (et & 7)) << 4)) = o;
                    }
                }
                // slots 2p and 2p+1 are drained; 2p+2 is re-read by the next pooled row (or drained
                // with the last one)
                tc_fence_before();
                if (p > 0 || !cont) mbar_arrive(&slot_empty[2 * p]);
                mbar_arrive(&slot_empty[2 * p + 1]);
                if (p == POOLED_PER_UNIT - 1) mbar_arrive(&slot_empty[2 * p + 2]);
                named_bar_sync(1, EPI_THREADS);
                // horizontal 3-tap max (cols 2pw-1, 2pw, 2pw+1) and coalesced store of the pooled row
                if (ph < POOL) {
                    const uint8_t* vr = vbuf + vb * VBUF_BYTES;
                    __nv_bfloat16* orow = out + ((1LL * b * POOL + ph) * POOL) * 64;
                    for (int task = etid; task < POOL * 8; task += EPI_THREADS) {
                        const int pw = task >> 3, c16 = task & 7;
                        const int c0 = 2 * pw;
                        uint4 a = *reinterpret_cast<const uint4*>(vr + c0 * 128 + ((c16 ^ (c0 & 7)) << 4));
                        const uint4 c = *reinterpret_cast<const uint4*>(vr + (c0 + 1) * 128 + ((c16 ^ ((c0 + 1) & 7)) << 4));
                        a.x = bf16x2_max(a.x, c.x); a.y = bf16x2_max(a.y, c.y);
                        a.z = bf16x2_max(a.z, c.z); a.w = bf16x2_max(a.w, c.w);
                        if (pw > 0) {
                            const uint4 l = *reinterpret_cast<const uint4*>(vr + (c0 - 1) * 128 + ((c16 ^ ((c0 - 1) & 7)) << 4));
                            a.x = bf16x2_max(a.x, l.x); a.y = bf16x2_max(a.y, l.y);
                            a.z = bf16x2_max(a.z, l.z); a.w = bf16x2_max(a.w, l.w);
                        }
                        *reinterpret_cast<uint4*>(orow + pw * 64 + c16 * 8) = a;
                    }
                }
                vb ^= 1;
            }
            if (!cont) ++n0;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

size_t stem_tc_packed_input_bytes(int B) { return 1ull * B * PAD_H * ROW_BYTES; }
size_t stem_tc_packed_weight_bytes() { return W_BYTES; }

cudaError_t launch_stem_tc_pack_weights(const float* w, const float* bn_w, const float* bn_b,
                                        const float* bn_m, const float* bn_v, void* wk, float* bias,
                                        cudaStream_t s) {
    stem_pack_weights_kernel<<<(28 * 64 * 8 + 255) / 256, 256, 0, s>>>(
        w, bn_w, bn_b, bn_m, bn_v, static_cast<__nv_bfloat16*>(wk), bias);
    return cudaGetLastError();
}

cudaError_t stem_tc_init() {
    return cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
}

// part 0: x fp32 NCHW [B,3,224,224] -> xp (scratch of stem_tc_packed_input_bytes(B));
// part 1: xp -> out NHWC bf16 [B,56,56,64].
cudaError_t launch_stem_tc_part(int part, const float* x, void* xp, const void* wk, const float* bias,
                                void* out, int B, cudaStream_t s) {
    if (part == 0) {
        const int64_t total = 1LL * B * PAD_H * (PAD_W / 4);
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
        if (blocks > cap) blocks = cap;
        stem_pack_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, static_cast<uint4*>(xp), B);
    } else {
        const int units = B * UNITS_PER_IMG;
        const int grid = units < num_sms() ? units : num_sms();
        return launch_pdl_small(stem_tc_kernel, dim3(grid), dim3(STEM_TC_THREADS), STEM_SMEM, s,
                                static_cast<const uint8_t*>(xp), static_cast<const uint8_t*>(wk), bias,
                                static_cast<__nv_bfloat16*>(out), B);
    }
    return cudaGetLastError();
}

// part 0 from decoded uint8 HWC images (see stem_pack_u8_kernel)
cudaError_t launch_stem_tc_pack_u8(const uint8_t* x, void* xp, int B, const float* mean, const float* std,
                                   cudaStream_t s) {
    StemNorm nm;
    for (int c = 0; c < 3; ++c) {
        nm.mean[c] = mean[c];
        nm.std[c] = std[c];
    }
    const int64_t total = 1LL * B * PAD_H * (PAD_W / 4);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    stem_pack_u8_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, static_cast<uint4*>(xp), B, nm);
    return cudaGetLastError();
}

cudaError_t launch_u8_hwc_to_f32_nchw(const uint8_t* x, float* out, int B, int H, int W, const float* mean,
                                      const float* std, cudaStream_t s) {
    StemNorm nm;
    for (int c = 0; c < 3; ++c) {
        nm.mean[c] = mean[c];
        nm.std[c] = std[c];
    }
    const int64_t pixels = 1LL * B * H * W;
    int64_t blocks = (pixels + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    u8_hwc_to_f32_nchw_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, out, pixels, H * W, nm);
    return cudaGetLastError();
}

cudaError_t launch_stem_tc(const float* x, void* xp, const void* wk, const float* bias, void* out, int B,
                           cudaStream_t s) {
    cudaError_t e = launch_stem_tc_part(0, x, xp, wk, bias, out, B, s);
    if (e != cudaSuccess) return e;
    return launch_stem_tc_part(1, x, xp, wk, bias, out, B, s);
}

}  // namespace rnb
