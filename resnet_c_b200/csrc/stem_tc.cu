// stem_tc.cu — the network stem on tcgen05 tensor cores (BF16 path):
//   conv 7x7/2 pad 3 (3 -> 64) + folded BN + ReLU + max-pool 3x3/2 pad 1, fused in ONE kernel that reads the FP32 image,
// replacing conv2dForwardKernel + batchNorm2dForwardKernel + reluForwardKernel + maxPool2dKernel as
// chained at /root/reference/cuda/inference/main.cu:176-192 (on CUDA cores the stem is
// FP32-compute-bound, 60 GFLOP per 256-batch; on tensor cores it is a ~0.1 ms pass).
//
// No im2col is ever built. The image lives in shared memory as zero-padded NHWC4 BF16 rows ("RGB0" pixels of 8 bytes,
// 232 pixels per row: 4 + 224 + 4). Two adjacent pixels form a 16-byte "super-pixel"; because the conv stride is 2,
// the window of output column ow starts at super-pixel ow and spans 4 of them (a zero-weight tap + the 7 real ones).
// For a fixed filter row the operand X[ow][j*8+e] = row[(ow+j)*8 + e] is a Hankel matrix that a NO-SWIZZLE K-major
// UMMA descriptor reads directly from the contiguous row: 16 bytes between consecutive rows (SBO = 128 per 8 rows)
// and 16 bytes between consecutive K core matrices (LBO = 16) — the descriptor simply overlaps itself.
//
// Form of the kernel (the third of round 2; profiles/stem_r2.md has the measurements behind every step):
//  * work unit = a PAIR of conv rows (2k, 2k+1) = one pooled row k. The WEIGHTS are the A operand and live in TMEM,
//    the image rows are the B operand:  D[(conv row a | b, oc)][ow] (+)= [W_t ; W_t-2][128 x 16] . row_t[112 x 16]^T
//    for the 9 input rows t of the pair and the 2 K steps: 18 x (M = 128, N = 112, K = 16) at the tensor pipe's floor
//    of 56 clocks (tools/mma_bench.cu: an MMA costs max(N/2, operand fetch) and only the 28 wavefronts of the image
//    slice come from shared memory; the first two forms fetched 48-64 per MMA and were bound by that pipe). A conv
//    row that does not read input row t has zero weights in its half of A, so there are no special cases.
//  * the input streams through a ring of eight 4-row chunks; three loader warps build the NHWC4 BF16 rows straight
//    from the caller's FP32 NCHW tensor (MODE 1; an L2 prefetch runs three chunks ahead) — there is no layout pre-pass.
//    MODE 2 serves the packed host paths (host_pack.cpp): the first `nb` images of the batch come from a BF16 NCHW
//    tensor — the FP32 image rounded to BF16 by the host cores, the rounding this loader would apply anyway, at half
//    the PCIe bytes — and the rest from an FP32 NCHW tensor, so that host cores and PCIe can share a batch.
//    uint8 input keeps its normalising pre-pass and feeds the same kernel through bulk copies (MODE 0).
//  * TWO MMA issuer warps take alternate pairs: the tensor pipe's queue holds about two MMAs, and one issuer's barrier
//    polls and bookkeeping (~430 clocks per pair) left the pipe dry 40 % of the time.
//  * TMEM lane = 32q + 16*rowsel + ocl (channel 16q + ocl), TMEM column = output column: a thread holds one conv row
//    of one channel, so the horizontal 3-tap max is register arithmetic; the vertical max needs the partner lane
//    (lane ^ 16) — values are exchanged AFTER bias + ReLU + BF16 rounding (monotone, they commute with max), two
//    pooled columns per register: 7 shuffles per thread and pair. The pooled row's third conv row (2k - 1) is the
//    previous pair's second row, carried in registers; a CTA whose contiguous range of pairs starts inside an image
//    first computes the pair before its range to obtain it.
//  * the eight epilogue warps never meet: each stages its [28 pooled columns][16 channels] block and sends it off with
//    its own TMA store; the next pair's accumulators are requested as soon as this pair's are reduced.
// Warps: 0, 2, 3 loaders (MODE 1, 2) / 0 bulk-copy producer (MODE 0); 1 and 12 MMA issuers; 2 also TMEM alloc; 4..11
// epilogue. TMEM: three accumulator slots of 112 columns + 144 columns of weights.
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "sm100_ptx.cuh"
#include "stem_tc_common.cuh"
#include "tensormap.h"

namespace rnb {

namespace {

using namespace stemtc;

constexpr int NCH = 8;                // ring depth in chunks; pair k reads chunks k, k+1, k+2
constexpr int RING_BYTES = NCH * CHUNK_BYTES + 512;  // + slack
constexpr int EPI_THREADS = 256;                 // 8 epilogue warps

// x [B,3,224,224] fp32 -> xp [B][235][232][4] bf16, zero border and zero 4th channel.
// One thread per group of 4 padded pixels (PAD_W = 232 = 58 groups): interior groups read one float4
// from each colour plane (image column 4g-4 .. 4g-1 for group g >= 1; 224 = 56 groups) and write 32 B.
// (Only the RNB_STEM_FUSED=0 fallback of the float path runs this: the fused kernel builds the rows itself.)
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
        const int c0 = 4 * gidx - 4;  // padded pixels 4g..4g+3 are image columns 4g-4..4g-1: one aligned float4 per plane
        if (ih >= 0 && ih < IMG && c0 >= 0 && c0 < IMG) {
            const float* row = x + (1LL * b * 3 * IMG + ih) * IMG + c0;
            const float4 a = __ldg(reinterpret_cast<const float4*>(row));
            const float4 d = __ldg(reinterpret_cast<const float4*>(row + 1LL * IMG * IMG));
            const float4 e = __ldg(reinterpret_cast<const float4*>(row + 2LL * IMG * IMG));
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
            g[0] = d.x; g[1] = d.y; g[2] = d.z; g[3] = d.w;
            bl[0] = e.x; bl[1] = e.y; bl[2] = e.z; bl[3] = e.w;
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
                const int iw = 4 * gidx - 4 + j;  // image column of padded pixel 4g + j
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

// One ring chunk (image rows 4c-4 .. 4c-1 of image b) built by ONE warp from the FP32 NCHW tensor: 4 rows x 56
// groups of 4 pixels = 7 groups per lane, all 21 float4 loads in flight, then RGB0 BF16 pixels, 32 bytes per group.
// Chunks 0 and 57 lie outside the image (zero rows). The 32-byte halo columns on either side are never written
// (zeroed once in the prologue).
__device__ __forceinline__ void stem_fill_chunk_f32(uint32_t dst, const float* __restrict__ x, int b, int c, int lane) {
    if (c == 0 || c == PAIRS + 1) {
        for (int i = lane; i < CHUNK_BYTES / 16; i += 32) st_shared_v4(dst + 16 * i, 0, 0, 0, 0);
        return;
    }
    const float* img = x + (1LL * b * 3 * IMG + (4 * c - 4)) * IMG;
    // the chunk this warp fills next (three chunks on) is requested into L2 now: a loader warp has one chunk of loads
    // in flight, and at HBM latency that was not enough once the MMA side stopped being the bottleneck
    if (lane < 3 && c + 3 <= PAIRS) {
        const float* nxt = img + 1LL * lane * IMG * IMG + 12 * IMG;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"(4 * IMG * 4) : "memory");
    }
    float4 v[7][3];
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const int id = it * 32 + lane, rr = id / 56, g = id - rr * 56;
        const float* src = img + rr * IMG + 4 * g;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            v[it][ch] = __ldg(reinterpret_cast<const float4*>(src + 1LL * ch * IMG * IMG));
    }
    // a lane's 32 bytes are two 16-byte halves; lanes 4..7 of every eight store the second half first, so that the
    // eight lanes of a quarter warp hit eight different 16-byte bank groups (the groups are 32 bytes apart)
    const int first = (lane >> 2) & 1;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const int id = it * 32 + lane, rr = id / 56, g = id - rr * 56;
        uint4 h[2];
        h[0].x = pack_bf16x2(v[it][0].x, v[it][1].x); h[0].y = pack_bf16x2(v[it][2].x, 0.f);
        h[0].z = pack_bf16x2(v[it][0].y, v[it][1].y); h[0].w = pack_bf16x2(v[it][2].y, 0.f);
        h[1].x = pack_bf16x2(v[it][0].z, v[it][1].z); h[1].y = pack_bf16x2(v[it][2].z, 0.f);
        h[1].z = pack_bf16x2(v[it][0].w, v[it][1].w); h[1].w = pack_bf16x2(v[it][2].w, 0.f);
        const uint32_t o = dst + rr * ROW_BYTES + 32 * g + 32;
        const uint4 h0 = first ? h[1] : h[0], h1 = first ? h[0] : h[1];
        st_shared_v4(o + 16 * first, h0.x, h0.y, h0.z, h0.w);
        st_shared_v4(o + 16 * (first ^ 1), h1.x, h1.y, h1.z, h1.w);
    }
}

// The same chunk from a BF16 NCHW tensor (MODE 2): 8-byte loads of 4 pixels per colour plane, RGB0 pixels assembled
// with byte permutes — values already rounded by the host, so the ring holds bit for bit what MODE 1 builds.
__device__ __forceinline__ void stem_fill_chunk_bf16(uint32_t dst, const uint16_t* __restrict__ x, int b, int c, int lane) {
    if (c == 0 || c == PAIRS + 1) {
        for (int i = lane; i < CHUNK_BYTES / 16; i += 32) st_shared_v4(dst + 16 * i, 0, 0, 0, 0);
        return;
    }
    const uint16_t* img = x + (1LL * b * 3 * IMG + (4 * c - 4)) * IMG;
    if (lane < 3 && c + 3 <= PAIRS) {
        const uint16_t* nxt = img + 1LL * lane * IMG * IMG + 12 * IMG;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"(4 * IMG * 2) : "memory");
    }
    uint2 v[7][3];
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const int id = it * 32 + lane, rr = id / 56, g = id - rr * 56;
        const uint16_t* src = img + rr * IMG + 4 * g;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            v[it][ch] = __ldg(reinterpret_cast<const uint2*>(src + 1LL * ch * IMG * IMG));
    }
    const int first = (lane >> 2) & 1;
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const int id = it * 32 + lane, rr = id / 56, g = id - rr * 56;
        uint4 h[2];
        h[0].x = __byte_perm(v[it][0].x, v[it][1].x, 0x5410); h[0].y = v[it][2].x & 0xffffu;
        h[0].z = __byte_perm(v[it][0].x, v[it][1].x, 0x7632); h[0].w = v[it][2].x >> 16;
        h[1].x = __byte_perm(v[it][0].y, v[it][1].y, 0x5410); h[1].y = v[it][2].y & 0xffffu;
        h[1].z = __byte_perm(v[it][0].y, v[it][1].y, 0x7632); h[1].w = v[it][2].y >> 16;
        const uint32_t o = dst + rr * ROW_BYTES + 32 * g + 32;
        const uint4 h0 = first ? h[1] : h[0], h1 = first ? h[0] : h[1];
        st_shared_v4(o + 16 * first, h0.x, h0.y, h0.z, h0.w);
        st_shared_v4(o + 16 * (first ^ 1), h1.x, h1.y, h1.z, h1.w);
    }
}

constexpr int STEM_T_THREADS = 128 + EPI_THREADS + 32;   // warps 0..3, eight epilogue warps, the second MMA issuer
constexpr int T_NBLK = 18;                 // A blocks: input row t = 0..8 of the pair x K step i
constexpr int T_NSLOT = 3;                 // TMEM: D slots of 112 columns at 0, 112, 224; weights at 336 .. 479
constexpr int T_D_PITCH = 112;
constexpr int T_A_COL = T_NSLOT * T_D_PITCH;
constexpr int T_WSTAGE_BYTES = 28 * 32;    // one epilogue warp's block: [28 pooled columns][16 ch] bf16
constexpr int T_STAGE_BYTES = 8 * T_WSTAGE_BYTES;   // x 2 buffers
constexpr int T_NBAR = 2 * NCH + 2 * T_NSLOT;
constexpr int STEM_T_SMEM = 1024 + RING_BYTES + 2 * T_STAGE_BYTES + T_NBAR * 8 + 16;
constexpr size_t T_W_BYTES = static_cast<size_t>(T_NBLK) * 128 * 8 * 4;

// w [64][3][7][7] fp32 + BN -> wt [block = 2t + i][m][8] u32: row m = 32q + 16*rowsel + ocl holds, as 16 BF16 (two per
// word, low half first), window pixels 4i .. 4i+3 x 4 channels of filter row t (rowsel 0, t <= 6) or t - 2 (rowsel 1,
// t >= 2) of output channel 16q + ocl; everything else is zero.
__global__ void stem_pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                         const float* __restrict__ bn_b, const float* __restrict__ bn_m,
                                         const float* __restrict__ bn_v, uint32_t* __restrict__ wt,
                                         float* __restrict__ bias) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T_NBLK * 128 * 8) return;
    const int col = idx & 7, m = (idx >> 3) & 127, blk = idx >> 10;
    const int t = blk >> 1, i = blk & 1;
    const int q = m >> 5, rs = (m >> 4) & 1, oc = 16 * q + (m & 15);
    const int kh = rs ? t - 2 : t;
    const bool row_ok = rs ? t >= 2 : t <= 6;
    double scale = 1.0, shift = 0.0;
    if (bn_w) {
        scale = static_cast<double>(bn_w[oc]) / sqrt(static_cast<double>(bn_v[oc]) + 1e-5);
        shift = static_cast<double>(bn_b[oc]) - static_cast<double>(bn_m[oc]) * scale;
    }
    if (blk == 0 && col == 0 && ((m >> 4) & 1) == 0) bias[oc] = static_cast<float>(shift);   // folded BN shift
    float v[2];
    for (int h = 0; h < 2; ++h) {
        const int e16 = 2 * col + h, pix = 4 * i + (e16 >> 2), c = e16 & 3, kw = pix - 1;
        v[h] = 0.f;
        if (row_ok && kw >= 0 && kw < 7 && c < 3)
            v[h] = static_cast<float>(static_cast<double>(w[((oc * 3 + c) * 7 + kh) * 7 + kw]) * scale);
    }
    wt[idx] = pack_bf16x2(v[0], v[1]);
}

template <int MODE>
__global__ void __launch_bounds__(STEM_T_THREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tm_out, const void* __restrict__ xin,
                 const uint32_t* __restrict__ wt, const float* __restrict__ bias, int B,
                 const float* __restrict__ xin_f32, int nb) {   // (MODE 2: images >= nb are read from xin_f32)
    using namespace ptx;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* stage = smem;                                // 2 x T_STAGE_BYTES (7 KB each: 1024-byte aligned, 128 B swizzle)
    uint8_t* ring = smem + 2 * T_STAGE_BYTES;             // NCH chunks + slack
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING_BYTES);
    uint64_t* ch_full = bars;                      // [NCH]
    uint64_t* ch_empty = bars + NCH;               // [NCH]
    uint64_t* acc_full = bars + 2 * NCH;           // [T_NSLOT]
    uint64_t* acc_empty = bars + 2 * NCH + T_NSLOT;  // [T_NSLOT]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + T_NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_pairs = B * PAIRS;
    const int ppc = num_pairs / static_cast<int>(gridDim.x), prem = num_pairs % static_cast<int>(gridDim.x);
    const int p_begin = static_cast<int>(blockIdx.x) * ppc + min(static_cast<int>(blockIdx.x), prem);
    const int p_end = p_begin + ppc + (static_cast<int>(blockIdx.x) < prem ? 1 : 0);

    if (threadIdx.x == 32) {
        for (int i = 0; i < NCH; ++i) {
            mbar_init(&ch_full[i], MODE == 0 ? 1 : 32);
            mbar_init(&ch_empty[i], 2);   // both MMA issuers release every chunk
        }
        for (int i = 0; i < T_NSLOT; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < RING_BYTES / 16; i += STEM_T_THREADS)   // halo columns, read-past slack
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp >= 4 && warp < 12) {
        // weights -> TMEM (constants): the two warps of a lane quarter write nine A blocks each, lane = row m
        const int q = warp & 3, m = 32 * q + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + T_A_COL;
        for (int blk = (warp - 4) >> 2; blk < T_NBLK; blk += 2) {
            const uint4* src = reinterpret_cast<const uint4*>(wt + (blk * 128 + m) * 8);
            const uint4 lo = __ldg(src), hi = __ldg(src + 1);
            const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            tmem_st_32x8(lane_base + blk * 8, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    griddep_launch_dependents();
    griddep_wait();

    if (warp == 0 || (MODE != 0 && (warp == 2 || warp == 3))) {
        // ===================================================== input: chunks into the ring, in stream order
        const int widx = warp == 0 ? 0 : warp - 1;
        for (StemSteps st(p_begin, p_end); !st.done(); st.next()) {
            const int b = st.b(), nnew = st.new_chunks();
            const int c_first = st.k() + 3 - nnew;
            for (int i = 0; i < nnew; ++i) {
                const int n = st.cn + i, c = c_first + i;
                if (MODE != 0 && (n % 3) != widx) continue;
                mbar_wait(&ch_empty[n & (NCH - 1)], ((n / NCH) & 1) ^ 1);
                uint8_t* dst = ring + (n & (NCH - 1)) * CHUNK_BYTES;
                if (MODE == 0) {
                    if (elect_one()) {
                        const uint8_t* src = static_cast<const uint8_t*>(xin) + (1LL * b * PAD_H + 4 * c + 1) * ROW_BYTES;
                        mbar_expect_tx(&ch_full[n & (NCH - 1)], CHUNK_BYTES);
                        bulk_copy_g2s(dst, src, CHUNK_BYTES, &ch_full[n & (NCH - 1)]);
                    }
                    __syncwarp();
                } else {
                    if (MODE == 1) stem_fill_chunk_f32(smem_u32(dst), static_cast<const float*>(xin), b, c, lane);
                    else if (b < nb) stem_fill_chunk_bf16(smem_u32(dst), static_cast<const uint16_t*>(xin), b, c, lane);
                    else stem_fill_chunk_f32(smem_u32(dst), xin_f32, b, c, lane);
                    fence_proxy_async_smem();
                    mbar_arrive(&ch_full[n & (NCH - 1)]);
                }
            }
        }
    } else if (warp == 1 || warp == 12) {
        // ===================================================== MMA issuers: 18 x (M = 128, N = 112, K = 16), A from TMEM
        // TWO issuing warps take alternate steps. Measured with clock probes on a single issuer (git history, commit
        // "Stem, third form"; profiles/stem_r2.md): per pair 1117 clocks inside the 18 MMA instructions (the tensor pipe's queue holds about two MMAs, so the
        // issuer is blocked for the pipe's 1008 clocks) and then 430 clocks of its own latency — two barrier polls of
        // ~100 clocks each even when the barrier is complete, commits, loop — during which the pipe ran dry: 60 %
        // active. With two issuers one warp's polls overlap the other warp's MMAs.
        // Chunks: a warp releases (tcgen05.commit on ch_empty, count 2) every chunk older than the first chunk of ITS
        // next step, including chunks only the other warp reads; the two warps are never more than two steps apart
        // (three accumulator slots), so a release cannot reach into the previous round of the ring.
        const int X = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, 112);
        const uint64_t b_desc0 = umma_smem_desc(smem_u32(ring), 16, 128, UMMA_LAYOUT_NONE);
        int rel = 0;   // next chunk this warp has to release
        for (StemSteps st(p_begin, p_end); !st.done(); st.next()) {
            const int j = st.step;
            if ((j & 1) != X) continue;
            const int cn_after = st.cn + st.new_chunks();
            // first chunk of this warp's next step (j + 2), or the total number of chunks if there is none
            StemSteps la = st;
            la.next();
            int rel_end = la.cn;
            if (!la.done()) {
                la.next();
                rel_end = la.done() ? la.cn : la.cn + la.new_chunks() - 3;
            }
            for (int m = cn_after - 3; m < cn_after; ++m) mbar_wait(&ch_full[m & (NCH - 1)], (m / NCH) & 1);
            mbar_wait(&acc_empty[j % T_NSLOT], ((j / T_NSLOT) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + (j % T_NSLOT) * T_D_PITCH;
                const uint32_t a0 = tmem_base + T_A_COL;
                // input row t of the pair = row (1 + t) & 3 of chunk cn_after - 3 + (1 + t) / 4
                const uint64_t c0 = b_desc0 + static_cast<uint64_t>((((cn_after - 3) & (NCH - 1)) * CHUNK_BYTES) >> 4);
                const uint64_t c1 = b_desc0 + static_cast<uint64_t>((((cn_after - 2) & (NCH - 1)) * CHUNK_BYTES) >> 4);
                const uint64_t c2 = b_desc0 + static_cast<uint64_t>((((cn_after - 1) & (NCH - 1)) * CHUNK_BYTES) >> 4);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint64_t row = (t < 3 ? c0 : t < 7 ? c1 : c2) + static_cast<uint64_t>((((1 + t) & 3) * ROW_BYTES) >> 4);
                    mma_f16_ts(d, a0 + 16 * t, row, idesc, t != 0);
                    mma_f16_ts(d, a0 + 16 * t + 8, row + 2, idesc, 1);
                }
                tc_commit(&acc_full[j % T_NSLOT]);
                for (int n = rel; n < rel_end; ++n) tc_commit(&ch_empty[n & (NCH - 1)]);
            }
            rel = rel_end;
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================================================== epilogue
        // warp quarter q = warp & 3: lanes 0..15 = conv row a (2k), lanes 16..31 = conv row b (2k+1) of channels
        // 16q .. 16q+15; the two warps of a quarter split the pooled columns (h = 0: 0..27, h = 1: 28..55).
        // The eight warps never meet: each stages its own [28 pooled columns][16 channels] block (32-byte rows: the
        // 32-bit stores of a warp fall into 32 different banks) and sends it off with its own TMA store; the next
        // pair's accumulators are requested as soon as this pair's have been reduced to 14 registers.
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int rs = lane >> 4, oc = 16 * q + (lane & 15);
        const float bias_c = __ldg(bias + oc);
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (h ? 48 : 0);
        uint8_t* const stg = stage + (warp - 4) * (2 * T_WSTAGE_BYTES);
        uint32_t carry[7];  // conv row 2k-1 (the previous pair's row b), this lane's 14 pooled columns, after bias + ReLU
#pragma unroll
        for (int u = 0; u < 7; ++u) carry[u] = 0u;
        int vb = 0;
        uint32_t c0[32], c1[32];   // 64 output columns of this lane's conv row: 0..63 (h = 0) or 48..111 (h = 1)
        StemSteps st(p_begin, p_end);
        if (!st.done()) {
            mbar_wait(&acc_full[0], 0);
            tc_fence_after();
            tmem_ld_32x32(lane_addr, c0);
            tmem_ld_32x32(lane_addr + 32, c1);
        }
        while (!st.done()) {
            const int b = st.b(), k = st.k(), j = st.step;
            const bool warm = st.warm;
            st.next();
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&acc_empty[j % T_NSLOT]);
            auto col = [&](int i) { return __uint_as_float(i < 32 ? c0[i] : c1[i - 32]); };
            // horizontal 3-tap max, bias, ReLU, BF16: pooled column 28h + jj reads columns 2jj-1, 2jj, 2jj+1 (+8 for h = 1)
            uint32_t yp[14];
            if (h == 0) {   // (warp-uniform: two fully unrolled bodies with compile-time register indices)
#pragma unroll
                for (int jj = 0; jj < 28; jj += 2) {
                    const int i0 = 2 * jj, i1 = i0 + 2;
                    const float y0 = fmaxf(fmaxf(col(i0), col(i0 + 1)), i0 > 0 ? col(i0 - 1) : col(i0)) + bias_c;
                    const float y1 = fmaxf(fmaxf(col(i1), col(i1 + 1)), col(i1 - 1)) + bias_c;
                    yp[jj >> 1] = pack_relu_bf16x2(y0, y1);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 28; jj += 2) {
                    const int i0 = 2 * jj + 8, i1 = i0 + 2;
                    const float y0 = fmaxf(fmaxf(col(i0), col(i0 + 1)), col(i0 - 1)) + bias_c;
                    const float y1 = fmaxf(fmaxf(col(i1), col(i1 + 1)), col(i1 - 1)) + bias_c;
                    yp[jj >> 1] = pack_relu_bf16x2(y0, y1);
                }
            }
            if (!st.done()) {   // the next pair's accumulators
                const int jn = st.step;
                mbar_wait(&acc_full[jn % T_NSLOT], (jn / T_NSLOT) & 1);
                tc_fence_after();
                tmem_ld_32x32(lane_addr + (jn % T_NSLOT) * T_D_PITCH, c0);
                tmem_ld_32x32(lane_addr + (jn % T_NSLOT) * T_D_PITCH + 32, c1);
            }
            if (k == 0) {
#pragma unroll
                for (int u = 0; u < 7; ++u) carry[u] = 0u;   // no conv row above the image (ReLU output is >= 0)
            }
            // vertical max with the partner lane (the pair's other conv row): lanes 0..15 finish pooled columns
            // 28h + 0..13, lanes 16..31 columns 28h + 14..27; each lane sends the half the other one finishes
            uint32_t res[7];
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                const uint32_t mine = rs ? yp[7 + u] : yp[u];
                const uint32_t other = __shfl_xor_sync(0xffffffffu, rs ? yp[u] : yp[7 + u], 16);
                const uint32_t rowb = rs ? mine : other;
                res[u] = bf16x2_max(bf16x2_max(mine, other), carry[u]);
                carry[u] = rowb;
            }
            if (warm) continue;  // the pair before this CTA's range: only its second conv row was wanted
            if (lane == 0) tma_store_wait_read<1>();   // the store that last read this staging buffer (two pairs ago)
            __syncwarp();
            {
                // channel pairs: the even lane of two neighbouring channels takes pooled column 2u of both, the odd
                // lane column 2u+1 — one 32-bit store each
                const int odd = lane & 1;
                const uint32_t sbase = smem_u32(stg) + vb * T_WSTAGE_BYTES + (14 * rs + odd) * 32 + ((lane & 15) >> 1) * 4;
#pragma unroll
                for (int u = 0; u < 7; ++u) {
                    const uint32_t recv = __shfl_xor_sync(0xffffffffu, res[u], 1);
                    const uint32_t word = odd ? __byte_perm(recv, res[u], 0x7632) : __byte_perm(res[u], recv, 0x5410);
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + 2 * u * 32), "r"(word) : "memory");
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_out, stg + vb * T_WSTAGE_BYTES, 16 * q, (b * POOL + k) * POOL + 28 * h);
                tma_store_commit();
            }
            vb ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

bool stem_fused_enabled() {
    static const bool on = !(getenv("RNB_STEM_FUSED") && atoi(getenv("RNB_STEM_FUSED")) == 0);
    return on;
}

size_t stem_tc_packed_input_bytes(int B) { return 1ull * B * PAD_H * ROW_BYTES; }
size_t stem_tc_packed_weight_bytes() { return T_W_BYTES; }

cudaError_t launch_stem_tc_pack_weights(const float* w, const float* bn_w, const float* bn_b,
                                        const float* bn_m, const float* bn_v, void* wk, float* bias,
                                        cudaStream_t s) {
    stem_pack_weights_kernel<<<(T_NBLK * 128 * 8 + 255) / 256, 256, 0, s>>>(w, bn_w, bn_b, bn_m, bn_v,
                                                                         static_cast<uint32_t*>(wk), bias);
    return cudaGetLastError();
}

cudaError_t stem_tc_init() {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_T_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_T_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_T_SMEM);
    return e;
}

static cudaError_t launch_stem_tc_kernel(int mode, const void* in, const void* wk, const float* bias, void* out, int B,
                                         cudaStream_t s, const float* in_f32 = nullptr, int nb = 0) {
    const int pairs = B * PAIRS;
    const int grid = pairs < num_sms() ? pairs : num_sms();
    // output [B*56*56 pooled pixels][64 ch] bf16; one store = 28 pooled pixels x 16 channels (an epilogue warp's block)
    CUtensorMap tm;
    const uint64_t dims[2] = {64, 1ull * B * POOL * POOL}, strides[1] = {128};
    const uint32_t box[2] = {16, 28};
    if (make_tiled_nd(&tm, TmDtype::BF16, out, 2, dims, strides, box, false) != 0) return cudaErrorInvalidValue;
    const uint32_t* wt = static_cast<const uint32_t*>(wk);
    if (mode == 0)
        return launch_pdl_small(stem_tc_kernel<0>, dim3(grid), dim3(STEM_T_THREADS), STEM_T_SMEM, s, tm, in, wt, bias, B,
                                in_f32, nb);
    if (mode == 2)
        return launch_pdl_small(stem_tc_kernel<2>, dim3(grid), dim3(STEM_T_THREADS), STEM_T_SMEM, s, tm, in, wt, bias, B,
                                in_f32, nb);
    return launch_pdl_small(stem_tc_kernel<1>, dim3(grid), dim3(STEM_T_THREADS), STEM_T_SMEM, s, tm, in, wt, bias, B,
                            in_f32, nb);
}

// Images [0, nb) from x_bf16 (BF16 NCHW: the FP32 image rounded to nearest-even by the host, host_pack.cpp), images
// [nb, B) from x_f32 (FP32 NCHW, indexed by the same image number) -> out: one launch
cudaError_t launch_stem_tc_from_mixed(const void* x_bf16, const float* x_f32, int nb, const void* wk, const float* bias,
                                      void* out, int B, cudaStream_t s) {
    return launch_stem_tc_kernel(2, x_bf16, wk, bias, out, B, s, x_f32, nb);
}

// xp (packed NHWC4 BF16, written by a pack kernel) -> out NHWC bf16 [B,56,56,64]
cudaError_t launch_stem_tc_from_packed(const void* xp, const void* wk, const float* bias, void* out, int B,
                                       cudaStream_t s) {
    return launch_stem_tc_kernel(0, xp, wk, bias, out, B, s);
}

// The float path in two "parts" (the callers time them separately). Fused form (default): part 0 does nothing and
// part 1 reads x fp32 NCHW [B,3,224,224] directly. RNB_STEM_FUSED=0: part 0 = layout pre-pass x -> xp (scratch of
// stem_tc_packed_input_bytes(B)), part 1 = xp -> out.
cudaError_t launch_stem_tc_part(int part, const float* x, void* xp, const void* wk, const float* bias,
                                void* out, int B, cudaStream_t s) {
    if (stem_fused_enabled()) return part == 0 ? cudaSuccess : launch_stem_tc_kernel(1, x, wk, bias, out, B, s);
    if (part == 0) {
        const int64_t total = 1LL * B * PAD_H * (PAD_W / 4);
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
        if (blocks > cap) blocks = cap;
        stem_pack_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, static_cast<uint4*>(xp), B);
        return cudaGetLastError();
    }
    return launch_stem_tc_kernel(0, xp, wk, bias, out, B, s);
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
