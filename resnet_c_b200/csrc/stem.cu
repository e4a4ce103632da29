// stem.cu — network stem on CUDA cores: conv 7x7/2 pad 3 (3 -> 64) + folded BN + ReLU, then the
// 3x3/2 pad 1 max-pool, replacing conv2dForwardKernel + batchNorm2dForwardKernel + reluForwardKernel
// + maxPool2dKernel as chained at /root/reference/cuda/inference/main.cu:176-192.
//
// Input is the reference's FP32 NCHW image tensor; output is NHWC in the activation type so the
// tensor-core layers can consume it directly.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"

namespace rnb {

namespace {

constexpr int TILE_H = 8;     // output rows per block
constexpr int TILE_W = 16;    // output cols per block
constexpr int IN_H = (TILE_H - 1) * 2 + 7;  // 21
constexpr int IN_W = (TILE_W - 1) * 2 + 7;  // 37
constexpr int IN_W_PAD = IN_W + 2;          // 39 (odd stride: stride-2 reads hit distinct banks)
constexpr int STEM_THREADS = 256;
constexpr int STEM_SMEM = (147 * 64 + 3 * IN_H * IN_W_PAD) * 4;

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Each thread: one output pixel x 32 output channels. Weights live in smem as [tap][oc] so a warp
// (32 pixels, same channel half) reads them as broadcast float4s.
template <typename T>
__global__ void __launch_bounds__(STEM_THREADS)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 T* __restrict__ out, int H, int W, int OH, int OW) {
    extern __shared__ float smem[];
    float* ws = smem;              // [147][64]
    float* xs = smem + 147 * 64;   // [3][IN_H][IN_W_PAD]
    const int b = blockIdx.z;
    const int oh0 = blockIdx.y * TILE_H;
    const int ow0 = blockIdx.x * TILE_W;
    const int ih0 = oh0 * 2 - 3;
    const int iw0 = ow0 * 2 - 3;

    for (int i = threadIdx.x; i < 147 * 64; i += STEM_THREADS) {
        const int oc = i & 63, tap = i >> 6;  // tap = (ic*7 + kh)*7 + kw, the reference's loop order
        ws[i] = w[oc * 147 + tap];
    }
    const float* xb = x + 1LL * b * 3 * H * W;
    for (int i = threadIdx.x; i < 3 * IN_H * IN_W; i += STEM_THREADS) {
        const int c = i / (IN_H * IN_W);
        const int rem = i - c * (IN_H * IN_W);
        const int r = rem / IN_W, col = rem - r * IN_W;
        const int ih = ih0 + r, iw = iw0 + col;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(xb + (1LL * c * H + ih) * W + iw);
        xs[(c * IN_H + r) * IN_W_PAD + col] = v;
    }
    __syncthreads();

    const int pix = threadIdx.x & 127;   // pixel within the tile
    const int half = threadIdx.x >> 7;   // channel half (0: 0..31, 1: 32..63)
    const int py = pix / TILE_W, px = pix % TILE_W;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int c = 0; c < 3; ++c) {
        for (int kh = 0; kh < 7; ++kh) {
            const float* xrow = xs + (c * IN_H + py * 2 + kh) * IN_W_PAD + px * 2;
            const float* wrow = ws + ((c * 7 + kh) * 7) * 64 + half * 32;
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
                const float v = xrow[kw];
                const float4* w4 = reinterpret_cast<const float4*>(wrow + kw * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 ww = w4[j];
                    acc[j * 4 + 0] = fmaf(v, ww.x, acc[j * 4 + 0]);
                    acc[j * 4 + 1] = fmaf(v, ww.y, acc[j * 4 + 1]);
                    acc[j * 4 + 2] = fmaf(v, ww.z, acc[j * 4 + 2]);
                    acc[j * 4 + 3] = fmaf(v, ww.w, acc[j * 4 + 3]);
                }
            }
        }
    }
    const int oh = oh0 + py, ow = ow0 + px;
    if (oh < OH && ow < OW) {
        T* op = out + ((1LL * b * OH + oh) * OW + ow) * 64 + half * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i] + __ldg(bias + half * 32 + i), 0.f);
        if constexpr (sizeof(T) == 2) {
            uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 o;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(acc[j * 8 + 0], acc[j * 8 + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(acc[j * 8 + 2], acc[j * 8 + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(acc[j * 8 + 4], acc[j * 8 + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(acc[j * 8 + 6], acc[j * 8 + 7]);
                o.x = *reinterpret_cast<uint32_t*>(&t0);
                o.y = *reinterpret_cast<uint32_t*>(&t1);
                o.z = *reinterpret_cast<uint32_t*>(&t2);
                o.w = *reinterpret_cast<uint32_t*>(&t3);
                o4[j] = o;
            }
        } else {
            float4* o4 = reinterpret_cast<float4*>(op);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o4[j] = make_float4(rna_tf32(acc[j * 4 + 0]), rna_tf32(acc[j * 4 + 1]),
                                    rna_tf32(acc[j * 4 + 2]), rna_tf32(acc[j * 4 + 3]));
        }
    }
}

// 3x3/2 pad 1 max-pool over NHWC; one thread per (pixel, 16-byte channel group).
// -inf init and skipped out-of-bounds taps as maxPool2dKernel (ops.cu:64-72).
template <typename T>
__global__ void maxpool_nhwc_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W,
                                    int C, int OH, int OW) {
    constexpr int VEC = 16 / sizeof(T);
    const int groups = C / VEC;
    const int64_t total = 1LL * B * OH * OW * groups;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int gidx = static_cast<int>(i % groups);
        int64_t t = i / groups;
        const int ow = static_cast<int>(t % OW);
        t /= OW;
        const int oh = static_cast<int>(t % OH);
        const int b = static_cast<int>(t / OH);
        float m[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) m[e] = -INFINITY;
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = oh * 2 - 1 + kh;
            if (ih < 0 || ih >= H) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = ow * 2 - 1 + kw;
                if (iw < 0 || iw >= W) continue;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(
                    x + ((1LL * b * H + ih) * W + iw) * C + gidx * VEC));
                if constexpr (sizeof(T) == 2) {
                    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        m[2 * e] = fmaxf(m[2 * e], __uint_as_float(u[e] << 16));
                        m[2 * e + 1] = fmaxf(m[2 * e + 1], __uint_as_float(u[e] & 0xFFFF0000u));
                    }
                } else {
                    m[0] = fmaxf(m[0], __uint_as_float(v.x));
                    m[1] = fmaxf(m[1], __uint_as_float(v.y));
                    m[2] = fmaxf(m[2], __uint_as_float(v.z));
                    m[3] = fmaxf(m[3], __uint_as_float(v.w));
                }
            }
        }
        uint4 o;
        if constexpr (sizeof(T) == 2) {
            // values are already bf16-representable: truncation is exact
            o.x = (__float_as_uint(m[0]) >> 16) | (__float_as_uint(m[1]) & 0xFFFF0000u);
            o.y = (__float_as_uint(m[2]) >> 16) | (__float_as_uint(m[3]) & 0xFFFF0000u);
            o.z = (__float_as_uint(m[4]) >> 16) | (__float_as_uint(m[5]) & 0xFFFF0000u);
            o.w = (__float_as_uint(m[6]) >> 16) | (__float_as_uint(m[7]) & 0xFFFF0000u);
        } else {
            o.x = __float_as_uint(m[0]);
            o.y = __float_as_uint(m[1]);
            o.z = __float_as_uint(m[2]);
            o.w = __float_as_uint(m[3]);
        }
        *reinterpret_cast<uint4*>(out + ((1LL * b * OH + oh) * OW + ow) * C + gidx * VEC) = o;
    }
}

}  // namespace

cudaError_t launch_stem_conv(const float* x, const float* w_folded, const float* bias, void* out,
                             int B, int H, int W, int esz, cudaStream_t s) {
    const int OH = (6 + H - 7) / 2 + 1, OW = (6 + W - 7) / 2 + 1;
    dim3 grid((OW + TILE_W - 1) / TILE_W, (OH + TILE_H - 1) / TILE_H, B);
    // per device and cheap; this kernel is off the hot path (non-224 inputs / RNB_NO_STEM_TC)
    cudaFuncSetAttribute(stem_conv_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    cudaFuncSetAttribute(stem_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    if (esz == 2)
        stem_conv_kernel<__nv_bfloat16><<<grid, STEM_THREADS, STEM_SMEM, s>>>(
            x, w_folded, bias, static_cast<__nv_bfloat16*>(out), H, W, OH, OW);
    else
        stem_conv_kernel<float><<<grid, STEM_THREADS, STEM_SMEM, s>>>(
            x, w_folded, bias, static_cast<float*>(out), H, W, OH, OW);
    return cudaGetLastError();
}

cudaError_t launch_maxpool_nhwc(const void* x, void* out, int B, int H, int W, int C, int esz,
                                cudaStream_t s) {
    const int OH = (2 + H - 3) / 2 + 1, OW = (2 + W - 3) / 2 + 1;
    const int64_t total = 1LL * B * OH * OW * (C * esz / 16);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
    if (blocks > cap) blocks = cap;
    if (esz == 2)
        maxpool_nhwc_kernel<__nv_bfloat16><<<static_cast<int>(blocks), 256, 0, s>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), B, H, W, C, OH, OW);
    else
        maxpool_nhwc_kernel<float><<<static_cast<int>(blocks), 256, 0, s>>>(
            static_cast<const float*>(x), static_cast<float*>(out), B, H, W, C, OH, OW);
    return cudaGetLastError();
}

}  // namespace rnb
