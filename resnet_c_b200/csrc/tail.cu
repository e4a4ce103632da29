// tail.cu — network tail: global average pool + fully-connected layer, replacing avgPool2dKernel
// (/root/reference/cuda/ops.cu:80-108) and linearForwardKernel (ops.cu:110-128) as called at
// cuda/inference/main.cu:213-224. Logits stay FP32. The arg-max (main.cu:243-251) is
// launch_argmax_f32 in ops_f32.cu.
//
// Layout choice: the pooled features are written TRANSPOSED, pooledT[C][n], so that in the FC kernel
// thread b reads pooledT[k][b] — a warp reads 128 contiguous bytes per k — while the 8 weights of
// the block's 8 classes for that k are one broadcast shared-memory read.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "internal.h"

namespace rnb {

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// [n][HW][C] T -> pooledT[C][n] fp32. One thread per (image, 16-byte channel group); consecutive
// threads read consecutive 16-byte groups of the same pixel (coalesced). The sum is sequential over
// pixels in FP32 as in avgPool2dKernel, and divided by k twice when HW = k*k (ops.cu:107).
template <typename T>
__global__ void avgpool_nhwc_kernel(const T* __restrict__ x, float* __restrict__ pooledT,
                                    __nv_bfloat16* __restrict__ pooled_bf16, int n, int HW, int C, int ksq) {
    constexpr int VEC = 16 / sizeof(T);
    const int groups = C / VEC;
    const int64_t total = 1LL * n * groups;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched with PDL: x is the previous kernel's output
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int gidx = static_cast<int>(i % groups);
        const int b = static_cast<int>(i / groups);
        const T* xp = x + 1LL * b * HW * C + gidx * VEC;
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
        auto add_pixel = [&](const uint4& v) {
            if constexpr (sizeof(T) == 2) {
                const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[2 * e] += __uint_as_float(u[e] << 16);
                    acc[2 * e + 1] += __uint_as_float(u[e] & 0xFFFF0000u);
                }
            } else {
                acc[0] += __uint_as_float(v.x);
                acc[1] += __uint_as_float(v.y);
                acc[2] += __uint_as_float(v.z);
                acc[3] += __uint_as_float(v.w);
            }
        };
        // Batches of NB independent 16-byte loads are issued before the first add (the adds stay in pixel order):
        // two latency rounds for the 7x7 map of the network tail (the kernel is latency-, not bandwidth-bound: 51 MB).
        constexpr int NB = 25;
        int p = 0;
        for (; p + NB <= HW; p += NB) {
            uint4 v[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(xp + 1LL * (p + j) * C));
#pragma unroll
            for (int j = 0; j < NB; ++j) add_pixel(v[j]);
        }
        {
            uint4 v[NB - 1];
#pragma unroll
            for (int j = 0; j < NB - 1; ++j)
                v[j] = p + j < HW ? __ldg(reinterpret_cast<const uint4*>(xp + 1LL * (p + j) * C)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < NB - 1; ++j)
                if (p + j < HW) add_pixel(v[j]);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            acc[e] = ksq > 0 ? acc[e] / static_cast<float>(ksq) / static_cast<float>(ksq)
                             : acc[e] / static_cast<float>(HW);
        if (sizeof(T) != 2 || !pooled_bf16) {
            // transposed FP32 for fc_kernel (scattered 4-byte stores: fine for the CUDA-core FC path only)
#pragma unroll
            for (int e = 0; e < VEC; ++e) pooledT[1LL * (gidx * VEC + e) * n + b] = acc[e];
        }
        if constexpr (sizeof(T) == 2) {
            if (pooled_bf16) {  // row-major BF16 copy: the A operand of the tensor-core FC (+ row-major FP32)
                float4* pr = reinterpret_cast<float4*>(pooledT + 1LL * b * C + gidx * VEC);
                pr[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                pr[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                uint4 o;
                o.x = pack2(acc[0], acc[1]); o.y = pack2(acc[2], acc[3]);
                o.z = pack2(acc[4], acc[5]); o.w = pack2(acc[6], acc[7]);
                *reinterpret_cast<uint4*>(pooled_bf16 + 1LL * b * C + gidx * VEC) = o;
            }
        }
    }
}

// FP8 (E4M3) input [n][HW][C] with per-tensor scale: thread = (image, 16 channels); row-major FP32 + BF16 outputs
// (the A operand of the tensor-core FC). Sum in FP32 in pixel order, then x scale, then the double division.
__global__ void avgpool_nhwc_fp8_kernel(const uint8_t* __restrict__ x, float* __restrict__ pooled,
                                        __nv_bfloat16* __restrict__ pooled_bf16, int n, int HW, int C, int ksq,
                                        float in_scale) {
    const int groups = C / 16;
    const int64_t total = 1LL * n * groups;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int gidx = static_cast<int>(i % groups);
        const int b = static_cast<int>(i / groups);
        const uint8_t* xp = x + 1LL * b * HW * C + gidx * 16;
        float acc[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        auto add_pixel = [&](const uint4& v) {
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t h01, h23;
                asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h01) : "h"(static_cast<uint16_t>(u[q] & 0xFFFFu)));
                asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h23) : "h"(static_cast<uint16_t>(u[q] >> 16)));
                const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
                const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
                acc[q * 4 + 0] += f01.x; acc[q * 4 + 1] += f01.y; acc[q * 4 + 2] += f23.x; acc[q * 4 + 3] += f23.y;
            }
        };
        // all loads of a batch of NB pixels are issued before the first add (two latency rounds for a 7 x 7 map)
        constexpr int NB = 25;
        for (int p0 = 0; p0 < HW; p0 += NB) {
            uint4 v[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j)
                v[j] = p0 + j < HW ? __ldg(reinterpret_cast<const uint4*>(xp + 1LL * (p0 + j) * C)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (p0 + j < HW) add_pixel(v[j]);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            acc[e] *= in_scale;
            acc[e] = ksq > 0 ? acc[e] / static_cast<float>(ksq) / static_cast<float>(ksq) : acc[e] / static_cast<float>(HW);
        }
        float4* pr = reinterpret_cast<float4*>(pooled + 1LL * b * C + gidx * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) pr[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        uint4 o0, o1;
        o0.x = pack2(acc[0], acc[1]); o0.y = pack2(acc[2], acc[3]); o0.z = pack2(acc[4], acc[5]); o0.w = pack2(acc[6], acc[7]);
        o1.x = pack2(acc[8], acc[9]); o1.y = pack2(acc[10], acc[11]); o1.z = pack2(acc[12], acc[13]); o1.w = pack2(acc[14], acc[15]);
        uint4* pb = reinterpret_cast<uint4*>(pooled_bf16 + 1LL * b * C + gidx * 16);
        pb[0] = o0;
        pb[1] = o1;
    }
}

// fc.weight [classes][C] fp32 -> [cpad][C] bf16 (rows >= classes zero), fc.bias -> [cpad] fp32 (zero padded)
__global__ void fc_pack_kernel(const float* __restrict__ w, const float* __restrict__ b,
                               __nv_bfloat16* __restrict__ wq, float* __restrict__ bq, int classes, int C,
                               int cpad) {
    const int64_t total = 1LL * cpad * C;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int o = static_cast<int>(i / C);
        wq[i] = __float2bfloat16_rn(o < classes ? w[i] : 0.f);
        if (i < cpad) bq[i] = i < classes ? b[i] : 0.f;
    }
}

// logits[b][o] = sum_k pooledT[k][b] * w[o][k] + bias[o].
// Block = 16 classes x 64 images; thread = 2 images x 16 classes (32 FP32 accumulators), so one k costs
// 2 coalesced loads of pooledT + 4 broadcast LDS.128 of weights for 32 FMAs (the first version, 1 image
// x 8 classes, was instruction-issue-bound at 14 instructions per 8 FMAs). The block's 8 warps split K
// (warp w takes k = w, w+8, ..) and meet in shared memory with a fixed summation order.
// FP32 FMA on CUDA cores: the layer is 2 MMAC per image.
constexpr int FC_OUT = 16;
constexpr int FC_IMG = 64;       // images per block (2 per lane)
constexpr int FC_WARPS = 8;
constexpr int FC_THREADS = FC_WARPS * 32;
constexpr int FC_KCHUNK = 512;   // weights staged per pass: 512 x 16 x 4 B = 32 KB
__global__ void __launch_bounds__(FC_THREADS)
fc_kernel(const float* __restrict__ pooledT, const float* __restrict__ w, const float* __restrict__ bias,
          float* __restrict__ out, int n, int C, int classes) {
    __shared__ __align__(16) float ws[FC_KCHUNK][FC_OUT];   // 32 KB, later reused for the partial sums
    const int o0 = blockIdx.x * FC_OUT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.y * FC_IMG + lane, b1 = b0 + 32;
    const bool act0 = b0 < n, act1 = b1 < n;
    float acc0[FC_OUT], acc1[FC_OUT];
#pragma unroll
    for (int j = 0; j < FC_OUT; ++j) acc0[j] = acc1[j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += FC_KCHUNK) {
        const int kc = min(FC_KCHUNK, C - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kc * FC_OUT; i += FC_THREADS) {
            const int j = i / kc, k = i - j * kc;  // consecutive threads read consecutive k of one row
            const int o = o0 + j;
            ws[k][j] = o < classes ? __ldg(w + 1LL * o * C + k0 + k) : 0.f;
        }
        __syncthreads();
        const float* xp = pooledT + 1LL * k0 * n;
#pragma unroll 16
        for (int k = warp; k < kc; k += FC_WARPS) {
            const float x0 = act0 ? __ldg(xp + 1LL * k * n + b0) : 0.f;
            const float x1 = act1 ? __ldg(xp + 1LL * k * n + b1) : 0.f;
#pragma unroll
            for (int q = 0; q < FC_OUT / 4; ++q) {
                const float4 wv = *reinterpret_cast<const float4*>(&ws[k][q * 4]);
                acc0[q * 4 + 0] = fmaf(x0, wv.x, acc0[q * 4 + 0]); acc1[q * 4 + 0] = fmaf(x1, wv.x, acc1[q * 4 + 0]);
                acc0[q * 4 + 1] = fmaf(x0, wv.y, acc0[q * 4 + 1]); acc1[q * 4 + 1] = fmaf(x1, wv.y, acc1[q * 4 + 1]);
                acc0[q * 4 + 2] = fmaf(x0, wv.z, acc0[q * 4 + 2]); acc1[q * 4 + 2] = fmaf(x1, wv.z, acc1[q * 4 + 2]);
                acc0[q * 4 + 3] = fmaf(x0, wv.w, acc0[q * 4 + 3]); acc1[q * 4 + 3] = fmaf(x1, wv.w, acc1[q * 4 + 3]);
            }
        }
    }
    // partial sums: part[warp][image 0..63][class 0..15] (+1 pad) = 8 x 64 x 17 floats = 34.8 KB > ws,
    // so reduce in two halves of 4 warps through the 32 KB of ws: [4][64][16] floats = 16 KB per half
    float* part = &ws[0][0];
    // this thread finally owns the outputs (image = tid >> 2 (0..63), classes 4*(tid&3)..+3)
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    for (int half = 0; half < 2; ++half) {
        __syncthreads();
        if ((warp >> 2) == half) {
            const int wl = warp & 3;
#pragma unroll
            for (int j = 0; j < FC_OUT; ++j) {
                part[(wl * FC_IMG + lane) * FC_OUT + j] = acc0[j];
                part[(wl * FC_IMG + lane + 32) * FC_OUT + j] = acc1[j];
            }
        }
        __syncthreads();
        const int img = threadIdx.x >> 2, c4 = (threadIdx.x & 3) * 4;
#pragma unroll
        for (int wl = 0; wl < 4; ++wl) {
            const float4 v = *reinterpret_cast<const float4*>(&part[(wl * FC_IMG + img) * FC_OUT + c4]);
            res[0] += v.x; res[1] += v.y; res[2] += v.z; res[3] += v.w;
        }
    }
    const int img = blockIdx.y * FC_IMG + (threadIdx.x >> 2);
    if (img < n) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + (threadIdx.x & 3) * 4 + j;
            if (o < classes) out[1LL * img * classes + o] = res[j] + (bias ? __ldg(bias + o) : 0.f);
        }
    }
}

// [rows][cols] -> [cols][rows], fp32 (debug readback of pooledT and the NCHW tail API).
__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                     int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[1LL * r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[1LL * c * rows + r] = tile[threadIdx.x][i];
    }
}

}  // namespace

cudaError_t launch_fc_pack(const float* w, const float* b, void* wq, float* bq, int classes, int C, int cpad,
                           cudaStream_t s) {
    fc_pack_kernel<<<1024, 256, 0, s>>>(w, b, static_cast<__nv_bfloat16*>(wq), bq, classes, C, cpad);
    return cudaGetLastError();
}

cudaError_t launch_avgpool_nhwc(const void* x, float* pooledT, void* pooled_bf16, int B, int HW, int C, int esz,
                                cudaStream_t s, float in_scale) {
    int ksq = 0;
    for (int k = 1; k * k <= HW; ++k)
        if (k * k == HW) ksq = k;
    if (esz == 1) {
        if (!pooled_bf16 || C % 16 != 0) return cudaErrorInvalidValue;
        const int64_t tot = 1LL * B * (C / 16);
        return launch_pdl_small(avgpool_nhwc_fp8_kernel, dim3(static_cast<int>((tot + 63) / 64)), dim3(64), 0, s,
                                static_cast<const uint8_t*>(x), pooledT, static_cast<__nv_bfloat16*>(pooled_bf16), B, HW, C,
                                ksq, in_scale);
    }
    const int64_t total = 1LL * B * (C * esz / 16);
    const int blocks = static_cast<int>((total + 127) / 128);
    if (esz == 2)  // 64-thread blocks: ~145 registers per thread, and the whole grid should be resident at once
        return launch_pdl_small(avgpool_nhwc_kernel<__nv_bfloat16>, dim3(static_cast<int>((total + 63) / 64)), dim3(64), 0, s,
                                static_cast<const __nv_bfloat16*>(x), pooledT,
                                static_cast<__nv_bfloat16*>(pooled_bf16), B, HW, C, ksq);
    return launch_pdl_small(avgpool_nhwc_kernel<float>, dim3(blocks), dim3(128), 0, s, static_cast<const float*>(x),
                            pooledT, static_cast<__nv_bfloat16*>(nullptr), B, HW, C, ksq);
}

cudaError_t launch_fc(const float* pooledT, const float* w, const float* bias, float* logits, int B,
                      int C, int classes, cudaStream_t s) {
    dim3 grid((classes + FC_OUT - 1) / FC_OUT, (B + FC_IMG - 1) / FC_IMG);
    fc_kernel<<<grid, FC_THREADS, 0, s>>>(pooledT, w, bias, logits, B, C, classes);
    return cudaGetLastError();
}

cudaError_t launch_transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_f32_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
    return cudaGetLastError();
}

}  // namespace rnb
