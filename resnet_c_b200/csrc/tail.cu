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
#include <cuda_runtime.h>

#include "internal.h"

namespace rnb {

namespace {

// [n][HW][C] T -> pooledT[C][n] fp32. One thread per (image, 16-byte channel group); consecutive
// threads read consecutive 16-byte groups of the same pixel (coalesced). The sum is sequential over
// pixels in FP32 as in avgPool2dKernel, and divided by k twice when HW = k*k (ops.cu:107).
template <typename T>
__global__ void avgpool_nhwc_kernel(const T* __restrict__ x, float* __restrict__ pooledT, int n, int HW,
                                    int C, int ksq) {
    constexpr int VEC = 16 / sizeof(T);
    const int groups = C / VEC;
    const int64_t total = 1LL * n * groups;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int gidx = static_cast<int>(i % groups);
        const int b = static_cast<int>(i / groups);
        const T* xp = x + 1LL * b * HW * C + gidx * VEC;
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
        for (int p = 0; p < HW; ++p) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(xp + 1LL * p * C));
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
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            pooledT[1LL * (gidx * VEC + e) * n + b] =
                ksq > 0 ? acc[e] / static_cast<float>(ksq) / static_cast<float>(ksq)
                        : acc[e] / static_cast<float>(HW);
    }
}

// logits[b][o] = sum_k pooledT[k][b] * w[o][k] + bias[o].
// Block = FC_OUT classes x 32 images; its FC_WARPS warps split K (warp w takes k = w, w+FC_WARPS, ..)
// so ~1000 blocks x 8 warps keep every SM busy; lane = image, so a warp reads 128 contiguous bytes
// of pooledT per k while the 8 class weights for that k are two broadcast LDS.128. Partial sums
// meet in shared memory. FP32 FMA on CUDA cores: the layer is 2 MMAC per image.
constexpr int FC_OUT = 8;
constexpr int FC_WARPS = 8;
constexpr int FC_THREADS = FC_WARPS * 32;
constexpr int FC_KCHUNK = 1024;  // weights staged per pass: 1024 x 8 x 4 B = 32 KB
__global__ void __launch_bounds__(FC_THREADS)
fc_kernel(const float* __restrict__ pooledT, const float* __restrict__ w, const float* __restrict__ bias,
          float* __restrict__ out, int n, int C, int classes) {
    __shared__ __align__(16) float ws[FC_KCHUNK][FC_OUT];
    __shared__ float part[FC_WARPS][32][FC_OUT + 1];
    const int o0 = blockIdx.x * FC_OUT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y * 32 + lane;
    const bool active = b < n;
    float acc[FC_OUT];
#pragma unroll
    for (int j = 0; j < FC_OUT; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += FC_KCHUNK) {
        const int kc = min(FC_KCHUNK, C - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kc * FC_OUT; i += FC_THREADS) {
            const int j = i / kc, k = i - j * kc;  // consecutive threads read consecutive k of one row
            const int o = o0 + j;
            ws[k][j] = o < classes ? __ldg(w + 1LL * o * C + k0 + k) : 0.f;
        }
        __syncthreads();
        if (active) {
            const float* xp = pooledT + 1LL * k0 * n + b;
#pragma unroll 8
            for (int k = warp; k < kc; k += FC_WARPS) {
                const float xv = __ldg(xp + 1LL * k * n);
                const float4 w0 = *reinterpret_cast<const float4*>(&ws[k][0]);
                const float4 w1 = *reinterpret_cast<const float4*>(&ws[k][4]);
                acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
                acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
                acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
                acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < FC_OUT; ++j) part[warp][lane][j] = acc[j];
    __syncthreads();
    // 32 images x 8 classes = 256 outputs, one per thread; fixed summation order over the warps
    const int ol = threadIdx.x & 7, bl = threadIdx.x >> 3;
    const int o = o0 + ol, bo = blockIdx.y * 32 + bl;
    if (o < classes && bo < n) {
        float sum = 0.f;
#pragma unroll
        for (int wi = 0; wi < FC_WARPS; ++wi) sum += part[wi][bl][ol];
        out[1LL * bo * classes + o] = sum + (bias ? __ldg(bias + o) : 0.f);
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

cudaError_t launch_avgpool_nhwc(const void* x, float* pooledT, int B, int HW, int C, int esz,
                                cudaStream_t s) {
    int ksq = 0;
    for (int k = 1; k * k <= HW; ++k)
        if (k * k == HW) ksq = k;
    const int64_t total = 1LL * B * (C * esz / 16);
    const int blocks = static_cast<int>((total + 127) / 128);
    if (esz == 2)
        avgpool_nhwc_kernel<__nv_bfloat16><<<blocks, 128, 0, s>>>(
            static_cast<const __nv_bfloat16*>(x), pooledT, B, HW, C, ksq);
    else
        avgpool_nhwc_kernel<float><<<blocks, 128, 0, s>>>(static_cast<const float*>(x), pooledT, B, HW,
                                                         C, ksq);
    return cudaGetLastError();
}

cudaError_t launch_fc(const float* pooledT, const float* w, const float* bias, float* logits, int B,
                      int C, int classes, cudaStream_t s) {
    dim3 grid((classes + FC_OUT - 1) / FC_OUT, (B + 31) / 32);
    fc_kernel<<<grid, FC_THREADS, 0, s>>>(pooledT, w, bias, logits, B, C, classes);
    return cudaGetLastError();
}

cudaError_t launch_transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t s) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_f32_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
    return cudaGetLastError();
}

}  // namespace rnb
