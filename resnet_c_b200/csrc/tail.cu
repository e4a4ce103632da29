// tail.cu — network tail: global average pool + fully-connected layer, replacing avgPool2dKernel
// (/root/reference/cuda/ops.cu:80-108) and linearForwardKernel (ops.cu:110-128) as called at
// cuda/inference/main.cu:213-224. Logits stay FP32. The arg-max (main.cu:243-251) is
// launch_argmax_f32 in ops_f32.cu.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"

namespace rnb {

namespace {

// [B][HW][C] T -> [B][C] fp32. One thread per (b, 16-byte channel group); consecutive threads read
// consecutive 16-byte groups of the same pixel (coalesced). Sum is sequential over pixels in FP32
// as in avgPool2dKernel, and divided by k twice when HW = k*k (ops.cu:107).
template <typename T>
__global__ void avgpool_nhwc_kernel(const T* __restrict__ x, float* __restrict__ pooled, int B, int HW,
                                    int C, int ksq) {
    constexpr int VEC = 16 / sizeof(T);
    const int groups = C / VEC;
    const int64_t total = 1LL * B * groups;
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
        float* op = pooled + 1LL * b * C + gidx * VEC;
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            op[e] = ksq > 0 ? acc[e] / static_cast<float>(ksq) / static_cast<float>(ksq)
                            : acc[e] / static_cast<float>(HW);
    }
}

// logits[b][o] = sum_i pooled[b][i] * w[o][i] + bias[o]; 32(b) x 64(o) tile per block, K chunks of 32
// through shared memory, 256 threads each owning a 2 x 4 micro-tile. FP32 FMA on CUDA cores: the
// layer is 2 MMAC per image, far below anything worth a tensor-core launch.
constexpr int FC_BM = 32, FC_BN = 64, FC_BK = 32;
__global__ void __launch_bounds__(256)
fc_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
          float* __restrict__ out, int B, int C, int classes) {
    __shared__ float as[FC_BK][FC_BM + 1];
    __shared__ float ws[FC_BK][FC_BN + 1];
    const int b0 = blockIdx.y * FC_BM;
    const int o0 = blockIdx.x * FC_BN;
    const int tx = threadIdx.x & 15;   // 16 column groups of 4 outputs
    const int ty = threadIdx.x >> 4;   // 16 row groups of 2 batch rows
    float acc[2][4] = {};
    for (int k0 = 0; k0 < C; k0 += FC_BK) {
        for (int i = threadIdx.x; i < FC_BM * FC_BK; i += 256) {
            const int r = i / FC_BK, k = i % FC_BK;
            const int b = b0 + r;
            as[k][r] = (b < B && k0 + k < C) ? __ldg(a + 1LL * b * C + k0 + k) : 0.f;
        }
        for (int i = threadIdx.x; i < FC_BN * FC_BK; i += 256) {
            const int r = i / FC_BK, k = i % FC_BK;
            const int o = o0 + r;
            ws[k][r] = (o < classes && k0 + k < C) ? __ldg(w + 1LL * o * C + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FC_BK; ++k) {
            const float a0 = as[k][ty * 2], a1 = as[k][ty * 2 + 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float wv = ws[k][tx * 4 + j];
                acc[0][j] = fmaf(a0, wv, acc[0][j]);
                acc[1][j] = fmaf(a1, wv, acc[1][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int b = b0 + ty * 2 + i;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + tx * 4 + j;
            if (o < classes) out[1LL * b * classes + o] = acc[i][j] + (bias ? __ldg(bias + o) : 0.f);
        }
    }
}

}  // namespace

cudaError_t launch_avgpool_nhwc(const void* x, float* pooled, int B, int HW, int C, int esz,
                                cudaStream_t s) {
    int ksq = 0;
    for (int k = 1; k * k <= HW; ++k)
        if (k * k == HW) ksq = k;
    const int64_t total = 1LL * B * (C * esz / 16);
    const int blocks = static_cast<int>((total + 127) / 128);
    if (esz == 2)
        avgpool_nhwc_kernel<__nv_bfloat16><<<blocks, 128, 0, s>>>(
            static_cast<const __nv_bfloat16*>(x), pooled, B, HW, C, ksq);
    else
        avgpool_nhwc_kernel<float><<<blocks, 128, 0, s>>>(static_cast<const float*>(x), pooled, B, HW,
                                                         C, ksq);
    return cudaGetLastError();
}

cudaError_t launch_fc(const float* pooled, const float* w, const float* bias, float* logits, int B,
                      int C, int classes, cudaStream_t s) {
    dim3 grid((classes + FC_BN - 1) / FC_BN, (B + FC_BM - 1) / FC_BM);
    fc_kernel<<<grid, 256, 0, s>>>(pooled, w, bias, logits, B, C, classes);
    return cudaGetLastError();
}

}  // namespace rnb
