// layout.cu — layout changes at the boundary of the tensor-core path and BN folding.
//
// The reference keeps everything FP32 NCHW (tensor.cuh, ops.cu idx4d :3-7). The tcgen05 path wants
// NHWC so that the GEMM K dimension (input channels) is contiguous for TMA. These kernels convert at
// the edges only: once per weight at load time, once per activation when a caller hands us NCHW.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"

namespace rnb {

namespace {

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <typename T>
__device__ __forceinline__ T to_act(float v);
template <>
__device__ __forceinline__ __nv_bfloat16 to_act<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ float to_act<float>(float v) {
    return rna_tf32(v);
}
__device__ __forceinline__ float from_act(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float from_act(float v) { return v; }

// [B][C][HW] fp32 -> [B][HW][C] T : 32x32 tiles through shared memory, both sides coalesced.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ out, int C, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    const int p0 = blockIdx.x * 32;
    const float* xb = x + 1LL * b * C * HW;
    T* ob = out + 1LL * b * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? xb[1LL * c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) ob[1LL * p * C + c] = to_act<T>(tile[threadIdx.x][i]);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ out, int C, int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    const int p0 = blockIdx.x * 32;
    const T* xb = x + 1LL * b * C * HW;
    float* ob = out + 1LL * b * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? from_act(xb[1LL * p * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) ob[1LL * c * HW + p] = tile[threadIdx.x][i];
    }
}

// Eval-mode BN as the reference evaluates it (ops.cu:149-150): y = (x-mean)/sqrt(var+1e-5)*w + b
// with the division/multiplication in double. Folded: scale = w/sqrt(var+1e-5), shift = b-mean*scale.
__device__ __forceinline__ void bn_fold(const float* bn_w, const float* bn_b, const float* bn_m,
                                        const float* bn_v, int oc, double& scale, double& shift) {
    if (bn_w) {
        scale = static_cast<double>(bn_w[oc]) / sqrt(static_cast<double>(bn_v[oc]) + 1e-5);
        shift = static_cast<double>(bn_b[oc]) - static_cast<double>(bn_m[oc]) * scale;
    } else {
        scale = 1.0;
        shift = 0.0;
    }
}

// OIHW fp32 -> [Cout][kh][kw][Cin] T with the BN scale folded in; one thread per packed element.
template <typename T>
__global__ void fold_pack_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                 const float* __restrict__ bn_b, const float* __restrict__ bn_m,
                                 const float* __restrict__ bn_v, T* __restrict__ packed,
                                 float* __restrict__ bias, int Cout, int Cin, int k) {
    const int64_t total = 1LL * Cout * k * k * Cin;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int ic = static_cast<int>(i % Cin);
        int64_t t = i / Cin;
        const int kw = static_cast<int>(t % k);
        t /= k;
        const int kh = static_cast<int>(t % k);
        const int oc = static_cast<int>(t / k);
        double scale, shift;
        bn_fold(bn_w, bn_b, bn_m, bn_v, oc, scale, shift);
        const float v = w[((1LL * oc * Cin + ic) * k + kh) * k + kw];
        packed[i] = to_act<T>(static_cast<float>(static_cast<double>(v) * scale));
        if (ic == 0 && kh == 0 && kw == 0) bias[oc] = static_cast<float>(shift);
    }
}

__global__ void fold_f32_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                const float* __restrict__ bn_b, const float* __restrict__ bn_m,
                                const float* __restrict__ bn_v, float* __restrict__ w_out,
                                float* __restrict__ bias, int Cout, int per_out) {
    const int64_t total = 1LL * Cout * per_out;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int oc = static_cast<int>(i / per_out);
        double scale, shift;
        bn_fold(bn_w, bn_b, bn_m, bn_v, oc, scale, shift);
        w_out[i] = static_cast<float>(static_cast<double>(w[i]) * scale);
        if (i % per_out == 0) bias[oc] = static_cast<float>(shift);
    }
}

}  // namespace

cudaError_t launch_nchw_to_nhwc(const float* x, void* out, int B, int C, int HW, int esz,
                                cudaStream_t s) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
    if (esz == 2)
        nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(
            x, static_cast<__nv_bfloat16*>(out), C, HW);
    else
        nchw_to_nhwc_kernel<float><<<grid, block, 0, s>>>(x, static_cast<float*>(out), C, HW);
    return cudaGetLastError();
}

cudaError_t launch_nhwc_to_nchw(const void* x, float* out, int B, int C, int HW, int esz,
                                cudaStream_t s) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
    if (esz == 2)
        nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(
            static_cast<const __nv_bfloat16*>(x), out, C, HW);
    else
        nhwc_to_nchw_kernel<float><<<grid, block, 0, s>>>(static_cast<const float*>(x), out, C, HW);
    return cudaGetLastError();
}

cudaError_t launch_fold_pack(const float* w, const float* bn_w, const float* bn_b, const float* bn_m,
                             const float* bn_v, void* packed, float* bias, int Cout, int Cin, int k,
                             int esz, cudaStream_t s) {
    const int64_t total = 1LL * Cout * k * k * Cin;
    const int blocks = static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    if (esz == 2)
        fold_pack_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(
            w, bn_w, bn_b, bn_m, bn_v, static_cast<__nv_bfloat16*>(packed), bias, Cout, Cin, k);
    else
        fold_pack_kernel<float><<<blocks, 256, 0, s>>>(w, bn_w, bn_b, bn_m, bn_v,
                                                      static_cast<float*>(packed), bias, Cout, Cin, k);
    return cudaGetLastError();
}

cudaError_t launch_fold_f32(const float* w, const float* bn_w, const float* bn_b, const float* bn_m,
                            const float* bn_v, float* w_out, float* bias, int Cout, int per_out,
                            cudaStream_t s) {
    const int64_t total = 1LL * Cout * per_out;
    const int blocks = static_cast<int>((total + 255) / 256);
    fold_f32_kernel<<<blocks, 256, 0, s>>>(w, bn_w, bn_b, bn_m, bn_v, w_out, bias, Cout, per_out);
    return cudaGetLastError();
}

}  // namespace rnb
