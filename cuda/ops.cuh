// ops.cuh — drop-in for olehskip/resnet.c cuda/ops.cuh: convOutputSize (:9-13) and the seven
// __global__ entry points (:15-32) with the reference's exact parameter lists, kept so that code
// which launches them directly (e.g. the reference's cuda/test.cu) still links.
//
// Launch contract (as in cuda/nn.cu): the conv / pool kernels take one block per output element with
// grid (spatial, B, C); linear takes grid (out_features, B); relu / add take one block per element;
// batch-norm takes a 3-D grid of 3-D blocks over (B, C, H*W). Extra threads in a block are ignored.
// These are compatibility entry points only — the fast path is the module forwards in nn.cuh and the
// ResNet engine, which run coalesced / tensor-core kernels from librnb.so instead.
#ifndef CUDA_OPS_CUH
#define CUDA_OPS_CUH

#include <cassert>
#include <cstdint>

#include "helpers.cuh"

// floor((2*padding + x - kernel_size) / stride) + 1 in unsigned arithmetic
__host__ __device__ inline uint64_t convOutputSize(uint64_t x, uint64_t kernel_size, uint64_t stride,
                                                   uint64_t padding)
{
    return (2 * padding + x - kernel_size) / stride + 1;
}

__global__ void conv2dForwardKernel(float* inp, float* out, float* weight, uint64_t kernel_size,
                                    uint64_t stride, uint64_t padding, uint64_t h_out, uint64_t w_out,
                                    uint64_t B, uint64_t in_channels, uint64_t out_channels, uint64_t H,
                                    uint64_t W);
__global__ void maxPool2dKernel(float* inp, float* out, uint64_t kernel_size, uint64_t stride,
                                uint64_t padding, uint64_t h_out, uint64_t w_out, uint64_t B,
                                uint64_t channels, uint64_t H, uint64_t W);
__global__ void avgPool2dKernel(float* inp, float* out, uint64_t kernel_size, uint64_t stride,
                                uint64_t padding, uint64_t h_out, uint64_t w_out, uint64_t B,
                                uint64_t channels, uint64_t H, uint64_t W);
__global__ void linearForwardKernel(float* inp, float* out, float* weight, float* bias, uint64_t B,
                                    uint64_t in_features, uint64_t out_features);
__global__ void reluForwardKernel(float* inp, float* out, uint64_t N);
__global__ void batchNorm2dForwardKernel(float* inp, float* out, float* weight, float* bias, float* mean,
                                         float* var, uint64_t B, uint64_t C, uint64_t N);
__global__ void addForwardKernel(float* inp1, float* inp2, float* out, uint64_t N);

#endif  // CUDA_OPS_CUH
