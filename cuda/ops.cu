// ops.cu — compatibility definitions of the seven __global__ entry points declared in ops.cuh.
//
// They honour the reference's launch contract (one block per output element, see ops.cuh) and its
// arithmetic (FP32, FMA-contracted sequential sums, the double detour in batch-norm), but are written
// around two small device helpers instead of seven copies of the same loop nest. Nothing on the fast
// path launches them: Conv2d::forward & co. (nn.cu) go through librnb.so.
#include <cmath>

#include "ops.cuh"

namespace
{

// Block -> (batch, channel, output row, output col) for the grid (spatial, B, C) contract.
struct OutputCoord
{
    uint64_t b, c, oh, ow;
    bool valid;
};

__device__ inline OutputCoord blockToOutput(uint64_t h_out, uint64_t w_out, uint64_t B, uint64_t C)
{
    OutputCoord o;
    o.oh = blockIdx.x / w_out;
    o.ow = blockIdx.x % w_out;
    o.b = blockIdx.y;
    o.c = blockIdx.z;
    o.valid = o.b < B && o.c < C && o.oh < h_out && o.ow < w_out;
    return o;
}

__device__ inline bool leaderThread()
{
    return threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0;
}

// Visits the in-image taps of a k x k window anchored at (oh*stride - padding, ow*stride - padding)
// in kh-major order and hands (kh, kw, flat offset inside the H x W plane) to `visit`.
template <class Visit>
__device__ inline void forEachTap(uint64_t oh, uint64_t ow, uint64_t k, uint64_t stride, uint64_t padding,
                                  uint64_t H, uint64_t W, Visit visit)
{
    const int64_t top = static_cast<int64_t>(oh * stride) - static_cast<int64_t>(padding);
    const int64_t left = static_cast<int64_t>(ow * stride) - static_cast<int64_t>(padding);
    for (uint64_t kh = 0; kh < k; ++kh) {
        const int64_t ih = top + static_cast<int64_t>(kh);
        if (ih < 0 || ih >= static_cast<int64_t>(H)) {
            continue;
        }
        for (uint64_t kw = 0; kw < k; ++kw) {
            const int64_t iw = left + static_cast<int64_t>(kw);
            if (iw < 0 || iw >= static_cast<int64_t>(W)) {
                continue;
            }
            visit(kh, kw, static_cast<uint64_t>(ih) * W + static_cast<uint64_t>(iw));
        }
    }
}

}  // namespace

__global__ void conv2dForwardKernel(float* inp, float* out, float* weight, uint64_t kernel_size,
                                    uint64_t stride, uint64_t padding, uint64_t h_out, uint64_t w_out,
                                    uint64_t B, uint64_t in_channels, uint64_t out_channels, uint64_t H,
                                    uint64_t W)
{
    const OutputCoord o = blockToOutput(h_out, w_out, B, out_channels);
    if (!o.valid || !leaderThread()) {
        return;
    }
    float acc = 0.f;
    for (uint64_t ic = 0; ic < in_channels; ++ic) {
        const float* plane = inp + (o.b * in_channels + ic) * H * W;
        const float* filt = weight + (o.c * in_channels + ic) * kernel_size * kernel_size;
        forEachTap(o.oh, o.ow, kernel_size, stride, padding, H, W,
                   [&](uint64_t kh, uint64_t kw, uint64_t off) {
                       acc = fmaf(plane[off], filt[kh * kernel_size + kw], acc);
                   });
    }
    out[((o.b * out_channels + o.c) * h_out + o.oh) * w_out + o.ow] = acc;
}

__global__ void maxPool2dKernel(float* inp, float* out, uint64_t kernel_size, uint64_t stride,
                                uint64_t padding, uint64_t h_out, uint64_t w_out, uint64_t B,
                                uint64_t channels, uint64_t H, uint64_t W)
{
    const OutputCoord o = blockToOutput(h_out, w_out, B, channels);
    if (!o.valid || !leaderThread()) {
        return;
    }
    const float* plane = inp + (o.b * channels + o.c) * H * W;
    float best = -INFINITY;
    forEachTap(o.oh, o.ow, kernel_size, stride, padding, H, W,
               [&](uint64_t, uint64_t, uint64_t off) { best = fmaxf(best, plane[off]); });
    out[((o.b * channels + o.c) * h_out + o.oh) * w_out + o.ow] = best;
}

__global__ void avgPool2dKernel(float* inp, float* out, uint64_t kernel_size, uint64_t stride,
                                uint64_t padding, uint64_t h_out, uint64_t w_out, uint64_t B,
                                uint64_t channels, uint64_t H, uint64_t W)
{
    const OutputCoord o = blockToOutput(h_out, w_out, B, channels);
    if (!o.valid || !leaderThread()) {
        return;
    }
    const float* plane = inp + (o.b * channels + o.c) * H * W;
    float total = 0.f;
    forEachTap(o.oh, o.ow, kernel_size, stride, padding, H, W,
               [&](uint64_t, uint64_t, uint64_t off) { total += plane[off]; });
    // divisor is the full window even when taps were clipped, applied as two divisions
    const float k = static_cast<float>(kernel_size);
    out[((o.b * channels + o.c) * h_out + o.oh) * w_out + o.ow] = total / k / k;
}

__global__ void linearForwardKernel(float* inp, float* out, float* weight, float* bias, uint64_t B,
                                    uint64_t in_features, uint64_t out_features)
{
    const uint64_t o = blockIdx.x, b = blockIdx.y;
    if (b >= B || o >= out_features || !leaderThread()) {
        return;
    }
    const float* x = inp + b * in_features;
    const float* w = weight + o * in_features;
    float acc = 0.f;
    for (uint64_t i = 0; i < in_features; ++i) {
        acc = fmaf(x[i], w[i], acc);
    }
    out[b * out_features + o] = bias ? acc + bias[o] : acc;
}

__global__ void reluForwardKernel(float* inp, float* out, uint64_t N)
{
    const uint64_t n = blockIdx.x;
    if (n < N && leaderThread()) {
        out[n] = fmaxf(inp[n], 0.f);
    }
}

__global__ void batchNorm2dForwardKernel(float* inp, float* out, float* weight, float* bias, float* mean,
                                         float* var, uint64_t B, uint64_t C, uint64_t N)
{
    const uint64_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t c = blockIdx.y * blockDim.y + threadIdx.y;
    const uint64_t n = blockIdx.z * blockDim.z + threadIdx.z;
    if (b >= B || c >= C || n >= N) {
        return;
    }
    const uint64_t i = (b * C + c) * N + n;
    const float centered = inp[i] - mean[c];
    const double scaled = static_cast<double>(centered) / sqrt(static_cast<double>(var[c]) + 1e-5);
    out[i] = static_cast<float>(scaled * static_cast<double>(weight[c]) + static_cast<double>(bias[c]));
}

__global__ void addForwardKernel(float* inp1, float* inp2, float* out, uint64_t N)
{
    const uint64_t n = blockIdx.x;
    if (n < N && leaderThread()) {
        out[n] = inp1[n] + inp2[n];
    }
}
