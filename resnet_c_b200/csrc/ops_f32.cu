// ops_f32.cu — FP32 NCHW per-op kernels behind the reference's module API (cuda/nn.cuh).
//
// These keep the *arithmetic* of the reference kernels (same accumulation order, same mixed
// float/double expressions, FMA contraction) so that a caller of Conv2d::forward etc. gets the
// reference's numbers, but none of its launch geometry: the reference runs one thread per block
// (nn.cu:9-10, 70-71, 82-83); here every kernel is a coalesced grid-stride kernel with the fastest
// tensor dimension on threadIdx.x.
#include <cstdint>
#include <cuda_runtime.h>

#include "internal.h"

namespace rnb {

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t n, int num_sms) {
    int64_t blocks = (n + kThreads - 1) / kThreads;
    const int64_t cap = static_cast<int64_t>(num_sms) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

// conv2dForwardKernel (ops.cu:14-48): out[b][oc][oh][ow] = sum_{ic,kh,kw} in * w, taps outside the
// image skipped, accumulation order ic -> kh -> kw in FP32 with FMA.
__global__ void conv2d_f32_kernel(const float* __restrict__ x, float* __restrict__ out,
                                  const float* __restrict__ w, int B, int Cin, int H, int W, int Cout,
                                  int k, int stride, int pad, int OH, int OW) {
    const int64_t total = 1LL * B * Cout * OH * OW;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int ow = static_cast<int>(i % OW);
        int64_t t = i / OW;
        const int oh = static_cast<int>(t % OH);
        t /= OH;
        const int oc = static_cast<int>(t % Cout);
        const int b = static_cast<int>(t / Cout);
        const int ih0 = oh * stride - pad;
        const int iw0 = ow * stride - pad;
        const float* wp = w + 1LL * oc * Cin * k * k;
        const float* xp = x + 1LL * b * Cin * H * W;
        float sum = 0.f;
        for (int ic = 0; ic < Cin; ++ic) {
            for (int kh = 0; kh < k; ++kh) {
                const int ih = ih0 + kh;
                if (ih < 0 || ih >= H) continue;
                for (int kw = 0; kw < k; ++kw) {
                    const int iw = iw0 + kw;
                    if (iw < 0 || iw >= W) continue;
                    sum = fmaf(__ldg(xp + (1LL * ic * H + ih) * W + iw),
                               __ldg(wp + (ic * k + kh) * k + kw), sum);
                }
            }
        }
        out[i] = sum;
    }
}

// batchNorm2dForwardKernel (ops.cu:139-151): float subtraction, then double for the rest because
// of the 1e-5 literal, narrowed on store.
__global__ void batchnorm2d_f32_kernel(const float* __restrict__ x, float* __restrict__ out,
                                       const float* __restrict__ weight,
                                       const float* __restrict__ bias,
                                       const float* __restrict__ mean,
                                       const float* __restrict__ var, int64_t total, int C, int HW) {
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int c = static_cast<int>((i / HW) % C);
        const float centered = x[i] - mean[c];
        const double inv = sqrt(static_cast<double>(var[c]) + 1e-5);
        out[i] = static_cast<float>(static_cast<double>(centered) / inv *
                                        static_cast<double>(weight[c]) +
                                    static_cast<double>(bias[c]));
    }
}

__global__ void relu_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n;
         i += 1LL * gridDim.x * blockDim.x)
        out[i] = fmaxf(x[i], 0.f);
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n;
         i += 1LL * gridDim.x * blockDim.x)
        out[i] = a[i] + b[i];
}

// maxPool2dKernel / avgPool2dKernel (ops.cu:50-108).
template <bool kMax>
__global__ void pool2d_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C,
                                  int H, int W, int k, int stride, int pad, int OH, int OW) {
    const int64_t total = 1LL * B * C * OH * OW;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int ow = static_cast<int>(i % OW);
        int64_t t = i / OW;
        const int oh = static_cast<int>(t % OH);
        const int64_t bc = t / OH;
        const float* xp = x + bc * H * W;
        const int ih0 = oh * stride - pad;
        const int iw0 = ow * stride - pad;
        float acc = kMax ? -INFINITY : 0.f;
        for (int kh = 0; kh < k; ++kh) {
            const int ih = ih0 + kh;
            if (ih < 0 || ih >= H) continue;
            for (int kw = 0; kw < k; ++kw) {
                const int iw = iw0 + kw;
                if (iw < 0 || iw >= W) continue;
                const float v = __ldg(xp + 1LL * ih * W + iw);
                acc = kMax ? fmaxf(acc, v) : acc + v;
            }
        }
        // the reference divides the float sum by the (integer) kernel size twice
        out[i] = kMax ? acc : acc / static_cast<float>(k) / static_cast<float>(k);
    }
}

// linearForwardKernel (ops.cu:110-128): sequential FP32 FMA over in_features, bias added last.
// One warp per output keeps the reference's accumulation order only if a single lane does the sum,
// so each thread owns one (b, o) pair; consecutive threads read consecutive rows of W, which the
// L1/L2 absorb (W is 8 MB at most).
__global__ void linear_f32_kernel(const float* __restrict__ x, float* __restrict__ out,
                                  const float* __restrict__ w, const float* __restrict__ bias, int B,
                                  int in_f, int out_f) {
    const int64_t total = 1LL * B * out_f;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int o = static_cast<int>(i % out_f);
        const int b = static_cast<int>(i / out_f);
        const float* xp = x + 1LL * b * in_f;
        const float* wp = w + 1LL * o * in_f;
        float acc = 0.f;
        for (int j = 0; j < in_f; ++j) acc = fmaf(__ldg(xp + j), __ldg(wp + j), acc);
        if (bias) acc += bias[o];
        out[i] = acc;
    }
}

// Row arg-max, lowest index wins ties (main.cu:243-251 uses strict '<'). One warp per row.
__global__ void argmax_f32_kernel(const float* __restrict__ x, int32_t* __restrict__ out, int B,
                                  int n) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched with PDL: x is the previous kernel's output
    if (warp >= B) return;
    const float* row = x + 1LL * warp * n;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    int j = lane;
    for (; j + 96 < n; j += 128) {  // four independent loads in flight; compared in index order
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = row[j + 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u] > best || best_i == 0x7fffffff) {
                best = v[u];
                best_i = j + 32 * u;
            }
    }
    for (; j < n; j += 32) {
        const float v = row[j];
        if (v > best || (v == best && j < best_i) || best_i == 0x7fffffff) {
            best = v;
            best_i = j;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        if (ov > best || (ov == best && oi < best_i)) {
            best = ov;
            best_i = oi;
        }
    }
    if (lane == 0) out[warp] = best_i == 0x7fffffff ? 0 : best_i;
}

// Row softmax + top-k (the step after the hot path: what a caller does with fc_out, main.cu:240-251, beyond the
// single arg-max). One warp per row: max, sum of exp(x - max) (expf, not the fast intrinsic), then k rounds of
// arg-max over the not-yet-taken entries (lowest index wins ties, like argmax_f32_kernel). probs_full (optional)
// receives the whole softmax row.
constexpr int kMaxTopK = 32;
__global__ void softmax_topk_kernel(const float* __restrict__ x, float* __restrict__ probs_full,
                                    float* __restrict__ top_p, int32_t* __restrict__ top_i, int B, int n, int k) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* row = x + 1LL * warp * n;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, row[j]);
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) sum += expf(row[j] - mx);
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (probs_full)
        for (int j = lane; j < n; j += 32) probs_full[1LL * warp * n + j] = expf(row[j] - mx) / sum;
    float last_v = INFINITY;
    int last_i = -1;
    for (int t = 0; t < k; ++t) {
        // next entry in (value desc, index asc) order strictly after (last_v, last_i)
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            const float v = row[j];
            const bool after = v < last_v || (v == last_v && j > last_i);
            if (after && (v > best || (v == best && j < best_i) || best_i == 0x7fffffff)) {
                best = v;
                best_i = j;
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
            if (oi != 0x7fffffff && (best_i == 0x7fffffff || ov > best || (ov == best && oi < best_i))) {
                best = ov;
                best_i = oi;
            }
        }
        if (lane == 0) {
            top_i[1LL * warp * k + t] = best_i == 0x7fffffff ? -1 : best_i;
            top_p[1LL * warp * k + t] = best_i == 0x7fffffff ? 0.f : expf(best - mx) / sum;
        }
        last_v = best;
        last_i = best_i;
    }
}

}  // namespace

cudaError_t launch_softmax_topk_f32(const float* x, float* probs_full, float* top_p, int32_t* top_i, int B, int n,
                                    int k, cudaStream_t s) {
    const int warps_per_block = kThreads / 32;
    softmax_topk_kernel<<<(B + warps_per_block - 1) / warps_per_block, kThreads, 0, s>>>(x, probs_full, top_p, top_i,
                                                                                       B, n, k);
    return cudaGetLastError();
}

cudaError_t launch_conv2d_f32(const float* x, float* out, const float* w, int B, int Cin, int H,
                              int W, int Cout, int k, int stride, int pad, cudaStream_t s) {
    const int OH = (2 * pad + H - k) / stride + 1;
    const int OW = (2 * pad + W - k) / stride + 1;
    const int64_t total = 1LL * B * Cout * OH * OW;
    conv2d_f32_kernel<<<grid_for(total, num_sms() * 4), kThreads, 0, s>>>(x, out, w, B, Cin, H, W,
                                                                         Cout, k, stride, pad, OH, OW);
    return cudaGetLastError();
}
cudaError_t launch_batchnorm2d_f32(const float* x, float* out, const float* weight,
                                   const float* bias, const float* mean, const float* var, int B,
                                   int C, int HW, cudaStream_t s) {
    const int64_t total = 1LL * B * C * HW;
    batchnorm2d_f32_kernel<<<grid_for(total, num_sms()), kThreads, 0, s>>>(x, out, weight, bias,
                                                                          mean, var, total, C, HW);
    return cudaGetLastError();
}
cudaError_t launch_relu_f32(const float* x, float* out, int64_t n, cudaStream_t s) {
    relu_f32_kernel<<<grid_for(n, num_sms()), kThreads, 0, s>>>(x, out, n);
    return cudaGetLastError();
}
cudaError_t launch_add_f32(const float* a, const float* b, float* out, int64_t n, cudaStream_t s) {
    add_f32_kernel<<<grid_for(n, num_sms()), kThreads, 0, s>>>(a, b, out, n);
    return cudaGetLastError();
}
cudaError_t launch_pool2d_f32(bool is_max, const float* x, float* out, int B, int C, int H, int W,
                              int k, int stride, int pad, cudaStream_t s) {
    const int OH = (2 * pad + H - k) / stride + 1;
    const int OW = (2 * pad + W - k) / stride + 1;
    const int64_t total = 1LL * B * C * OH * OW;
    if (is_max)
        pool2d_f32_kernel<true><<<grid_for(total, num_sms()), kThreads, 0, s>>>(x, out, B, C, H, W, k,
                                                                               stride, pad, OH, OW);
    else
        pool2d_f32_kernel<false><<<grid_for(total, num_sms()), kThreads, 0, s>>>(
            x, out, B, C, H, W, k, stride, pad, OH, OW);
    return cudaGetLastError();
}
cudaError_t launch_linear_f32(const float* x, float* out, const float* w, const float* bias, int B,
                              int in_f, int out_f, cudaStream_t s) {
    const int64_t total = 1LL * B * out_f;
    linear_f32_kernel<<<grid_for(total, num_sms()), kThreads, 0, s>>>(x, out, w, bias, B, in_f,
                                                                     out_f);
    return cudaGetLastError();
}
cudaError_t launch_argmax_f32(const float* x, int32_t* out, int B, int n, cudaStream_t s) {
    const int warps_per_block = 2;  // 128 blocks for 256 rows: spread over the SMs (the kernel is latency-bound)
    return launch_pdl_small(argmax_f32_kernel, dim3((B + warps_per_block - 1) / warps_per_block),
                            dim3(32 * warps_per_block), 0, s, x, out, B, n);
}

}  // namespace rnb
