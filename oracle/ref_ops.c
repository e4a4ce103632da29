/* ref_ops.c — ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C CPU restatement of the seven device kernels of olehskip/resnet.c (cuda/ops.cu) plus the
 * host arg-max of cuda/inference/main.cu. One C function per reference kernel; each keeps the
 * reference's arithmetic exactly: FP32 storage, the same accumulation order, FMA contraction where
 * nvcc contracts (`sum += a * b` -> fmaf), the FP64 detour in batch-norm, division of the pooled sum
 * by the integer kernel size twice. Only the iteration over *output* elements differs (OpenMP loop
 * instead of one CUDA block per output), which cannot change any result.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -mfma -shared -fPIC -o libref_ops.so ref_ops.c -lm
 * (contraction is spelled out with fmaf/fma where the device compiler contracts; see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* convOutputSize, cuda/ops.cuh:9-13 */
uint64_t ref_conv_output_size(uint64_t x, uint64_t kernel_size, uint64_t stride, uint64_t padding) {
    return (2 * padding + x - kernel_size) / stride + 1;
}

/* conv2dForwardKernel, cuda/ops.cu:14-48.
 * inp [B][Cin][H][W], weight [Cout][Cin][k][k], out [B][Cout][OH][OW]; taps that fall outside the
 * image are skipped (:35-37); accumulation order ic -> kh -> kw (:30-32) in a float initialised to 0
 * (:29); the device compiler contracts `sum += inp*weight` (:42) into one FMA. */
void ref_conv2d_forward(const float* inp, float* out, const float* weight, uint64_t k, uint64_t stride,
                        uint64_t padding, uint64_t out_h, uint64_t out_w, uint64_t B, uint64_t Cin,
                        uint64_t Cout, uint64_t H, uint64_t W) {
    const int64_t jobs = (int64_t)(B * Cout);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t job = 0; job < jobs; ++job) {
        const uint64_t b = (uint64_t)job / Cout, oc = (uint64_t)job % Cout;
        for (uint64_t oh = 0; oh < out_h; ++oh) {
            for (uint64_t ow = 0; ow < out_w; ++ow) {
                const int64_t ih0 = (int64_t)(oh * stride) - (int64_t)padding;
                const int64_t iw0 = (int64_t)(ow * stride) - (int64_t)padding;
                float sum = 0.0f;
                for (uint64_t ic = 0; ic < Cin; ++ic) {
                    const float* plane = inp + (b * Cin + ic) * H * W;
                    const float* wk = weight + (oc * Cin + ic) * k * k;
                    for (uint64_t kh = 0; kh < k; ++kh) {
                        const int64_t ih = ih0 + (int64_t)kh;
                        if (ih < 0 || ih >= (int64_t)H) continue;
                        for (uint64_t kw = 0; kw < k; ++kw) {
                            const int64_t iw = iw0 + (int64_t)kw;
                            if (iw < 0 || iw >= (int64_t)W) continue;
                            sum = fmaf(plane[ih * (int64_t)W + iw], wk[kh * k + kw], sum);
                        }
                    }
                }
                out[((b * Cout + oc) * out_h + oh) * out_w + ow] = sum;
            }
        }
    }
}

/* maxPool2dKernel, cuda/ops.cu:50-78: running max starts at -inf (:64), out-of-image taps skipped. */
void ref_maxpool2d_forward(const float* inp, float* out, uint64_t k, uint64_t stride, uint64_t padding,
                           uint64_t out_h, uint64_t out_w, uint64_t B, uint64_t C, uint64_t H,
                           uint64_t W) {
    const int64_t planes = (int64_t)(B * C);
#pragma omp parallel for
    for (int64_t pl = 0; pl < planes; ++pl) {
        const float* plane = inp + (uint64_t)pl * H * W;
        float* oplane = out + (uint64_t)pl * out_h * out_w;
        for (uint64_t oh = 0; oh < out_h; ++oh)
            for (uint64_t ow = 0; ow < out_w; ++ow) {
                float mx = -INFINITY;
                for (uint64_t kh = 0; kh < k; ++kh)
                    for (uint64_t kw = 0; kw < k; ++kw) {
                        const int64_t ih = (int64_t)(oh * stride + kh) - (int64_t)padding;
                        const int64_t iw = (int64_t)(ow * stride + kw) - (int64_t)padding;
                        if (ih < 0 || ih >= (int64_t)H || iw < 0 || iw >= (int64_t)W) continue;
                        mx = fmaxf(mx, plane[ih * (int64_t)W + iw]);
                    }
                oplane[oh * out_w + ow] = mx;
            }
    }
}

/* avgPool2dKernel, cuda/ops.cu:80-108: FP32 running sum over in-image taps, then
 * `sum / kernel_size / kernel_size` (:107) — two divisions by the integer converted to float,
 * i.e. the full k*k divisor even when taps were skipped. */
void ref_avgpool2d_forward(const float* inp, float* out, uint64_t k, uint64_t stride, uint64_t padding,
                           uint64_t out_h, uint64_t out_w, uint64_t B, uint64_t C, uint64_t H,
                           uint64_t W) {
    const int64_t planes = (int64_t)(B * C);
#pragma omp parallel for
    for (int64_t pl = 0; pl < planes; ++pl) {
        const float* plane = inp + (uint64_t)pl * H * W;
        float* oplane = out + (uint64_t)pl * out_h * out_w;
        for (uint64_t oh = 0; oh < out_h; ++oh)
            for (uint64_t ow = 0; ow < out_w; ++ow) {
                float sum = 0.0f;
                for (uint64_t kh = 0; kh < k; ++kh)
                    for (uint64_t kw = 0; kw < k; ++kw) {
                        const int64_t ih = (int64_t)(oh * stride + kh) - (int64_t)padding;
                        const int64_t iw = (int64_t)(ow * stride + kw) - (int64_t)padding;
                        if (ih < 0 || ih >= (int64_t)H || iw < 0 || iw >= (int64_t)W) continue;
                        sum += plane[ih * (int64_t)W + iw];
                    }
                oplane[oh * out_w + ow] = sum / (float)k / (float)k;
            }
    }
}

/* linearForwardKernel, cuda/ops.cu:110-128: out[b][o] = sum_i inp[b][i]*weight[o][i] sequentially in
 * FP32 (FMA-contracted), bias added afterwards when the pointer is non-null (:124-126). */
void ref_linear_forward(const float* inp, float* out, const float* weight, const float* bias,
                        uint64_t B, uint64_t in_features, uint64_t out_features) {
    const int64_t jobs = (int64_t)(B * out_features);
#pragma omp parallel for
    for (int64_t job = 0; job < jobs; ++job) {
        const uint64_t b = (uint64_t)job / out_features, o = (uint64_t)job % out_features;
        const float* x = inp + b * in_features;
        const float* w = weight + o * in_features;
        float curr = 0.0f;
        for (uint64_t i = 0; i < in_features; ++i) curr = fmaf(x[i], w[i], curr);
        if (bias) curr += bias[o];
        out[b * out_features + o] = curr;
    }
}

/* reluForwardKernel, cuda/ops.cu:130-137 */
void ref_relu_forward(const float* inp, float* out, uint64_t N) {
#pragma omp parallel for
    for (int64_t n = 0; n < (int64_t)N; ++n) out[n] = fmaxf(inp[n], 0.0f);
}

/* batchNorm2dForwardKernel, cuda/ops.cu:139-151. The expression
 *   (inp - mean[c]) / sqrt(var[c] + 1e-5) * weight[c] + bias[c]
 * subtracts in float, then the double literal 1e-5 promotes sqrt, the division, the multiplication
 * and the addition to double; the store narrows to float (:149-150). In place is allowed. */
void ref_batchnorm2d_forward(const float* inp, float* out, const float* weight, const float* bias,
                             const float* mean, const float* var, uint64_t B, uint64_t C, uint64_t N) {
    const int64_t planes = (int64_t)(B * C);
#pragma omp parallel for
    for (int64_t pl = 0; pl < planes; ++pl) {
        const uint64_t c = (uint64_t)pl % C;
        const double denom = sqrt((double)var[c] + 1e-5);
        const double g = (double)weight[c], sh = (double)bias[c];
        const float mu = mean[c];
        const float* x = inp + (uint64_t)pl * N;
        float* y = out + (uint64_t)pl * N;
        for (uint64_t n = 0; n < N; ++n) {
            const float centered = x[n] - mu;
            /* nvcc (default -fmad=true) contracts the double multiply-add as well */
            y[n] = (float)fma((double)centered / denom, g, sh);
        }
    }
}

/* addForwardKernel, cuda/ops.cu:153-160 */
void ref_add_forward(const float* a, const float* b, float* out, uint64_t N) {
#pragma omp parallel for
    for (int64_t n = 0; n < (int64_t)N; ++n) out[n] = a[n] + b[n];
}

/* Host arg-max of cuda/inference/main.cu:243-251: strict '<' so the first maximum wins. */
void ref_argmax_rows(const float* x, int32_t* out, uint64_t B, uint64_t classes) {
    for (uint64_t b = 0; b < B; ++b) {
        uint64_t mx = 0;
        for (uint64_t i = 1; i < classes; ++i)
            if (x[b * classes + mx] < x[b * classes + i]) mx = i;
        out[b] = (int32_t)mx;
    }
}
