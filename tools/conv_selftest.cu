// conv_selftest.cu — standalone GPU check of the implicit-GEMM conv kernel against a host loop
// (same loop nest as /root/reference/cuda/ops.cu:14-48, NHWC indexing, double accumulation).
// Build: see resnet_c_b200/build.py (target "selftest").  Run on a B200:  ./conv_selftest [bf16|tf32]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "../resnet_c_b200/csrc/conv_plan.h"

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                 \
        }                                                                            \
    } while (0)

struct Shape {
    int B, H, W, Cin, Cout, k, stride, pad, relu, res, bn;
};

static uint32_t lcg_state = 12345u;
static float frand() {
    lcg_state = lcg_state * 1664525u + 1013904223u;
    return ((lcg_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float to_bf16f(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float to_tf32f(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;  // round-to-nearest (ties away), matches cvt.rna
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main(int argc, char** argv) {
    const bool tf32 = argc > 1 && !strcmp(argv[1], "tf32");
    const int esz = tf32 ? 4 : 2;
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s, %d SMs, smem/block optin %zu\n", prop.name, prop.multiProcessorCount,
           prop.sharedMemPerBlockOptin);
    CK(rnb::conv_kernels_init());

    std::vector<Shape> shapes = {
        // B  H   W   Cin  Cout k  s  p relu res bn
        {1, 8, 8, 64, 64, 1, 1, 0, 0, 0, 0},       // single tile, single k-block
        {1, 8, 8, 64, 64, 1, 1, 0, 1, 1, 0},       // + residual + relu
        {2, 12, 12, 128, 128, 1, 1, 0, 1, 0, 0},   // 2 k-blocks, M=288 (tail tile)
        {2, 12, 12, 64, 64, 3, 1, 1, 1, 0, 0},     // 3x3 pad 1
        {3, 14, 14, 128, 256, 3, 1, 1, 1, 1, 0},   // 3x3 + residual, wraps images
        {2, 28, 28, 128, 128, 3, 2, 1, 1, 0, 0},   // 3x3 stride 2
        {2, 28, 28, 256, 512, 1, 2, 0, 0, 0, 0},   // 1x1 stride 2 downsample
        {2, 56, 56, 64, 256, 1, 1, 0, 1, 1, 0},    // layer1 conv3
        {2, 56, 56, 256, 64, 1, 1, 0, 1, 0, 0},    // layer1 conv1
        {2, 56, 56, 64, 64, 3, 1, 1, 1, 0, 0},     // layer1 conv2
        {4, 7, 7, 512, 512, 3, 1, 1, 1, 0, 0},     // layer4 conv2
        {4, 7, 7, 512, 2048, 1, 1, 0, 1, 1, 0},    // layer4 conv3
        {4, 14, 14, 1024, 2048, 1, 2, 0, 0, 0, 0}, // layer4 downsample
        {2, 56, 56, 64, 256, 1, 1, 0, 1, 1, 64},   // forced BN=64
        {32, 56, 56, 64, 256, 1, 1, 0, 1, 1, 0},   // many tiles per CTA (persistence, ring wrap)
        // CTA-pair kernel (force_bn 1128 / 1256)
        {1, 16, 16, 64, 256, 1, 1, 0, 0, 0, 1256},     // one pair tile, one k-block
        {1, 16, 16, 64, 256, 1, 1, 0, 1, 1, 1256},     // + residual + relu
        {1, 12, 12, 128, 128, 3, 1, 1, 1, 0, 1128},    // M=144: peer half mostly out of range
        {3, 14, 14, 128, 256, 3, 1, 1, 1, 1, 1256},    // wraps images, residual
        {2, 28, 28, 256, 512, 1, 2, 0, 0, 0, 1256},    // strided 1x1, 2 n-tiles
        {2, 28, 28, 128, 128, 3, 2, 1, 1, 0, 1128},    // 3x3 stride 2
        {4, 7, 7, 512, 2048, 1, 1, 0, 1, 1, 1256},     // M=196 (tail), 8 n-tiles, residual
        {32, 56, 56, 64, 256, 1, 1, 0, 1, 1, 1256},    // many tiles per pair
        {32, 28, 28, 128, 128, 3, 1, 1, 1, 0, 1128},   // compute-bound 3x3, BN=128 pairs
        {32, 14, 14, 256, 256, 3, 1, 1, 1, 0, 1256},   // compute-bound 3x3, BN=256 pairs
        {32, 14, 14, 256, 256, 3, 1, 1, 1, 0, 128},    // same, single-CTA tiles (for comparison)
        // halo-resident 3x3 kernel (force_bn 3064)
        {1, 8, 8, 64, 64, 3, 1, 1, 0, 0, 3064},        // small: 4 tiles, borders everywhere
        {2, 56, 56, 64, 64, 3, 1, 1, 1, 0, 3064},      // layer1 conv2 shape
        {3, 28, 28, 64, 64, 3, 1, 1, 1, 0, 3064},      // narrower rows
        {32, 56, 56, 64, 64, 3, 1, 1, 1, 0, 3064},     // many tiles per CTA
        {32, 56, 56, 64, 64, 3, 1, 1, 1, 0, 64},       // same through the im2col kernel (for comparison)
    };

    int failures = 0;
    for (const Shape& s : shapes) {
        if (tf32 && s.bn == 3064) continue;  // the BF16 halo-resident kernel (its TF32 sibling is force code 4064)
        const int OH = (2 * s.pad + s.H - s.k) / s.stride + 1;
        const int OW = (2 * s.pad + s.W - s.k) / s.stride + 1;
        const size_t n_in = 1ull * s.B * s.H * s.W * s.Cin;
        const size_t K = 1ull * s.k * s.k * s.Cin;
        const size_t n_w = s.Cout * K;
        const size_t M = 1ull * s.B * OH * OW;
        const size_t n_out = M * s.Cout;
        std::vector<float> h_in(n_in), h_w(n_w), h_b(s.Cout), h_res(n_out);
        auto rnd = [&](float x) { return tf32 ? to_tf32f(x) : to_bf16f(x); };
        for (auto& v : h_in) v = rnd(frand() * 2.f);
        for (auto& v : h_w) v = rnd(frand() * 0.25f);
        for (auto& v : h_b) v = frand();
        for (auto& v : h_res) v = rnd(frand() * 2.f);

        void *d_in, *d_w, *d_res, *d_out;
        float* d_b;
        CK(cudaMalloc(&d_in, n_in * esz));
        CK(cudaMalloc(&d_w, n_w * esz));
        CK(cudaMalloc(&d_res, n_out * esz));
        CK(cudaMalloc(&d_out, n_out * esz));
        CK(cudaMalloc(&d_b, s.Cout * 4));
        CK(cudaMemset(d_out, 0xFF, n_out * esz));
        auto upload = [&](void* dst, const std::vector<float>& src) {
            if (tf32) {
                CK(cudaMemcpy(dst, src.data(), src.size() * 4, cudaMemcpyHostToDevice));
            } else {
                std::vector<__nv_bfloat16> t(src.size());
                for (size_t i = 0; i < src.size(); ++i) t[i] = __float2bfloat16(src[i]);
                CK(cudaMemcpy(dst, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
            }
        };
        upload(d_in, h_in);
        upload(d_w, h_w);
        upload(d_res, h_res);
        CK(cudaMemcpy(d_b, h_b.data(), s.Cout * 4, cudaMemcpyHostToDevice));

        rnb::ConvDesc d{};
        d.B = s.B; d.H = s.H; d.W = s.W; d.Cin = s.Cin; d.Cout = s.Cout;
        d.ksize = s.k; d.stride = s.stride; d.pad = s.pad; d.relu = s.relu;
        d.act = tf32 ? rnb::ActType::TF32 : rnb::ActType::BF16;
        d.in = d_in; d.weight = d_w; d.bias = d_b; d.residual = s.res ? d_res : nullptr; d.out = d_out;
        rnb::ConvPlan plan;
        char err[256];
        int rc = rnb::conv_plan_init(&plan, d, prop.multiProcessorCount, s.bn, err, sizeof(err));
        if (rc) {
            printf("plan failed: %s\n", err);
            return 3;
        }
        CK(rnb::conv_plan_launch(plan, 0));
        cudaError_t se = cudaDeviceSynchronize();
        if (se != cudaSuccess) {
            printf("kernel failed: %s\n", cudaGetErrorString(se));
            return 4;
        }
        // timing (3 more launches)
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; ++i) CK(rnb::conv_plan_launch(plan, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 3;

        std::vector<float> h_out(n_out);
        if (tf32) {
            CK(cudaMemcpy(h_out.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost));
        } else {
            std::vector<__nv_bfloat16> t(n_out);
            CK(cudaMemcpy(t.data(), d_out, n_out * 2, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n_out; ++i) h_out[i] = __bfloat162float(t[i]);
        }

        // host reference
        double max_err = 0, max_ref = 0;
        long long bad = 0;
        long long first_bad = -1;
#pragma omp parallel for schedule(dynamic, 16) reduction(max : max_err, max_ref) reduction(+ : bad)
        for (long long m = 0; m < static_cast<long long>(M); ++m) {
            const int b = m / (OH * OW);
            const int rem = m % (OH * OW);
            const int oh = rem / OW, ow = rem % OW;
            for (int oc = 0; oc < s.Cout; ++oc) {
                double acc = 0;
                for (int kh = 0; kh < s.k; ++kh)
                    for (int kw = 0; kw < s.k; ++kw) {
                        const int ih = oh * s.stride - s.pad + kh;
                        const int iw = ow * s.stride - s.pad + kw;
                        if (ih < 0 || ih >= s.H || iw < 0 || iw >= s.W) continue;
                        const float* ip = &h_in[((1ull * b * s.H + ih) * s.W + iw) * s.Cin];
                        const float* wp = &h_w[(1ull * oc * s.k * s.k + kh * s.k + kw) * s.Cin];
                        for (int ic = 0; ic < s.Cin; ++ic) acc += (double)ip[ic] * wp[ic];
                    }
                acc += h_b[oc];
                if (s.res) acc += h_res[m * s.Cout + oc];
                if (s.relu && acc < 0) acc = 0;
                const double got = h_out[m * s.Cout + oc];
                const double e = fabs(got - acc);
                const double tol = (tf32 ? 2e-3 : 1e-2) * fabs(acc) + (tf32 ? 2e-3 : 2e-2);
                if (!(e <= tol)) {
                    ++bad;
#pragma omp critical
                    if (first_bad < 0 || m * s.Cout + oc < first_bad) first_bad = m * s.Cout + oc;
                }
                if (e > max_err) max_err = e;
                if (fabs(acc) > max_ref) max_ref = fabs(acc);
            }
        }
        const double tflops = plan.flops / (ms * 1e-3) / 1e12;
        printf("[%s] B%d %dx%d Cin%d Cout%d k%d s%d p%d relu%d res%d bn%d grid%d : max_err %.4g (max |ref| %.4g) bad %lld/%zu  %.3f ms  %.1f TFLOP/s %s\n",
               tf32 ? "tf32" : "bf16", s.B, s.H, s.W, s.Cin, s.Cout, s.k, s.stride, s.pad, s.relu,
               s.res, plan.bn * (plan.ctas == 2 ? -1 : 1), plan.grid, max_err, max_ref, bad, n_out, ms, tflops,
               bad ? "FAIL" : "ok");
        if (bad) {
            ++failures;
            const long long m = first_bad / s.Cout, oc = first_bad % s.Cout;
            printf("   first bad at m=%lld oc=%lld got %g ; row dump (first 8 ch): ", m, oc,
                   h_out[first_bad]);
            for (int c = 0; c < 8; ++c) printf("%g ", h_out[m * s.Cout + c]);
            printf("\n");
        }
        cudaFree(d_in); cudaFree(d_w); cudaFree(d_res); cudaFree(d_out); cudaFree(d_b);
    }
    printf("conv_selftest: %d failing shapes\n", failures);
    return failures ? 1 : 0;
}
