// ref_cuda_driver.cu — ORACLE (test infrastructure). An extern "C" face over the reference's OWN CUDA
// implementation: this file is compiled together with $(REFERENCE)/cuda/ops.cu and
// $(REFERENCE)/cuda/nn.cu (used where they lie, never copied) against the reference's own headers,
// into oracle/_ref/libref_cuda.so. The -m gpu tests call it as a second opinion on the CPU oracle:
// same inputs through the reference's Conv2d / BatchNorm2d / Pool2d / Linear / reluForward /
// addForward modules (cuda/nn.cuh, cuda/nn.cu) on the GPU.
//
// All pointers are HOST pointers; every call uploads, runs the reference module, downloads.
#include <cstring>

#include "nn.cuh"      // the REFERENCE's headers (-I$(REFERENCE)/cuda)
#include "ops.cuh"
#include "tensor.cuh"

namespace
{
FloatTensor toGpu(const float* host, Shape shape)
{
    FloatTensor cpu(shape, Device::CPU);
    std::memcpy(cpu.data(), host, cpu.size());
    return cpu.cuda();
}
void toHost(FloatTensor& gpu, float* host)
{
    FloatTensor cpu = gpu.cpu();
    std::memcpy(host, cpu.data(), cpu.size());
}
}  // namespace

extern "C" {

void refcuda_conv2d(const float* x, const float* w, float* out, uint64_t B, uint64_t Cin, uint64_t H,
                    uint64_t W, uint64_t Cout, uint64_t k, uint64_t stride, uint64_t pad)
{
    FloatTensor xg = toGpu(x, Shape({B, Cin, H, W}));
    Conv2d conv(toGpu(w, Shape({Cout, Cin, k, k})), Cin, Cout, k, stride, pad);
    FloatTensor og(conv.getOutShape(xg.shape()), Device::GPU);
    conv.forward(xg, og);
    toHost(og, out);
}

void refcuda_batchnorm2d(const float* x, const float* weight, const float* bias, const float* mean,
                         const float* var, float* out, uint64_t B, uint64_t C, uint64_t H, uint64_t W)
{
    FloatTensor xg = toGpu(x, Shape({B, C, H, W}));
    BatchNorm2d bn(toGpu(weight, Shape({C})), toGpu(bias, Shape({C})), toGpu(mean, Shape({C})),
                   toGpu(var, Shape({C})), C);
    bn.forward(xg, xg);  // in place, as main.cu:145 does
    toHost(xg, out);
}

void refcuda_relu(const float* x, float* out, uint64_t n)
{
    FloatTensor xg = toGpu(x, Shape({n}));
    reluForward(xg, xg);
    toHost(xg, out);
}

void refcuda_add(const float* a, const float* b, float* out, uint64_t n)
{
    FloatTensor ag = toGpu(a, Shape({n}));
    FloatTensor bg = toGpu(b, Shape({n}));
    addForward(ag, bg, ag);
    toHost(ag, out);
}

void refcuda_pool2d(int is_max, const float* x, float* out, uint64_t B, uint64_t C, uint64_t H, uint64_t W,
                    uint64_t k, uint64_t stride, uint64_t pad)
{
    FloatTensor xg = toGpu(x, Shape({B, C, H, W}));
    Pool2d pool(C, k, stride, pad);
    FloatTensor og(pool.getOutShape(xg.shape()), Device::GPU);
    if (is_max) {
        pool.maxforward(xg, og);
    } else {
        pool.avgforward(xg, og);
    }
    toHost(og, out);
}

void refcuda_linear(const float* x, const float* w, const float* bias, float* out, uint64_t B,
                    uint64_t in_features, uint64_t out_features)
{
    FloatTensor xg = toGpu(x, Shape({B, in_features}));
    Linear fc(toGpu(w, Shape({out_features, in_features})), toGpu(bias, Shape({out_features})), in_features,
              out_features);
    FloatTensor og(fc.getOutShape(xg.shape()), Device::GPU);
    fc.forward(xg, og);
    toHost(og, out);
}

}  // extern "C"
