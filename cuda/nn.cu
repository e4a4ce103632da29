// nn.cu — host side of the drop-in modules (reference: cuda/nn.cu:3-87). Every forward hands raw
// device pointers to the C ABI of librnb.so (include/rnb.h), then restores the reference's calling
// convention: synchronous on return, abort on any error.
#include "nn.cuh"

#include <cstdlib>
#include <vector>

#include "../include/rnb.h"

namespace
{

// The reference aborts on every failure (helpers.cuh:13-22); keep that at the module boundary.
void rnbCheck(int rc, const char* what)
{
    if (rc != RNB_OK) {
        std::cerr << "rnb: " << what << " failed (" << rc << "): " << rnb_last_error() << "\n";
        std::abort();
    }
}

void ensureInit()
{
    static bool done = false;
    if (!done) {
        int device = 0;
        gpuErrchk(cudaGetDevice(&device));
        rnbCheck(rnb_init(device), "rnb_init");
        done = true;
    }
}

void syncAndCheck()
{
    gpuErrchk(cudaDeviceSynchronize());
    gpuErrchk(cudaGetLastError());
}

int asInt(uint64_t v)
{
    assert(v <= 0x7fffffffULL);
    return static_cast<int>(v);
}

bool tensorCoreEligible(const Conv2d& c)
{
    return (c.kernel_size == 1 || c.kernel_size == 3) && c.in_channels % 64 == 0 && c.out_channels % 64 == 0;
}

rnb_conv_params_t convParams(Conv2d& c, BatchNorm2d* bn)
{
    rnb_conv_params_t p{};
    p.w = c.weight.data();
    if (bn) {
        p.bn_weight = bn->weight.data();
        p.bn_bias = bn->bias.data();
        p.bn_mean = bn->mean.data();
        p.bn_var = bn->var.data();
    }
    p.Cin = asInt(c.in_channels);
    p.Cout = asInt(c.out_channels);
    p.k = asInt(c.kernel_size);
    p.stride = asInt(c.stride);
    p.pad = asInt(c.padding);
    return p;
}

BlockPlanPtr makePlan(int kind, Precision precision, const std::vector<rnb_conv_params_t>& convs)
{
    rnb_block_t* b = nullptr;
    rnbCheck(rnb_block_create(kind, static_cast<int>(precision), convs.data(), static_cast<int>(convs.size()), &b),
             "rnb_block_create");
    return BlockPlanPtr(b, [](rnb_block* p) { rnb_block_destroy(p); });
}

// RNB_MODULE_TC=tf32 | bf16: Conv2d::forward on tensor cores where the shape allows (default: FP32 CUDA cores)
int moduleTcPrecision()
{
    static const int mode = [] {
        const char* e = std::getenv("RNB_MODULE_TC");
        if (!e) return -1;
        const std::string v(e);
        return v == "tf32" ? RNB_DTYPE_TF32 : (v == "bf16" ? RNB_DTYPE_BF16 : -1);
    }();
    return mode;
}

// conv -> bn -> (+residual) -> relu through the fused tcgen05 path when the shape allows it,
// otherwise through the per-op FP32 kernels (e.g. a 3-channel input).
void convBnAct(Conv2d& conv, BatchNorm2d& bn, FloatTensor& x, FloatTensor* residual, bool relu,
               Precision precision, FloatTensor& out)
{
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    if (tensorCoreEligible(conv)) {
        rnbCheck(rnb_conv_bn_act_forward(x.data(), conv.weight.data(), bn.weight.data(), bn.bias.data(),
                                         bn.mean.data(), bn.var.data(), residual ? residual->data() : nullptr,
                                         out.data(), asInt(B), asInt(C), asInt(H), asInt(W),
                                         asInt(conv.out_channels), asInt(conv.kernel_size), asInt(conv.stride),
                                         asInt(conv.padding), relu ? 1 : 0, static_cast<int>(precision),
                                         nullptr),
                 "rnb_conv_bn_act_forward");
        syncAndCheck();
        return;
    }
    conv.forward(x, out);
    bn.forward(out, out);
    if (residual) {
        addForward(out, *residual, out);
    }
    if (relu) {
        reluForward(out, out);
    }
}

std::optional<std::pair<Conv2d, BatchNorm2d>> loadDownsample(const std::string& prefix, uint64_t in_channels,
                                                             uint64_t out_channels, uint64_t stride,
                                                             bool with_downsample)
{
    std::optional<std::pair<Conv2d, BatchNorm2d>> ds;
    if (with_downsample) {
        ds.emplace(Conv2d::loadWeightToCuda(prefix + "downsample.0", in_channels, out_channels, 1, stride),
                   BatchNorm2d::loadWeightToCuda(prefix + "downsample.1", out_channels));
    }
    return ds;
}

}  // namespace

void Conv2d::forward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    assert(C == in_channels);
    assert(out.shape() == getOutShape(x.shape()));
    if (moduleTcPrecision() >= 0 && tensorCoreEligible(*this)) {
        if (!plan_) {
            std::vector<rnb_conv_params_t> convs;
            convs.push_back(convParams(*this, nullptr));
            plan_ = makePlan(RNB_BLOCK_CONV, static_cast<Precision>(moduleTcPrecision()), convs);
        }
        rnbCheck(rnb_block_forward(plan_.get(), x.data(), asInt(B), asInt(H), asInt(W), nullptr, 0, out.data(), nullptr),
                 "rnb_block_forward");
        syncAndCheck();
        return;
    }
    rnbCheck(rnb_conv2d_forward(x.data(), out.data(), weight.data(), asInt(B), asInt(C), asInt(H), asInt(W),
                                asInt(out_channels), asInt(kernel_size), asInt(stride), asInt(padding),
                                nullptr),
             "rnb_conv2d_forward");
    syncAndCheck();
}

void BatchNorm2d::forward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    assert(C == channels_num);
    rnbCheck(rnb_batchnorm2d_forward(x.data(), out.data(), weight.data(), bias.data(), mean.data(), var.data(),
                                     asInt(B), asInt(C), asInt(H * W), nullptr),
             "rnb_batchnorm2d_forward");
    syncAndCheck();
}

void Pool2d::maxforward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    rnbCheck(rnb_maxpool2d_forward(x.data(), out.data(), asInt(B), asInt(C), asInt(H), asInt(W),
                                   asInt(kernel_size), asInt(stride), asInt(padding), nullptr),
             "rnb_maxpool2d_forward");
    syncAndCheck();
}

void Pool2d::avgforward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    rnbCheck(rnb_avgpool2d_forward(x.data(), out.data(), asInt(B), asInt(C), asInt(H), asInt(W),
                                   asInt(kernel_size), asInt(stride), asInt(padding), nullptr),
             "rnb_avgpool2d_forward");
    syncAndCheck();
}

void Linear::forward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    const uint64_t B = x.shape().at(0);
    rnbCheck(rnb_linear_forward(x.data(), out.data(), weight.data(), bias.data(), asInt(B), asInt(in_features),
                                asInt(out_features), nullptr),
             "rnb_linear_forward");
    syncAndCheck();
}

void reluForward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    assert(x.shape() == out.shape());
    rnbCheck(rnb_relu_forward(x.data(), out.data(), static_cast<int64_t>(x.numel()), nullptr),
             "rnb_relu_forward");
    syncAndCheck();
}

void addForward(FloatTensor& a, FloatTensor& b, FloatTensor& out)
{
    ensureInit();
    assert(a.shape() == b.shape());
    assert(a.shape() == out.shape());
    rnbCheck(rnb_add_forward(a.data(), b.data(), out.data(), static_cast<int64_t>(a.numel()), nullptr),
             "rnb_add_forward");
    syncAndCheck();
}

// ------------------------------------------------------------------------------------------------
Bottleneck Bottleneck::loadWeightToCuda(std::string prefix, uint64_t in_channels, uint64_t inter_channels,
                                        uint64_t out_channels, uint64_t stride, bool with_downsample,
                                        Precision precision)
{
    return Bottleneck(Conv2d::loadWeightToCuda(prefix + "conv1", in_channels, inter_channels, 1),
                      BatchNorm2d::loadWeightToCuda(prefix + "bn1", inter_channels),
                      Conv2d::loadWeightToCuda(prefix + "conv2", inter_channels, inter_channels, 3, stride, 1),
                      BatchNorm2d::loadWeightToCuda(prefix + "bn2", inter_channels),
                      Conv2d::loadWeightToCuda(prefix + "conv3", inter_channels, out_channels, 1),
                      BatchNorm2d::loadWeightToCuda(prefix + "bn3", out_channels),
                      loadDownsample(prefix, in_channels, out_channels, stride, with_downsample), precision);
}

void Bottleneck::forward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    assert(out.shape() == getOutShape(x.shape()));
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    if (tensorCoreEligible(conv1) && tensorCoreEligible(conv2) && tensorCoreEligible(conv3) &&
        (!downsample || tensorCoreEligible(downsample->first))) {
        // planned block: folded / packed weights cached, NHWC intermediates, fused layer1 tail where it applies
        if (!plan_) {
            std::vector<rnb_conv_params_t> convs;
            convs.push_back(convParams(conv1, &bn1));
            convs.push_back(convParams(conv2, &bn2));
            convs.push_back(convParams(conv3, &bn3));
            if (downsample) {
                convs.push_back(convParams(downsample->first, &downsample->second));
            }
            plan_ = makePlan(RNB_BLOCK_BOTTLENECK, precision, convs);
        }
        rnbCheck(rnb_block_forward(plan_.get(), x.data(), asInt(B), asInt(H), asInt(W), nullptr, 1, out.data(), nullptr),
                 "rnb_block_forward");
        syncAndCheck();
        return;
    }
    FloatTensor shortcut(Device::GPU);
    if (downsample) {
        shortcut = FloatTensor(downsample->first.getOutShape(x.shape()), Device::GPU);
        convBnAct(downsample->first, downsample->second, x, nullptr, false, precision, shortcut);
    }
    FloatTensor t1(conv1.getOutShape(x.shape()), Device::GPU);
    convBnAct(conv1, bn1, x, nullptr, true, precision, t1);
    FloatTensor t2(conv2.getOutShape(t1.shape()), Device::GPU);
    convBnAct(conv2, bn2, t1, nullptr, true, precision, t2);
    convBnAct(conv3, bn3, t2, downsample ? &shortcut : &x, true, precision, out);
}

BasicBlock BasicBlock::loadWeightToCuda(std::string prefix, uint64_t in_channels, uint64_t out_channels,
                                        uint64_t stride, bool with_downsample, Precision precision)
{
    return BasicBlock(Conv2d::loadWeightToCuda(prefix + "conv1", in_channels, out_channels, 3, stride, 1),
                      BatchNorm2d::loadWeightToCuda(prefix + "bn1", out_channels),
                      Conv2d::loadWeightToCuda(prefix + "conv2", out_channels, out_channels, 3, 1, 1),
                      BatchNorm2d::loadWeightToCuda(prefix + "bn2", out_channels),
                      loadDownsample(prefix, in_channels, out_channels, stride, with_downsample), precision);
}

void BasicBlock::forward(FloatTensor& x, FloatTensor& out)
{
    ensureInit();
    assert(out.shape() == getOutShape(x.shape()));
    const auto [B, C, H, W] = x.shape().as_tuple<4>();
    if (tensorCoreEligible(conv1) && tensorCoreEligible(conv2) && (!downsample || tensorCoreEligible(downsample->first))) {
        if (!plan_) {
            std::vector<rnb_conv_params_t> convs;
            convs.push_back(convParams(conv1, &bn1));
            convs.push_back(convParams(conv2, &bn2));
            if (downsample) {
                convs.push_back(convParams(downsample->first, &downsample->second));
            }
            plan_ = makePlan(RNB_BLOCK_BASIC, precision, convs);
        }
        rnbCheck(rnb_block_forward(plan_.get(), x.data(), asInt(B), asInt(H), asInt(W), nullptr, 1, out.data(), nullptr),
                 "rnb_block_forward");
        syncAndCheck();
        return;
    }
    FloatTensor shortcut(Device::GPU);
    if (downsample) {
        shortcut = FloatTensor(downsample->first.getOutShape(x.shape()), Device::GPU);
        convBnAct(downsample->first, downsample->second, x, nullptr, false, precision, shortcut);
    }
    FloatTensor t1(conv1.getOutShape(x.shape()), Device::GPU);
    convBnAct(conv1, bn1, x, nullptr, true, precision, t1);
    convBnAct(conv2, bn2, t1, downsample ? &shortcut : &x, true, precision, out);
}

// ------------------------------------------------------------------------------------------------
ResNet::ResNet(const std::string& arch, Precision precision, const std::string& weights_dir,
               uint64_t max_batch)
    : max_batch_(max_batch)
{
    ensureInit();
    rnbCheck(rnb_model_create(arch.c_str(), static_cast<int>(precision), weights_dir.c_str(), asInt(max_batch),
                              0, &handle_),
             "rnb_model_create");
}

ResNet::ResNet(Packed, const std::string& path, uint64_t max_batch) : max_batch_(max_batch)
{
    ensureInit();
    rnbCheck(rnb_model_create_packed(path.c_str(), asInt(max_batch), 0, &handle_), "rnb_model_create_packed");
}

void ResNet::savePacked(const std::string& path)
{
    rnbCheck(rnb_model_save_packed(handle_, path.c_str()), "rnb_model_save_packed");
}

std::vector<int32_t> ResNet::predictU8(const uint8_t* x_u8_dev, uint64_t B, FloatTensor& logits)
{
    assert(logits.device == Device::GPU && logits.shape() == Shape({B, numClasses()}));
    int32_t* top1_dev = static_cast<int32_t*>(safeCudaMalloc(B * sizeof(int32_t)));
    rnbCheck(rnb_model_forward_u8(handle_, x_u8_dev, asInt(B), logits.data(), top1_dev, nullptr),
             "rnb_model_forward_u8");
    syncAndCheck();
    std::vector<int32_t> top1(B);
    gpuErrchk(cudaMemcpy(top1.data(), top1_dev, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
    gpuErrchk(cudaFree(top1_dev));
    return top1;
}

ResNet::~ResNet()
{
    rnb_model_destroy(handle_);
}

uint64_t ResNet::numClasses() const
{
    return static_cast<uint64_t>(rnb_model_num_classes(handle_));
}

void ResNet::forward(FloatTensor& x, FloatTensor& logits)
{
    assert(x.device == Device::GPU && logits.device == Device::GPU);
    assert(logits.shape() == getOutShape(x.shape()));
    rnbCheck(rnb_model_forward(handle_, x.data(), asInt(x.shape().at(0)), logits.data(), nullptr, nullptr),
             "rnb_model_forward");
    syncAndCheck();
}

std::vector<int32_t> ResNet::predictHost(FloatTensor& x_cpu, FloatTensor& logits_cpu)
{
    assert(x_cpu.device == Device::CPU && logits_cpu.device == Device::CPU);
    assert(logits_cpu.shape() == getOutShape(x_cpu.shape()));
    const uint64_t B = x_cpu.shape().at(0);
    std::vector<int32_t> top1(B);
    rnbCheck(rnb_model_forward_host(handle_, x_cpu.data(), asInt(B), logits_cpu.data(), top1.data()),
             "rnb_model_forward_host");   // synchronous: results are on the host when it returns
    return top1;
}

std::vector<int32_t> ResNet::predict(FloatTensor& x, FloatTensor& logits)
{
    const uint64_t B = x.shape().at(0);
    int32_t* top1_dev = static_cast<int32_t*>(safeCudaMalloc(B * sizeof(int32_t)));
    rnbCheck(rnb_model_forward(handle_, x.data(), asInt(B), logits.data(), top1_dev, nullptr),
             "rnb_model_forward");
    syncAndCheck();
    std::vector<int32_t> top1(B);
    gpuErrchk(cudaMemcpy(top1.data(), top1_dev, B * sizeof(int32_t), cudaMemcpyDeviceToHost));
    gpuErrchk(cudaFree(top1_dev));
    return top1;
}
