// nn.cuh — drop-in for olehskip/resnet.c cuda/nn.cuh: Conv2d (:8-40), BatchNorm2d (:42-72),
// Pool2d (:74-98), Linear (:100-133), reluForward / addForward (:135-136) with the same constructors,
// public members, weight-file naming ("weights_bin/<name>.<suffix>") and forward(x, out) semantics
// (caller pre-allocates `out` with getOutShape, call is synchronous, aborts on CUDA error, in-place
// aliasing allowed for BN / ReLU / add).
//
// Added (BASELINE.json north_star): ReLU, MaxPool, Bottleneck, BasicBlock and a whole-model ResNet.
// The block and model objects are where the B200-native path lives: BatchNorm folded into the
// convolution, bias + residual + ReLU fused into the tcgen05 implicit-GEMM epilogue, NHWC BF16/TF32
// activations — all behind librnb.so (include/rnb.h). Everything still takes and returns the
// reference's FP32 NCHW FloatTensor.
#ifndef CUDA_NN_CUH
#define CUDA_NN_CUH

#include <memory>
#include <optional>
#include <string>

#include "helpers.cuh"
#include "ops.cuh"
#include "tensor.cuh"

// Arithmetic type of the tensor-core path (values match RNB_DTYPE_* in include/rnb.h).
enum class Precision : int
{
    BF16 = 0,
    TF32 = 1
};

// Planned block behind a module (include/rnb.h rnb_block_*): BN folded + weights packed once, NHWC staging and launch
// descriptors cached per input shape. Created lazily by the first forward() from the module's CURRENT weights (a copy:
// assign new weights -> construct a new module).
struct rnb_block;
using BlockPlanPtr = std::shared_ptr<rnb_block>;

class Conv2d
{
public:
    Conv2d(FloatTensor weight, uint64_t in_channels, uint64_t out_channels, uint64_t kernel_size,
           uint64_t stride = 1, uint64_t padding = 0)
        : weight(std::move(weight)), in_channels(in_channels), out_channels(out_channels),
          kernel_size(kernel_size), stride(stride), padding(padding)
    {
    }

    // Reads weights_bin/<name>.weight (raw float32, OIHW) onto the GPU.
    static Conv2d loadWeightToCuda(std::string name, uint64_t in_channels, uint64_t out_channels,
                                   uint64_t kernel_size, uint64_t stride = 1, uint64_t padding = 0)
    {
        FloatTensor flat = FloatTensor::loadToCuda("weights_bin/" + name + ".weight");
        return Conv2d(flat.view(Shape({out_channels, in_channels, kernel_size, kernel_size})), in_channels,
                      out_channels, kernel_size, stride, padding);
    }

    FloatTensor weight;
    const uint64_t in_channels, out_channels, kernel_size, stride, padding;

    Shape getOutShape(Shape x_shape)
    {
        assert(x_shape.size() == 4);
        assert(x_shape[1] == in_channels);
        return Shape({x_shape[0], out_channels, convOutputSize(x_shape[2], kernel_size, stride, padding),
                      convOutputSize(x_shape[3], kernel_size, stride, padding)});
    }

    // FP32 CUDA-core kernel with the reference's arithmetic order (bit-exact against ops.cu:14-48). With
    // RNB_MODULE_TC=tf32 | bf16 in the environment, shapes the tensor cores take (k in {1,3}, channels % 64 == 0) run
    // the tcgen05 implicit-GEMM kernel instead (<= 1e-3 / 2e-2 relative, BASELINE.json north_star).
    void forward(FloatTensor& x, FloatTensor& out);

private:
    BlockPlanPtr plan_;
};

class BatchNorm2d
{
public:
    BatchNorm2d(FloatTensor&& weight, FloatTensor&& bias, FloatTensor&& mean, FloatTensor&& var,
                uint64_t channels_num)
        : weight(std::move(weight)), bias(std::move(bias)), mean(std::move(mean)), var(std::move(var)),
          channels_num(channels_num)
    {
        assert(this->weight.shape() == Shape({channels_num}));
        assert(this->bias.shape() == Shape({channels_num}));
        assert(this->mean.shape() == Shape({channels_num}));
        assert(this->var.shape() == Shape({channels_num}));
    }

    // Reads weights_bin/<name>.{weight,bias,running_mean,running_var}.
    static BatchNorm2d loadWeightToCuda(std::string name, uint64_t channels_num)
    {
        const std::string base = "weights_bin/" + name;
        return BatchNorm2d(FloatTensor::loadToCuda(base + ".weight"), FloatTensor::loadToCuda(base + ".bias"),
                           FloatTensor::loadToCuda(base + ".running_mean"),
                           FloatTensor::loadToCuda(base + ".running_var"), channels_num);
    }

    FloatTensor weight;
    FloatTensor bias;
    FloatTensor mean;
    FloatTensor var;
    const uint64_t channels_num;

    void forward(FloatTensor& x, FloatTensor& out);
};

class Pool2d
{
public:
    Pool2d(uint64_t channels, uint64_t kernel_size, uint64_t stride = 1, uint64_t padding = 0)
        : channels(channels), kernel_size(kernel_size), stride(stride), padding(padding)
    {
    }
    const uint64_t channels, kernel_size, stride, padding;

    uint64_t outSideSize(uint64_t side_size)
    {
        return convOutputSize(side_size, kernel_size, stride, padding);
    }

    Shape getOutShape(Shape x_shape)
    {
        assert(x_shape.size() == 4);
        assert(x_shape[1] == channels);
        return Shape({x_shape[0], channels, outSideSize(x_shape[2]), outSideSize(x_shape[3])});
    }

    void maxforward(FloatTensor& x, FloatTensor& out);
    void avgforward(FloatTensor& x, FloatTensor& out);
};

class Linear
{
public:
    Linear(FloatTensor weight, FloatTensor bias, uint64_t in_features, uint64_t out_features)
        : weight(std::move(weight)), bias(std::move(bias)), in_features(in_features),
          out_features(out_features)
    {
        assert(this->weight.shape() == Shape({out_features, in_features}));
        assert(this->bias.shape() == Shape({out_features}));
    }

    // Reads weights_bin/<name>.weight ([out, in]) and weights_bin/<name>.bias.
    static Linear loadWeightToCuda(std::string name, uint64_t in_features, uint64_t out_features)
    {
        FloatTensor w = FloatTensor::loadToCuda("weights_bin/" + name + ".weight");
        FloatTensor b = FloatTensor::loadToCuda("weights_bin/" + name + ".bias");
        return Linear(w.view(Shape({out_features, in_features})), b.view(Shape({out_features})), in_features,
                      out_features);
    }

    FloatTensor weight, bias;
    const uint64_t in_features, out_features;

    Shape getOutShape(Shape x_shape)
    {
        assert(x_shape.size() == 2);
        assert(x_shape[1] == in_features);
        return Shape({x_shape.at(0), out_features});
    }

    void forward(FloatTensor& x, FloatTensor& out);
};

void reluForward(FloatTensor& x, FloatTensor& out);
void addForward(FloatTensor& a, FloatTensor& b, FloatTensor& out);

// ------------------------------------------------------------------------------------------------
// Additions in the reference's style.

// Module face of reluForward.
class ReLU
{
public:
    void forward(FloatTensor& x, FloatTensor& out)
    {
        reluForward(x, out);
    }
};

// Module face of Pool2d::maxforward.
class MaxPool : public Pool2d
{
public:
    using Pool2d::Pool2d;
    void forward(FloatTensor& x, FloatTensor& out)
    {
        maxforward(x, out);
    }
};

// conv1x1 -> bn -> relu -> conv3x3(stride) -> bn -> relu -> conv1x1 -> bn -> (+shortcut) -> relu,
// shortcut = bn(conv1x1/stride(x)) when the block changes shape, else x. This is the reference's
// ResnetBlock + layerForward body (cuda/inference/main.cu:18-46, 130-164) as ONE module whose forward
// runs three (four with a shortcut conv) fused tensor-core launches.
class Bottleneck
{
public:
    Bottleneck(Conv2d&& conv1, BatchNorm2d&& bn1, Conv2d&& conv2, BatchNorm2d&& bn2, Conv2d&& conv3,
               BatchNorm2d&& bn3, std::optional<std::pair<Conv2d, BatchNorm2d>> downsample = {},
               Precision precision = Precision::BF16)
        : conv1(std::move(conv1)), bn1(std::move(bn1)), conv2(std::move(conv2)), bn2(std::move(bn2)),
          conv3(std::move(conv3)), bn3(std::move(bn3)), downsample(std::move(downsample)),
          precision(precision)
    {
    }

    // Weight names follow createLayer (main.cu:59-75): "<prefix>conv1", "<prefix>bn1", ...,
    // "<prefix>downsample.0" / "<prefix>downsample.1"; prefix e.g. "layer2.0.".
    static Bottleneck loadWeightToCuda(std::string prefix, uint64_t in_channels, uint64_t inter_channels,
                                       uint64_t out_channels, uint64_t stride, bool with_downsample,
                                       Precision precision = Precision::BF16);

    Conv2d conv1;
    BatchNorm2d bn1;
    Conv2d conv2;
    BatchNorm2d bn2;
    Conv2d conv3;
    BatchNorm2d bn3;
    std::optional<std::pair<Conv2d, BatchNorm2d>> downsample;
    const Precision precision;

    Shape getOutShape(Shape x_shape)
    {
        return conv3.getOutShape(conv2.getOutShape(conv1.getOutShape(x_shape)));
    }

    void forward(FloatTensor& x, FloatTensor& out);

private:
    BlockPlanPtr plan_;
};

// conv3x3(stride) -> bn -> relu -> conv3x3 -> bn -> (+shortcut) -> relu (torchvision BasicBlock; the
// reference has no such block — ResNet-18/34 need it).
class BasicBlock
{
public:
    BasicBlock(Conv2d&& conv1, BatchNorm2d&& bn1, Conv2d&& conv2, BatchNorm2d&& bn2,
               std::optional<std::pair<Conv2d, BatchNorm2d>> downsample = {},
               Precision precision = Precision::BF16)
        : conv1(std::move(conv1)), bn1(std::move(bn1)), conv2(std::move(conv2)), bn2(std::move(bn2)),
          downsample(std::move(downsample)), precision(precision)
    {
    }

    static BasicBlock loadWeightToCuda(std::string prefix, uint64_t in_channels, uint64_t out_channels,
                                       uint64_t stride, bool with_downsample,
                                       Precision precision = Precision::BF16);

    Conv2d conv1;
    BatchNorm2d bn1;
    Conv2d conv2;
    BatchNorm2d bn2;
    std::optional<std::pair<Conv2d, BatchNorm2d>> downsample;
    const Precision precision;

    Shape getOutShape(Shape x_shape)
    {
        return conv2.getOutShape(conv1.getOutShape(x_shape));
    }

    void forward(FloatTensor& x, FloatTensor& out);

private:
    BlockPlanPtr plan_;
};

// Whole network (ResnetModel + createResnet152 + resnet152Forward + the CPU arg-max of
// cuda/inference/main.cu:91-125, 168-226, 243-251) for "resnet18|34|50|101|152". Loads every file
// of `weights_dir` (save_weights.py format) once, folds BN, plans the fused launches, and replays
// them from a CUDA graph.
class ResNet
{
public:
    ResNet(const std::string& arch, Precision precision = Precision::BF16,
           const std::string& weights_dir = "weights_bin", uint64_t max_batch = 1);
    ~ResNet();
    ResNet(const ResNet&) = delete;
    ResNet& operator=(const ResNet&) = delete;

    uint64_t numClasses() const;
    Shape getOutShape(Shape x_shape)
    {
        assert(x_shape.size() == 4);
        return Shape({x_shape[0], numClasses()});
    }

    // x: [B,3,224,224] on the GPU; logits: [B,numClasses()] on the GPU (pre-allocated). Synchronous.
    void forward(FloatTensor& x, FloatTensor& logits);
    // Same, and also returns the per-image arg-max (first maximum wins) on the host.
    std::vector<int32_t> predict(FloatTensor& x, FloatTensor& logits);
    // The reference's whole sequence from HOST tensors (Tensor::load -> loadToCuda -> forward -> cpu(),
    // cuda/inference/main.cu:233-251) in one call: x and logits live on the CPU. Input upload, forward pass and
    // read-back are pipelined inside (rnb_model_forward_host); on the BF16 path the host cores may round part of the
    // batch to BF16 so that half of its bytes cross PCIe — the logits are bit-identical to predict() either way.
    std::vector<int32_t> predictHost(FloatTensor& x_cpu, FloatTensor& logits_cpu);
    // Decoded images: x_u8_dev is uint8 HWC [B,224,224,3] on the GPU; the /255 + mean/std normalisation of
    // convert_imgs_to_bin.py:18 runs inside the stem pre-pass (bit-identical to predict() on the float tensor).
    std::vector<int32_t> predictU8(const uint8_t* x_u8_dev, uint64_t B, FloatTensor& logits);

    // Pre-packed weight blob: one checksummed file instead of one raw file per state_dict key.
    void savePacked(const std::string& path);
    struct Packed {};  // tag
    ResNet(Packed, const std::string& path, uint64_t max_batch = 1);

private:
    struct rnb_model* handle_ = nullptr;
    uint64_t max_batch_;
};

#endif  // CUDA_NN_CUH
