// resnet_infer.cu — the reference's driver (cuda/inference/main.cu:228-254) rebuilt on the
// whole-model object: load weights_bin/ + an image .bin, run the network, print the arg-max.
//
//   resnet_infer [arch=resnet152] [bf16|tf32] [batch=1] [weights_dir=weights_bin] [image.bin]
//
// The image file is the [1,3,224,224] float32 NCHW tensor written by convert_imgs_to_bin.py; for
// batch > 1 it is replicated. Prints "max index is N" per image exactly like the reference.
#include <chrono>
#include <cstring>
#include <string>

#include "nn.cuh"
#include "tensor.cuh"

int main(int argc, char** argv)
{
    const std::string arch = argc > 1 ? argv[1] : "resnet152";
    const Precision precision = (argc > 2 && !strcmp(argv[2], "tf32")) ? Precision::TF32 : Precision::BF16;
    const uint64_t B = argc > 3 ? std::stoull(argv[3]) : 1;
    const std::string weights_dir = argc > 4 ? argv[4] : "weights_bin";
    const std::string image = argc > 5 ? argv[5] : "test_bins/ILSVRC2012_val_00004749.bin";
    const uint64_t C = 3, H = 224, W = 224;
    std::cout << "Started\n";

    ResNet model(arch, precision, weights_dir, B);
    std::cout << "created model\n";

    FloatTensor one = FloatTensor::loadToCpu(image);
    assert(one.numel() == C * H * W);
    FloatTensor batch_cpu(Shape({B, C, H, W}), Device::CPU);
    for (uint64_t b = 0; b < B; ++b) {
        std::memcpy(batch_cpu.data() + b * C * H * W, one.data(), C * H * W * sizeof(float));
    }
    FloatTensor inp = batch_cpu.cuda();
    FloatTensor logits(model.getOutShape(inp.shape()), Device::GPU);

    std::vector<int32_t> top1 = model.predict(inp, logits);  // first call plans + captures the graph
    const auto t0 = std::chrono::steady_clock::now();
    top1 = model.predict(inp, logits);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Finished (" << ms << " ms for " << B << " image(s))\n";

    for (uint64_t b = 0; b < B; ++b) {
        std::cout << "max index is " << top1[b] << std::endl;
    }
    if (std::getenv("RNB_CHECK_HOST_PATH")) {
        // the same batch through host tensors (upload + forward + read-back in one call): must give the same answer
        FloatTensor logits_cpu(model.getOutShape(batch_cpu.shape()), Device::CPU);
        const std::vector<int32_t> top1_host = model.predictHost(batch_cpu, logits_cpu);
        FloatTensor logits_dev_cpu = logits.cpu();
        const bool same = top1_host == top1 &&
                          !std::memcmp(logits_cpu.data(), logits_dev_cpu.data(), logits_cpu.numel() * sizeof(float));
        std::cout << "host path: " << (same ? "identical" : "DIFFERENT") << std::endl;
        if (!same) return 1;
    }
    if (const char* dump = std::getenv("RNB_DUMP_LOGITS")) {
        logits.cpu().save(dump);  // raw float32 [B, classes], the reference's "cuda_out.bin" idea
    }
    return 0;
}
