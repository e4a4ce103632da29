// block_check.cu — exercises the Bottleneck / BasicBlock / Conv2d MODULE CLASSES of cuda/nn.cuh the way a user of the
// reference's host API would (weights from weights_bin/<prefix>..., FloatTensor in / out), for tests/test_gpu_modules.py:
//
//   block_check bottleneck|basic|conv <prefix> <in> <mid> <out> <stride> <downsample 0|1> <bf16|tf32> <B> <H> <x.bin> <y.bin>
//
// (conv: <prefix> is the conv's name, <mid> is the kernel size, <downsample> the padding.) Runs forward three times
// (the first plans), prints the wall time of the last synchronous call, saves the output like Tensor::save.
#include <chrono>
#include <cstring>
#include <string>

#include "nn.cuh"
#include "tensor.cuh"

int main(int argc, char** argv)
{
    if (argc < 13) {
        std::cerr << "usage: block_check kind prefix in mid out stride ds dtype B H x.bin y.bin\n";
        return 1;
    }
    const std::string kind = argv[1], prefix = argv[2];
    const uint64_t in = std::stoull(argv[3]), mid = std::stoull(argv[4]), outc = std::stoull(argv[5]);
    const uint64_t stride = std::stoull(argv[6]);
    const uint64_t ds = std::stoull(argv[7]);
    const Precision prec = !strcmp(argv[8], "tf32") ? Precision::TF32 : Precision::BF16;
    const uint64_t B = std::stoull(argv[9]), H = std::stoull(argv[10]);
    FloatTensor x = FloatTensor::loadToCuda(argv[11]).view(Shape({B, in, H, H}));
    double ms = 0;
    auto timed = [&](auto&& fwd) {
        for (int i = 0; i < 3; ++i) {
            const auto t0 = std::chrono::steady_clock::now();
            fwd();
            ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    };
    if (kind == "bottleneck") {
        Bottleneck blk = Bottleneck::loadWeightToCuda(prefix, in, mid, outc, stride, ds != 0, prec);
        FloatTensor y(blk.getOutShape(x.shape()), Device::GPU);
        timed([&] { blk.forward(x, y); });
        y.cpu().save(argv[12]);
    } else if (kind == "basic") {
        BasicBlock blk = BasicBlock::loadWeightToCuda(prefix, in, outc, stride, ds != 0, prec);
        FloatTensor y(blk.getOutShape(x.shape()), Device::GPU);
        timed([&] { blk.forward(x, y); });
        y.cpu().save(argv[12]);
    } else {
        Conv2d conv = Conv2d::loadWeightToCuda(prefix, in, outc, mid, stride, ds);
        FloatTensor y(conv.getOutShape(x.shape()), Device::GPU);
        timed([&] { conv.forward(x, y); });
        y.cpu().save(argv[12]);
    }
    std::cout << "forward_ms " << ms << "\n";
    return 0;
}
