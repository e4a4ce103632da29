// tensor.cuh — drop-in for olehskip/resnet.c cuda/tensor.cuh: the same public names and semantics
// (Device :15, Shape :21-57, Tensor<T> :59-245, FloatTensor :247), re-implemented.
//
// Kept on purpose (callers such as cuda/inference/main.cu rely on them):
//   * Tensor(Device) is an empty, falsy tensor of shape {0}; Tensor(Shape, Device) allocates
//     UNINITIALISED storage (malloc / cudaMalloc).
//   * "move" construction shares the storage with the source (the source stays usable);
//     move-assignment empties the source; copy-assignment is deleted.
//   * view() aliases the same storage; toDevice()/cuda()/cpu() always copy and throw
//     std::runtime_error when source and destination devices are equal.
//   * loadToCpu() sizes the tensor from the file (raw float32, no header) and aborts if the file
//     cannot be opened; save() writes the raw bytes of a CPU tensor.
// Changed on purpose:
//   * Shape::numel() accumulates in uint64_t (the reference's std::accumulate seed is an int and
//     silently wraps past 2^31-1 elements, tensor.cuh:28).
#ifndef CUDA_TENSOR_CUH
#define CUDA_TENSOR_CUH

#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "helpers.cuh"

enum class Device
{
    CPU,
    GPU
};

class Shape : public std::vector<uint64_t>
{
public:
    using std::vector<uint64_t>::vector;

    uint64_t numel() const
    {
        assert(!empty());
        uint64_t n = 1;
        for (uint64_t d : *this) {
            n *= d;
        }
        return n;
    }

    // Shape as a std::tuple of exactly N extents (aborts when the rank differs), for
    // `const auto [b, c, h, w] = shape.as_tuple<4>();`
    template <std::size_t N>
    auto as_tuple() const
    {
        static_assert(N > 0, "Tuple size must be positive");
        if (size() != N) {
            std::abort();
        }
        return unpack(std::make_index_sequence<N>{});
    }

    friend std::ostream& operator<<(std::ostream& os, const Shape& s)
    {
        os << "(";
        const char* sep = "";
        for (uint64_t d : s) {
            os << sep << d;
            sep = ", ";
        }
        return os << ")";
    }

private:
    template <std::size_t... I>
    auto unpack(std::index_sequence<I...>) const
    {
        return std::make_tuple((*this)[I]...);
    }
};

template <class T>
struct Tensor
{
    // Empty tensor: no storage, shape {0}, converts to false.
    Tensor(Device device) : device(device), shape_({0}) {}

    // Allocates numel()*sizeof(T) bytes on `device`; contents are uninitialised.
    Tensor(Shape shape, Device device = Device::CPU) : device(device), shape_(std::move(shape))
    {
        assert(!shape_.empty());
        if (numel() == 0) {
            return;
        }
        if (device == Device::CPU) {
            storage_ = std::shared_ptr<T>(static_cast<T*>(std::malloc(size())), [](T* p) { std::free(p); });
        } else {
            storage_ = std::shared_ptr<T>(static_cast<T*>(safeCudaMalloc(size())), [](T* p) {
                if (p) {
                    cudaFree(p);
                }
            });
        }
    }

    // Shares the storage of `other` (reference semantics: tensor.cuh:96-102).
    Tensor(Tensor<T>&& other) : device(other.device), shape_(other.shape_), storage_(other.storage_)
    {
        assert(!shape_.empty());
    }

    static Tensor<T> arange_cpu(Shape shape)
    {
        Tensor<T> t(shape, Device::CPU);
        const uint64_t n = t.numel();
        for (uint64_t i = 0; i < n; ++i) {
            t.data()[i] = static_cast<T>(i);
        }
        return t;
    }

    static Tensor<T> ones_cpu(Shape shape)
    {
        Tensor<T> t(shape, Device::CPU);
        const uint64_t n = t.numel();
        for (uint64_t i = 0; i < n; ++i) {
            t.data()[i] = static_cast<T>(1);
        }
        return t;
    }

    // Raw file of T (no header) -> 1-D CPU tensor sized from the file.
    static Tensor<T> loadToCpu(std::string file_name)
    {
        std::ifstream file(file_name, std::ios::binary | std::ios::ate);
        if (!file.is_open()) {
            std::cerr << "Can't open " << file_name << std::endl;
            std::abort();
        }
        const std::streamsize bytes = file.tellg();
        const uint64_t n = static_cast<uint64_t>(bytes) / sizeof(T);
        assert(n > 0);
        Tensor<T> out(Shape({n}), Device::CPU);
        file.seekg(0, std::ios::beg);
        file.read(reinterpret_cast<char*>(out.data()), static_cast<std::streamsize>(n * sizeof(T)));
        assert(!file.fail());
        return out;
    }

    static Tensor<T> loadToCuda(std::string file_name)
    {
        return loadToCpu(file_name).cuda();
    }

    void save(std::string file_name)
    {
        assert(device == Device::CPU);
        std::ofstream file(file_name, std::ios::binary);
        assert(file.is_open());
        file.write(reinterpret_cast<const char*>(data()), static_cast<std::streamsize>(size()));
        assert(!file.fail());
    }

    // Same storage, different shape (element counts must match).
    Tensor<T> view(Shape new_shape)
    {
        assert(!new_shape.empty());
        assert(new_shape.numel() == shape_.numel());
        return Tensor<T>(storage_, std::move(new_shape), device);
    }

    uint64_t numel() const
    {
        return shape_.numel();
    }

    uint64_t size()
    {
        return numel() * sizeof(T);
    }

    const Device device = Device::CPU;

    // Always a fresh copy on the other device; same-device "transfers" throw.
    Tensor<T> toDevice(Device target)
    {
        if (target == device) {
            throw std::runtime_error("Unsupported device transfer combination");
        }
        gpuErrchk(cudaDeviceSynchronize());
        Tensor<T> out(shape_, target);
        const cudaMemcpyKind kind = target == Device::GPU ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        if (numel() != 0) {
            gpuErrchk(cudaMemcpy(out.data(), data(), size(), kind));
        }
        gpuErrchk(cudaDeviceSynchronize());
        return out;
    }

    Tensor<T> cuda()
    {
        return toDevice(Device::GPU);
    }

    Tensor<T> cpu()
    {
        return toDevice(Device::CPU);
    }

    void operator=(const Tensor<T>&) = delete;
    void operator=(Tensor<T>&& other)
    {
        assert(device == other.device);
        storage_ = std::move(other.storage_);
        shape_ = std::move(other.shape_);
        assert(!shape_.empty());
        other.storage_ = nullptr;
        other.shape_ = Shape({0});
    }

    explicit operator bool() const
    {
        return static_cast<bool>(storage_);
    }

    const Shape& shape() const
    {
        return shape_;
    }

    T* data() const
    {
        return storage_.get();
    }

private:
    Tensor(std::shared_ptr<T> storage, Shape shape, Device device)
        : device(device), shape_(std::move(shape)), storage_(std::move(storage))
    {
    }

    Shape shape_;
    std::shared_ptr<T> storage_;
};

using FloatTensor = Tensor<float>;

#endif  // CUDA_TENSOR_CUH
