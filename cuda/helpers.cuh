// helpers.cuh — drop-in for olehskip/resnet.c cuda/helpers.cuh (:6-35): CEIL, gpuErrchk / gpuAssert
// (print + abort on a CUDA error) and safeCudaMalloc (checked cudaMalloc, allocation log under
// -DDEBUG). Same names and behaviour, so code written against the reference header builds unchanged.
#ifndef CUDA_HELPERS_CUH
#define CUDA_HELPERS_CUH

#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <iostream>

#define CEIL(a, b) (((a) + (b)-1) / (b))

inline void gpuAssert(cudaError_t status, const char* file, int line, bool abort = true)
{
    if (status == cudaSuccess) {
        return;
    }
    std::cerr << "GPUassert: " << cudaGetErrorString(status) << " " << file << " " << line << "\n";
    if (abort) {
        std::abort();
    }
}

#define gpuErrchk(ans)                          \
    do {                                        \
        gpuAssert((ans), __FILE__, __LINE__);   \
    } while (0)

inline void* safeCudaMalloc(uint64_t size)
{
    void* ptr = nullptr;
    gpuErrchk(cudaMalloc(&ptr, size));
#ifdef DEBUG
    static uint64_t running_total = 0;
    running_total += size;
    std::cerr << "GPU allocate ptr: " << ptr << ". Size: " << size << " bytes. Total: " << running_total
              << " bytes" << std::endl;
#endif
    return ptr;
}

#endif  // CUDA_HELPERS_CUH
