// tensormap.cu — see tensormap.h
#include "tensormap.h"

#include <cuda_runtime.h>
#include <mutex>

namespace rnb {

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                    cuuint32_t, cuuint32_t, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_tiled = nullptr;
EncodeIm2colFn g_im2col = nullptr;
int g_driver_version = 0;
int g_small_patch_mode = -1;  // see set_im2col_small_patch()
std::once_flag g_once;

void resolve() {
    std::call_once(g_once, [] {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_tiled = reinterpret_cast<EncodeTiledFn>(fn);
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
        cudaDriverGetVersion(&g_driver_version);
    });
}

CUtensorMapDataType to_cu(TmDtype d) {
    return d == TmDtype::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                              : (d == TmDtype::U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}
uint32_t esize(TmDtype d) { return d == TmDtype::BF16 ? 2u : (d == TmDtype::U8 ? 1u : 4u); }

}  // namespace

int make_tiled_2d(CUtensorMap* out, TmDtype dtype, const void* base, uint64_t rows, uint64_t cols,
                  uint32_t box_rows) {
    resolve();
    if (!g_tiled) return -1;
    const uint32_t es = esize(dtype);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * es};
    cuuint32_t box[2] = {128u / es, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return static_cast<int>(g_tiled(out, to_cu(dtype), 2, const_cast<void*>(base), dims, strides,
                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
}

int make_im2col_nhwc(CUtensorMap* out, TmDtype dtype, const void* base, uint64_t N, uint64_t H,
                     uint64_t W, uint64_t C, int ksize, int stride, int pad, uint32_t pixels) {
    resolve();
    if (!g_im2col) return -1;
    const uint32_t es = esize(dtype);
    cuuint64_t dims[4] = {C, W, H, N};
    cuuint64_t strides[3] = {C * es, W * C * es, H * W * C * es};
    // Bounding box of filter-origin positions: [-pad, dim + pad - ksize] per spatial dim, expressed
    // as offsets from the tensor's lower / upper corners.
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
    CUresult r = g_im2col(out, to_cu(dtype), 4, const_cast<void*>(base), dims, strides, lower,
                          upper, 128u / es, pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return static_cast<int>(r);
    // Drivers up to 13.1 encode im2col descriptors of tensors smaller than 128 KiB with a flag the
    // hardware then mis-handles; clearing bit 21 of the second descriptor word restores the
    // documented behaviour (same work-around as CUTLASS' make_im2col_tma_copy_desc).
    // The version cut-off is only the first guess: rnb_init() runs a sub-128-KiB im2col convolution against the FP32
    // kernel and switches the work-around on or off according to what the hardware actually does (api.cu).
    const bool patch = g_small_patch_mode < 0 ? g_driver_version <= 13010 : g_small_patch_mode != 0;
    if (patch && N * H * W * C * es < 131072) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
    return 0;
}

void set_im2col_small_patch(int mode) { g_small_patch_mode = mode; }
int im2col_small_patch() {
    resolve();
    return g_small_patch_mode < 0 ? (g_driver_version <= 13010 ? 1 : 0) : (g_small_patch_mode != 0 ? 1 : 0);
}

int make_tiled_nd(CUtensorMap* out, TmDtype dtype, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
    resolve();
    if (!g_tiled) return -1;
    cuuint64_t d[5];
    cuuint64_t s[4];
    cuuint32_t b[5];
    cuuint32_t e[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = 1;
        if (i < rank - 1) s[i] = strides_bytes[i];
    }
    return static_cast<int>(g_tiled(
        out, to_cu(dtype), static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, s, b, e,
        CU_TENSOR_MAP_INTERLEAVE_NONE,
        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
}

}  // namespace rnb
