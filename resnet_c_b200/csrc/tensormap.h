// tensormap.h — host-side construction of the TMA descriptors used by the conv kernels.
// The driver entry points are resolved at run time (cudaGetDriverEntryPoint), so librnb.so only
// links the CUDA runtime.
#pragma once
#include <cstdint>
#include <cuda.h>

namespace rnb {

enum class TmDtype { BF16, F32, U8 };  // U8: FP8 (E4M3) tensors travel as bytes

// Returns 0 on success, otherwise a CUresult (or -1 if the driver entry points are unavailable).
// All maps use 128-byte swizzle; the inner box extent is always 128 bytes.

// 2-D row-major matrix [rows][cols] of `dtype`; box = (128 bytes of cols) x box_rows.
int make_tiled_2d(CUtensorMap* out, TmDtype dtype, const void* base, uint64_t rows, uint64_t cols,
                  uint32_t box_rows);

// NHWC activation tensor [N][H][W][C] viewed through an im2col window: `channels` (= 128 bytes)
// per pixel, `pixels` output pixels per load, filter `ksize` x `ksize`, `stride`, `pad`.
int make_im2col_nhwc(CUtensorMap* out, TmDtype dtype, const void* base, uint64_t N, uint64_t H,
                     uint64_t W, uint64_t C, int ksize, int stride, int pad, uint32_t pixels);

// Work-around for im2col descriptors of tensors smaller than 128 KiB (see tensormap.cu): -1 = decide by driver
// version (default until rnb_init()'s self-test has run), 0 = off, 1 = on. im2col_small_patch() = what is in force.
void set_im2col_small_patch(int mode);
int im2col_small_patch();

// Generic tiled map (up to 5-D, no swizzle option) used by the stem.
int make_tiled_nd(CUtensorMap* out, TmDtype dtype, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, bool swizzle128);

}  // namespace rnb
