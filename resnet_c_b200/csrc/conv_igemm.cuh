// conv_igemm.cuh — implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces, for every non-stem convolution of the network, the reference's chain
//   conv2dForwardKernel (/root/reference/cuda/ops.cu:14-48)      — direct conv, 1 thread / output
//   batchNorm2dForwardKernel (ops.cu:139-151)                    — folded into weights + bias
//   addForwardKernel (ops.cu:153-160)                            — residual, fused in the epilogue
//   reluForwardKernel (ops.cu:130-137)                           — fused in the epilogue
// as issued by layerForward (/root/reference/cuda/inference/main.cu:127-166).
//
// GEMM view (per launch):  D[M, N] = A[M, K] * B[N, K]^T
//   M = batch*OH*OW output pixels (NHWC rows), N = Cout, K = kh*kw*Cin (tap-major, Cin innermost)
//   A is never materialised: each 128-pixel x 128-byte K-slab is fetched by ONE im2col-mode TMA
//   (cuTensorMapEncodeIm2col) straight from the NHWC activation tensor, zero-filling the padding
//   halo and wrapping across rows/images in hardware.
//   B is the BN-folded weight matrix [Cout][kh][kw][Cin], fetched by a tiled TMA.
//   D accumulates in TMEM (fp32), double-buffered so the epilogue of tile i overlaps the MMAs of
//   tile i+1. The epilogue adds bias (+ residual, TMA-prefetched into the staging buffer), applies
//   ReLU, rounds to the activation type and leaves through a swizzled smem tile + TMA store.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):
//   warp 0        : TMA producer (A im2col + B tiled) over an NSTAGE ring — whole warp in the loop,
//   warp 1        : tcgen05.mma issuer                                      one elected lane issues
//   warp 2        : TMEM allocator / deallocator
//   warp 3        : store warp — TMA stores of finished tiles, residual prefetch, staging recycling
//   warps 4..11   : epilogue. Warps w and w+4 share TMEM lane quarter (w & 3) = 32 output pixels and
//                   split the tile's columns, thread t owns one pixel row of the staging tile.
// The epilogue is the critical path of every layer whose K loop is short (1x1 convs into wide
// tensors, i.e. most of the HBM traffic of the network): it never waits for a store or issues one —
// it only signals c_full[buf] and moves on; warp 3 does the rest.
#pragma once
#include <cuda_fp16.h>

#include "sm100_ptx.cuh"

namespace rnb {

struct ConvGeom {
    int M;                // batch * OH * OW
    int OH, OW;           // output spatial size
    int Cout;             // N of the GEMM
    int stride;           // conv stride (traversal stride of the im2col map)
    int lower;            // -padding: coordinate of the first filter tap for output pixel 0
    int ksize;            // square filter size (1 or 3)
    int kblocks_per_tap;  // Cin / BK
    int num_kblocks;      // ksize*ksize*kblocks_per_tap
    int m_tiles, n_tiles;
    int relu;     // apply max(x, 0)
    int has_res;  // add residual[M][Cout] before the ReLU
    // Tile traversal direction. Consecutive launches of the network alternate it, so that a kernel
    // starts with the part of its input the previous kernel wrote LAST — the part still in the
    // 126 MB L2 — instead of the part that has already been evicted to HBM.
    int reverse;
    // CTA-pair kernel only (conv_igemm2.cuh): tiles [0, split_from) are computed whole, tiles [split_from, tiles)
    // as two independent N halves each (work units of half the duration) so that the last, partially filled
    // wave of a persistent grid costs half a tile time. split_from is a multiple of the pair count (or 0);
    // split_from == m_tiles * n_tiles: no split.
    int split_from;
    // FP8 (E4M3) path only, ESZ == 1. Real value = stored value x scale. With the per-tensor
    // activation scales s_in, s_res, s_out fixed by calibration and one weight scale per output channel w[c]:
    //   q_out = RN_e4m3( relu( acc * chan_scale[c] + bias[c] + q_res * res_mul ) )
    //   chan_scale[c] = w[c] * s_in / s_out, bias[c] = shift[c] / s_out (the `bias` kernel argument), res_mul = s_res / s_out
    // — the per-channel vectors are pre-multiplied on the host side (conv_plan.cu::fp8_premultiply) so that the epilogue
    // spends one FMA per output. amax != nullptr: calibration launch (s_out = 1) — the kernel also max-reduces the
    // value before rounding (after the ReLU, rows < M only) into *amax (non-negative float bit pattern, atomicMax).
    const float* chan_scale;
    float res_mul;
    float* amax;
};

template <int BN_, int ESZ_, int NSTAGE_, int NCBUF_, int OSZ_ = ESZ_, int EPI_WARPS_ = 8>
struct ConvCfg {
    static constexpr int BM = 128;
    static constexpr int BN = BN_;
    static constexpr int ESZ = ESZ_;                 // bytes per activation/weight element (2 = bf16, 4 = tf32)
    static constexpr int OSZ = OSZ_;                 // bytes per OUTPUT element; OSZ != ESZ: unrounded FP32 out (FC)
    static constexpr int BK = 128 / ESZ_;            // one 128-byte swizzle row of K per stage
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr int NCBUF = NCBUF_;             // epilogue staging buffers (>= 2)
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_BYTES = BN_ * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BOX_COLS = 128 / OSZ_;      // output columns per 128-byte staging row
    static constexpr int NBOX = BN_ / BOX_COLS;
    static_assert(BN_ % BOX_COLS == 0, "tile N must be a whole number of staging boxes (FP8: BN >= 128)");
    static constexpr int BOX_BYTES = BM * 128;
    static constexpr int CBUF_BYTES = NBOX * BOX_BYTES;
    static constexpr int TMEM_COLS = 2 * BN_;        // two accumulator stages
    static constexpr int NBAR = 2 * NSTAGE_ + 4 + 3 * NCBUF_;
    static constexpr int SMEM_BYTES =
        1024 /*align slack*/ + NSTAGE_ * STAGE_BYTES + NCBUF_ * CBUF_BYTES + NBAR * 8 + 16;
    // epilogue warps: 8 (two per TMEM lane quarter, each takes every second 32-column chunk) or 16 (four per quarter:
    // a 128-column tile is ONE chunk per warp — for layers whose time is the epilogue's instruction stream)
    static constexpr int EPI_WARPS = EPI_WARPS_;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
    static_assert(EPI_WARPS_ == 8 || EPI_WARPS_ == 16, "8 or 16 epilogue warps");
    static_assert((BN_ / 32) % (EPI_WARPS_ / 4) == 0, "the warps of a lane quarter split the 32-column chunks evenly");
    static_assert(NCBUF_ >= 2, "need at least two staging buffers");
    static_assert((BN_ / 32) % 2 == 0, "the two warps of a lane quarter split the 32-column chunks");
    static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512,
                  "TMEM allocation must be a power of two >= 32 columns");
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// relu + round-to-nearest BF16 of two floats in one instruction
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// (a0, a1) += (b0, b1) as ONE packed FP32 instruction (FADD2, sm_100): IEEE round-to-nearest per lane,
// i.e. bit-identical to two FADDs at half the issue slots — the epilogues are issue-bound
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
    asm("{\n\t"
        ".reg .b64 ra, rb;\n\t"
        "mov.b64 ra, {%0, %1};\n\t"
        "mov.b64 rb, {%2, %3};\n\t"
        "add.rn.f32x2 ra, ra, rb;\n\t"
        "mov.b64 {%0, %1}, ra;\n\t"
        "}"
        : "+f"(a0), "+f"(a1)
        : "f"(b0), "f"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// One 32-column chunk of the accumulator: + bias (+ residual read from the staging row) -> ReLU ->
// round to the activation type -> write back to the same swizzled staging row.
// All loads (residual from shared memory, bias) are issued before the first store so that their
// latencies overlap (the staging row is addressed in the shared window: the compiler cannot prove
// that a generic store does not alias the next generic load and would serialise them).
template <int ESZ, bool RELU>
__device__ __forceinline__ void epilogue_chunk_t(const uint32_t (&v)[32], uint8_t* row, uint32_t c16_base,
                                                 uint32_t swz, const float* __restrict__ bias32, int has_res) {
    constexpr bool relu = RELU;  // compile-time: a run-time flag made ptxas issue BOTH converts, predicated
    const uint32_t row_addr = ptx::smem_u32(row);
    if (ESZ == 2) {
        uint4 rr[4];
        if (has_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rr[j] = lds128(row_addr + (((c16_base + j) ^ swz) << 4));
        }
        float4 b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = __ldg(reinterpret_cast<const float4*>(bias32 + j * 4));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[j * 8 + e]);
            add2(x[0], x[1], b[2 * j].x, b[2 * j].y);
            add2(x[2], x[3], b[2 * j].z, b[2 * j].w);
            add2(x[4], x[5], b[2 * j + 1].x, b[2 * j + 1].y);
            add2(x[6], x[7], b[2 * j + 1].z, b[2 * j + 1].w);
            if (has_res) {
                add2(x[0], x[1], bf16_lo(rr[j].x), bf16_hi(rr[j].x));
                add2(x[2], x[3], bf16_lo(rr[j].y), bf16_hi(rr[j].y));
                add2(x[4], x[5], bf16_lo(rr[j].z), bf16_hi(rr[j].z));
                add2(x[6], x[7], bf16_lo(rr[j].w), bf16_hi(rr[j].w));
            }
            const uint32_t dst = row_addr + (((c16_base + j) ^ swz) << 4);
            if (relu)
                sts128(dst, pack_bf16x2_relu(x[0], x[1]), pack_bf16x2_relu(x[2], x[3]),
                       pack_bf16x2_relu(x[4], x[5]), pack_bf16x2_relu(x[6], x[7]));
            else
                sts128(dst, pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                       pack_bf16x2(x[6], x[7]));
        }
    } else {
        // 32 fp32 columns = 8 16-byte pieces of the row; residual loaded four pieces at a time
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint4 rr[4];
            if (has_res) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rr[j] = lds128(row_addr + (((c16_base + half * 4 + j) ^ swz) << 4));
            }
            float4 b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = __ldg(reinterpret_cast<const float4*>(bias32 + (half * 4 + j) * 4));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = __uint_as_float(v[(half * 4 + j) * 4 + e]);
                x[0] += b[j].x; x[1] += b[j].y; x[2] += b[j].z; x[3] += b[j].w;
                if (has_res) {
                    x[0] += __uint_as_float(rr[j].x); x[1] += __uint_as_float(rr[j].y);
                    x[2] += __uint_as_float(rr[j].z); x[3] += __uint_as_float(rr[j].w);
                }
                if (relu) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
                }
                sts128(row_addr + (((c16_base + half * 4 + j) ^ swz) << 4), __float_as_uint(round_tf32(x[0])),
                       __float_as_uint(round_tf32(x[1])), __float_as_uint(round_tf32(x[2])),
                       __float_as_uint(round_tf32(x[3])));
            }
        }
    }
}

// run-time ReLU flag -> one warp-uniform branch per chunk around two specialised bodies
template <int ESZ>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], uint8_t* row, uint32_t c16_base,
                                               uint32_t swz, const float* __restrict__ bias32,
                                               int has_res, int relu) {
    if (relu)
        epilogue_chunk_t<ESZ, true>(v, row, c16_base, swz, bias32, has_res);
    else
        epilogue_chunk_t<ESZ, false>(v, row, c16_base, swz, bias32, has_res);
}

// ---- FP8 (E4M3) epilogue
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    uint16_t lo, hi;  // cvt packs its FIRST source into the upper byte: the element with the lower address goes second
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
    return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}
__device__ __forceinline__ void unpack_e4m3x4(uint32_t v, float (&o)[4]) {
    uint32_t h01, h23;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h01) : "h"(static_cast<uint16_t>(v & 0xFFFFu)));
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h23) : "h"(static_cast<uint16_t>(v >> 16)));
    const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
    const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
    o[0] = f01.x; o[1] = f01.y; o[2] = f23.x; o[3] = f23.y;
}

__device__ __forceinline__ uint32_t pack_e4m3x4_relu(float a, float b, float c, float d) {
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.relu.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
    asm("cvt.rn.satfinite.relu.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
    return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}

// One 32-column chunk of the accumulator on the FP8 path: 32 output bytes = two 16-byte pieces of the staging row.
// scale32 / bias32 are PRE-MULTIPLIED per channel (ConvGeom): q = RN_e4m3(relu(acc * scale'[c] + bias'[c] + r * res'))
// — one FMA (two with a residual) and half a convert per output. `track`: calibration launch, returns max|q value
// before rounding| of the chunk (0 otherwise).
template <bool RELU, bool HAS_RES>
__device__ __forceinline__ float epilogue_chunk_fp8_t(const uint32_t (&v)[32], uint8_t* row, uint32_t c16_base,
                                                      uint32_t swz, const float* __restrict__ scale32,
                                                      const float* __restrict__ bias32, float res_mul, bool track) {
    constexpr bool relu = RELU, has_res = HAS_RES;
    const uint32_t row_addr = ptx::smem_u32(row);
    uint4 rr[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (has_res) {
#pragma unroll
        for (int j = 0; j < 2; ++j) rr[j] = lds128(row_addr + (((c16_base + j) ^ swz) << 4));
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        // per-channel vectors of this 16-column piece (loaded per piece: the 16-warp variant has 96 registers per thread)
        float4 sc[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sc[q] = __ldg(reinterpret_cast<const float4*>(scale32 + (j * 4 + q) * 4));
            b[q] = __ldg(reinterpret_cast<const float4*>(bias32 + (j * 4 + q) * 4));
        }
        const uint32_t rw[4] = {rr[j].x, rr[j].y, rr[j].z, rr[j].w};
        uint32_t ow[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int g4 = j * 4 + q;  // group of four columns
            float x[4];
            x[0] = fmaf(__uint_as_float(v[g4 * 4 + 0]), sc[q].x, b[q].x);
            x[1] = fmaf(__uint_as_float(v[g4 * 4 + 1]), sc[q].y, b[q].y);
            x[2] = fmaf(__uint_as_float(v[g4 * 4 + 2]), sc[q].z, b[q].z);
            x[3] = fmaf(__uint_as_float(v[g4 * 4 + 3]), sc[q].w, b[q].w);
            if (has_res) {
                float r[4];
                unpack_e4m3x4(rw[q], r);
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = fmaf(r[e], res_mul, x[e]);
            }
            if (track) {
#pragma unroll
                for (int e = 0; e < 4; ++e) amax = fmaxf(amax, relu ? x[e] : fabsf(x[e]));
            }
            ow[q] = relu ? pack_e4m3x4_relu(x[0], x[1], x[2], x[3]) : pack_e4m3x4(x[0], x[1], x[2], x[3]);
        }
        sts128(row_addr + (((c16_base + j) ^ swz) << 4), ow[0], ow[1], ow[2], ow[3]);
    }
    return amax;
}

__device__ __forceinline__ float epilogue_chunk_fp8(const uint32_t (&v)[32], uint8_t* row, uint32_t c16_base,
                                                    uint32_t swz, const float* __restrict__ scale32,
                                                    const float* __restrict__ bias32, float res_mul, int has_res,
                                                    int relu, bool track) {
    if (has_res)
        return relu ? epilogue_chunk_fp8_t<true, true>(v, row, c16_base, swz, scale32, bias32, res_mul, track)
                    : epilogue_chunk_fp8_t<false, true>(v, row, c16_base, swz, scale32, bias32, res_mul, track);
    return relu ? epilogue_chunk_fp8_t<true, false>(v, row, c16_base, swz, scale32, bias32, res_mul, track)
                : epilogue_chunk_fp8_t<false, false>(v, row, c16_base, swz, scale32, bias32, res_mul, track);
}

// calibration launches: fold this thread's maximum into *amax (warp reduce, one atomic per warp)
__device__ __forceinline__ void amax_commit(float* amax, float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(v));
}

// FP32-output variant (the FC layer: logits stay FP32): 32 columns = one full 128-byte staging row,
// + bias, optional ReLU, NO rounding, no residual.
__device__ __forceinline__ void epilogue_chunk_f32out(const uint32_t (&v)[32], uint8_t* row, uint32_t swz,
                                                      const float* __restrict__ bias32, int relu) {
    const uint32_t row_addr = ptx::smem_u32(row);
    float4 b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = __ldg(reinterpret_cast<const float4*>(bias32 + j * 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float x[4];
        x[0] = __uint_as_float(v[j * 4 + 0]) + b[j].x;
        x[1] = __uint_as_float(v[j * 4 + 1]) + b[j].y;
        x[2] = __uint_as_float(v[j * 4 + 2]) + b[j].z;
        x[3] = __uint_as_float(v[j * 4 + 3]) + b[j].w;
        if (relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
        }
        sts128(row_addr + ((static_cast<uint32_t>(j) ^ swz) << 4), __float_as_uint(x[0]), __float_as_uint(x[1]),
               __float_as_uint(x[2]), __float_as_uint(x[3]));
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ CUtensorMap tmRes, const float* __restrict__ bias,
                  const ConvGeom g) {
    using namespace ptx;
    constexpr int BN = Cfg::BN;
    constexpr int NSTAGE = Cfg::NSTAGE;
    constexpr int NCBUF = Cfg::NCBUF;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_stage = smem;
    uint8_t* smem_c = smem + NSTAGE * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + NCBUF * Cfg::CBUF_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + NSTAGE;
    uint64_t* tmem_full = bars + 2 * NSTAGE;
    uint64_t* tmem_empty = bars + 2 * NSTAGE + 2;
    uint64_t* res_full = bars + 2 * NSTAGE + 4;           // residual tile landed in staging buffer
    uint64_t* c_full = res_full + NCBUF;                  // staging buffer holds a finished tile
    uint64_t* c_free = c_full + NCBUF;                    // staging buffer may be overwritten
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = g.m_tiles * g.n_tiles;
    const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        if (g.has_res) tma_prefetch_desc(&tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], Cfg::EPI_WARPS);
        }
        for (int i = 0; i < NCBUF; ++i) {
            mbar_init(&res_full[i], 1);
            mbar_init(&c_full[i], Cfg::EPI_WARPS);
            mbar_init(&c_free[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // Everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail;
    // from here on we touch tensors it produced (or still reads), so wait for it to finish.
    griddep_launch_dependents();
    griddep_wait();

    // local tile index -> (m_blk, n_blk)
    auto tile_coords = [&](int it_local, int& m_blk, int& n_blk) {
        const int t = blockIdx.x + it_local * gridDim.x;
        const int tt = g.reverse ? num_tiles - 1 - t : t;
        n_blk = tt % g.n_tiles;
        m_blk = tt / g.n_tiles;
    };

    if (warp == 0) {
        // ===================================================== TMA producer
        int stage = 0;
        uint32_t phase = 0;
        const int ohw = g.OH * g.OW;
        for (int it = 0; it < my_tiles; ++it) {
            int m_blk, n_blk;
            tile_coords(it, m_blk, n_blk);
            const int m0 = m_blk * Cfg::BM;
            const int img = m0 / ohw;
            const int rem = m0 - img * ohw;
            const int p = rem / g.OW;
            const int q = rem - p * g.OW;
            const int w0 = g.lower + q * g.stride;
            const int h0 = g.lower + p * g.stride;
            int tap_r = 0, tap_s = 0, cblk = 0;
            for (int kb = 0; kb < g.num_kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem_stage + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_im2col_4d(sa, &tmA, &full_bar[stage], cblk * Cfg::BK, w0, h0, img,
                                       static_cast<uint16_t>(tap_s), static_cast<uint16_t>(tap_r));
                    tma_load_2d(sb, &tmB, &full_bar[stage], kb * Cfg::BK, n_blk * BN);
                }
                __syncwarp();
                if (++cblk == g.kblocks_per_tap) {
                    cblk = 0;
                    if (++tap_s == g.ksize) {
                        tap_s = 0;
                        ++tap_r;
                    }
                }
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // Descriptors are a per-kernel constant plus (stage offset + 32-byte K step) >> 4 in the
        // address field.
        constexpr uint32_t idesc = umma_instr_desc(
            Cfg::ESZ == 2 ? UMMA_FMT_BF16 : (Cfg::ESZ == 4 ? UMMA_FMT_TF32 : UMMA_FMT_E4M3), Cfg::BM, BN);
        const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_stage), 0, 1024, UMMA_LAYOUT_SW128);
        const uint64_t b_desc0 =
            umma_smem_desc(smem_u32(smem_stage) + Cfg::A_BYTES, 0, 1024, UMMA_LAYOUT_SW128);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            for (int kb = 0; kb < g.num_kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t soff = static_cast<uint64_t>((stage * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // 32 bytes of K per instruction (16 bf16 / 8 tf32) inside the 128-byte swizzle row
                        const uint64_t ad = a_desc0 + soff + static_cast<uint64_t>(k * 2);
                        const uint64_t bd = b_desc0 + soff + static_cast<uint64_t>(k * 2);
                        if (Cfg::ESZ == 2)
                            mma_f16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
                        else if (Cfg::ESZ == 4)
                            mma_tf32_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
                        else
                            mma_f8_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
                    }
                    tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (kb == g.num_kblocks - 1) tc_commit(&tmem_full[as]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == NSTAGE) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== store warp
        // tile it lives in staging buffer it % NCBUF. When its store has been read out of shared
        // memory the buffer is recycled: with a residual, the residual tile of tile it+NCBUF is
        // prefetched straight into it (NCBUF tiles ahead of the epilogue); without, c_free is raised.
        auto issue_residual = [&](int it_local) {
            int m_blk, n_blk;
            tile_coords(it_local, m_blk, n_blk);
            const int cs = it_local % NCBUF;
            uint8_t* buf = smem_c + cs * Cfg::CBUF_BYTES;
            mbar_expect_tx(&res_full[cs], Cfg::CBUF_BYTES);
#pragma unroll
            for (int b = 0; b < Cfg::NBOX; ++b)
                tma_load_2d(buf + b * Cfg::BOX_BYTES, &tmRes, &res_full[cs],
                            n_blk * BN + b * Cfg::BOX_COLS, m_blk * Cfg::BM);
        };
        if (g.has_res && elect_one()) {
            for (int i = 0; i < NCBUF && i < my_tiles; ++i) issue_residual(i);
        }
        __syncwarp();
        for (int it = 0; it < my_tiles; ++it) {
            const int cs = it % NCBUF;
            mbar_wait(&c_full[cs], (it / NCBUF) & 1);
            if (elect_one()) {
                int m_blk, n_blk;
                tile_coords(it, m_blk, n_blk);
                const uint8_t* cbuf = smem_c + cs * Cfg::CBUF_BYTES;
#pragma unroll
                for (int b = 0; b < Cfg::NBOX; ++b)
                    tma_store_2d(&tmOut, cbuf + b * Cfg::BOX_BYTES, n_blk * BN + b * Cfg::BOX_COLS,
                                 m_blk * Cfg::BM);
                tma_store_commit();
                tma_store_wait_read<0>();  // smem of this buffer has been read: recycle it
                if (it + NCBUF < my_tiles) {
                    if (g.has_res)
                        issue_residual(it + NCBUF);
                    else
                        mbar_arrive(&c_free[cs]);
                }
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================== epilogue
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int h = (warp - 4) >> 2;                // which half of the 32-column chunks
        const int row_in_tile = q * 32 + lane;        // TMEM lane == pixel row of the tile
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        for (int it = 0; it < my_tiles; ++it) {
            int m_blk, n_blk;
            tile_coords(it, m_blk, n_blk);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int cs = it % NCBUF;
            uint8_t* cbuf = smem_c + cs * Cfg::CBUF_BYTES;
            if (g.has_res)
                mbar_wait(&res_full[cs], (it / NCBUF) & 1);
            else if (it >= NCBUF)
                mbar_wait(&c_free[cs], ((it / NCBUF) - 1) & 1);
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            const float* bias_n = bias + n_blk * BN;
            float amax = 0.f;
#pragma unroll 1
            for (int chunk = h; chunk < BN / 32; chunk += Cfg::EPI_WARPS / 4) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(taddr + chunk * 32, v);
                tmem_ld_wait();
                const int byte_off = chunk * 32 * Cfg::OSZ;
                uint8_t* row = cbuf + (byte_off >> 7) * Cfg::BOX_BYTES + row_in_tile * 128;
                if (Cfg::OSZ == 1)
                    amax = fmaxf(amax, epilogue_chunk_fp8(v, row, (byte_off & 127) >> 4, swz,
                                                          g.chan_scale + n_blk * BN + chunk * 32, bias_n + chunk * 32,
                                                          g.res_mul, g.has_res, g.relu, g.amax != nullptr));
                else if (Cfg::OSZ != Cfg::ESZ)
                    epilogue_chunk_f32out(v, row, swz, bias_n + chunk * 32, g.relu);
                else
                    epilogue_chunk<Cfg::OSZ == 1 ? 2 : Cfg::ESZ>(v, row, (byte_off & 127) >> 4, swz, bias_n + chunk * 32,
                                                                 g.has_res, g.relu);
            }
            if (Cfg::OSZ == 1 && g.amax) amax_commit(g.amax, m_blk * Cfg::BM + row_in_tile < g.M ? amax : 0.f);
            // accumulator drained by this warp: hand the TMEM stage back; publish the staged rows
            // to the async proxy and tell the store warp
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tmem_empty[as]);
                mbar_arrive(&c_full[cs]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace rnb
