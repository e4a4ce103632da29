// conv3x3_halo.cuh — 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels, BF16, with the input halo
// tile kept in shared memory and the weights resident (the layer1 conv2 of every Bottleneck ResNet
// and layer1 conv1 of the BasicBlock ones; replaces the same reference chain as conv_igemm.cuh:
// conv2dForwardKernel + batchNorm2dForwardKernel + reluForwardKernel, ops.cu:14-48,139-151,130-137).
//
// Why a special kernel: with the generic im2col kernel this layer re-reads its activation tile from
// L2 once per filter tap (9 x 16 KB per 128 x 64 output tile) plus the weights (9 x 8 KB): the
// layer runs at the L2 -> SM fabric limit (~13 TB/s), 2.5x above its HBM time. Here
//   * the 72 KB weight matrix is loaded ONCE per CTA and stays in shared memory;
//   * an output tile is 2 image rows (x 64-pixel pitch = 128 GEMM rows, 112 valid for W = 56); for
//     each filter ROW r one tiled TMA fetches the input rows (h0+r-1, h0+r) x pixels [-1, 63) with
//     hardware zero fill of the halo — 3 loads per tile instead of 9;
//   * the three taps s = 0,1,2 of a filter row are the SAME shared-memory tile read through UMMA
//     descriptors whose start address is shifted by s x 128 bytes (one pixel): the 128-byte swizzle
//     is address-based, so a row-shifted view of a TMA-written tile is still consistent.
// Rows w >= W of the 64-pixel pitch are garbage and are never stored (the TMA store box is W wide).
// Roles / epilogue / store warp are those of conv_igemm.cuh.
#pragma once
#include "conv_igemm.cuh"

namespace rnb {

struct HaloGeom {
    int N, H, W;       // images, spatial size (W <= 62, H even)
    int tiles;         // N * H / 2
    int rows_per_img;  // H / 2 tiles per image
    int relu;
    int reverse;
};

struct HaloCfg {
    static constexpr int BM = 128, BN = 64, PITCH = 64;
    static constexpr int NSLOT = 6;                      // A ring: one slot per (tile, filter row)
    static constexpr int A_SLOT_BYTES = 17 * 1024;       // 16 KB tile + read-past for the shifted views
    static constexpr int A_TX_BYTES = 2 * PITCH * 128;   // bytes one TMA load delivers
    static constexpr int B_TAP_BYTES = 64 * 128;         // one tap: 64 out-channels x 64 in-channels
    static constexpr int B_BYTES = 9 * B_TAP_BYTES;      // 72 KB resident
    static constexpr int NCBUF = 3;
    static constexpr int CBUF_BYTES = BM * 128;          // 128 rows x 64 bf16
    static constexpr int TMEM_COLS = 128;                // two 64-column accumulator stages
    static constexpr int NBAR = 2 * NSLOT + 4 + 2 * NCBUF + 1;
    static constexpr int SMEM_BYTES = 1024 + B_BYTES + NSLOT * A_SLOT_BYTES + NCBUF * CBUF_BYTES + NBAR * 8 + 16;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
};
static_assert(HaloCfg::SMEM_BYTES <= 232448, "smem budget");

namespace ptx {
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
          "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int32_t c0,
                                             int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        :
        : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
}  // namespace ptx

// tmA: 4-D tiled map over the input  [C=64, W, H, N], box {64, 64, 2, 1}, 128B swizzle
// tmB: 2-D tiled map over the weights [64][9*64] (tap-major K), box {64, 64}
// tmOut: 4-D tiled map over the output [C=64, W, H, N], box {64, W, 1, 1}
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ bias,
                    const HaloGeom g) {
    using namespace ptx;
    constexpr int NSLOT = Cfg::NSLOT, NCBUF = Cfg::NCBUF;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_b = smem;
    uint8_t* smem_a = smem_b + Cfg::B_BYTES;
    uint8_t* smem_c = smem_a + NSLOT * Cfg::A_SLOT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + NCBUF * Cfg::CBUF_BYTES);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + NSLOT;
    uint64_t* tmem_full = bars + 2 * NSLOT;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* c_full = tmem_empty + 2;
    uint64_t* c_free = c_full + NCBUF;
    uint64_t* b_full = c_free + NCBUF;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int my_tiles = (g.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], Cfg::EPI_WARPS);
        }
        for (int i = 0; i < NCBUF; ++i) {
            mbar_init(&c_full[i], Cfg::EPI_WARPS);
            mbar_init(&c_free[i], 1);
        }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    // the read-past pad of every A slot is read by the shifted views of garbage rows only, but keep
    // it finite (NaN x 0 weights never happens, yet uninitialised smem could hold NaN patterns)
    for (int i = threadIdx.x; i < NSLOT * 64; i += Cfg::THREADS)
        reinterpret_cast<uint4*>(smem_a + (i / 64) * Cfg::A_SLOT_BYTES + Cfg::A_TX_BYTES)[i % 64] =
            make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();  // see conv_igemm.cuh
    griddep_wait();

    auto tile_coords = [&](int it_local, int& img, int& h0) {
        const int t = blockIdx.x + it_local * gridDim.x;
        const int tt = g.reverse ? g.tiles - 1 - t : t;
        img = tt / g.rows_per_img;
        h0 = (tt - img * g.rows_per_img) * 2;
    };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            mbar_expect_tx(b_full, Cfg::B_BYTES);
            for (int tap = 0; tap < 9; ++tap)
                tma_load_2d(smem_b + tap * Cfg::B_TAP_BYTES, &tmB, b_full, tap * 64, 0);
        }
        __syncwarp();
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            int img, h0;
            tile_coords(it, img, h0);
            for (int r = 0; r < 3; ++r) {
                mbar_wait(&a_empty[slot], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[slot], Cfg::A_TX_BYTES);
                    // pixels [-1, 63) of input rows h0+r-1 and h0+r; out-of-image parts are zero-filled
                    tma_load_4d(smem_a + slot * Cfg::A_SLOT_BYTES, &tmA, &a_full[slot], 0, -1, h0 + r - 1, img);
                }
                __syncwarp();
                if (++slot == NSLOT) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, Cfg::BM, Cfg::BN);
        const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), 0, 1024, UMMA_LAYOUT_SW128);
        const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), 0, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(b_full, 0);
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * Cfg::BN;
            for (int r = 0; r < 3; ++r) {
                mbar_wait(&a_full[slot], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_slot = a_desc0 + static_cast<uint64_t>((slot * Cfg::A_SLOT_BYTES) >> 4);
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        // tap (r, s): the same tile, start address shifted by s pixels (s x 128 B). The
                        // 128-byte swizzle is a function of the shared-memory ADDRESS bits, so the shifted
                        // view reads exactly what TMA wrote; the descriptor's base_offset stays 0
                        // (measured: setting it to s breaks the result).
                        const uint64_t a_tap = a_slot + static_cast<uint64_t>((s * 128) >> 4);
                        const uint64_t b_tap = b_desc0 + static_cast<uint64_t>(((r * 3 + s) * Cfg::B_TAP_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss(d_tmem, a_tap + static_cast<uint64_t>(k * 2),
                                       b_tap + static_cast<uint64_t>(k * 2), idesc, (r | s | k) != 0);
                    }
                    tc_commit(&a_empty[slot]);
                    if (r == 2) tc_commit(&tmem_full[as]);
                }
                __syncwarp();
                if (++slot == NSLOT) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== store warp
        for (int it = 0; it < my_tiles; ++it) {
            const int cs = it % NCBUF;
            mbar_wait(&c_full[cs], (it / NCBUF) & 1);
            if (elect_one()) {
                int img, h0;
                tile_coords(it, img, h0);
                const uint8_t* cbuf = smem_c + cs * Cfg::CBUF_BYTES;
                tma_store_4d(&tmOut, cbuf, 0, 0, h0, img);                              // GEMM rows 0..W-1
                tma_store_4d(&tmOut, cbuf + Cfg::PITCH * 128, 0, 0, h0 + 1, img);       // rows 64..64+W-1
                tma_store_commit();
                tma_store_wait_read<0>();
                if (it + NCBUF < my_tiles) mbar_arrive(&c_free[cs]);
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================== epilogue
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;             // 32-column chunk handled by this warp
        const int row_in_tile = q * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            const int cs = it % NCBUF;
            uint8_t* cbuf = smem_c + cs * Cfg::CBUF_BYTES;
            if (it >= NCBUF) mbar_wait(&c_free[cs], ((it / NCBUF) - 1) & 1);
            mbar_wait(&tmem_full[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::BN;
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(taddr + h * 32, v);
            tmem_ld_wait();
            epilogue_chunk<2>(v, cbuf + row_in_tile * 128, static_cast<uint32_t>(h * 4), swz, bias + h * 32, 0,
                              g.relu);
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
