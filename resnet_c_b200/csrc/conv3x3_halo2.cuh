// conv3x3_halo2.cuh — 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels, TF32 (FP32 storage), as a CTA PAIR
// with resident weights and a halo-resident input ring: layer1 of the BasicBlock networks on the TF32 path
// (ResNet-18/34; BASELINE configs[1]). Same reference chain as conv_igemm.cuh: conv2dForwardKernel +
// batchNorm2dForwardKernel (+ addForwardKernel) + reluForwardKernel, /root/reference/cuda/ops.cu:14-48,139-151,
// 153-160,130-137.
//
// Why: with the generic im2col kernel a 128 x 64 TF32 tile pulls 18 K-blocks x 24 KB through L2 for 4.7 M MACs
// (about 190 B/clk per SM at the TF32 rate) — the layer runs at the L2 -> SM fabric limit, 200 us per launch at
// 294 TFLOP/s. Here (cf. conv3x3_halo.cuh for BF16, bneck_l1.cuh for the pair structure)
//   * the 147 KB weight matrix is loaded ONCE, split across the pair (72 KB per CTA, 32 output channels each);
//   * a tile is 4 image rows x 64-pixel pitch (2 rows per CTA); per filter ROW one TMA load per 32-channel K block
//     fetches the two input rows it needs (zero halo by hardware OOB fill) into a 3-slot ring;
//   * the three taps of a filter row are the same slot read through UMMA descriptors shifted by one pixel (128 B).
// K order (tap row, tap column, channel block) and rounding are those of the generic kernel: results are
// bit-identical to it. Epilogue: bias (+ residual, TMA-prefetched into the staging tile) -> ReLU -> cvt.rna.tf32
// -> swizzled staging -> TMA store (boxes W pixels wide: the garbage rows w >= W are never stored).
// Warps (384 threads): 0 TMA producer, 1 MMA issuer (leader CTA), 2 TMEM alloc, 3 store warp, 4..11 epilogue.
#pragma once
#include "bneck_l1.cuh"

namespace rnb {

struct Halo2Geom {
    int N, H, W;        // images, spatial size (8 <= W <= 62, H % 4 == 0)
    int tiles;          // N * H / 4 pair tiles
    int tiles_per_img;
    int relu, has_res;
    int reverse;
};

struct Halo2Cfg {
    static constexpr int PITCH = 64;
    static constexpr int NSLOT = 3;
    static constexpr int KB_BYTES = 2 * PITCH * 128;     // one 32-channel K block of a slot: 2 rows x 64 px x 128 B
    static constexpr int SLOT_BYTES = 2 * KB_BYTES;      // 32 KB
    static constexpr int RING_BYTES = NSLOT * SLOT_BYTES + 1024;
    static constexpr int W_BLK_BYTES = 32 * 128;         // this CTA's 32 output channels x one K block of one tap
    static constexpr int W_BYTES = 18 * W_BLK_BYTES;     // 72 KB
    static constexpr int BOX_BYTES = 128 * 128;          // 128 rows x 32 fp32
    static constexpr int STAGE_BYTES = 2 * BOX_BYTES;    // 64 output channels
    static constexpr int TMEM_COLS = 128;                // two 64-column accumulator stages
    static constexpr int NBAR = 1 + 2 * NSLOT + 4 + 2;
    static constexpr int SMEM_BYTES = 1024 + W_BYTES + RING_BYTES + STAGE_BYTES + NBAR * 8 + 16;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
};
static_assert(Halo2Cfg::SMEM_BYTES <= 232448, "smem budget");

// tmA  : input  [C=64 fp32, W, H, N], box {32, 64, 2, 1}  (loaded at w = -1)
// tmB  : weights [64][576] fp32 (tap-major K, TF32-rounded), box {32, 32}
// tmRes: residual [C=64, W, H, N], box {32, 64, 2, 1}      (has_res; else unused)
// tmOut: output [C=64, W, H, N], box {32, W, 1, 1}
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
conv3x3_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOut,
                     const float* __restrict__ bias, const Halo2Geom g) {
    using namespace ptx;
    constexpr int NSLOT = Cfg::NSLOT;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_w = smem;
    uint8_t* smem_ring = smem_w + Cfg::W_BYTES;
    uint8_t* smem_stage = smem_ring + Cfg::RING_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::STAGE_BYTES);
    uint64_t* w_full = bars;                // leader
    uint64_t* a_full = w_full + 1;          // leader, [NSLOT]
    uint64_t* a_empty = a_full + NSLOT;     // per CTA (multicast commit), [NSLOT]
    uint64_t* d_full = a_empty + NSLOT;     // per CTA (multicast commit), [2]
    uint64_t* d_empty = d_full + 2;         // leader, 16 arrivals, [2]
    uint64_t* stage_ready = d_empty + 2;    // per CTA: staging tile free (or the residual tile has landed in it)
    uint64_t* c_full = stage_ready + 1;     // per CTA, 8 arrivals: staging tile holds the finished output
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int T = (g.tiles - pair + num_pairs - 1) / num_pairs;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        if (g.has_res) tma_prefetch_desc(&tmRes);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_empty[i], 2 * Cfg::EPI_WARPS);
        }
        mbar_init(stage_ready, 1);
        mbar_init(c_full, Cfg::EPI_WARPS);
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    for (int i = threadIdx.x; i < 64; i += Cfg::THREADS)  // read-past pad of the ring (garbage rows only)
        reinterpret_cast<uint4*>(smem_ring + NSLOT * Cfg::SLOT_BYTES)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();  // see conv_igemm.cuh
    griddep_wait();

    auto tile_coords = [&](int it_local, int& img, int& h0) {
        const int t = pair + it_local * num_pairs;
        const int tt = g.reverse ? g.tiles - 1 - t : t;
        img = tt / g.tiles_per_img;
        h0 = (tt - img * g.tiles_per_img) * 4 + static_cast<int>(rank) * 2;
    };

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(w_full, 2 * Cfg::W_BYTES);
            for (int kbi = 0; kbi < 18; ++kbi)
                tma_load_2d_2sm(smem_w + kbi * Cfg::W_BLK_BYTES, &tmB, w_full, kbi * 32, static_cast<int>(rank) * 32);
        }
        __syncwarp();
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; it < T; ++it) {
            int img, h0;
            tile_coords(it, img, h0);
            for (int r = 0; r < 3; ++r) {
                mbar_wait(&a_empty[slot], phase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&a_full[slot], 2 * Cfg::SLOT_BYTES);
                    uint8_t* dst = smem_ring + slot * Cfg::SLOT_BYTES;
                    // pixels [-1, 63) of input rows h0+r-1, h0+r, one load per 32-channel K block
                    tma_load_4d_2sm(dst, &tmA, &a_full[slot], 0, -1, h0 + r - 1, img);
                    tma_load_4d_2sm(dst + Cfg::KB_BYTES, &tmA, &a_full[slot], 32, -1, h0 + r - 1, img);
                }
                __syncwarp();
                if (++slot == NSLOT) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_TF32, 256, 64);
            const uint64_t ring_desc = umma_smem_desc(smem_u32(smem_ring), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t w_desc = umma_smem_desc(smem_u32(smem_w), 0, 1024, UMMA_LAYOUT_SW128);
            mbar_wait(w_full, 0);
            tc_fence_after();
            int slot = 0;
            uint32_t phase = 0;
            for (int i = 0; i < T; ++i) {
                const int buf = i & 1;
                mbar_wait(&d_empty[buf], ((i >> 1) & 1) ^ 1);
                const uint32_t d_tmem = tmem_base + buf * 64;
                for (int r = 0; r < 3; ++r) {
                    mbar_wait(&a_full[slot], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_slot = ring_desc + static_cast<uint64_t>((slot * Cfg::SLOT_BYTES) >> 4);
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb) {
                                // tap (r, s), channel block kb: the slot's K-block slab shifted by s pixels
                                const uint64_t a = a_slot + static_cast<uint64_t>((kb * Cfg::KB_BYTES + s * 128) >> 4);
                                const uint64_t b = w_desc + static_cast<uint64_t>((((r * 3 + s) * 2 + kb) * Cfg::W_BLK_BYTES) >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k)  // 8 TF32 of K (32 bytes) per instruction
                                    mma_tf32_ss_2sm(d_tmem, a + static_cast<uint64_t>(k * 2), b + static_cast<uint64_t>(k * 2),
                                                    idesc, (r | s | kb | k) != 0);
                            }
                        }
                        tc_commit_2sm(&a_empty[slot]);
                        if (r == 2) tc_commit_2sm(&d_full[buf]);
                    }
                    __syncwarp();
                    if (++slot == NSLOT) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== store warp (both CTAs)
        // one staging tile: before tile t may be written the store of tile t-1 must have left shared memory;
        // with a residual, the residual tile of t is then TMA-loaded INTO the staging tile (updated in place)
        auto make_ready = [&](int it_local) {
            if (g.has_res) {
                int img, h0;
                tile_coords(it_local, img, h0);
                mbar_expect_tx(stage_ready, Cfg::STAGE_BYTES);
                tma_load_4d(smem_stage, &tmRes, stage_ready, 0, 0, h0, img);
                tma_load_4d(smem_stage + Cfg::BOX_BYTES, &tmRes, stage_ready, 32, 0, h0, img);
            } else {
                mbar_arrive(stage_ready);
            }
        };
        if (T > 0 && elect_one()) make_ready(0);
        __syncwarp();
        for (int it = 0; it < T; ++it) {
            mbar_wait(c_full, it & 1);
            if (elect_one()) {
                int img, h0;
                tile_coords(it, img, h0);
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    const uint8_t* box = smem_stage + hb * Cfg::BOX_BYTES;
                    tma_store_4d(&tmOut, box, hb * 32, 0, h0, img);
                    tma_store_4d(&tmOut, box + Cfg::PITCH * 128, hb * 32, 0, h0 + 1, img);
                }
                tma_store_commit();
                tma_store_wait_read<0>();
                if (it + 1 < T) make_ready(it + 1);
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================== epilogue (both CTAs)
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;  // 32-channel half = staging box
        const int row_in_tile = q * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int i = 0; i < T; ++i) {
            const int buf = i & 1;
            mbar_wait(stage_ready, i & 1);
            mbar_wait(&d_full[buf], (i >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(lane_base + buf * 64 + h * 32, v);
            tmem_ld_wait();
            epilogue_chunk<4>(v, smem_stage + h * Cfg::BOX_BYTES + row_in_tile * 128, 0u, swz, bias + h * 32, g.has_res,
                              g.relu);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_leader(&d_empty[buf]);
                mbar_arrive(c_full);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace rnb
