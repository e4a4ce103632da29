// bneck_c3n1.cuh — conv3 (1x1, 128 -> 512) + BN + shortcut + ReLU of a layer2-shaped Bottleneck block AND the
// next block's conv1 (1x1, 512 -> 128) + BN + ReLU in ONE launch (BF16):
//
//     y   = relu(bn3(conv3_1x1(t2)) + x)           t2 [M,128], x / y [M,512]     (layerForward, main.cu:153-163)
//     t1' = relu(bn1'(conv1'_1x1(y)))              t1' [M,128]                    (next block, main.cu:138-146)
//
// replacing two launches of the conv2dForwardKernel / batchNorm2dForwardKernel / addForwardKernel /
// reluForwardKernel chain (/root/reference/cuda/ops.cu:14-48,139-151,153-160,130-137). At 28x28 both are
// HBM-bound; fused, y is consumed from the staging tiles it is stored from, so the 206 MB (per 256 images)
// re-read of y by conv1' disappears: 51 + 206 + 206 + 51 MB instead of 463 + 257 MB.
//
// Structure (the machinery of bneck_l1.cuh without its 3x3 stage): a CTA PAIR (cta_group::2) per tile of
// 256 consecutive pixel rows (128 per CTA); 1x1 convs need no halo, so all tensors are plain 2-D [M][C] TMA
// maps. Both weight matrices are resident, split across the pair (64 + 64 KB per CTA).
//   conv3 : A2 tile (128 rows x 128 K, one TMA pair per tile) x W3_q^T for the four 128-channel quarters q of
//           y -> D2 (two 128-column halves, quarter q in half q & 1)
//   E2    : D2 + bias3 + residual (TMA-prefetched into the staging box) -> ReLU -> BF16 -> staging boxes
//           (64 channels each, 8 per tile) -> TMA store of y; each box is also the K block of
//   conv1': box_b x W1n_b^T accumulated over b = 0..7 -> D3 (128 columns, double-buffered)
//   E3    : D3 + bias1' -> ReLU -> BF16 -> two staging boxes -> TMA store of t1'
// Warps (384 threads): 0 TMA producer (weights once, A2 tiles), 1 conv3 issuer, 2 TMEM alloc + conv1'
// issuer, 3 store warp (pipelined stores, residual prefetch, box recycling), 4..11 epilogue.
// Rounding points and K order are those of the separate kernels: the result is bit-identical to them.
#pragma once
#include "bneck_l1.cuh"

namespace rnb {

// In-kernel timeline (tools/c3n1s_timeline.py builds a separate library with -DRNB_TIMELINE; the product build
// compiles all of this away): one thread of each of four warp roles of CTA 0 records clock64() at its hand-off points
// while it works on its SECOND tile, into thread-local storage, and dumps it to global memory when the kernel ends.
#ifdef RNB_TIMELINE
constexpr int kTlMax = 160;
__device__ long long g_c3n1s_tl[4][kTlMax];
__device__ int g_c3n1s_tl_n[4];
#define TL_DECL(role_)                                                                                  \
    long long tl_buf[kTlMax];                                                                           \
    int tl_n = 0;                                                                                       \
    const int tl_role = (role_);                                                                        \
    const bool tl_me = blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ((role_) != 0 || threadIdx.x == 128); \
    int tl_tile = -1;
#define TL_TILE(i_) tl_tile = (i_);
#define TL(tag_)                                                                             \
    if (tl_me && tl_tile == 1 && tl_n < kTlMax) tl_buf[tl_n++] = (clock64() << 8) | (tag_);
#define TL_DUMP()                                                        \
    if (tl_me) {                                                         \
        for (int i_ = 0; i_ < tl_n; ++i_) g_c3n1s_tl[tl_role][i_] = tl_buf[i_]; \
        g_c3n1s_tl_n[tl_role] = tl_n;                                    \
    }
#else
#define TL_DECL(role_)
#define TL_TILE(i_)
#define TL(tag_)
#define TL_DUMP()
#endif

struct C3n1Geom {
    int M;        // pixel rows (batch * H * W)
    int tiles;    // ceil(M / 256) pair tiles
    int reverse;  // tile traversal direction (see ConvGeom::reverse)
};

struct C3n1Cfg {
    static constexpr int K3 = 128, N3 = 512, N1 = 128;        // conv3: K3 -> N3; conv1': N3 -> N1
    static constexpr int W3_BLK_BYTES = 64 * 128;             // this CTA's 64 rows of one (quarter, k-block)
    static constexpr int W3_BYTES = 4 * 2 * W3_BLK_BYTES;     // 64 KB
    static constexpr int W1N_BLK_BYTES = 64 * 128;            // this CTA's 64 rows of one 64-wide K block
    static constexpr int W1N_BYTES = 8 * W1N_BLK_BYTES;       // 64 KB
    static constexpr int BOX_BYTES = 16384;                   // 128 rows x 64 bf16, 128-byte swizzled
    static constexpr int A2_BYTES = 2 * BOX_BYTES;            // two K blocks
    static constexpr int NPOOL = 4;
    static constexpr int IPT = 10;                            // staging items per tile: 8 boxes of y + 2 of t1'
    static constexpr int TMEM_COLS = 512;
    static constexpr int D2_COL = 0, D3_COL = 256;
    static constexpr int NBAR = 1 + 2 + 4 + 4 + 4 * NPOOL;
    static constexpr int SMEM_BYTES = 1024 + W3_BYTES + W1N_BYTES + A2_BYTES + NPOOL * BOX_BYTES + NBAR * 8 + 16;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
};
static_assert(C3n1Cfg::SMEM_BYTES <= 232448, "smem budget");

struct C3n1Params {
    const float* bias3;   // [512]
    const float* bias1n;  // [128]
};

// Tensor maps (BF16, 128-byte swizzle, 2-D [rows][cols], box = 64 cols x box_rows):
//   tmA   t2  [M][128]   box rows 128      tmW3  [512][128] box rows 64      tmW1n [128][512] box rows 64
//   tmRes x   [M][512]   box rows 128      tmY   y [M][512] box rows 128     tmT1n t1' [M][128] box rows 128
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
bneck_c3n1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW3,
                  const __grid_constant__ CUtensorMap tmW1n, const __grid_constant__ CUtensorMap tmRes,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmT1n,
                  const C3n1Params prm, const C3n1Geom g) {
    using namespace ptx;
    constexpr int NPOOL = Cfg::NPOOL, IPT = Cfg::IPT;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_w3 = smem;
    uint8_t* smem_w1n = smem_w3 + Cfg::W3_BYTES;
    uint8_t* smem_a2 = smem_w1n + Cfg::W1N_BYTES;
    uint8_t* smem_pool = smem_a2 + Cfg::A2_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_pool + NPOOL * Cfg::BOX_BYTES);
    uint64_t* w_full = bars;                 // leader: weights of both CTAs landed
    uint64_t* a_full = w_full + 1;           // leader: A2 tile landed in both CTAs
    uint64_t* a_empty = a_full + 1;          // per CTA (multicast commit)
    uint64_t* d2_full = a_empty + 1;         // per CTA (multicast commit), [2 halves]
    uint64_t* d2_empty = d2_full + 2;        // leader, 16 arrivals, [2 halves]
    uint64_t* d3_full = d2_empty + 2;        // per CTA (multicast commit), [2]
    uint64_t* d3_empty = d3_full + 2;        // leader, 16 arrivals, [2]
    uint64_t* box_ready = d3_empty + 2;      // per CTA: staging box free (or its residual tile landed)
    uint64_t* c_full = box_ready + NPOOL;    // per CTA, 8 arrivals
    uint64_t* cx_full = c_full + NPOOL;      // leader, 16 arrivals (conv1' operand complete in both CTAs)
    uint64_t* c_mma_done = cx_full + NPOOL;  // per CTA (multicast commit)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int T = (g.tiles - pair + num_pairs - 1) / num_pairs;
    const int items = T * IPT;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW3);
        tma_prefetch_desc(&tmW1n);
        tma_prefetch_desc(&tmRes);
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmT1n);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(w_full, 1);
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 2 * Cfg::EPI_WARPS);
            mbar_init(&d3_full[i], 1);
            mbar_init(&d3_empty[i], 2 * Cfg::EPI_WARPS);
        }
        for (int i = 0; i < NPOOL; ++i) {
            mbar_init(&box_ready[i], 1);
            mbar_init(&c_full[i], Cfg::EPI_WARPS);
            mbar_init(&cx_full[i], 2 * Cfg::EPI_WARPS);
            mbar_init(&c_mma_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();  // see conv_igemm.cuh
    griddep_wait();

    // local tile index -> first pixel row owned by THIS CTA
    auto tile_row0 = [&](int it_local) {
        const int t = pair + it_local * num_pairs;
        const int tt = g.reverse ? g.tiles - 1 - t : t;
        return tt * 256 + static_cast<int>(rank) * 128;
    };

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (elect_one()) {
            const int r64 = static_cast<int>(rank) * 64;
            if (rank == 0) mbar_expect_tx(w_full, 2 * (Cfg::W3_BYTES + Cfg::W1N_BYTES));
            for (int q = 0; q < 4; ++q)
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d_2sm(smem_w3 + (q * 2 + kb) * Cfg::W3_BLK_BYTES, &tmW3, w_full, kb * 64, q * 128 + r64);
            for (int kb = 0; kb < 8; ++kb)
                tma_load_2d_2sm(smem_w1n + kb * Cfg::W1N_BLK_BYTES, &tmW1n, w_full, kb * 64, r64);
        }
        __syncwarp();
        for (int it = 0; it < T; ++it) {
            const int row0 = tile_row0(it);
            mbar_wait(a_empty, (it & 1) ^ 1);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(a_full, 2 * Cfg::A2_BYTES);
                tma_load_2d_2sm(smem_a2, &tmA, a_full, 0, row0);
                tma_load_2d_2sm(smem_a2 + Cfg::BOX_BYTES, &tmA, a_full, 64, row0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================================================== conv3 MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc128 = umma_instr_desc(UMMA_FMT_BF16, 256, 128);
            const uint64_t a2_desc = umma_smem_desc(smem_u32(smem_a2), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t w3_desc = umma_smem_desc(smem_u32(smem_w3), 0, 1024, UMMA_LAYOUT_SW128);
            mbar_wait(w_full, 0);
            tc_fence_after();
            for (int i = 0; i < T; ++i) {
                mbar_wait(a_full, i & 1);
#pragma unroll 1
                for (int hq = 0; hq < 2; ++hq) {      // quarter q = 2 * hq + hf lives in D2 half hf
#pragma unroll 1
                    for (int hf = 0; hf < 2; ++hf) {
                        const int q = 2 * hq + hf;
                        // use number 2i + hq of this half
                        mbar_wait(hf == 0 ? &d2_empty[0] : &d2_empty[1], (hq & 1) ^ 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d_tmem = tmem_base + Cfg::D2_COL + hf * 128;
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb) {
                                const uint64_t a = a2_desc + static_cast<uint64_t>((kb * Cfg::BOX_BYTES) >> 4);
                                const uint64_t b = w3_desc + static_cast<uint64_t>(((q * 2 + kb) * Cfg::W3_BLK_BYTES) >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma_f16_ss_2sm(d_tmem, a + static_cast<uint64_t>(k * 2),
                                                   b + static_cast<uint64_t>(k * 2), idesc128, (kb | k) != 0);
                            }
                            if (q == 3) tc_commit_2sm(a_empty);  // the A2 tile is fully consumed
                            tc_commit_2sm(hf == 0 ? &d2_full[0] : &d2_full[1]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================== conv1' MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc128 = umma_instr_desc(UMMA_FMT_BF16, 256, 128);
            const uint64_t w1n_desc = umma_smem_desc(smem_u32(smem_w1n), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t pool_desc = umma_smem_desc(smem_u32(smem_pool), 0, 1024, UMMA_LAYOUT_SW128);
            uint32_t cx_phase_bits = 0;  // bit cs = parity of the next cx_full[cs] wait
            mbar_wait(w_full, 0);
            tc_fence_after();
            for (int i = 0; i < T; ++i) {
                const int buf = i & 1;
                mbar_wait(buf == 0 ? &d3_empty[0] : &d3_empty[1], ((i >> 1) & 1) ^ 1);
#pragma unroll 1
                for (int b = 0; b < 8; ++b) {
                    const int cs = (i * IPT + b) % NPOOL;
                    mbar_wait(&cx_full[cs], (cx_phase_bits >> cs) & 1);
                    cx_phase_bits ^= 1u << cs;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d_tmem = tmem_base + Cfg::D3_COL + buf * 128;
                        const uint64_t a = pool_desc + static_cast<uint64_t>((cs * Cfg::BOX_BYTES) >> 4);
                        const uint64_t bd = w1n_desc + static_cast<uint64_t>((b * Cfg::W1N_BLK_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss_2sm(d_tmem, a + static_cast<uint64_t>(k * 2), bd + static_cast<uint64_t>(k * 2),
                                           idesc128, (b | k) != 0);
                        tc_commit_2sm(&c_mma_done[cs]);
                        if (b == 7) tc_commit_2sm(buf == 0 ? &d3_full[0] : &d3_full[1]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== store warp (both CTAs), see bneck_l1.cuh
        // item = tile * IPT + sub: sub 0..7 = 64-channel boxes of y (residual prefetched), 8..9 = boxes of t1'
        auto prepare = [&](int item) {
            const int cs = item % NPOOL;
            const int it_local = item / IPT, sub = item - it_local * IPT;
            if (sub < 8) {
                mbar_expect_tx(&box_ready[cs], Cfg::BOX_BYTES);
                tma_load_2d(smem_pool + cs * Cfg::BOX_BYTES, &tmRes, &box_ready[cs], sub * 64, tile_row0(it_local));
            } else {
                mbar_arrive(&box_ready[cs]);
            }
        };
        if (elect_one()) {
            for (int i = 0; i < NPOOL && i < items; ++i) prepare(i);
        }
        __syncwarp();
        uint32_t md_phase_bits = 0;
        auto recycle = [&](int item) {
            const int cs = item % NPOOL;
            if (item % IPT < 8) {  // conv1' must have consumed the box as well
                mbar_wait(&c_mma_done[cs], (md_phase_bits >> cs) & 1);
                md_phase_bits ^= 1u << cs;
            }
            if (item + NPOOL < items && elect_one()) prepare(item + NPOOL);
            __syncwarp();
        };
        for (int item = 0; item < items; ++item) {
            const int cs = item % NPOOL;
            const int it_local = item / IPT, sub = item - it_local * IPT;
            mbar_wait(&c_full[cs], (item / NPOOL) & 1);
            if (elect_one()) {
                const uint8_t* box = smem_pool + cs * Cfg::BOX_BYTES;
                if (sub < 8)
                    tma_store_2d(&tmY, box, sub * 64, tile_row0(it_local));
                else
                    tma_store_2d(&tmT1n, box, (sub - 8) * 64, tile_row0(it_local));
                tma_store_commit();
                tma_store_wait_read<1>();
            }
            __syncwarp();
            if (item > 0) recycle(item - 1);
        }
        if (elect_one()) tma_store_wait_read<0>();
        __syncwarp();
        if (items > 0) recycle(items - 1);
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
    } else {
        // ===================================================== epilogue (both CTAs)
        const int q4 = warp & 3;
        const int h = (warp - 4) >> 2;  // which 32-column half of a 64-column box
        const int row_in_tile = q4 * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
        const uint32_t row_off = static_cast<uint32_t>(row_in_tile) * 128;

        auto publish = [&]() {
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
        };
        // one 64-channel box: accumulator columns [col, col + 64) -> staging box of `item`
        auto box_step = [&](int item, uint32_t col, const float* bias64, int has_res, bool to_mma) {
            const int cs = item % NPOOL;
            mbar_wait(&box_ready[cs], (item / NPOOL) & 1);
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(lane_base + col + h * 32, v);
            tmem_ld_wait();
            epilogue_chunk<2>(v, smem_pool + cs * Cfg::BOX_BYTES + row_off, static_cast<uint32_t>(h * 4), swz,
                              bias64 + h * 32, has_res, 1);
            publish();
            if (lane == 0) {
                mbar_arrive(&c_full[cs]);
                if (to_mma) mbar_arrive_leader(&cx_full[cs]);
            }
        };
        // quarter q = 2 * hq + hf of y for tile i (D2 half hf)
        auto E2 = [&](int i, int hq, int hf) {
            const int q = 2 * hq + hf;
            mbar_wait(hf == 0 ? &d2_full[0] : &d2_full[1], hq & 1);
            tc_fence_after();
            box_step(i * IPT + 2 * q, Cfg::D2_COL + hf * 128, prm.bias3 + q * 128, 1, true);
            box_step(i * IPT + 2 * q + 1, Cfg::D2_COL + hf * 128 + 64, prm.bias3 + q * 128 + 64, 1, true);
            if (lane == 0) mbar_arrive_leader(hf == 0 ? &d2_empty[0] : &d2_empty[1]);
        };
        auto E3 = [&](int i) {
            const int buf = i & 1;
            mbar_wait(buf == 0 ? &d3_full[0] : &d3_full[1], (i >> 1) & 1);
            tc_fence_after();
            box_step(i * IPT + 8, Cfg::D3_COL + buf * 128, prm.bias1n, 0, false);
            box_step(i * IPT + 9, Cfg::D3_COL + buf * 128 + 64, prm.bias1n + 64, 0, false);
            if (lane == 0) mbar_arrive_leader(buf == 0 ? &d3_empty[0] : &d3_empty[1]);
        };
        for (int i = 0; i < T; ++i) {
            E2(i, 0, 0);
            E2(i, 0, 1);
            E2(i, 1, 0);
            E2(i, 1, 1);
            E3(i);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's smem/TMEM must stay alive until the leader's MMAs have retired
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------
// Streamed-weights variant for wider blocks (layer3: conv3 256 -> 1024, conv1' 1024 -> 256, 2 x 512 KB of
// weights): same dataflow, but the weight matrices pass through two small shared-memory rings instead of being
// resident. Per CTA: A2 tile (128 rows x K3, single buffer), W3 ring (2 stages, one 128-channel chunk of y =
// 64 rows x K3 per stage), W1n ring (2 stages, one 64-wide K block of conv1' = N1/2 rows x 128 B per stage),
// 4 staging boxes. TMEM: D2 2 x 128 columns + D3 N1 (= 256) columns.
// Warps (416 threads): 0 producer (A2 + W3 ring), 1 conv3 issuer, 2 TMEM alloc + conv1' issuer, 3 store warp,
// 4..11 epilogue, 12 W1n-ring producer.
template <int K3_, int N3_, int N1_, int NW3_ = 4, int NW1_ = 2, int NPOOL_ = 4>
struct C3n1sCfg {
    static constexpr int K3 = K3_, N3 = N3_, N1 = N1_;
    static constexpr int KB3 = K3_ / 64;                      // K blocks of conv3
    static constexpr int NCHUNK = N3_ / 128;                  // 128-channel chunks of y
    static constexpr int NBOX_Y = N3_ / 64;                   // staging boxes of y per tile (= K blocks of conv1')
    static constexpr int NBOX_T = N1_ / 64;                   // staging boxes of t1' per tile
    static constexpr int IPT = NBOX_Y + NBOX_T;
    static constexpr int BOX_BYTES = 16384;
    static constexpr int A2_BYTES = KB3 * BOX_BYTES;
    static constexpr int W3_BLK_BYTES = 64 * 128;             // this CTA's 64 rows of one K block of a chunk
    static constexpr int KBS = 2;                             // K blocks per W3 ring stage (fine-grained ring:
    static constexpr int W3_STAGE_BYTES = KBS * W3_BLK_BYTES; //  more stages in flight for the same bytes)
    static constexpr int NW3 = NW3_;
    static_assert(KB3 % KBS == 0, "stage granularity");
    static constexpr int W1N_STAGE_BYTES = (N1_ / 2) * 128;   // this CTA's N1/2 rows of one K block
    static constexpr int NW1 = NW1_;
    static constexpr int NPOOL = NPOOL_;
    static constexpr int TMEM_COLS = 512;
    static constexpr int D2_COL = 0, D3_COL = 256;
    static constexpr int NBAR = 2 + 2 * NW3 + 2 * NW1 + 4 + 2 + 4 * NPOOL;
    static constexpr int SMEM_BYTES = 1024 + A2_BYTES + NW3 * W3_STAGE_BYTES + NW1 * W1N_STAGE_BYTES +
                                      NPOOL * BOX_BYTES + NBAR * 8 + 16;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32 + 32;
    static_assert(N1_ <= 256 && N1_ % 64 == 0 && N3_ % 128 == 0 && K3_ % 64 == 0, "shape");
    static_assert((N3_ / 128) % 4 == 0, "each D2 half must be used an even number of times per tile");
};
using C3n1sL3 = C3n1sCfg<256, 1024, 256>;
static_assert(C3n1sL3::SMEM_BYTES <= 232448, "smem budget");
// The same 224 KB split differently (RNB_C3N1S_RINGS=1/2/3, tools/stem_ab.py A/B): the in-kernel timeline
// (profiles/c3n1s_timeline_r1.md) shows the epilogue waiting on box_ready — residual loads take 2-4 k clocks under
// load and only two of four boxes can be in flight — while the conv3 issuer waits on the epilogue, not on its ring.
using C3n1sL3P5 = C3n1sCfg<256, 1024, 256, 3, 2, 5>;    // W3 ring 3 x 16 KB, five staging boxes
using C3n1sL3W3P5 = C3n1sCfg<256, 1024, 256, 2, 3, 5>;  // W3 ring 2, W1n ring 3, five boxes
using C3n1sL3P6 = C3n1sCfg<256, 1024, 256, 2, 2, 6>;    // W3 ring 2, six boxes
static_assert(C3n1sL3P5::SMEM_BYTES <= 232448 && C3n1sL3W3P5::SMEM_BYTES <= 232448 && C3n1sL3P6::SMEM_BYTES <= 232448,
              "smem budget");

// Tensor maps: tmA t2 [M][K3] box rows 128; tmW3 [N3][K3] box rows 64; tmW1n [N1][N3] box rows N1/2;
//              tmRes / tmY [M][N3] box rows 128; tmT1n [M][N1] box rows 128
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
bneck_c3n1s_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW3,
                   const __grid_constant__ CUtensorMap tmW1n, const __grid_constant__ CUtensorMap tmRes,
                   const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmT1n,
                   const C3n1Params prm, const C3n1Geom g) {
    using namespace ptx;
    constexpr int NPOOL = Cfg::NPOOL, IPT = Cfg::IPT, NW3 = Cfg::NW3, NW1 = Cfg::NW1;
    constexpr int NCHUNK = Cfg::NCHUNK, NBOX_Y = Cfg::NBOX_Y, NBOX_T = Cfg::NBOX_T, KB3 = Cfg::KB3;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a2 = smem;
    uint8_t* smem_w3 = smem_a2 + Cfg::A2_BYTES;
    uint8_t* smem_w1n = smem_w3 + NW3 * Cfg::W3_STAGE_BYTES;
    uint8_t* smem_pool = smem_w1n + NW1 * Cfg::W1N_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_pool + NPOOL * Cfg::BOX_BYTES);
    uint64_t* a_full = bars;                 // leader
    uint64_t* a_empty = a_full + 1;          // per CTA (multicast commit)
    uint64_t* w3_full = a_empty + 1;         // leader, [NW3]
    uint64_t* w3_empty = w3_full + NW3;      // per CTA (multicast commit), [NW3]
    uint64_t* w1_full = w3_empty + NW3;      // leader, [NW1]
    uint64_t* w1_empty = w1_full + NW1;      // per CTA (multicast commit), [NW1]
    uint64_t* d2_full = w1_empty + NW1;      // per CTA, [2 halves]
    uint64_t* d2_empty = d2_full + 2;        // leader, 16 arrivals, [2 halves]
    uint64_t* d3_full = d2_empty + 2;        // per CTA
    uint64_t* d3_empty = d3_full + 1;        // leader, 16 arrivals
    uint64_t* box_ready = d3_empty + 1;
    uint64_t* c_full = box_ready + NPOOL;
    uint64_t* cx_full = c_full + NPOOL;
    uint64_t* c_mma_done = cx_full + NPOOL;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int T = (g.tiles - pair + num_pairs - 1) / num_pairs;
    const int items = T * IPT;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW3);
        tma_prefetch_desc(&tmW1n);
        tma_prefetch_desc(&tmRes);
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmT1n);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < NW3; ++i) {
            mbar_init(&w3_full[i], 1);
            mbar_init(&w3_empty[i], 1);
        }
        for (int i = 0; i < NW1; ++i) {
            mbar_init(&w1_full[i], 1);
            mbar_init(&w1_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 2 * Cfg::EPI_WARPS);
        }
        mbar_init(d3_full, 1);
        mbar_init(d3_empty, 2 * Cfg::EPI_WARPS);
        for (int i = 0; i < NPOOL; ++i) {
            mbar_init(&box_ready[i], 1);
            mbar_init(&c_full[i], Cfg::EPI_WARPS);
            mbar_init(&cx_full[i], 2 * Cfg::EPI_WARPS);
            mbar_init(&c_mma_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();  // see conv_igemm.cuh
    griddep_wait();

    auto tile_row0 = [&](int it_local) {
        const int t = pair + it_local * num_pairs;
        const int tt = g.reverse ? g.tiles - 1 - t : t;
        return tt * 256 + static_cast<int>(rank) * 128;
    };

    if (warp == 0) {
        // ===================================================== producer: A2 tile + W3 ring (both CTAs)
        const int r64 = static_cast<int>(rank) * 64;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < T; ++it) {
            const int row0 = tile_row0(it);
            mbar_wait(a_empty, (it & 1) ^ 1);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(a_full, 2 * Cfg::A2_BYTES);
                for (int kb = 0; kb < KB3; ++kb)
                    tma_load_2d_2sm(smem_a2 + kb * Cfg::BOX_BYTES, &tmA, a_full, kb * 64, row0);
            }
            __syncwarp();
            for (int c = 0; c < NCHUNK; ++c) {
                for (int ks = 0; ks < KB3 / Cfg::KBS; ++ks) {
                    mbar_wait(&w3_empty[stage], phase ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&w3_full[stage], 2 * Cfg::W3_STAGE_BYTES);
                        for (int kk = 0; kk < Cfg::KBS; ++kk)
                            tma_load_2d_2sm(smem_w3 + stage * Cfg::W3_STAGE_BYTES + kk * Cfg::W3_BLK_BYTES, &tmW3,
                                            &w3_full[stage], (ks * Cfg::KBS + kk) * 64, c * 128 + r64);
                    }
                    __syncwarp();
                    if (++stage == NW3) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 12) {
        // ===================================================== producer: W1n ring (both CTAs)
        const int rrow = static_cast<int>(rank) * (Cfg::N1 / 2);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < T; ++it) {
            for (int b = 0; b < NBOX_Y; ++b) {
                mbar_wait(&w1_empty[stage], phase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&w1_full[stage], 2 * Cfg::W1N_STAGE_BYTES);
                    tma_load_2d_2sm(smem_w1n + stage * Cfg::W1N_STAGE_BYTES, &tmW1n, &w1_full[stage], b * 64, rrow);
                }
                __syncwarp();
                if (++stage == NW1) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== conv3 MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc128 = umma_instr_desc(UMMA_FMT_BF16, 256, 128);
            const uint64_t a2_desc = umma_smem_desc(smem_u32(smem_a2), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t w3_desc = umma_smem_desc(smem_u32(smem_w3), 0, 1024, UMMA_LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0;
            TL_DECL(1)
            for (int i = 0; i < T; ++i) {
                TL_TILE(i)
                TL(1)
                mbar_wait(a_full, i & 1);
                TL(2)
#pragma unroll 1
                for (int hc = 0; hc < NCHUNK / 2; ++hc) {  // chunk c = 2 * hc + hf lives in D2 half hf
#pragma unroll 1
                    for (int hf = 0; hf < 2; ++hf) {
                        const int c = 2 * hc + hf;
                        // use number (NCHUNK / 2) * i + hc of this half (NCHUNK / 2 is even)
                        TL(10 + c)
                        mbar_wait(hf == 0 ? &d2_empty[0] : &d2_empty[1], (hc & 1) ^ 1);
                        TL(20 + c)
                        const uint32_t d_tmem = tmem_base + Cfg::D2_COL + hf * 128;
#pragma unroll 1
                        for (int ks = 0; ks < KB3 / Cfg::KBS; ++ks) {
                            mbar_wait(&w3_full[stage], phase);
                            TL(30 + c)
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t bs = w3_desc + static_cast<uint64_t>((stage * Cfg::W3_STAGE_BYTES) >> 4);
#pragma unroll
                                for (int kk = 0; kk < Cfg::KBS; ++kk) {
                                    const int kb = ks * Cfg::KBS + kk;
                                    const uint64_t a = a2_desc + static_cast<uint64_t>((kb * Cfg::BOX_BYTES) >> 4);
                                    const uint64_t b = bs + static_cast<uint64_t>((kk * Cfg::W3_BLK_BYTES) >> 4);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma_f16_ss_2sm(d_tmem, a + static_cast<uint64_t>(k * 2),
                                                       b + static_cast<uint64_t>(k * 2), idesc128, (kb | k) != 0);
                                }
                                tc_commit_2sm(&w3_empty[stage]);
                                if (ks == KB3 / Cfg::KBS - 1) {
                                    if (c == NCHUNK - 1) tc_commit_2sm(a_empty);  // the A2 tile is fully consumed
                                    tc_commit_2sm(hf == 0 ? &d2_full[0] : &d2_full[1]);
                                }
                            }
                            __syncwarp();
                            if (++stage == NW3) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
            TL_DUMP()
        }
    } else if (warp == 2) {
        // ===================================================== conv1' MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idescn1 = umma_instr_desc(UMMA_FMT_BF16, 256, Cfg::N1);
            const uint64_t w1n_desc = umma_smem_desc(smem_u32(smem_w1n), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t pool_desc = umma_smem_desc(smem_u32(smem_pool), 0, 1024, UMMA_LAYOUT_SW128);
            uint32_t cx_phase_bits = 0;
            int stage = 0;
            uint32_t phase = 0;
            TL_DECL(2)
            for (int i = 0; i < T; ++i) {
                TL_TILE(i)
                TL(1)
                mbar_wait(d3_empty, (i & 1) ^ 1);
                TL(2)
#pragma unroll 1
                for (int b = 0; b < NBOX_Y; ++b) {
                    const int cs = (i * IPT + b) % NPOOL;
                    mbar_wait(&w1_full[stage], phase);
                    TL(10 + b)
                    mbar_wait(&cx_full[cs], (cx_phase_bits >> cs) & 1);
                    TL(30 + b)
                    cx_phase_bits ^= 1u << cs;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d_tmem = tmem_base + Cfg::D3_COL;
                        const uint64_t a = pool_desc + static_cast<uint64_t>((cs * Cfg::BOX_BYTES) >> 4);
                        const uint64_t bd = w1n_desc + static_cast<uint64_t>((stage * Cfg::W1N_STAGE_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss_2sm(d_tmem, a + static_cast<uint64_t>(k * 2), bd + static_cast<uint64_t>(k * 2),
                                           idescn1, (b | k) != 0);
                        tc_commit_2sm(&c_mma_done[cs]);
                        tc_commit_2sm(&w1_empty[stage]);
                        if (b == NBOX_Y - 1) tc_commit_2sm(d3_full);
                    }
                    __syncwarp();
                    if (++stage == NW1) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            TL_DUMP()
        }
    } else if (warp == 3) {
        // ===================================================== store warp (both CTAs), see bneck_l1.cuh
        auto prepare = [&](int item) {
            const int cs = item % NPOOL;
            const int it_local = item / IPT, sub = item - it_local * IPT;
            if (sub < NBOX_Y) {
                mbar_expect_tx(&box_ready[cs], Cfg::BOX_BYTES);
                tma_load_2d(smem_pool + cs * Cfg::BOX_BYTES, &tmRes, &box_ready[cs], sub * 64, tile_row0(it_local));
            } else {
                mbar_arrive(&box_ready[cs]);
            }
        };
        if (elect_one()) {
            for (int i = 0; i < NPOOL && i < items; ++i) prepare(i);
        }
        __syncwarp();
        uint32_t md_phase_bits = 0;
        TL_DECL(3)
        auto recycle = [&](int item) {
            const int cs = item % NPOOL;
            if (item % IPT < NBOX_Y) {
                mbar_wait(&c_mma_done[cs], (md_phase_bits >> cs) & 1);
                md_phase_bits ^= 1u << cs;
            }
            TL(90 + item % IPT)
            if (item + NPOOL < items && elect_one()) prepare(item + NPOOL);
            __syncwarp();
        };
        for (int item = 0; item < items; ++item) {
            const int cs = item % NPOOL;
            const int it_local = item / IPT, sub = item - it_local * IPT;
            TL_TILE(it_local)
            TL(10 + sub)
            mbar_wait(&c_full[cs], (item / NPOOL) & 1);
            TL(40 + sub)
            if (elect_one()) {
                const uint8_t* box = smem_pool + cs * Cfg::BOX_BYTES;
                if (sub < NBOX_Y)
                    tma_store_2d(&tmY, box, sub * 64, tile_row0(it_local));
                else
                    tma_store_2d(&tmT1n, box, (sub - NBOX_Y) * 64, tile_row0(it_local));
                tma_store_commit();
                tma_store_wait_read<1>();
            }
            __syncwarp();
            TL(65 + sub)
            if (item > 0) recycle(item - 1);
        }
        if (elect_one()) tma_store_wait_read<0>();
        __syncwarp();
        if (items > 0) recycle(items - 1);
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
        TL_DUMP()
    } else if (warp >= 4 && warp < 12) {
        // ===================================================== epilogue (both CTAs)
        const int q4 = warp & 3;
        const int h = (warp - 4) >> 2;
        const int row_in_tile = q4 * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
        const uint32_t row_off = static_cast<uint32_t>(row_in_tile) * 128;

        TL_DECL(0)
        auto box_step = [&](int item, uint32_t col, const float* bias64, int has_res, bool to_mma) {
            const int cs = item % NPOOL;
            TL(20 + item % IPT)
            mbar_wait(&box_ready[cs], (item / NPOOL) & 1);
            TL(50 + item % IPT)
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(lane_base + col + h * 32, v);
            tmem_ld_wait();
            TL(80 + item % IPT)
            epilogue_chunk<2>(v, smem_pool + cs * Cfg::BOX_BYTES + row_off, static_cast<uint32_t>(h * 4), swz,
                              bias64 + h * 32, has_res, 1);
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&c_full[cs]);
                if (to_mma) mbar_arrive_leader(&cx_full[cs]);
            }
            TL(110 + item % IPT)
        };
        for (int i = 0; i < T; ++i) {
            TL_TILE(i)
#pragma unroll 1
            for (int hc = 0; hc < NCHUNK / 2; ++hc) {
                // half 0: chunk 2 hc
                TL(1)
                mbar_wait(&d2_full[0], hc & 1);
                TL(2)
                tc_fence_after();
                box_step(i * IPT + 4 * hc, Cfg::D2_COL, prm.bias3 + hc * 256, 1, true);
                box_step(i * IPT + 4 * hc + 1, Cfg::D2_COL + 64, prm.bias3 + hc * 256 + 64, 1, true);
                if (lane == 0) mbar_arrive_leader(&d2_empty[0]);
                // half 1: chunk 2 hc + 1
                TL(3)
                mbar_wait(&d2_full[1], hc & 1);
                TL(4)
                tc_fence_after();
                box_step(i * IPT + 4 * hc + 2, Cfg::D2_COL + 128, prm.bias3 + hc * 256 + 128, 1, true);
                box_step(i * IPT + 4 * hc + 3, Cfg::D2_COL + 192, prm.bias3 + hc * 256 + 192, 1, true);
                if (lane == 0) mbar_arrive_leader(&d2_empty[1]);
            }
            TL(5)
            mbar_wait(d3_full, i & 1);
            TL(6)
            tc_fence_after();
#pragma unroll 1
            for (int jb = 0; jb < NBOX_T; ++jb)
                box_step(i * IPT + NBOX_Y + jb, Cfg::D3_COL + jb * 64, prm.bias1n + jb * 64, 0, false);
            if (lane == 0) mbar_arrive_leader(d3_empty);
            TL(7)
        }
        TL_DUMP()
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
