// conv_igemm2.cuh — the 2-CTA (cta_group::2) variant of the implicit-GEMM convolution.
//
// Same math and same fused epilogue as conv_igemm.cuh (the replacement of conv2dForwardKernel +
// batchNorm2dForwardKernel + addForwardKernel + reluForwardKernel, /root/reference/cuda/ops.cu:14-48,
// 139-151,153-160,130-137), but each tile is computed by a PAIR of CTAs on the two SMs of a TPC:
//
//   tile = 256 pixels x BN channels (BN = 128 or 256); CTA r of the pair owns pixels [128r, 128r+128)
//   A: each CTA im2col-loads its own 128-pixel K-slab        (16 KB / stage)
//   B: each CTA loads HALF of the weight tile (BN/2 rows)    (8 or 16 KB / stage)
//   one tcgen05.mma.cta_group::2 (M = 256), issued by the leader CTA, reads A from both SMs and the
//   two B halves from both SMs; each SM's TMEM receives its own 128 x BN half of the accumulator.
//
// Why: with 128 x 128 single-CTA tiles every 1 M MACs pull 32 KB through L2 (32 MAC/B) and the
// compute-bound layers saturate the L2 -> SM fabric (~13 TB/s) at ~40 % of tensor peak. The pair
// halves the bytes per MAC (64 MAC/B at BN = 256) and halves each SM's shared-memory operand reads.
//
// Synchronisation (barriers have the same offsets in both CTAs):
//   full[s]       lives in the LEADER: the leader's producer arms it with the byte count of BOTH
//                 CTAs; both CTAs' TMA loads complete_tx on it (.cta_group::2 form, peer bit cleared)
//   empty[s]      one per CTA, released by the leader's multicast tcgen05.commit
//   tmem_full[a]  one per CTA, multicast commit after the last K block of a tile
//   tmem_empty[a] in the leader, 16 arrivals: one per epilogue warp of both CTAs (remote arrive)
//   res_full[c] / c_full[c] / c_free[c]   per CTA (each CTA stores / prefetches the rows it owns)
// Tail split (ConvGeom::split_from): the tiles of the last, partially filled wave are issued as two work units
// of BN/2 columns each (own accumulator stage, B box of BN/4 rows per CTA through tmBh, instruction descriptor
// with N = BN/2). Per column the K order is unchanged, so the result is bit-identical to the unsplit launch.
// The epilogue works on 32 KB sub-tiles (128 rows x 256 bytes) through NCBUF rotating staging
// buffers handed to a dedicated store warp, exactly as in the single-CTA kernel.
#pragma once
#include "conv_igemm.cuh"

namespace rnb {

// RESB_KB_ > 0: RESIDENT weights — the layer's whole weight matrix (RESB_KB_ K blocks, one N tile) is loaded into
// shared memory ONCE per CTA and only the A operand streams. For the 128-channel 3x3 convs of layer2 (18 K blocks,
// 144 KB per CTA) the per-tile weight reloads were a third of an L2 -> SM traffic that sits at the fabric limit
// (profiles/l2conv3x3_r1.md: 231 of 694 MB); paid for with a 3-deep A ring and ONE staging buffer.
template <int BN_, int ESZ_, int NSTAGE_, int NCBUF_, int OSZ_ = ESZ_, int RESB_KB_ = 0>
struct Conv2Cfg {
    static constexpr int BM = 256;                    // pixels per CTA pair
    static constexpr int BM_CTA = 128;                // pixels per CTA
    static constexpr int BN = BN_;
    static constexpr int ESZ = ESZ_;
    static constexpr int OSZ = OSZ_;                   // output (and residual) element bytes; OSZ = 1 with ESZ = 2:
                                                       // BF16 operands, E4M3 output (hand-over of a mixed FP8 plan)
    static constexpr int BK = 128 / ESZ_;
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr int NCBUF = NCBUF_;
    static constexpr int A_BYTES = BM_CTA * 128;
    static constexpr int B_BYTES = (BN_ / 2) * 128;    // this CTA's half of the weight tile
    static constexpr int RESB_KB = RESB_KB_;
    static constexpr int RES_BYTES = RESB_KB_ * B_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + (RESB_KB_ > 0 ? 0 : B_BYTES);
    static constexpr int BOX_COLS = 128 / OSZ_;
    static constexpr int SUB_BOXES = OSZ_ == 1 ? 1 : 2;  // 128-byte boxes per epilogue sub-tile (FP8: one box = 128 columns)
    static constexpr int EPI_N = SUB_BOXES * BOX_COLS;   // columns per epilogue sub-tile
    static constexpr int NSUB = BN_ / EPI_N;
    static constexpr int BOX_BYTES = BM_CTA * 128;
    static constexpr int CBUF_BYTES = SUB_BOXES * BOX_BYTES;   // 32 KB (16 KB on the FP8 path)
    static constexpr int TMEM_COLS = 2 * BN_;
    static constexpr int NBAR = 2 * NSTAGE_ + 4 + 3 * NCBUF_ + 1;   // + b_full (resident weights)
    static constexpr int SMEM_BYTES =
        1024 + RES_BYTES + NSTAGE_ * STAGE_BYTES + NCBUF_ * CBUF_BYTES + NBAR * 8 + 16;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32;
    static_assert(NCBUF_ >= 1, "need a staging buffer");
    static_assert(RES_BYTES % 1024 == 0, "the stage ring behind the resident weights must stay 1024-byte aligned");
    static_assert(BN_ % EPI_N == 0, "tile N must be a multiple of the epilogue sub-tile");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
};

namespace ptx {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in the pair's leader CTA (works from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                int32_t x, int32_t y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)),
          "r"(smem_u32(bar) & kPeerBitMask), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(void* smem_dst, const CUtensorMap* m,
                                                       uint64_t* bar, int32_t c, int32_t w, int32_t h,
                                                       int32_t n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)),
          "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// commit: arrive on the barrier at this offset in BOTH CTAs of the pair once prior MMAs retire
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_f8_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ss_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

}  // namespace ptx

// Tensor maps: tmA im2col over the input (128 pixels x 128 bytes per load), tmB tiled over the packed
// weights with a box of BN/2 rows, tmOut / tmRes tiled over [M][Cout] with boxes of 128 rows.
// Warp roles per CTA (384 threads): 0 TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM alloc,
// 3 store warp (TMA stores, residual prefetch, staging recycling), 4..11 epilogue (two warps per
// TMEM lane quarter, splitting the 32-column chunks of each sub-tile).
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
conv_igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmRes, const float* __restrict__ bias,
                   const ConvGeom g) {
    using namespace ptx;
    constexpr int BN = Cfg::BN;
    constexpr int NSTAGE = Cfg::NSTAGE;
    constexpr int NCBUF = Cfg::NCBUF;
    constexpr int NSUB = Cfg::NSUB;
    constexpr int HSUB = NSUB >= 2 ? NSUB / 2 : 1;  // sub-tiles of a half unit (the planner splits only if NSUB is even)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_res = smem;                       // resident weights (RESB_KB K blocks of this CTA's N half), or empty
    uint8_t* smem_stage = smem + Cfg::RES_BYTES;
    uint8_t* smem_c = smem_stage + NSTAGE * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + NCBUF * Cfg::CBUF_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + NSTAGE;
    uint64_t* tmem_full = bars + 2 * NSTAGE;
    uint64_t* tmem_empty = bars + 2 * NSTAGE + 2;
    uint64_t* res_full = bars + 2 * NSTAGE + 4;
    uint64_t* c_full = res_full + NCBUF;
    uint64_t* c_free = c_full + NCBUF;
    uint64_t* b_full = c_free + NCBUF;              // leader: the resident weights of BOTH CTAs have landed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_tiles = g.m_tiles * g.n_tiles;       // m_tiles counts 256-pixel tiles here
    // work units: whole tiles [0, F), then the two N halves of every tile in [F, num_tiles) (see ConvGeom)
    const int F = g.split_from;
    const int num_units = F + 2 * (num_tiles - F);
    const int my_tiles = (num_units - pair + num_pairs - 1) / num_pairs;   // units of this pair, whole ones first
    const int my_full = F > pair ? (F - pair + num_pairs - 1) / num_pairs : 0;
    const int my_items = my_full * NSUB + (my_tiles - my_full) * HSUB;      // (unit, sub-tile) epilogue work items

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (F < num_tiles) tma_prefetch_desc(&tmBh);
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
            mbar_init(&tmem_empty[i], 2 * Cfg::EPI_WARPS);
        }
        for (int i = 0; i < NCBUF; ++i) {
            mbar_init(&res_full[i], 1);
            mbar_init(&c_full[i], Cfg::EPI_WARPS);
            mbar_init(&c_free[i], 1);
        }
        mbar_init(b_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();  // see conv_igemm.cuh
    griddep_wait();

    // local unit index -> (m_blk, n_blk, half): half = -1 for a whole tile, 0 / 1 for the N halves of a split tile
    auto tile_coords = [&](int it_local, int& m_blk, int& n_blk, int& half) {
        const int u = pair + it_local * num_pairs;
        int t = u;
        half = -1;
        if (u >= F) {
            t = F + ((u - F) >> 1);
            half = (u - F) & 1;
        }
        const int tt = g.reverse ? num_tiles - 1 - t : t;
        n_blk = tt % g.n_tiles;
        m_blk = tt / g.n_tiles;
    };
    // epilogue work item -> first output column / first pixel row of this CTA's half
    auto item_coords = [&](int item, int& col0, int& row0) {
        int it_local, sub;
        if (item < my_full * NSUB) {
            it_local = item / NSUB;
            sub = item - it_local * NSUB;
        } else {
            const int j = item - my_full * NSUB;
            it_local = my_full + j / HSUB;
            sub = j - (j / HSUB) * HSUB;
        }
        int m_blk, n_blk, half;
        tile_coords(it_local, m_blk, n_blk, half);
        col0 = n_blk * BN + (half > 0 ? BN / 2 : 0) + sub * Cfg::EPI_N;
        row0 = m_blk * Cfg::BM + static_cast<int>(rank) * Cfg::BM_CTA;
    };

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (Cfg::RESB_KB > 0) {
            // resident weights: every K block of this CTA's half of the (single) N tile, once
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(b_full, 2 * Cfg::RES_BYTES);
                for (int kb = 0; kb < Cfg::RESB_KB; ++kb)
                    tma_load_2d_2sm(smem_res + kb * Cfg::B_BYTES, &tmB, b_full, kb * Cfg::BK,
                                    static_cast<int>(rank) * (BN / 2));
            }
            __syncwarp();
        }
        int stage = 0;
        uint32_t phase = 0;
        const int ohw = g.OH * g.OW;
        for (int it = 0; it < my_tiles; ++it) {
            int m_blk, n_blk, half;
            tile_coords(it, m_blk, n_blk, half);
            const int m0 = m_blk * Cfg::BM + static_cast<int>(rank) * Cfg::BM_CTA;
            const int img = m0 / ohw;
            const int rem = m0 - img * ohw;
            const int p = rem / g.OW;
            const int q = rem - p * g.OW;
            const int w0 = g.lower + q * g.stride;
            const int h0 = g.lower + p * g.stride;
            const int nrow0 = half < 0 ? n_blk * BN + static_cast<int>(rank) * (BN / 2)
                                       : n_blk * BN + half * (BN / 2) + static_cast<int>(rank) * (BN / 4);
            const uint32_t tx_bytes =
                Cfg::RESB_KB > 0 ? 2 * Cfg::A_BYTES : 2 * (Cfg::A_BYTES + (half < 0 ? Cfg::B_BYTES : Cfg::B_BYTES / 2));
            const CUtensorMap* tmBu = half < 0 ? &tmB : &tmBh;
            int tap_r = 0, tap_s = 0, cblk = 0;
            for (int kb = 0; kb < g.num_kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem_stage + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], tx_bytes);
                    tma_load_im2col_4d_2sm(sa, &tmA, &full_bar[stage], cblk * Cfg::BK, w0, h0, img,
                                           static_cast<uint16_t>(tap_s), static_cast<uint16_t>(tap_r));
                    if (Cfg::RESB_KB == 0) tma_load_2d_2sm(sb, tmBu, &full_bar[stage], kb * Cfg::BK, nrow0);
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
        // ===================================================== MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t fmt = Cfg::ESZ == 2 ? UMMA_FMT_BF16 : (Cfg::ESZ == 4 ? UMMA_FMT_TF32 : UMMA_FMT_E4M3);
            constexpr uint32_t idesc_full = umma_instr_desc(fmt, Cfg::BM, BN);
            constexpr uint32_t idesc_half = umma_instr_desc(fmt, Cfg::BM, BN / 2);
            const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_stage), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t b_desc0 =
                umma_smem_desc(smem_u32(smem_stage) + Cfg::A_BYTES, 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t b_res0 = umma_smem_desc(smem_u32(smem_res), 0, 1024, UMMA_LAYOUT_SW128);
            if (Cfg::RESB_KB > 0) {
                mbar_wait(b_full, 0);
                tc_fence_after();
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                const uint32_t idesc = it < my_full ? idesc_full : idesc_half;
                for (int kb = 0; kb < g.num_kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t soff = static_cast<uint64_t>((stage * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = a_desc0 + soff + static_cast<uint64_t>(k * 2);
                            const uint64_t bd = Cfg::RESB_KB > 0
                                                    ? b_res0 + static_cast<uint64_t>((kb * Cfg::B_BYTES) >> 4) + static_cast<uint64_t>(k * 2)
                                                    : b_desc0 + soff + static_cast<uint64_t>(k * 2);
                            if (Cfg::ESZ == 2)
                                mma_f16_ss_2sm(d_tmem, ad, bd, idesc, (kb | k) != 0);
                            else if (Cfg::ESZ == 4)
                                mma_tf32_ss_2sm(d_tmem, ad, bd, idesc, (kb | k) != 0);
                            else
                                mma_f8_ss_2sm(d_tmem, ad, bd, idesc, (kb | k) != 0);
                        }
                        tc_commit_2sm(&empty_bar[stage]);
                        if (kb == g.num_kblocks - 1) tc_commit_2sm(&tmem_full[as]);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===================================================== store warp (both CTAs)
        auto issue_residual = [&](int item) {
            int col0, row0;
            item_coords(item, col0, row0);
            const int cs = item % NCBUF;
            uint8_t* buf = smem_c + cs * Cfg::CBUF_BYTES;
            mbar_expect_tx(&res_full[cs], Cfg::CBUF_BYTES);
            tma_load_2d(buf, &tmRes, &res_full[cs], col0, row0);
            if (Cfg::SUB_BOXES == 2)
                tma_load_2d(buf + Cfg::BOX_BYTES, &tmRes, &res_full[cs], col0 + Cfg::BOX_COLS, row0);
        };
        if (g.has_res && elect_one()) {
            for (int i = 0; i < NCBUF && i < my_items; ++i) issue_residual(i);
        }
        __syncwarp();
        for (int item = 0; item < my_items; ++item) {
            const int cs = item % NCBUF;
            mbar_wait(&c_full[cs], (item / NCBUF) & 1);
            if (elect_one()) {
                int col0, row0;
                item_coords(item, col0, row0);
                const uint8_t* cbuf = smem_c + cs * Cfg::CBUF_BYTES;
                tma_store_2d(&tmOut, cbuf, col0, row0);
                if (Cfg::SUB_BOXES == 2) tma_store_2d(&tmOut, cbuf + Cfg::BOX_BYTES, col0 + Cfg::BOX_COLS, row0);
                tma_store_commit();
                tma_store_wait_read<0>();
                if (item + NCBUF < my_items) {
                    if (g.has_res)
                        issue_residual(item + NCBUF);
                    else
                        mbar_arrive(&c_free[cs]);
                }
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================== epilogue (both CTAs)
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int row_in_tile = q * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        int item = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            const int nsub = it < my_full ? NSUB : HSUB;
#pragma unroll 1
            for (int sub = 0; sub < nsub; ++sub, ++item) {
                int col0, row0;
                item_coords(item, col0, row0);
                const int cs = item % NCBUF;
                uint8_t* cbuf = smem_c + cs * Cfg::CBUF_BYTES;
                if (g.has_res)
                    mbar_wait(&res_full[cs], (item / NCBUF) & 1);
                else if (item >= NCBUF)
                    mbar_wait(&c_free[cs], ((item / NCBUF) - 1) & 1);
                float amax = 0.f;
#pragma unroll 1
                for (int chunk = h; chunk < Cfg::EPI_N / 32; chunk += 2) {
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32x32(taddr + sub * Cfg::EPI_N + chunk * 32, v);
                    tmem_ld_wait();
                    const int byte_off = chunk * 32 * Cfg::OSZ;
                    uint8_t* row = cbuf + (byte_off >> 7) * Cfg::BOX_BYTES + row_in_tile * 128;
                    if (Cfg::OSZ == 1)
                        amax = fmaxf(amax, epilogue_chunk_fp8(v, row, (byte_off & 127) >> 4, swz,
                                                              g.chan_scale + col0 + chunk * 32, bias + col0 + chunk * 32,
                                                              g.res_mul, g.has_res, g.relu, g.amax != nullptr));
                    else
                        epilogue_chunk<Cfg::OSZ == 1 ? 2 : Cfg::ESZ>(v, row, (byte_off & 127) >> 4, swz,
                                                                     bias + col0 + chunk * 32, g.has_res, g.relu);
                }
                if (Cfg::OSZ == 1 && g.amax) amax_commit(g.amax, row0 + row_in_tile < g.M ? amax : 0.f);
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    // last sub-tile: this warp has drained its part of the accumulator -> tell the leader
                    if (sub == nsub - 1) mbar_arrive_leader(&tmem_empty[as]);
                    mbar_arrive(&c_full[cs]);
                }
            }
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

}  // namespace rnb
