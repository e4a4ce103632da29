// bneck_l1.cuh — the tail of a layer1 Bottleneck block as ONE kernel (BF16, 64 -> 64 -> 256 channels):
//
//     t2 = relu(bn2(conv2_3x3(t1)))                 t1 = this block's conv1 output        [N,H,W,64]
//     y  = relu(bn3(conv3_1x1(t2)) + shortcut)      shortcut = x  or  bn_ds(conv_ds_1x1(x))
//     t1' = relu(bn1'(conv1'_1x1(y)))               optional: the NEXT block's conv1      [N,H,W,64]
//
// i.e. the part of layerForward (/root/reference/cuda/inference/main.cu:138-163) after the first
// conv/bn/relu, plus the first conv/bn/relu of the following block, replacing 3-4 launches of
// conv2dForwardKernel / batchNorm2dForwardKernel / reluForwardKernel / addForwardKernel
// (/root/reference/cuda/ops.cu:14-48,139-151,130-137,153-160) per block.
//
// Why: at 56x56 these layers are HBM-bound. Layer by layer a block moves 103 (t1 out) + 103 + 103
// (t2) + 103 + 411 (residual) + 411 (y) + 411 (y re-read by the next conv1) + 103 MB per 256 images;
// fused, t2 never leaves the SM and y is consumed from the staging tile it is stored from:
// 103 + 411 + 411 + 103 MB. With the downsample branch folded in (block 0: the shortcut is a second
// K block of the SAME accumulator, `D2 += x_tile * Wds^T`) the 411 MB shortcut tensor is neither
// written nor read.
//
// Structure: a CTA PAIR (cta_group::2) per tile of 4 image rows x 64-pixel pitch (256 GEMM rows, 2
// rows per CTA). All weights are resident in shared memory, split across the pair (each CTA holds
// half of the output channels of every matrix: 36 + 16 (+16) + 16 KB).
//   conv2 : halo-resident 3x3 (cf. conv3x3_halo.cuh) — ONE TMA load per tile fetches the 4 input rows
//           x 64-pixel pitch a CTA needs (zero halo by hardware OOB fill); the nine taps are the same
//           tile read through UMMA descriptors shifted by r rows (8 KB) + s pixels (128 B);
//           accumulator D1 (64 cols, double-buffered), input tile double-buffered
//   E1    : D1 + bias2 -> ReLU -> BF16 pairs -> tcgen05.st back into TMEM (A2, 32 cols, double-buffered)
//   conv3 : A2 (A operand FROM TMEM) x W3^T (K = 64) [+ x_tile x Wds^T] -> D2 (256 cols, two halves)
//   E2    : D2 + bias3 (+ residual, TMA-prefetched into the staging box) -> ReLU -> BF16 -> four
//           64-channel staging boxes -> TMA store to y; each box is ALSO the K block of
//   conv1': box_j x W1n_j^T accumulated over j = 0..3 -> D3 (64 cols)
//   E3    : D3 + bias1' -> ReLU -> BF16 -> staging box -> TMA store to t1'
// Warp roles per CTA (416 threads): 0 TMA producer (weights once, input tiles), 1 conv2 MMA issuer
// (leader CTA only), 2 TMEM allocator + shortcut-input loader (DS mode), 3 store warp (stores, residual
// prefetch, box recycling), 4..11 epilogue (two warps per TMEM lane quarter, 32 columns each), 12
// conv3 / conv1' MMA issuer (leader only). Two issuers on purpose: conv2 waits for HBM, conv3 / conv1'
// wait for the epilogue — one in-order issuer makes each stall the other.
// The epilogue is software-pipelined one tile ahead: iteration i runs E1(i+1), E3(i-1), E2(i), so
// every MMA group has a full epilogue phase to complete before its result is needed.
// Rows w >= W of the 64-pixel pitch are garbage end to end (GEMM rows are independent) and are never
// stored (the TMA store boxes are W wide).
#pragma once
#include "conv3x3_halo.cuh"
#include "conv_igemm2.cuh"

namespace rnb {

struct BneckGeom {
    int N, H, W;        // images, spatial size (8 <= W <= 62, H % 4 == 0)
    int tiles;          // N * H / 4 pair tiles
    int tiles_per_img;  // H / 4
    int has_next;       // compute and store the next block's conv1 output
    int reverse;        // tile traversal direction (see ConvGeom::reverse)
    int debug;          // timing experiments: 1 = no residual loads, 2 = no stores, 8 = all stores to one tile
                        // (results wrong); 4 = L2 prefetch of residual tiles (results right)
};

// N1_ = output channels of the fused next conv1: 64 (next block of the same layer) or 128 (first block of
// the NEXT layer, 256 -> 128 at the same resolution). N1 = 128 needs a 128-column D3, paid for with a
// single-buffered D1 (the epilogue runs one tile ahead, so conv2 still has a whole tile of slack).
// NABUF_ = input tiles in flight. The folded-downsample form (DS) is short of shared memory: with two input
// tiles it has room for three staging boxes only; with ONE input tile (the next load starts when conv2 has consumed
// the tile — the epilogue runs a whole tile behind conv2, which hides the load) it gets five (measured: 2.597 ->
// 2.586 ms per ResNet-50 step; a second shortcut-input tile instead of the fifth box: no gain).
template <bool DS_, int N1_ = 64, int NABUF_ = 2>
struct BneckCfg {
    static constexpr bool DS = DS_;  // shortcut = downsample conv of x, fused as a second K block
    static constexpr int N1 = N1_;
    static constexpr int NT = N1_ / 64;             // staging boxes of t1' per tile
    static constexpr int ND1 = N1_ == 64 ? 2 : 1;   // D1 buffers
    static_assert(N1_ == 64 || N1_ == 128, "next conv1 width");
    static_assert(!(DS_ && N1_ != 64), "the folded-downsample form is only built with N1 = 64");
    static constexpr int PITCH = 64;
    static constexpr int NABUF = NABUF_;            // input tiles in flight
    static constexpr int ATILE_BYTES = 32768;       // 4 rows x 64 pixels x 128 B
    static constexpr int AROW_BYTES = 8192;         // one input row of the tile
    static constexpr int W2_TAP_BYTES = 32 * 128;   // this CTA's 32 output channels of one tap
    static constexpr int W2_BYTES = 9 * W2_TAP_BYTES;
    static constexpr int W3_HALF_BYTES = 64 * 128;  // this CTA's 64 channels of one 128-channel half
    static constexpr int W3_BYTES = 2 * W3_HALF_BYTES;
    static constexpr int WDS_BYTES = DS_ ? W3_BYTES : 0;
    static constexpr int W1N_KB_BYTES = (N1_ / 2) * 128;  // this CTA's N1/2 channels x one 64-wide K block
    static constexpr int W1N_BYTES = 4 * W1N_KB_BYTES;
    static constexpr int W_BYTES = W2_BYTES + W3_BYTES + WDS_BYTES + W1N_BYTES;
    static constexpr int RING_BYTES = NABUF * ATILE_BYTES + 1024;  // + read-past pad of the shifted views
    static constexpr int BOX_BYTES = 16384;         // 128 rows x 64 bf16, 128-byte swizzled
    static constexpr int P_BYTES = DS_ ? BOX_BYTES : 0;
    static constexpr int NPOOL = DS_ ? (NABUF_ == 1 ? 5 : 3) : (N1_ == 64 ? 5 : 4);  // staging boxes (R mode: also the residual prefetch depth)
    static constexpr int TMEM_COLS = 512;
    static constexpr int D1_COL = 0, D2_COL = 64 * ND1, D3_COL = D2_COL + 256, A2_COL = D3_COL + N1_;
    static_assert(A2_COL + 64 <= 512, "TMEM budget");
    static constexpr int NBAR = 1 + 2 * NABUF + 4 + 2 + 4 + 2 + 2 + 4 * NPOOL;
    static constexpr int SMEM_BYTES =
        1024 + W_BYTES + RING_BYTES + P_BYTES + NPOOL * BOX_BYTES + NBAR * 8 + 16;
    static constexpr int EPI_WARPS = 8;
    static constexpr int THREADS = 128 + EPI_WARPS * 32 + 32;
};
static_assert(BneckCfg<false>::SMEM_BYTES <= 232448, "smem budget");
static_assert(BneckCfg<true>::SMEM_BYTES <= 232448, "smem budget");
static_assert(BneckCfg<true, 64, 1>::SMEM_BYTES <= 232448, "smem budget");
static_assert(BneckCfg<false, 128>::SMEM_BYTES <= 232448, "smem budget");

namespace ptx {
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)),
          "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// pull one box of a 4-D tiled tensor into L2 (no shared-memory destination, no completion signal)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int32_t c0, int32_t c1, int32_t c2,
                                                int32_t c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// cluster-scope release arrive on the leader's barrier: orders this CTA's shared-memory writes
// (already fenced to the async proxy) before the leader's MMAs that read them through the pair
__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) &
                                                                                      kPeerBitMask)
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T over a CTA pair: A = 128 rows per CTA, 16-bit elements packed two
// per 32-bit column (element k of row m: lane m, column k / 2, low half first)
__device__ __forceinline__ void mma_f16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (thread t writes lane t)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0],"
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
}  // namespace ptx

struct BneckParams {
    const float* bias2;   // [64]   conv2 folded BN shift
    const float* bias3;   // [256]  conv3 shift (DS mode: conv3 shift + downsample shift)
    const float* bias1n;  // [64]   next block's conv1 shift (has_next)
};

// Tensor maps (all BF16, 128-byte swizzle):
//   tmA   input t1      [C=64,  W, H, N]  box {64, 64, 4, 1}   (loaded at w = -1, h = h0 - 1: zero halo)
//   tmW2  conv2 weights [64][576]         box {64, 32}
//   tmW3  conv3 weights [256][64]         box {64, 64}
//   tmWds downsample weights [256][64]    box {64, 64}         (DS mode; else unused)
//   tmW1n next conv1 weights [64][256]    box {64, 32}         (has_next; else unused)
//   tmRes R mode: shortcut x [C=256, W, H, N] box {64, 64, 2, 1}; DS mode: block input x [C=64, ...] same box
//   tmY   output y      [C=256, W, H, N]  box {64, W, 1, 1}
//   tmT1n output t1'    [C=64,  W, H, N]  box {64, W, 1, 1}    (has_next; else unused)
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
bneck_l1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW2,
                const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmWds,
                const __grid_constant__ CUtensorMap tmW1n, const __grid_constant__ CUtensorMap tmRes,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmT1n,
                const BneckParams prm, const BneckGeom g) {
    using namespace ptx;
    constexpr bool DS = Cfg::DS;
    constexpr int NABUF = Cfg::NABUF, NPOOL = Cfg::NPOOL;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_w2 = smem;
    uint8_t* smem_w3 = smem_w2 + Cfg::W2_BYTES;
    uint8_t* smem_wds = smem_w3 + Cfg::W3_BYTES;
    uint8_t* smem_w1n = smem_wds + Cfg::WDS_BYTES;
    uint8_t* smem_ring = smem_w1n + Cfg::W1N_BYTES;
    uint8_t* smem_p = smem_ring + Cfg::RING_BYTES;
    uint8_t* smem_pool = smem_p + Cfg::P_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_pool + NPOOL * Cfg::BOX_BYTES);
    uint64_t* w_full = bars;                 // leader: all resident weights of both CTAs landed
    uint64_t* a_full = w_full + 1;           // leader: input tile landed in both CTAs
    uint64_t* a_empty = a_full + NABUF;      // per CTA (multicast commit)
    uint64_t* d1_full = a_empty + NABUF;     // per CTA (multicast commit), [2]
    uint64_t* d1_empty = d1_full + 2;        // leader, 16 arrivals, [2]
    uint64_t* a2_full = d1_empty + 2;        // leader, 16 arrivals, [2]
    uint64_t* d2_full = a2_full + 2;         // per CTA (multicast commit), [2 halves]
    uint64_t* d2_empty = d2_full + 2;        // leader, 16 arrivals, [2 halves]
    uint64_t* p_full = d2_empty + 2;         // leader: shortcut-input tile landed in both CTAs
    uint64_t* p_empty = p_full + 1;          // per CTA (multicast commit)
    uint64_t* d3_full = p_empty + 1;         // per CTA (multicast commit)
    uint64_t* d3_empty = d3_full + 1;        // leader, 16 arrivals
    uint64_t* box_ready = d3_empty + 1;      // per CTA: staging box free (or its residual tile landed)
    uint64_t* c_full = box_ready + NPOOL;    // per CTA, 8 arrivals: box holds finished output
    uint64_t* cx_full = c_full + NPOOL;      // leader, 16 arrivals: box finished in both CTAs (conv1' operand)
    uint64_t* c_mma_done = cx_full + NPOOL;  // per CTA (multicast commit): conv1' has read the box
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int T = (g.tiles - pair + num_pairs - 1) / num_pairs;   // tiles of this pair
    const int IPT = g.has_next ? 4 + Cfg::NT : 4;                  // staging items per tile
    const int items = T * IPT;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmW3);
        if (DS) tma_prefetch_desc(&tmWds);
        tma_prefetch_desc(&tmRes);
        tma_prefetch_desc(&tmY);
        if (g.has_next) {
            tma_prefetch_desc(&tmW1n);
            tma_prefetch_desc(&tmT1n);
        }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < NABUF; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 2 * Cfg::EPI_WARPS);
            mbar_init(&a2_full[i], 2 * Cfg::EPI_WARPS);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], 2 * Cfg::EPI_WARPS);
        }
        mbar_init(p_full, 1);
        mbar_init(p_empty, 1);
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
    // read-past pad of the ring (garbage rows only; keep it defined)
    for (int i = threadIdx.x; i < 64; i += Cfg::THREADS)
        reinterpret_cast<uint4*>(smem_ring + NABUF * Cfg::ATILE_BYTES)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    griddep_launch_dependents();  // see conv_igemm.cuh
    griddep_wait();

    // local tile index -> image, first image row owned by THIS CTA
    auto tile_coords = [&](int it_local, int& img, int& h0) {
        const int t = pair + it_local * num_pairs;
        const int tt = g.reverse ? g.tiles - 1 - t : t;
        img = tt / g.tiles_per_img;
        h0 = (tt - img * g.tiles_per_img) * 4 + static_cast<int>(rank) * 2;
    };

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (elect_one()) {
            const int r32 = static_cast<int>(rank) * 32, r64 = static_cast<int>(rank) * 64;
            const uint32_t wbytes = Cfg::W2_BYTES + Cfg::W3_BYTES + Cfg::WDS_BYTES +
                                    (g.has_next ? Cfg::W1N_BYTES : 0);
            if (rank == 0) mbar_expect_tx(w_full, 2 * wbytes);
            for (int tap = 0; tap < 9; ++tap)
                tma_load_2d_2sm(smem_w2 + tap * Cfg::W2_TAP_BYTES, &tmW2, w_full, tap * 64, r32);
            for (int hf = 0; hf < 2; ++hf) {
                tma_load_2d_2sm(smem_w3 + hf * Cfg::W3_HALF_BYTES, &tmW3, w_full, 0, hf * 128 + r64);
                if (DS) tma_load_2d_2sm(smem_wds + hf * Cfg::W3_HALF_BYTES, &tmWds, w_full, 0, hf * 128 + r64);
            }
            if (g.has_next)
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d_2sm(smem_w1n + kb * Cfg::W1N_KB_BYTES, &tmW1n, w_full, kb * 64,
                                    static_cast<int>(rank) * (Cfg::N1 / 2));
        }
        __syncwarp();
        for (int it = 0; it < T; ++it) {
            int img, h0;
            tile_coords(it, img, h0);
            const int ab = it % NABUF;
            mbar_wait(&a_empty[ab], ((it / NABUF) & 1) ^ 1);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&a_full[ab], 2 * Cfg::ATILE_BYTES);
                // pixels [-1, 63) of input rows h0-1 .. h0+2; out-of-image parts are zero-filled
                tma_load_4d_2sm(smem_ring + ab * Cfg::ATILE_BYTES, &tmA, &a_full[ab], 0, -1, h0 - 1, img);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================================================== conv2 MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc64 = umma_instr_desc(UMMA_FMT_BF16, 256, 64);
            const uint64_t ring_desc = umma_smem_desc(smem_u32(smem_ring), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t w2_desc = umma_smem_desc(smem_u32(smem_w2), 0, 1024, UMMA_LAYOUT_SW128);
            mbar_wait(w_full, 0);
            tc_fence_after();
            for (int i = 0; i < T; ++i) {
                const int buf = i % Cfg::ND1;
                const uint32_t ph = (i / Cfg::ND1) & 1;
                mbar_wait(&d1_empty[buf], ph ^ 1);
                const int ab = i % NABUF;
                mbar_wait(&a_full[ab], (i / NABUF) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + Cfg::D1_COL + buf * 64;
                    const uint64_t a_tile = ring_desc + static_cast<uint64_t>((ab * Cfg::ATILE_BYTES) >> 4);
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            // tap (r, s): the same tile, shifted by r input rows and s pixels (the 128-byte
                            // swizzle is a function of the address bits, see conv3x3_halo.cuh)
                            const uint64_t a_tap = a_tile + static_cast<uint64_t>((r * Cfg::AROW_BYTES + s * 128) >> 4);
                            const uint64_t b_tap = w2_desc + static_cast<uint64_t>(((r * 3 + s) * Cfg::W2_TAP_BYTES) >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_f16_ss_2sm(d_tmem, a_tap + static_cast<uint64_t>(k * 2),
                                               b_tap + static_cast<uint64_t>(k * 2), idesc64, (r | s | k) != 0);
                        }
                    }
                    tc_commit_2sm(&a_empty[ab]);
                    tc_commit_2sm(&d1_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 12) {
        // ===================================================== conv3 / conv1' MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idescn1 = umma_instr_desc(UMMA_FMT_BF16, 256, Cfg::N1);
            constexpr uint32_t idesc128 = umma_instr_desc(UMMA_FMT_BF16, 256, 128);
            const uint64_t w3_desc = umma_smem_desc(smem_u32(smem_w3), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t wds_desc = umma_smem_desc(smem_u32(smem_wds), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t w1n_desc = umma_smem_desc(smem_u32(smem_w1n), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t p_desc = umma_smem_desc(smem_u32(smem_p), 0, 1024, UMMA_LAYOUT_SW128);
            const uint64_t pool_desc = umma_smem_desc(smem_u32(smem_pool), 0, 1024, UMMA_LAYOUT_SW128);
            uint32_t cx_phase_bits = 0;  // bit cs = parity of the next cx_full[cs] wait
            mbar_wait(w_full, 0);
            tc_fence_after();

            // conv3 of tile i, 128-channel half hf -> D2 half hf
            auto conv3 = [&](int i, int hf) {
                const int buf = i & 1;
                if (hf == 0) {
                    mbar_wait(&a2_full[buf], (i >> 1) & 1);
                    if (DS) mbar_wait(p_full, i & 1);
                }
                mbar_wait(&d2_empty[hf], (i & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + Cfg::D2_COL + hf * 128;
                    const uint32_t a_tmem = tmem_base + Cfg::A2_COL + buf * 32;
                    const uint64_t b = w3_desc + static_cast<uint64_t>((hf * Cfg::W3_HALF_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 16 BF16 of K = 8 TMEM columns per instruction
                        mma_f16_ts_2sm(d_tmem, a_tmem + k * 8, b + static_cast<uint64_t>(k * 2), idesc128, k != 0);
                    if (DS) {
                        const uint64_t bd = wds_desc + static_cast<uint64_t>((hf * Cfg::W3_HALF_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss_2sm(d_tmem, p_desc + static_cast<uint64_t>(k * 2),
                                           bd + static_cast<uint64_t>(k * 2), idesc128, 1);
                        if (hf == 1) tc_commit_2sm(p_empty);
                    }
                    tc_commit_2sm(&d2_full[hf]);
                }
                __syncwarp();
            };
            // conv1' K block j of tile i: staging box (item i*IPT + j) x W1n_j -> D3
            auto conv1n = [&](int i, int j) {
                const int cs = (i * IPT + j) % NPOOL;
                if (j == 0) mbar_wait(&d3_empty[0], (i & 1) ^ 1);
                mbar_wait(&cx_full[cs], (cx_phase_bits >> cs) & 1);
                cx_phase_bits ^= 1u << cs;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + Cfg::D3_COL;
                    const uint64_t a = pool_desc + static_cast<uint64_t>((cs * Cfg::BOX_BYTES) >> 4);
                    const uint64_t b = w1n_desc + static_cast<uint64_t>((j * Cfg::W1N_KB_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_f16_ss_2sm(d_tmem, a + static_cast<uint64_t>(k * 2), b + static_cast<uint64_t>(k * 2),
                                       idescn1, (j | k) != 0);
                    tc_commit_2sm(&c_mma_done[cs]);
                    if (j == 3) tc_commit_2sm(d3_full);
                }
                __syncwarp();
            };

            conv3(0, 0);
            conv3(0, 1);
            for (int i = 0; i < T; ++i) {
                for (int j = 0; j < 4; ++j) {
                    if (g.has_next) conv1n(i, j);
                    if (j == 1 && i + 1 < T) conv3(i + 1, 0);
                    if (j == 3 && i + 1 < T) conv3(i + 1, 1);
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================== shortcut-input loader (DS mode, both CTAs)
        if (DS) {
            for (int i = 0; i < T; ++i) {
                int img, h0;
                tile_coords(i, img, h0);
                mbar_wait(p_empty, (i & 1) ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(p_full, 2 * Cfg::BOX_BYTES);
                    tma_load_4d_2sm(smem_p, &tmRes, p_full, 0, 0, h0, img);
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ===================================================== store warp (both CTAs)
        // item = tile * IPT + sub: sub 0..3 = 64-channel boxes of y, sub 4 = t1' (has_next).
        // Stores are pipelined: the box of item k is recycled only after the store of item k+1 has been
        // issued (cp.async.bulk.wait_group.read 1), so one store is always in flight behind the one
        // being waited for. Optional (debug bit 4; measured SLOWER on B200, 167 -> 190 us, so off):
        // residual tiles pulled into L2 PF_TILES tiles ahead with cp.async.bulk.prefetch.tensor.
        constexpr int PF_TILES = 3;
        auto prefetch_tile = [&](int it_local) {  // whole 128-row x 256-channel residual tile -> L2
            int img, h0;
            tile_coords(it_local, img, h0);
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) tma_prefetch_4d(&tmRes, sub * 64, 0, h0, img);
        };
        auto prepare = [&](int item) {  // make box (item % NPOOL) ready for `item`
            const int cs = item % NPOOL;
            const int it_local = item / IPT, sub = item - it_local * IPT;
            if (!DS && sub < 4 && !(g.debug & 1)) {
                int img, h0;
                tile_coords(it_local, img, h0);
                mbar_expect_tx(&box_ready[cs], Cfg::BOX_BYTES);
                tma_load_4d(smem_pool + cs * Cfg::BOX_BYTES, &tmRes, &box_ready[cs], sub * 64, 0, h0, img);
                if (sub == 0 && it_local + PF_TILES < T && (g.debug & 4)) prefetch_tile(it_local + PF_TILES);
            } else {
                mbar_arrive(&box_ready[cs]);
            }
        };
        if (elect_one()) {
            if (!DS && (g.debug & 4))
                for (int t = 1; t < PF_TILES && t < T; ++t) prefetch_tile(t);
            for (int i = 0; i < NPOOL && i < items; ++i) prepare(i);
        }
        __syncwarp();
        uint32_t md_phase_bits = 0;  // bit cs = parity of the next c_mma_done[cs] wait
        // recycle the box of `item` (its store has been read out of shared memory)
        auto recycle = [&](int item) {
            const int cs = item % NPOOL;
            const int sub = item % IPT;
            if (g.has_next && sub < 4) {  // conv1' must have consumed the box as well
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
                int img, h0;
                tile_coords(it_local, img, h0);
                if (g.debug & 8) {  // timing experiment: every store hits the same (L2-resident) tile
                    img = static_cast<int>(blockIdx.x) % g.N;
                    h0 = 0;
                }
                const uint8_t* box = smem_pool + cs * Cfg::BOX_BYTES;
                if (g.debug & 2) {
                } else if (sub < 4) {
                    tma_store_4d(&tmY, box, sub * 64, 0, h0, img);
                    tma_store_4d(&tmY, box + Cfg::PITCH * 128, sub * 64, 0, h0 + 1, img);
                } else {
                    tma_store_4d(&tmT1n, box, (sub - 4) * 64, 0, h0, img);
                    tma_store_4d(&tmT1n, box + Cfg::PITCH * 128, (sub - 4) * 64, 0, h0 + 1, img);
                }
                tma_store_commit();
                if (g.debug & 32)
                    tma_store_wait_read<0>();
                else
                    tma_store_wait_read<1>();  // every store but the one just issued has left shared memory
            }
            __syncwarp();
            if (g.debug & 32)
                recycle(item);
            else if (item > 0)
                recycle(item - 1);
        }
        if (elect_one()) tma_store_wait_read<0>();
        __syncwarp();
        if (items > 0 && !(g.debug & 32)) recycle(items - 1);
        if (elect_one()) tma_store_wait_all<0>();
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ===================================================== epilogue (both CTAs)
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;  // which 32-column half of a 64-column group
        const int row_in_tile = q * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row_in_tile & 7);
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t row_off = static_cast<uint32_t>(row_in_tile) * 128;

        auto publish = [&]() {
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
        };
        // D1[i & 1] + bias2 -> ReLU -> BF16 pairs -> A2[i & 1] in TMEM (the A operand of conv3)
        auto E1 = [&](int i) {
            const int buf = i & 1;            // A2 buffer
            const int dbuf = i % Cfg::ND1;    // D1 buffer
            mbar_wait(&d1_full[dbuf], (i / Cfg::ND1) & 1);
            tc_fence_after();
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(lane_base + Cfg::D1_COL + dbuf * 64 + h * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
            const float* b32 = prm.bias2 + h * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(b32 + j * 4));
                float x0 = __uint_as_float(v[j * 4 + 0]), x1 = __uint_as_float(v[j * 4 + 1]);
                float x2 = __uint_as_float(v[j * 4 + 2]), x3 = __uint_as_float(v[j * 4 + 3]);
                add2(x0, x1, b.x, b.y);
                add2(x2, x3, b.z, b.w);
                pk[j * 2 + 0] = pack_bf16x2_relu(x0, x1);
                pk[j * 2 + 1] = pack_bf16x2_relu(x2, x3);
            }
            tmem_st_32x16(lane_base + Cfg::A2_COL + buf * 32 + h * 16, pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_leader(&d1_empty[dbuf]);
                mbar_arrive_leader(&a2_full[buf]);
            }
        };
        // D2 + bias3 (+ residual) -> ReLU -> staging boxes of items i*IPT + 0..3
        auto E2 = [&](int i) {
            // (two nested loops on purpose: indexing the half barriers with j >> 1 inside one loop over
            // j = 0..3 was strength-reduced by nvcc 12.9 to base + 4*j — misaligned for odd j)
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                mbar_wait(&d2_full[hf], i & 1);
                tc_fence_after();
#pragma unroll 1
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = hf * 2 + jj;
                    const int item = i * IPT + j;
                    const int cs = item % NPOOL;
                    mbar_wait(&box_ready[cs], (item / NPOOL) & 1);
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32x32(lane_base + Cfg::D2_COL + j * 64 + h * 32, v);
                    tmem_ld_wait();
                    epilogue_chunk<2>(v, smem_pool + cs * Cfg::BOX_BYTES + row_off, static_cast<uint32_t>(h * 4), swz,
                                      prm.bias3 + j * 64 + h * 32, DS ? 0 : 1, 1);
                    publish();
                    if (lane == 0) {
                        mbar_arrive(&c_full[cs]);
                        if (g.has_next) mbar_arrive_leader(&cx_full[cs]);
                    }
                }
                if (lane == 0) mbar_arrive_leader(&d2_empty[hf]);  // this warp has drained half hf
            }
        };
        // D3 + bias1' -> ReLU -> staging boxes of items i*IPT + 4 .. (NT boxes of 64 channels)
        auto E3 = [&](int i) {
            mbar_wait(d3_full, i & 1);
            tc_fence_after();
#pragma unroll 1
            for (int jb = 0; jb < Cfg::NT; ++jb) {
                const int item = i * IPT + 4 + jb;
                const int cs = item % NPOOL;
                mbar_wait(&box_ready[cs], (item / NPOOL) & 1);
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(lane_base + Cfg::D3_COL + jb * 64 + h * 32, v);
                tmem_ld_wait();
                epilogue_chunk<2>(v, smem_pool + cs * Cfg::BOX_BYTES + row_off, static_cast<uint32_t>(h * 4), swz,
                                  prm.bias1n + jb * 64 + h * 32, 0, 1);
                publish();
                if (lane == 0) mbar_arrive(&c_full[cs]);
            }
            if (lane == 0) mbar_arrive_leader(d3_empty);
        };

        if (T > 0) E1(0);
        for (int i = 0; i < T; ++i) {
            if (i + 1 < T) E1(i + 1);
            if (g.has_next && i > 0) E3(i - 1);
            E2(i);
        }
        if (g.has_next && T > 0) E3(T - 1);
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
