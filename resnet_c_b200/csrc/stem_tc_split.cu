// stem_tc_split.cu — the network stem on tensor cores for the TF32 path:
//   conv 7x7/2 pad 3 (3 -> 64) + folded BN + ReLU + max-pool 3x3/2 pad 1  (main.cu:176-192)
// with FP32-class accuracy out of BF16 MMAs.
//
// The TF32 model must keep its logits within 1e-3 of the FP32 reference, so its stem cannot run in
// plain BF16 (2e-3 at the stem output alone). kind::tf32 cannot use the Hankel descriptor of
// stem_tc.cu either (16-byte core-matrix rows are 4 fp32 = one pixel, but the conv stride is two
// pixels). Instead every operand is split into two BF16 terms, x = x_hi + x_lo, w = w_hi + w_lo, and
//   x*w  ~=  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi          (dropped term ~2^-16 relative)
// is accumulated by THREE tcgen05.mma per K step into the same FP32 TMEM accumulator.
//
// Structure = stem_tc.cu's (third form of round 2): weights (hi and lo: 36 A blocks = 288 TMEM columns) as the A
// operand in TMEM, the image rows (two rings: hi and lo terms) as the B operand, conv-row pairs, two alternating MMA
// issuer warps, loader warps that build the NHWC4 rows from the caller's FP32 NCHW tensor, independent per-warp
// epilogues with their own TMA stores. Differences from the BF16 kernel:
//   * 3 x 18 MMAs per pair; two accumulator slots (2 x 112 + 288 = 512 TMEM columns);
//   * the output is FP32 NHWC rounded to TF32 (the activation type of the TF32 path): the partner-lane exchange moves
//     one value per register (lanes 0..15 finish the even pooled columns of their half, lanes 16..31 the odd ones).
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "sm100_ptx.cuh"
#include "stem_tc_common.cuh"
#include "tensormap.h"

namespace rnb {

namespace {

using namespace stemtc;

constexpr int NCH = 8;                                 // ring depth in chunks (pair k reads chunks k, k+1, k+2)
constexpr int RING_BYTES = NCH * CHUNK_BYTES + 512;    // one ring (hi or lo) + slack
constexpr int EPI_THREADS = 256;
constexpr int THREADS = 128 + EPI_THREADS + 32;        // warps 0..3, eight epilogue warps, the second MMA issuer
constexpr int S_NBLK = 36;                             // A blocks: (input row t, K step i) x (hi, lo)
constexpr int S_NSLOT = 2;                             // TMEM: D slots of 112 columns at 0 and 112; weights at 224 .. 511
constexpr int S_D_PITCH = 112;
constexpr int S_A_COL = S_NSLOT * S_D_PITCH;
constexpr int S_WSTAGE_BYTES = 28 * 64;                // one epilogue warp's block: [28 pooled columns][16 ch] fp32
constexpr int S_STAGE_BYTES = 8 * S_WSTAGE_BYTES;      // x 2 buffers
constexpr int NBAR = 2 * NCH + 2 * S_NSLOT;
constexpr int SMEM = 1024 + 2 * S_STAGE_BYTES + 2 * RING_BYTES + NBAR * 8 + 16;
constexpr size_t S_W_BYTES = static_cast<size_t>(S_NBLK) * 128 * 8 * 4;
static_assert(SMEM <= 232448, "smem budget");
static_assert(S_A_COL + S_NBLK * 8 <= 512, "TMEM budget");

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// w [64][3][7][7] fp32 + BN -> wt [block = (2t + i)*2 + part][m][8] u32 (part 0 = hi, 1 = lo BF16 term): row
// m = 32q + 16*rowsel + ocl holds, as 16 BF16 (two per word, low half first), window pixels 4i .. 4i+3 x 4 channels of
// filter row t (rowsel 0, t <= 6) or t - 2 (rowsel 1, t >= 2) of output channel 16q + ocl; everything else is zero.
__global__ void stem_pack_weights_split_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                               const float* __restrict__ bn_b, const float* __restrict__ bn_m,
                                               const float* __restrict__ bn_v, uint32_t* __restrict__ wt,
                                               float* __restrict__ bias) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S_NBLK * 128 * 8) return;
    const int col = idx & 7, m = (idx >> 3) & 127, blk = idx >> 10;
    const int part = blk & 1, t = blk >> 2, i = (blk >> 1) & 1;
    const int q = m >> 5, rs = (m >> 4) & 1, oc = 16 * q + (m & 15);
    const int kh = rs ? t - 2 : t;
    const bool row_ok = rs ? t >= 2 : t <= 6;
    double scale = 1.0, shift = 0.0;
    if (bn_w) {
        scale = static_cast<double>(bn_w[oc]) / sqrt(static_cast<double>(bn_v[oc]) + 1e-5);
        shift = static_cast<double>(bn_b[oc]) - static_cast<double>(bn_m[oc]) * scale;
    }
    if (blk == 0 && col == 0 && rs == 0) bias[oc] = static_cast<float>(shift);
    float v[2];
    for (int h = 0; h < 2; ++h) {
        const int e16 = 2 * col + h, pix = 4 * i + (e16 >> 2), c = e16 & 3, kw = pix - 1;
        float x = 0.f;
        if (row_ok && kw >= 0 && kw < 7 && c < 3)
            x = static_cast<float>(static_cast<double>(w[((oc * 3 + c) * 7 + kh) * 7 + kw]) * scale);
        const float hi = __bfloat162float(__float2bfloat16_rn(x));
        v[h] = part ? x - hi : hi;
    }
    wt[idx] = pack_bf16x2(v[0], v[1]);
}

// One chunk (image rows 4c-4 .. 4c-1 of image b) of BOTH rings, built by one warp from the FP32 NCHW tensor:
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (the difference is exact in FP32). See stem_fill_chunk_f32.
__device__ __forceinline__ void stem_fill_chunk_split(uint32_t dst_hi, uint32_t dst_lo, const float* __restrict__ x,
                                                      int b, int c, int lane) {
    if (c == 0 || c == PAIRS + 1) {
        for (int i = lane; i < CHUNK_BYTES / 16; i += 32) {
            st_shared_v4(dst_hi + 16 * i, 0, 0, 0, 0);
            st_shared_v4(dst_lo + 16 * i, 0, 0, 0, 0);
        }
        return;
    }
    const float* img = x + (1LL * b * 3 * IMG + (4 * c - 4)) * IMG;
    if (lane < 3 && c + 3 <= PAIRS) {   // the chunk this warp fills next, requested into L2 now (see stem_tc.cu)
        const float* nxt = img + 1LL * lane * IMG * IMG + 12 * IMG;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"(4 * IMG * 4) : "memory");
    }
    float4 v[7][3];
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const int id = it * 32 + lane, rr = id / 56, g = id - rr * 56;
        const float* src = img + rr * IMG + 4 * g;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            v[it][ch] = __ldg(reinterpret_cast<const float4*>(src + 1LL * ch * IMG * IMG));
    }
    const int first = (lane >> 2) & 1;  // bank-conflict-free store order, as in stem_fill_chunk_f32
#pragma unroll
    for (int it = 0; it < 7; ++it) {
        const int id = it * 32 + lane, rr = id / 56, g = id - rr * 56;
        const uint32_t off = rr * ROW_BYTES + 32 * g + 32;
        float px[4][3], hi[4][3], lo[4][3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            px[0][ch] = v[it][ch].x; px[1][ch] = v[it][ch].y; px[2][ch] = v[it][ch].z; px[3][ch] = v[it][ch].w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                hi[p][ch] = __bfloat162float(__float2bfloat16_rn(px[p][ch]));
                lo[p][ch] = px[p][ch] - hi[p][ch];
            }
        uint32_t wh[8], wl[8];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            wh[2 * p] = pack_bf16x2(hi[p][0], hi[p][1]); wh[2 * p + 1] = pack_bf16x2(hi[p][2], 0.f);
            wl[2 * p] = pack_bf16x2(lo[p][0], lo[p][1]); wl[2 * p + 1] = pack_bf16x2(lo[p][2], 0.f);
        }
        const int f4 = 4 * first, s4 = 4 * (first ^ 1);
        st_shared_v4(dst_hi + off + 4 * f4, first ? wh[4] : wh[0], first ? wh[5] : wh[1], first ? wh[6] : wh[2], first ? wh[7] : wh[3]);
        st_shared_v4(dst_hi + off + 4 * s4, first ? wh[0] : wh[4], first ? wh[1] : wh[5], first ? wh[2] : wh[6], first ? wh[3] : wh[7]);
        st_shared_v4(dst_lo + off + 4 * f4, first ? wl[4] : wl[0], first ? wl[5] : wl[1], first ? wl[6] : wl[2], first ? wl[7] : wl[3]);
        st_shared_v4(dst_lo + off + 4 * s4, first ? wl[0] : wl[4], first ? wl[1] : wl[5], first ? wl[2] : wl[6], first ? wl[3] : wl[7]);
    }
}

__global__ void __launch_bounds__(THREADS, 1)
stem_tc_split_kernel(const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ x,
                     const uint32_t* __restrict__ wt, const float* __restrict__ bias, int B) {
    using namespace ptx;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* stage = smem;                          // 2 x S_STAGE_BYTES
    uint8_t* ring = smem + 2 * S_STAGE_BYTES;       // [hi ring][lo ring]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * RING_BYTES);
    uint64_t* ch_full = bars;                       // [NCH]
    uint64_t* ch_empty = bars + NCH;                // [NCH]
    uint64_t* acc_full = bars + 2 * NCH;            // [S_NSLOT]
    uint64_t* acc_empty = bars + 2 * NCH + S_NSLOT; // [S_NSLOT]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_pairs = B * PAIRS;
    const int ppc = num_pairs / static_cast<int>(gridDim.x), prem = num_pairs % static_cast<int>(gridDim.x);
    const int p_begin = static_cast<int>(blockIdx.x) * ppc + min(static_cast<int>(blockIdx.x), prem);
    const int p_end = p_begin + ppc + (static_cast<int>(blockIdx.x) < prem ? 1 : 0);

    if (threadIdx.x == 32) {
        for (int i = 0; i < NCH; ++i) {
            mbar_init(&ch_full[i], 32);
            mbar_init(&ch_empty[i], 2);   // both MMA issuers release every chunk
        }
        for (int i = 0; i < S_NSLOT; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < 2 * RING_BYTES / 16; i += THREADS)   // halo columns, slack
        reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp >= 4 && warp < 12) {
        // weights -> TMEM (constants): the two warps of a lane quarter write 18 A blocks each, lane = row m
        const int q = warp & 3, m = 32 * q + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + S_A_COL;
        for (int blk = (warp - 4) >> 2; blk < S_NBLK; blk += 2) {
            const uint4* src = reinterpret_cast<const uint4*>(wt + (blk * 128 + m) * 8);
            const uint4 lo = __ldg(src), hi = __ldg(src + 1);
            const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            tmem_st_32x8(lane_base + blk * 8, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0 || warp == 2 || warp == 3) {
        // ===================================================== loaders: chunk n belongs to loader n % 3
        const int widx = warp == 0 ? 0 : warp - 1;
        for (StemSteps st(p_begin, p_end); !st.done(); st.next()) {
            const int b = st.b(), nnew = st.new_chunks();
            const int c_first = st.k() + 3 - nnew;
            for (int i = 0; i < nnew; ++i) {
                const int n = st.cn + i, c = c_first + i;
                if ((n % 3) != widx) continue;
                mbar_wait(&ch_empty[n & (NCH - 1)], ((n / NCH) & 1) ^ 1);
                const uint32_t dst = smem_u32(ring) + (n & (NCH - 1)) * CHUNK_BYTES;
                stem_fill_chunk_split(dst, dst + RING_BYTES, x, b, c, lane);
                fence_proxy_async_smem();
                mbar_arrive(&ch_full[n & (NCH - 1)]);
            }
        }
    } else if (warp == 1 || warp == 12) {
        // ===================================================== MMA issuers (alternate steps; see stem_tc.cu)
        const int X = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, 112);
        const uint64_t b_desc0 = umma_smem_desc(smem_u32(ring), 16, 128, UMMA_LAYOUT_NONE);
        constexpr uint64_t X_LO = RING_BYTES >> 4;   // hi -> lo ring
        int rel = 0;   // next chunk this warp has to release
        for (StemSteps st(p_begin, p_end); !st.done(); st.next()) {
            const int j = st.step;
            if ((j & 1) != X) continue;
            const int cn_after = st.cn + st.new_chunks();
            StemSteps la = st;   // first chunk of this warp's next step (j + 2), or the total number of chunks
            la.next();
            int rel_end = la.cn;
            if (!la.done()) {
                la.next();
                rel_end = la.done() ? la.cn : la.cn + la.new_chunks() - 3;
            }
            for (int m = cn_after - 3; m < cn_after; ++m) mbar_wait(&ch_full[m & (NCH - 1)], (m / NCH) & 1);
            mbar_wait(&acc_empty[j % S_NSLOT], ((j / S_NSLOT) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + (j % S_NSLOT) * S_D_PITCH;
                const uint32_t a0 = tmem_base + S_A_COL;
                const uint64_t c0 = b_desc0 + static_cast<uint64_t>((((cn_after - 3) & (NCH - 1)) * CHUNK_BYTES) >> 4);
                const uint64_t c1 = b_desc0 + static_cast<uint64_t>((((cn_after - 2) & (NCH - 1)) * CHUNK_BYTES) >> 4);
                const uint64_t c2 = b_desc0 + static_cast<uint64_t>((((cn_after - 1) & (NCH - 1)) * CHUNK_BYTES) >> 4);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint64_t row = (t < 3 ? c0 : t < 7 ? c1 : c2) + static_cast<uint64_t>((((1 + t) & 3) * ROW_BYTES) >> 4);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint32_t a_hi = a0 + ((2 * t + i) * 2) * 8, a_lo = a_hi + 8;
                        mma_f16_ts(d, a_hi, row + 2 * i, idesc, (t | i) != 0);      // w_hi * x_hi
                        mma_f16_ts(d, a_lo, row + 2 * i, idesc, 1);                  // w_lo * x_hi
                        mma_f16_ts(d, a_hi, row + 2 * i + X_LO, idesc, 1);           // w_hi * x_lo
                    }
                }
                tc_commit(&acc_full[j % S_NSLOT]);
                for (int n = rel; n < rel_end; ++n) tc_commit(&ch_empty[n & (NCH - 1)]);
            }
            rel = rel_end;
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================================================== epilogue (FP32 / TF32-rounded output)
        // warp quarter q = warp & 3: lanes 0..15 = conv row a (2k), lanes 16..31 = conv row b (2k+1) of channels
        // 16q .. 16q+15; the two warps of a quarter split the pooled columns (h = 0: 0..27, h = 1: 28..55)
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int rs = lane >> 4, oc = 16 * q + (lane & 15);
        const float bias_c = __ldg(bias + oc);
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (h ? 48 : 0);
        uint8_t* const stg = stage + (warp - 4) * (2 * S_WSTAGE_BYTES);
        float carry[14];  // conv row 2k-1 (the previous pair's row b) at the 14 pooled columns this lane finishes
#pragma unroll
        for (int u = 0; u < 14; ++u) carry[u] = 0.f;
        int vb = 0;
        uint32_t c0[32], c1[32];   // 64 output columns of this lane's conv row: 0..63 (h = 0) or 48..111 (h = 1)
        StemSteps st(p_begin, p_end);
        if (!st.done()) {
            mbar_wait(&acc_full[0], 0);
            tc_fence_after();
            tmem_ld_32x32(lane_addr, c0);
            tmem_ld_32x32(lane_addr + 32, c1);
        }
        while (!st.done()) {
            const int b = st.b(), k = st.k(), j = st.step;
            const bool warm = st.warm;
            st.next();
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&acc_empty[j % S_NSLOT]);
            auto col = [&](int i) { return __uint_as_float(i < 32 ? c0[i] : c1[i - 32]); };
            // horizontal 3-tap max, bias, ReLU, TF32 rounding (all monotone: they commute with the vertical max)
            float y[28];
            if (h == 0) {
#pragma unroll
                for (int jj = 0; jj < 28; ++jj) {
                    const int i0 = 2 * jj;
                    y[jj] = rna_tf32(fmaxf(fmaxf(fmaxf(col(i0), col(i0 + 1)), i0 > 0 ? col(i0 - 1) : col(i0)) + bias_c, 0.f));
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 28; ++jj) {
                    const int i0 = 2 * jj + 8;
                    y[jj] = rna_tf32(fmaxf(fmaxf(fmaxf(col(i0), col(i0 + 1)), col(i0 - 1)) + bias_c, 0.f));
                }
            }
            if (!st.done()) {   // the next pair's accumulators
                const int jn = st.step;
                mbar_wait(&acc_full[jn % S_NSLOT], (jn / S_NSLOT) & 1);
                tc_fence_after();
                tmem_ld_32x32(lane_addr + (jn % S_NSLOT) * S_D_PITCH, c0);
                tmem_ld_32x32(lane_addr + (jn % S_NSLOT) * S_D_PITCH + 32, c1);
            }
            if (k == 0) {
#pragma unroll
                for (int u = 0; u < 14; ++u) carry[u] = 0.f;   // no conv row above the image (ReLU output is >= 0)
            }
            // vertical max with the partner lane (the pair's other conv row): lanes 0..15 finish the even pooled
            // columns of this warp's 28, lanes 16..31 the odd ones; each lane sends what the other one finishes
            float res[14];
#pragma unroll
            for (int u = 0; u < 14; ++u) {
                const float mine = rs ? y[2 * u + 1] : y[2 * u];
                const float other = __shfl_xor_sync(0xffffffffu, rs ? y[2 * u] : y[2 * u + 1], 16);
                const float rowb = rs ? mine : other;
                res[u] = fmaxf(fmaxf(mine, other), carry[u]);
                carry[u] = rowb;
            }
            if (warm) continue;  // the pair before this CTA's range: only its second conv row was wanted
            if (lane == 0) tma_store_wait_read<1>();   // the store that last read this staging buffer (two pairs ago)
            __syncwarp();
            {
                // [28 pooled columns][16 ch] fp32, 64-byte rows: lanes 0..15 write row 2u, lanes 16..31 row 2u+1 —
                // 32 consecutive words per instruction
                const uint32_t sbase = smem_u32(stg) + vb * S_WSTAGE_BYTES + rs * 64 + (lane & 15) * 4;
#pragma unroll
                for (int u = 0; u < 14; ++u)
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + 2 * u * 64), "r"(__float_as_uint(res[u])) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_out, stg + vb * S_WSTAGE_BYTES, 16 * q, (b * POOL + k) * POOL + 28 * h);
                tma_store_commit();
            }
            vb ^= 1;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// no layout pre-pass: the scratch tensor of the two-launch form is not used (a token size keeps the callers'
// allocation paths unchanged)
size_t stem_tc_split_packed_input_bytes(int) { return 256; }
size_t stem_tc_split_packed_weight_bytes() { return S_W_BYTES; }

cudaError_t stem_tc_split_init() {
    return cudaFuncSetAttribute(stem_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
}

cudaError_t launch_stem_tc_split_pack_weights(const float* w, const float* bn_w, const float* bn_b,
                                              const float* bn_m, const float* bn_v, void* wk, float* bias,
                                              cudaStream_t s) {
    stem_pack_weights_split_kernel<<<(S_NBLK * 128 * 8 + 255) / 256, 256, 0, s>>>(w, bn_w, bn_b, bn_m, bn_v,
                                                                               static_cast<uint32_t*>(wk), bias);
    return cudaGetLastError();
}

// part 0: nothing (kept so that the callers' two-part timing stays uniform); part 1: x fp32 NCHW -> out NHWC fp32
// [B,56,56,64].
cudaError_t launch_stem_tc_split_part(int part, const float* x, void* /*xp*/, const void* wk, const float* bias,
                                      void* out, int B, cudaStream_t s) {
    if (part == 0) return cudaSuccess;
    const int pairs = B * PAIRS;
    const int grid = pairs < num_sms() ? pairs : num_sms();
    // output [B*56*56 pooled pixels][64 ch] fp32; one store = 28 pooled pixels x 16 channels (an epilogue warp's block)
    CUtensorMap tm;
    const uint64_t dims[2] = {64, 1ull * B * POOL * POOL}, strides[1] = {256};
    const uint32_t box[2] = {16, 28};
    if (make_tiled_nd(&tm, TmDtype::F32, out, 2, dims, strides, box, false) != 0) return cudaErrorInvalidValue;
    stem_tc_split_kernel<<<grid, THREADS, SMEM, s>>>(tm, x, static_cast<const uint32_t*>(wk), bias, B);
    return cudaGetLastError();
}

}  // namespace rnb
