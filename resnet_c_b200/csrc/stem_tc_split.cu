// stem_tc_split.cu — the network stem on tensor cores for the TF32 path:
//   conv 7x7/2 pad 3 (3 -> 64) + folded BN + ReLU + max-pool 3x3/2 pad 1  (main.cu:176-192)
// with FP32-class accuracy out of BF16 MMAs.
//
// The TF32 model must keep its logits within 1e-3 of the FP32 reference, so its stem cannot run in
// plain BF16 (2e-3 at the stem output alone). kind::tf32 cannot use the Hankel descriptor of
// stem_tc.cu either (16-byte core-matrix rows are 4 fp32 = one pixel, but the conv stride is two
// pixels). Instead every operand is split into two BF16 terms, x = x_hi + x_lo, w = w_hi + w_lo, and
//   x*w  ~=  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi          (dropped term ~2^-16 relative)
// is accumulated by THREE tcgen05.mma per K step into the same FP32 TMEM accumulator. Input layout,
// Hankel descriptors, work units and warp roles are exactly those of stem_tc.cu; differences:
//   * the pre-pass writes two packed images (hi, lo); a unit is two 35 KB bulk copies;
//   * weights are two 28 KB matrices;
//   * the output is FP32 NHWC rounded to TF32 (the activation type of the TF32 path), so the
//     horizontal-pool staging row is 256 bytes and there is room for only one staging buffer.
// The stem MMAs are cheap (the layer is 3 % of the network's FLOPs), tripling them costs ~0.15 ms per
// 256-batch against 2.2 ms for the FP32 CUDA-core stem it replaces.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "internal.h"
#include "sm100_ptx.cuh"

namespace rnb {

namespace {

constexpr int IMG = 224;
constexpr int PAD_W = 232;
constexpr int PAD_H = 235;
constexpr int ROW_BYTES = PAD_W * 8;
constexpr int CONV = 112, POOL = 56;
constexpr int ROWS_PER_UNIT = 7;
constexpr int POOLED_PER_UNIT = 3;
constexpr int UNITS_PER_IMG = (POOL + POOLED_PER_UNIT - 1) / POOLED_PER_UNIT;
constexpr int IN_ROWS = 2 * ROWS_PER_UNIT + 5;
constexpr int IN_BYTES = IN_ROWS * ROW_BYTES;                            // 35264
constexpr int IN_SLOT_BYTES = ((IN_BYTES + 512 + 1023) / 1024) * 1024;   // 35840 (hi or lo)
constexpr int W_BYTES = 28 * 1024;                                       // hi or lo
constexpr int VROW = 256;                                                // 64 ch x fp32
constexpr int VBUF_BYTES = 112 * VROW;
constexpr int NBAR = 4 + 2 * ROWS_PER_UNIT;
constexpr int EPI_THREADS = 256;
constexpr int THREADS = 128 + EPI_THREADS;
constexpr int SMEM = 1024 + 2 * 2 * IN_SLOT_BYTES + 2 * W_BYTES + VBUF_BYTES + NBAR * 8 + 16;
static_assert(SMEM <= 232448, "smem budget");

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// x [B,3,224,224] fp32 -> xp[0] = hi, xp[1] = lo, each [B][235][232][4] bf16 (see stem_tc.cu).
__global__ void stem_pack_split_kernel(const float* __restrict__ x, uint2* __restrict__ xp_hi,
                                       uint2* __restrict__ xp_lo, int B) {
    const int64_t total = 1LL * B * PAD_H * PAD_W;
    for (int64_t i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total;
         i += 1LL * gridDim.x * blockDim.x) {
        const int pw = static_cast<int>(i % PAD_W);
        int64_t t = i / PAD_W;
        const int pr = static_cast<int>(t % PAD_H);
        const int b = static_cast<int>(t / PAD_H);
        const int ih = pr - 5, iw = pw - 3;
        uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
        if (ih >= 0 && ih < IMG && iw >= 0 && iw < IMG) {
            const float* p = x + (1LL * b * 3 * IMG + ih) * IMG + iw;
            float v[3], h[3], l[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c] = __ldg(p + 1LL * c * IMG * IMG);
                h[c] = __bfloat162float(__float2bfloat16_rn(v[c]));
                l[c] = v[c] - h[c];  // exact in fp32
            }
            __nv_bfloat162 a = __floats2bfloat162_rn(h[0], h[1]), c2 = __floats2bfloat162_rn(h[2], 0.f);
            hi.x = *reinterpret_cast<uint32_t*>(&a);
            hi.y = *reinterpret_cast<uint32_t*>(&c2);
            a = __floats2bfloat162_rn(l[0], l[1]);
            c2 = __floats2bfloat162_rn(l[2], 0.f);
            lo.x = *reinterpret_cast<uint32_t*>(&a);
            lo.y = *reinterpret_cast<uint32_t*>(&c2);
        }
        xp_hi[i] = hi;
        xp_lo[i] = lo;
    }
}

// Folded weights split into hi / lo BF16 matrices, layout [kh*4+j][oc][e] (see stem_tc.cu).
__global__ void stem_pack_weights_split_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                               const float* __restrict__ bn_b, const float* __restrict__ bn_m,
                                               const float* __restrict__ bn_v, __nv_bfloat16* __restrict__ wk_hi,
                                               __nv_bfloat16* __restrict__ wk_lo, float* __restrict__ bias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 28 * 64 * 8) return;
    const int e = i & 7, oc = (i >> 3) & 63, chunk = i >> 9;
    const int kh = chunk >> 2, j = chunk & 3;
    const int kw = 2 * j + (e >> 2), c = e & 3;
    double scale = 1.0, shift = 0.0;
    if (bn_w) {
        scale = static_cast<double>(bn_w[oc]) / sqrt(static_cast<double>(bn_v[oc]) + 1e-5);
        shift = static_cast<double>(bn_b[oc]) - static_cast<double>(bn_m[oc]) * scale;
    }
    float v = 0.f;
    if (kw < 7 && c < 3) v = static_cast<float>(static_cast<double>(w[((oc * 3 + c) * 7 + kh) * 7 + kw]) * scale);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    wk_hi[i] = h;
    wk_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    if (chunk == 0 && e == 0) bias[oc] = static_cast<float>(shift);
}

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            ptx::smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
stem_tc_split_kernel(const uint8_t* __restrict__ xp_hi, const uint8_t* __restrict__ xp_lo,
                     const uint8_t* __restrict__ wk_hi, const uint8_t* __restrict__ wk_lo,
                     const float* __restrict__ bias, float* __restrict__ out, int B) {
    using namespace ptx;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    // input slots: [slot 0 hi][slot 0 lo][slot 1 hi][slot 1 lo]
    uint8_t* in_slot = smem;
    uint8_t* wsm = smem + 4 * IN_SLOT_BYTES;   // [hi 28 KB][lo 28 KB]
    uint8_t* vbuf = wsm + 2 * W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(vbuf + VBUF_BYTES);
    uint64_t* in_full = bars;
    uint64_t* in_empty = bars + 2;
    uint64_t* slot_full = bars + 4;
    uint64_t* slot_empty = bars + 4 + ROWS_PER_UNIT;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5;
    const int num_units = B * UNITS_PER_IMG;
    // contiguous unit range per CTA; the conv row shared with the previous unit of the same image is carried in
    // registers by the epilogue instead of being recomputed (see stem_tc.cu)
    const int upc = num_units / static_cast<int>(gridDim.x), urem = num_units % static_cast<int>(gridDim.x);
    const int u_begin = static_cast<int>(blockIdx.x) * upc + min(static_cast<int>(blockIdx.x), urem);
    const int u_end = u_begin + upc + (static_cast<int>(blockIdx.x) < urem ? 1 : 0);
    auto continues = [&](int u) { return u > u_begin && (u % UNITS_PER_IMG) != 0; };

    if (threadIdx.x == 32) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&in_full[i], 1);
            mbar_init(&in_empty[i], 1);
        }
        for (int i = 0; i < ROWS_PER_UNIT; ++i) {
            mbar_init(&slot_full[i], 1);
            mbar_init(&slot_empty[i], EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < W_BYTES / 16; i += THREADS) {
        reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(wk_hi) + i);
        reinterpret_cast<uint4*>(wsm + W_BYTES)[i] = __ldg(reinterpret_cast<const uint4*>(wk_lo) + i);
    }
    constexpr int SLACK16 = (IN_SLOT_BYTES - IN_BYTES) / 16;
    for (int i = threadIdx.x; i < 4 * SLACK16; i += THREADS)
        reinterpret_cast<uint4*>(in_slot + (i / SLACK16) * IN_SLOT_BYTES + IN_BYTES)[i % SLACK16] =
            make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================== producer: two bulk copies per unit
        int it = 0;
        for (int u = u_begin; u < u_end; ++u, ++it) {
            const int s = it & 1;
            mbar_wait(&in_empty[s], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                const int b = u / UNITS_PER_IMG, v = u - b * UNITS_PER_IMG;
                const int64_t off = (1LL * b * PAD_H + 12 * v) * ROW_BYTES;
                mbar_expect_tx(&in_full[s], 2 * IN_BYTES);
                bulk_copy_g2s(in_slot + (2 * s) * IN_SLOT_BYTES, xp_hi + off, IN_BYTES, &in_full[s]);
                bulk_copy_g2s(in_slot + (2 * s + 1) * IN_SLOT_BYTES, xp_lo + off, IN_BYTES, &in_full[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: 3 split products per K step
        constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, 64);
        const uint64_t a_desc0 = umma_smem_desc(smem_u32(in_slot), 16, 128, UMMA_LAYOUT_NONE);
        const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), 1024, 128, UMMA_LAYOUT_NONE);
        constexpr uint64_t A_LO = IN_SLOT_BYTES >> 4;   // hi -> lo image inside a slot
        constexpr uint64_t B_LO = W_BYTES >> 4;         // hi -> lo weights
        int it = 0;
        uint32_t n0 = 0;  // uses of slot 0 so far (continuing units skip it)
        for (int u = u_begin; u < u_end; ++u, ++it) {
            const int s = it & 1;
            mbar_wait(&in_full[s], (it >> 1) & 1);
            const bool cont = continues(u);
            if (!cont) ++n0;
            for (int r = cont ? 1 : 0; r < ROWS_PER_UNIT; ++r) {
                mbar_wait(&slot_empty[r], r == 0 ? (n0 & 1) : (it & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + r * 64;
#pragma unroll
                    for (int kh = 0; kh < 7; ++kh) {
                        const uint64_t a_row = a_desc0 + static_cast<uint64_t>(
                                                             (2 * s * IN_SLOT_BYTES + (2 * r + kh) * ROW_BYTES) >> 4);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint64_t ad = a_row + static_cast<uint64_t>(i * 2);
                            const uint64_t bd = b_desc0 + static_cast<uint64_t>(((kh * 4 + 2 * i) * 1024) >> 4);
                            mma_f16_ss(d_tmem, ad, bd, idesc, (kh | i) != 0);   // x_hi * w_hi
                            mma_f16_ss(d_tmem, ad, bd + B_LO, idesc, 1);        // x_hi * w_lo
                            mma_f16_ss(d_tmem, ad + A_LO, bd, idesc, 1);        // x_lo * w_hi
                        }
                    }
                    tc_commit(&slot_full[r]);
                    if (r == ROWS_PER_UNIT - 1) tc_commit(&in_empty[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue (FP32 / TF32-rounded output)
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int et = q * 32 + (threadIdx.x & 31);
        const int etid = threadIdx.x - 128;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 32;
        int it = 0;
        uint32_t n0 = 0;  // uses of slot 0 so far
        float carry[32];  // the last conv row of the previous pooled row (and of the previous unit)
#pragma unroll
        for (int i = 0; i < 32; ++i) carry[i] = -INFINITY;
        for (int u = u_begin; u < u_end; ++u, ++it) {
            const int b = u / UNITS_PER_IMG, v = u - b * UNITS_PER_IMG;
            const uint32_t par = it & 1;
            const bool cont = continues(u);
            if (!cont) ++n0;
            for (int p = 0; p < POOLED_PER_UNIT; ++p) {
                const int ph = v * POOLED_PER_UNIT + p;
                if (p == 0 && !cont) mbar_wait(&slot_full[0], (n0 & 1) ^ 1);
                mbar_wait(&slot_full[2 * p + 1], par);
                mbar_wait(&slot_full[2 * p + 2], par);
                tc_fence_after();
                float m[32];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k == 0 && (p > 0 || cont)) {  // read once, as row k = 2 of the previous pooled row
#pragma unroll
                        for (int i = 0; i < 32; ++i) m[i] = carry[i];
                        continue;
                    }
                    const int oh = 6 * v - 1 + 2 * p + k;
                    uint32_t raw[32];
                    __syncwarp();
                    tmem_ld_32x32(lane_addr + (2 * p + k) * 64, raw);
                    tmem_ld_wait();
                    const bool valid = oh >= 0 && oh < CONV;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float x = valid ? __uint_as_float(raw[i]) : -INFINITY;
                        m[i] = k == 0 ? x : fmaxf(m[i], x);
                        if (k == 2) carry[i] = x;
                    }
                }
                tc_fence_before();
                if (p > 0 || !cont) mbar_arrive(&slot_empty[2 * p]);
                mbar_arrive(&slot_empty[2 * p + 1]);
                if (p == POOLED_PER_UNIT - 1) mbar_arrive(&slot_empty[2 * p + 2]);
                // single staging buffer: the previous pooled row's horizontal pass must be over
                named_bar_sync(1, EPI_THREADS);
                if (et < CONV) {
                    uint8_t* vrow = vbuf + et * VROW + half * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = rna_tf32(fmaxf(m[j * 4 + 0] + __ldg(bias + half * 32 + j * 4 + 0), 0.f));
                        o.y = rna_tf32(fmaxf(m[j * 4 + 1] + __ldg(bias + half * 32 + j * 4 + 1), 0.f));
                        o.z = rna_tf32(fmaxf(m[j * 4 + 2] + __ldg(bias + half * 32 + j * 4 + 2), 0.f));
                        o.w = rna_tf32(fmaxf(m[j * 4 + 3] + __ldg(bias + half * 32 + j * 4 + 3), 0.f));
                        *reinterpret_cast<float4*>(vrow + ((j ^ (et & 7)) << 4)) = o;
                    }
                }
                named_bar_sync(2, EPI_THREADS);
                if (ph < POOL) {
                    float* orow = out + ((1LL * b * POOL + ph) * POOL) * 64;
                    for (int task = etid; task < POOL * 16; task += EPI_THREADS) {
                        const int pw = task >> 4, c16 = task & 15;   // 16 chunks of 4 floats per pixel
                        const int hh = c16 >> 3, cj = c16 & 7;
                        const int c0 = 2 * pw;
                        auto ld = [&](int row) {
                            return *reinterpret_cast<const float4*>(vbuf + row * VROW + hh * 128 + ((cj ^ (row & 7)) << 4));
                        };
                        float4 a = ld(c0);
                        const float4 c = ld(c0 + 1);
                        a.x = fmaxf(a.x, c.x); a.y = fmaxf(a.y, c.y); a.z = fmaxf(a.z, c.z); a.w = fmaxf(a.w, c.w);
                        if (pw > 0) {
                            const float4 l = ld(c0 - 1);
                            a.x = fmaxf(a.x, l.x); a.y = fmaxf(a.y, l.y); a.z = fmaxf(a.z, l.z); a.w = fmaxf(a.w, l.w);
                        }
                        *reinterpret_cast<float4*>(orow + pw * 64 + c16 * 4) = a;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

size_t stem_tc_split_packed_input_bytes(int B) { return 2ull * B * PAD_H * ROW_BYTES; }
size_t stem_tc_split_packed_weight_bytes() { return 2 * W_BYTES; }

cudaError_t stem_tc_split_init() {
    return cudaFuncSetAttribute(stem_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
}

cudaError_t launch_stem_tc_split_pack_weights(const float* w, const float* bn_w, const float* bn_b,
                                              const float* bn_m, const float* bn_v, void* wk, float* bias,
                                              cudaStream_t s) {
    __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(wk);
    stem_pack_weights_split_kernel<<<(28 * 64 * 8 + 255) / 256, 256, 0, s>>>(w, bn_w, bn_b, bn_m, bn_v, hi,
                                                                          hi + W_BYTES / 2, bias);
    return cudaGetLastError();
}

// part 0: x fp32 NCHW -> xp (hi image followed by lo image); part 1: xp -> out NHWC fp32 [B,56,56,64].
cudaError_t launch_stem_tc_split_part(int part, const float* x, void* xp, const void* wk, const float* bias,
                                      void* out, int B, cudaStream_t s) {
    uint8_t* hi = static_cast<uint8_t*>(xp);
    uint8_t* lo = hi + 1ull * B * PAD_H * ROW_BYTES;
    if (part == 0) {
        const int64_t total = 1LL * B * PAD_H * PAD_W;
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
        if (blocks > cap) blocks = cap;
        stem_pack_split_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, reinterpret_cast<uint2*>(hi),
                                                                      reinterpret_cast<uint2*>(lo), B);
    } else {
        const int units = B * UNITS_PER_IMG;
        const int grid = units < num_sms() ? units : num_sms();
        const uint8_t* w_hi = static_cast<const uint8_t*>(wk);
        stem_tc_split_kernel<<<grid, THREADS, SMEM, s>>>(hi, lo, w_hi, w_hi + W_BYTES, bias,
                                                        static_cast<float*>(out), B);
    }
    return cudaGetLastError();
}

}  // namespace rnb
