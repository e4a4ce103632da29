// stem_tc_common.cuh — what the two tensor-core stems (stem_tc.cu: BF16; stem_tc_split.cu: TF32 accuracy out of
// split BF16 MMAs) share: the padded NHWC4 row geometry, the order of the weight blocks, the sequence of conv-row
// pairs / input chunks a CTA walks, and a few PTX helpers. See stem_tc.cu for the scheme.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

namespace rnb {
namespace stemtc {

constexpr int IMG = 224;
constexpr int PAD_W = 232;            // 4 + 224 + 4 pixels (padded pixel = image column + 4)
constexpr int PAD_H = 235;            // 5 + 224 + 6 rows (padded row pr = ih + 5)
constexpr int ROW_BYTES = PAD_W * 8;  // 1856
constexpr int POOL = 56;
constexpr int PAIRS = POOL;           // conv-row pairs (= pooled rows) per image
constexpr int CHUNK_ROWS = 4;         // input rows per ring chunk: chunk c of an image = image rows 4c-4 .. 4c-1
constexpr int CHUNK_BYTES = CHUNK_ROWS * ROW_BYTES;  // 7424
constexpr int NSLOT = 4;              // TMEM pair slots (128 columns each)

// position of filter row kh inside a K chunk's group of weight blocks: [W6 W4 W2 W0 | W5 W3 W1], so that the block
// after W_kh is W_kh-2 and an N = 128 MMA starting at W_kh covers both conv rows of a pair
__host__ __device__ constexpr int wpos(int kh) { return (kh & 1) ? 4 + (5 - kh) / 2 : (6 - kh) / 2; }

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                              uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            ptx::smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

// D[tmem] (+)= A[tmem] * B[smem]^T: A = 128 rows x 16 BF16, packed two per 32-bit column (lane m, 8 columns)
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :
                 : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// The sequence of work every role of a CTA walks in lock step: the CTA's contiguous range of pairs [p, p_end),
// preceded by the pair before it when the range starts inside an image (`warm`: computed only for the conv row it
// hands to the first real pair). `cn` counts the input chunks issued so far: a step that starts a segment (the first
// step, or the first pair of an image) brings three new chunks (k, k+1, k+2), every other step one (k+2), and a
// step reads the last three.
struct StemSteps {
    int p, p_end, cn, step, bb, kk;   // (bb, kk) = image and pair-in-image of p, kept incrementally (no divisions per step)
    bool warm;
    __device__ StemSteps(int p_begin, int p_end_)
        : p(p_begin), p_end(p_end_), cn(0), step(0), bb(p_begin / PAIRS), kk(p_begin % PAIRS),
          warm(p_begin < p_end_ && (p_begin % PAIRS) != 0) {}
    __device__ bool done() const { return p >= p_end; }
    __device__ int b() const { return bb; }
    __device__ int k() const { return kk - (warm ? 1 : 0); }
    __device__ bool seg_start() const { return step == 0 || k() == 0; }
    __device__ int new_chunks() const { return seg_start() ? 3 : 1; }
    // the next step starts a new segment (another image) or does not exist: this step's chunks all die with it
    __device__ bool seg_ends() const { return !warm && (p + 1 >= p_end || kk + 1 == PAIRS); }
    __device__ void next() {
        cn += new_chunks();
        if (warm) {
            warm = false;
        } else {
            ++p;
            if (++kk == PAIRS) {
                kk = 0;
                ++bb;
            }
        }
        ++step;
    }
};

}  // namespace stemtc
}  // namespace rnb
