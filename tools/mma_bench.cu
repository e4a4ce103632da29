// mma_bench.cu — how many clocks does ONE tcgen05.mma (M = 128, K = 16, BF16) cost per SM as a function of N and of the
// shared-memory layout of its operands?  Built to settle why the stem's Hankel MMAs (no-swizzle K-major descriptors,
// LBO = 16 B / SBO = 128 B for A, LBO = 1024 B / SBO = 128 B for B) run at less than half of the tensor pipe's floor
// (DESIGN.md 3.2).  One CTA per SM, one issuing lane, `iters` back-to-back MMAs into one accumulator, one commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I resnet_c_b200/csrc tools/mma_bench.cu -o build/mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace rnb::ptx;

struct Op {
    uint32_t layout, lbo, sbo;   // descriptor fields
    uint32_t step[4];            // byte offsets the descriptor start cycles through (K advance / row shifts)
};

__global__ void __launch_bounds__(128, 1) k(Op a, Op b, int N, int iters, int lsu, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* A = smem;               // 64 KB
    uint8_t* B = smem + 65536;       // 64 KB
    uint8_t* V = smem + 131072;      // 32 KB of LSU traffic
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (131072 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 2) { __syncwarp(); tmem_alloc(&tptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    if (warp == 1) {
        const uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, static_cast<uint32_t>(N));
        const uint64_t ad0 = umma_smem_desc(smem_u32(A), a.lbo, a.sbo, a.layout);
        const uint64_t bd0 = umma_smem_desc(smem_u32(B), b.lbo, b.sbo, b.layout);
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    mma_f16_ss(tb, ad0 + (a.step[j] >> 4), bd0 + (b.step[j] >> 4), idesc, 1);
            }
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
    } else if (warp == 3 && lsu) {
        // epilogue-like LSU traffic beside the operand fetch: 16-byte stores + loads, conflict-free
        uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
        uint4 acc = v;
        for (int i = 0; i < iters * lsu; ++i) {
            uint4* p = reinterpret_cast<uint4*>(V) + ((i * 32 + (threadIdx.x & 31)) & 2047);
            *p = v;
            uint4 r = *(reinterpret_cast<uint4*>(V) + ((i * 32 + 64 + (threadIdx.x & 31)) & 2047));
            acc.x ^= r.x;
        }
        if (acc.x == 0x1234567u) out[1] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tb, 512); }
}

// TS form: A (128 x 16 BF16 = 8 TMEM columns) read from TMEM, B from shared memory
__global__ void __launch_bounds__(128, 1) kts(Op b, int N, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* B = smem + 65536;
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (131072 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 2) { __syncwarp(); tmem_alloc(&tptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    if (warp == 1) {
        const uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, static_cast<uint32_t>(N));
        const uint64_t bd0 = umma_smem_desc(smem_u32(B), b.lbo, b.sbo, b.layout);
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t a = tb + 256 + 8 * j, acc = 1;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb), "r"(a),
                                 "l"(bd0 + (b.step[j] >> 4)), "r"(idesc), "r"(acc) : "memory");
                }
            }
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tb, 512); }
}
static void runts(const char* name, Op b, int N, long long* out) {
    const int iters = 4096;
    cudaMemset(out, 0, 16);
    kts<<<148, 128, 1024 + 131072 + 32768>>>(b, N, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-58s N=%3d      : %-8s %7.1f clk per MMA (floor %3d) = %.2fx\n", name, N, cudaGetErrorString(e),
           double(h) / iters, N / 2, double(h) / iters / (N / 2));
    if (e != cudaSuccess) exit(1);
}

static void run(const char* name, Op a, Op b, int N, int lsu, long long* out) {
    const int iters = 4096;
    cudaMemset(out, 0, 16);
    k<<<148, 128, 1024 + 131072 + 32768>>>(a, b, N, iters, lsu, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-58s N=%3d lsu=%d: %-8s %7.1f clk per MMA (floor %3d) = %.2fx\n", name, N, lsu, cudaGetErrorString(e),
           double(h) / iters, N / 2, double(h) / iters / (N / 2));
    if (e != cudaSuccess) exit(1);
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 131072 + 32768);
    // A operands
    const Op a_hankel = {UMMA_LAYOUT_NONE, 16, 128, {0, 32, 1856, 1856 + 32}};          // the stem's overlapping descriptor
    const Op a_none = {UMMA_LAYOUT_NONE, 2048, 128, {0, 4096, 8192, 12288}};              // canonical no-swizzle, K chunks 2 KB apart
    const Op a_sw128 = {UMMA_LAYOUT_SW128, 16, 1024, {0, 32, 64, 96}};
    const Op a_sw32 = {UMMA_LAYOUT_SW32, 16, 256, {0, 4096, 8192, 12288}};
    const Op a_sw64 = {UMMA_LAYOUT_SW64, 16, 512, {0, 32, 8192, 8192 + 32}};
    // B operands (N rows)
    const Op b_stem = {UMMA_LAYOUT_NONE, 1024, 128, {0, 2048, 4096, 6144}};               // the stem's [chunk][64 oc][8 e]
    const Op b_none4k = {UMMA_LAYOUT_NONE, 4096, 128, {0, 8192, 16384, 24576}};           // same, N up to 256
    const Op b_sw128 = {UMMA_LAYOUT_SW128, 16, 1024, {0, 32, 64, 96}};
    const Op b_sw32 = {UMMA_LAYOUT_SW32, 16, 256, {0, 8192, 16384, 24576}};
    const Op b_sw64 = {UMMA_LAYOUT_SW64, 16, 512, {0, 32, 16384, 16384 + 32}};
    for (int N : {64, 128, 256}) {
        run("A sw128 / B sw128 (GEMM reference)", a_sw128, b_sw128, N, 0, out);
        run("A hankel none / B none (stem today)", a_hankel, N == 64 ? b_stem : b_none4k, N, 0, out);
        run("A hankel none / B sw128", a_hankel, b_sw128, N, 0, out);
        run("A hankel none / B sw64", a_hankel, b_sw64, N, 0, out);
        run("A hankel none / B sw32", a_hankel, b_sw32, N, 0, out);
        run("A none canonical / B none", a_none, b_none4k, N, 0, out);
        run("A sw128 / B none", a_sw128, b_none4k, N, 0, out);
        run("A sw32 / B sw32", a_sw32, b_sw32, N, 0, out);
        run("A sw64 / B sw64", a_sw64, b_sw64, N, 0, out);
    }
    cudaFuncSetAttribute(kts, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 131072 + 32768);
    {
        const Op b_hankel = {UMMA_LAYOUT_NONE, 16, 128, {0, 32, 1856, 1856 + 32}};
        for (int N : {64, 112, 128, 256}) {
            runts("A from TMEM / B hankel none (stem, third form)", b_hankel, N, out);
            runts("A from TMEM / B sw128", b_sw128, N, out);
        }
    }
    for (int lsu : {1, 2, 4}) {
        run("A sw128 / B sw128 + LSU traffic", a_sw128, b_sw128, 64, lsu, out);
        run("A hankel none / B none + LSU traffic", a_hankel, b_stem, 64, lsu, out);
    }
    return 0;
}
