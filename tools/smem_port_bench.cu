// smem_port_bench.cu — do TMA (bulk-copy) WRITES into shared memory slow down the tensor core's operand READS from it?
// One CTA per SM: warp 1 issues back-to-back tcgen05.mma (M = 128, N given, K = 16, SW128 operands at fixed addresses),
// warp 0 streams 16 KB bulk copies from an L2-resident global buffer into a 4-stage ring elsewhere in shared memory,
// warp 3 optionally hammers a third region with 16-byte LSU stores + loads.  Reports clocks per MMA and the bytes per
// clock the copies achieved, alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I resnet_c_b200/csrc tools/smem_port_bench.cu -o build/smem_port_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace rnb::ptx;

constexpr int STAGE = 16384, NST = 4;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* g, int N, int mma_iters, int copies, int lsu, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* A = smem;                 // 32 KB
    uint8_t* B = smem + 32768;         // 32 KB
    uint8_t* R = smem + 65536;         // NST x 16 KB ring
    uint8_t* V = R + NST * STAGE;      // 16 KB LSU area
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar, cbar[NST];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (65536 + NST * STAGE + 16384) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < NST; ++i) mbar_init(&cbar[i], 1); fence_mbar_init(); }
    if (warp == 2) { __syncwarp(); tmem_alloc(&tptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    if (warp == 1 && mma_iters > 0) {
        const uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, static_cast<uint32_t>(N));
        const uint64_t ad0 = umma_smem_desc(smem_u32(A), 16, 1024, UMMA_LAYOUT_SW128);
        const uint64_t bd0 = umma_smem_desc(smem_u32(B), 16, 1024, UMMA_LAYOUT_SW128);
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < mma_iters; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_f16_ss(tb, ad0 + 2 * j, bd0 + 2 * j, idesc, 1);
            }
            tc_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
    } else if (warp == 0 && copies > 0) {
        const uint8_t* src = g + static_cast<size_t>(blockIdx.x) * NST * STAGE;
        long long t0 = clock64();
        if (elect_one()) {
            for (int c = 0; c < copies; ++c) {
                const int s = c % NST;
                if (c >= NST) mbar_wait(&cbar[s], ((c / NST) - 1) & 1);
                mbar_expect_tx(&cbar[s], STAGE);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(R + s * STAGE)), "l"(src + s * STAGE), "r"(STAGE), "r"(smem_u32(&cbar[s])) : "memory");
            }
            for (int c = copies; c < copies + NST; ++c) { const int s = c % NST; mbar_wait(&cbar[s], ((c / NST) - 1) & 1); }
        }
        __syncwarp();
        long long t1 = clock64();
        if (elect_one() && blockIdx.x == 0) out[1] = t1 - t0;
    } else if (warp == 3 && lsu > 0) {
        uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int i = 0; i < lsu; ++i) {
            uint4* p = reinterpret_cast<uint4*>(V) + ((i * 32 + (threadIdx.x & 31)) & 1023);
            *p = v;
            uint4 r = *(reinterpret_cast<uint4*>(V) + ((i * 32 + 512 + (threadIdx.x & 31)) & 1023));
            acc ^= r.x;
        }
        long long t1 = clock64();
        if (acc == 0x1234567u) out[3] = 1;
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[2] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tb, 512); }
}

// copies only, `nst` stages of `stage` bytes in flight (is the rate above a latency limit: bytes in flight / round trip?)
__global__ void __launch_bounds__(128, 1) kc(const uint8_t* g, int nst, int stage, int copies, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* R = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __shared__ uint64_t cbar[16];
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&cbar[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint8_t* src = g + static_cast<size_t>(blockIdx.x) * 65536;
        long long t0 = clock64();
        if (elect_one()) {
            for (int c = 0; c < copies; ++c) {
                const int s = c % nst;
                if (c >= nst) mbar_wait(&cbar[s], ((c / nst) - 1) & 1);
                mbar_expect_tx(&cbar[s], stage);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(R + s * stage)), "l"(src + (s * stage) % 65536), "r"(stage), "r"(smem_u32(&cbar[s])) : "memory");
            }
            for (int c = copies; c < copies + nst; ++c) { const int s = c % nst; mbar_wait(&cbar[s], ((c / nst) - 1) & 1); }
        }
        __syncwarp();
        long long t1 = clock64();
        if (elect_one() && blockIdx.x == 0) out[0] = t1 - t0;
    }
}
static void runc(const uint8_t* g, int nst, int stage, long long* out) {
    const int copies = (64 << 20) / stage / 16;
    cudaMemset(out, 0, 32);
    kc<<<148, 128, 1024 + 200 * 1024>>>(g, nst, stage, copies, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("copies only: %2d stages x %5d B = %3d KB in flight: %-8s %6.1f B/clk/SM\n", nst, stage, nst * stage / 1024,
           cudaGetErrorString(e), double(copies) * stage / double(h));
}

static void run(const uint8_t* g, int N, int mma_iters, int copies, int lsu, long long* out, int grid = 148) {
    cudaMemset(out, 0, 32);
    k<<<grid, 128, 1024 + 65536 + NST * STAGE + 16384>>>(g, N, mma_iters, copies, lsu, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4] = {0, 0, 0, 0};
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("grid=%3d N=%3d mma=%5d copies=%5d lsu=%6d: %-8s", grid, N, mma_iters, copies, lsu, cudaGetErrorString(e));
    if (mma_iters) printf("  %6.1f clk per MMA (floor %d)", double(h[0]) / mma_iters, N / 2);
    if (copies) printf("  copies %5.1f B/clk/SM", double(copies) * STAGE / double(h[1]));
    if (lsu) printf("  lsu %5.1f clk per st+ld pair", double(h[2]) / lsu);
    printf("\n");
    if (e != cudaSuccess) exit(1);
}

int main() {
    long long* out;
    uint8_t* g;
    cudaMalloc(&out, 32);
    cudaMalloc(&g, 148ull * NST * STAGE);
    cudaMemset(g, 0, 148ull * NST * STAGE);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 65536 + NST * STAGE + 16384);
    for (int N : {64, 128, 256}) {
        run(g, N, 8192, 0, 0, out);
        run(g, N, 0, 2048, 0, out);
        // copies sized so that both streams run for about the same time
        run(g, N, 8192, N == 64 ? 1024 : N == 128 ? 1536 : 3072, 0, out);
        run(g, N, 8192, 0, 8192 * 4, out);
        run(g, N, 8192, N == 64 ? 1024 : N == 128 ? 1536 : 3072, 8192 * 2, out);
    }
    cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 200 * 1024);
    for (int nst : {2, 4, 8, 12}) runc(g, nst, 16384, out);
    for (int nst : {4, 8, 16}) runc(g, nst, 8192, out);
    for (int nst : {2, 4, 6}) runc(g, nst, 32768, out);
    // is the 65 B/clk an SM ingest limit or an aggregate (L2 / crossbar) one? fewer SMs pulling:
    for (int grid : {8, 16, 37, 74, 111, 148}) run(g, 128, 0, 4096, 0, out, grid);
    return 0;
}
