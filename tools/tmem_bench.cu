// tmem_bench.cu — microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, alone and with the tensor
// pipe busy (tcgen05.mma streaming into OTHER TMEM columns). Answers what bounds the short-K fused epilogues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I resnet_c_b200/csrc tools/tmem_bench.cu -o build/tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace rnb::ptx;

// nld warps (4..4+nld-1) each issue `iters` x (tcgen05.ld.32x32b.x32 + wait); warp 1 issues `mma_iters` MMAs
// (M=128, N=256, K=16, bf16, operands = zeroed smem) back to back.
__global__ void __launch_bounds__(384, 1) k(int nld, int iters, int mma_iters, long long* out, float* sink, int fence_mode = 0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 49152 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 2) { __syncwarp(); tmem_alloc(&tptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    long long t0 = clock64();
    if (warp == 1) {
        const uint64_t ad = umma_smem_desc(smem_u32(smem), 0, 1024, UMMA_LAYOUT_SW128);
        const uint64_t bd = umma_smem_desc(smem_u32(smem) + 16384, 0, 1024, UMMA_LAYOUT_SW128);
        constexpr uint32_t idesc = umma_instr_desc(UMMA_FMT_BF16, 128, 256);
        if (mma_iters > 0) {
            if (elect_one()) {
                for (int i = 0; i < mma_iters; ++i) mma_f16_ss(tb + 256, ad, bd, idesc, 1);  // columns 256..511
                tc_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, 0);
            if (lane == 0 && blockIdx.x == 0) out[1] = clock64() - t0;
        }
    } else if (warp >= 4 && warp < 4 + nld) {
        const uint32_t lb = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        float acc = 0.f;
        for (int i = 0; i < iters; ++i) {
            uint32_t v[32];
            tmem_ld_32x32(lb + ((i & 3) * 32) + ((warp - 4) >> 2) * 128, v);   // columns 0..255
            if (fence_mode == 1) {          // a shared-memory store + proxy fence while the load is in flight
                reinterpret_cast<uint32_t*>(smem)[12288 + threadIdx.x] = i;
                fence_proxy_async_smem();
            }
            tmem_ld_wait();
            if (fence_mode == 2) {          // the same after the load has completed
                reinterpret_cast<uint32_t*>(smem)[12288 + threadIdx.x] = i;
                fence_proxy_async_smem();
            }
            acc += __uint_as_float(v[i & 31]);
        }
        if (acc == 123.f) sink[0] = acc;
        if (lane == 0 && blockIdx.x == 0 && warp == 4) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tb, 512); }
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 16); cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    const int iters = 2000;
    for (int mma : {0, 4000}) {
        for (int nld : {0, 1, 4, 8}) {
            if (nld == 0 && mma == 0) continue;
            cudaMemset(out, 0, 16);
            k<<<148, 384, 60000>>>(nld, iters, mma, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("ld warps %d, mma %d: %s  ld: %.1f clk per tcgen05.ld.x32 per warp (%.1f B/clk/SM aggregate)   mma: %.1f clk per M128N256K16\n",
                   nld, mma, cudaGetErrorString(e), nld ? double(h[0]) / iters : 0.0,
                   nld && h[0] ? 4096.0 * nld * iters / double(h[0]) : 0.0, mma ? double(h[1]) / mma : 0.0);
        }
    }
    // does fence.proxy.async (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) wait for a tcgen05.ld in flight?
    for (int fm : {1, 2}) {
        cudaMemset(out, 0, 16);
        k<<<148, 384, 60000>>>(1, iters, 0, out, sink, fm);
        cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("1 ld warp, st.shared + fence.proxy.async %s the load: %.1f clk per iteration\n",
               fm == 1 ? "WHILE IN FLIGHT (before wait::ld of)" : "AFTER wait::ld of", double(h[0]) / iters);
    }
    return 0;
}
