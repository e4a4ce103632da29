// tmem_bench.cu — tcgen05.ld throughput per SM for different load widths (32x32b .x16 / .x32 / .x64 / .x128),
// 8 epilogue-style warps (two per lane quarter) each reading its quarter in a loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace rnb::ptx;

template <int N> struct Ld;
#define REGS16(v, o) "=r"(v[o+0]),"=r"(v[o+1]),"=r"(v[o+2]),"=r"(v[o+3]),"=r"(v[o+4]),"=r"(v[o+5]),"=r"(v[o+6]),"=r"(v[o+7]),"=r"(v[o+8]),"=r"(v[o+9]),"=r"(v[o+10]),"=r"(v[o+11]),"=r"(v[o+12]),"=r"(v[o+13]),"=r"(v[o+14]),"=r"(v[o+15])
template <> struct Ld<16> { static __device__ __forceinline__ void go(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : REGS16(v, 0) : "r"(a) : "memory"); } };
template <> struct Ld<32> { static __device__ __forceinline__ void go(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : REGS16(v, 0), REGS16(v, 16) : "r"(a) : "memory"); } };
template <> struct Ld<64> { static __device__ __forceinline__ void go(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : REGS16(v, 0), REGS16(v, 16), REGS16(v, 32), REGS16(v, 48) : "r"(a) : "memory"); } };

template <int N>
__global__ void __launch_bounds__(384, 1) k(int iters, long long* out, float* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 2) { __syncwarp(); tmem_alloc(&tptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    long long t0 = clock64();
    if (warp >= 4) {
        const uint32_t lb = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 256;
        uint32_t acc = 0;
        for (int i = 0; i < iters; ++i) {
            uint32_t v[N];
            Ld<N>::go(lb + ((i * N) & 255 & ~(N - 1)), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < N; ++j) acc ^= v[j];
        }
        if (acc == 0x12345u) sink[0] = 1.f;
        __syncwarp();
        if (lane == 0 && blockIdx.x == 0 && warp == 4) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tb, 512); }
}

// Does fence.proxy.async (the fence an epilogue needs before a TMA store / UMMA may read what it wrote) wait for a
// tcgen05.ld in flight?  mode 0: ld; wait::ld; work; st.shared; fence     mode 1: ld(next); work(cur); st.shared; fence; wait::ld
__global__ void __launch_bounds__(384, 1) kf(int iters, int mode, long long* out, float* sink) {
    __shared__ uint32_t tptr;
    __shared__ uint32_t buf[256 * 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 2) { __syncwarp(); tmem_alloc(&tptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    long long t0 = clock64();
    if (warp >= 4) {
        const uint32_t lb = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 256;
        uint32_t acc = 0;
        uint32_t va[32], vb[32];
        auto work = [&](uint32_t (&v)[32]) {   // ~100 dependent-ish ALU ops + 4 shared stores + proxy fence
            uint32_t x = acc;
#pragma unroll
            for (int j = 0; j < 32; ++j) x = x * 1664525u + v[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) buf[(threadIdx.x - 128) * 4 + j] = x + j;
            fence_proxy_async_smem();
            acc = x;
        };
        if (mode == 0) {
            for (int i = 0; i < iters; ++i) {
                Ld<32>::go(lb + ((i * 32) & 224), va);
                tmem_ld_wait();
                work(va);
            }
        } else {
            Ld<32>::go(lb, va);
            for (int i = 0; i < iters; i += 2) {
                tmem_ld_wait();
                Ld<32>::go(lb + (((i + 1) * 32) & 224), vb);
                work(va);
                tmem_ld_wait();
                Ld<32>::go(lb + (((i + 2) * 32) & 224), va);
                work(vb);
            }
            tmem_ld_wait();
        }
        if (acc == 0x12345u) sink[0] = 1.f;
        __syncwarp();
        if (lane == 0 && blockIdx.x == 0 && warp == 4) out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tb, 512); }
}

template <int N> void run(long long* out, float* sink) {
    const int iters = 4000;
    cudaMemset(out, 0, 16);
    k<N><<<148, 384>>>(iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("32x32b.x%-3d 8 warps: %s  %.1f clk per load per warp, %.1f B/clk/SM aggregate\n", N, cudaGetErrorString(e),
           double(h) / iters, 8.0 * 128 * N * iters / double(h));
}
int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 16); cudaMalloc(&sink, 4);
    run<16>(out, sink); run<32>(out, sink); run<64>(out, sink);
    for (int mode = 0; mode < 2; ++mode) {
        const int iters = 4000;
        cudaMemset(out, 0, 16);
        kf<<<148, 384>>>(iters, mode, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("ld.x32 + work + st.shared + fence.proxy.async, %s: %s  %.1f clk per step (8 warps)\n",
               mode ? "next load issued BEFORE the work (double-buffered)" : "load; wait; work (serial)",
               cudaGetErrorString(e), double(h) / iters);
    }
    return 0;
}
