// host_pack.cpp — see host_pack.h. FP32 -> BF16 rounding on the host cores for the host-input paths.
#include "host_pack.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <pthread.h>
#include <sched.h>

namespace rnb {

namespace {

inline uint16_t bf16_rne_scalar(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;               // NaN (cvt.rn.bf16.f32: canonical NaN)
    return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);  // ties to even; overflow rounds to inf
}

void convert_scalar(const float* src, uint16_t* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = bf16_rne_scalar(src[i]);
}

// 16 floats per iteration in integer arithmetic (no denormal flushing, unlike vcvtneps2bf16); the result is written
// with non-temporal stores when dst is 32-byte aligned: it is read next by the DMA engine, not by this core.
__attribute__((target("avx2"))) void convert_avx2(const float* src, uint16_t* dst, size_t n) {
    const __m256i bias = _mm256_set1_epi32(0x7fff), one = _mm256_set1_epi32(1), absmask = _mm256_set1_epi32(0x7fffffff),
                  inf = _mm256_set1_epi32(0x7f800000), qnan = _mm256_set1_epi32(0x7fff);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i u0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i u1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
        __m256i r0 = _mm256_srli_epi32(
            _mm256_add_epi32(_mm256_add_epi32(u0, bias), _mm256_and_si256(_mm256_srli_epi32(u0, 16), one)), 16);
        __m256i r1 = _mm256_srli_epi32(
            _mm256_add_epi32(_mm256_add_epi32(u1, bias), _mm256_and_si256(_mm256_srli_epi32(u1, 16), one)), 16);
        r0 = _mm256_blendv_epi8(r0, qnan, _mm256_cmpgt_epi32(_mm256_and_si256(u0, absmask), inf));
        r1 = _mm256_blendv_epi8(r1, qnan, _mm256_cmpgt_epi32(_mm256_and_si256(u1, absmask), inf));
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(r0, r1), 0xD8);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), p);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), p);
    }
    _mm_sfence();
    convert_scalar(src + i, dst + i, n - i);
}

// The same with 512-bit registers (32 floats per iteration, one down-converting move instead of pack + permute).
__attribute__((target("avx512f,avx512bw"))) void convert_avx512(const float* src, uint16_t* dst, size_t n) {
    const __m512i bias = _mm512_set1_epi32(0x7fff), one = _mm512_set1_epi32(1), absmask = _mm512_set1_epi32(0x7fffffff),
                  inf = _mm512_set1_epi32(0x7f800000), qnan = _mm512_set1_epi32(0x7fff);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 63) == 0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m512i u0 = _mm512_loadu_si512(reinterpret_cast<const void*>(src + i));
        const __m512i u1 = _mm512_loadu_si512(reinterpret_cast<const void*>(src + i + 16));
        __m512i r0 = _mm512_srli_epi32(
            _mm512_add_epi32(_mm512_add_epi32(u0, bias), _mm512_and_si512(_mm512_srli_epi32(u0, 16), one)), 16);
        __m512i r1 = _mm512_srli_epi32(
            _mm512_add_epi32(_mm512_add_epi32(u1, bias), _mm512_and_si512(_mm512_srli_epi32(u1, 16), one)), 16);
        r0 = _mm512_mask_mov_epi32(r0, _mm512_cmpgt_epi32_mask(_mm512_and_si512(u0, absmask), inf), qnan);
        r1 = _mm512_mask_mov_epi32(r1, _mm512_cmpgt_epi32_mask(_mm512_and_si512(u1, absmask), inf), qnan);
        const __m512i p = _mm512_inserti64x4(_mm512_castsi256_si512(_mm512_cvtepi32_epi16(r0)), _mm512_cvtepi32_epi16(r1), 1);
        if (aligned) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i), p);
        else _mm512_storeu_si512(reinterpret_cast<void*>(dst + i), p);
    }
    _mm_sfence();
    convert_scalar(src + i, dst + i, n - i);
}

// 1 = AVX2 (default where the CPU has it), 0 = scalar, 2 = AVX-512 F + BW. RNB_HOST_ISA = scalar | avx2 | avx512
// selects (tests, A/Bs); the 512-bit form stays opt-in: measured within noise of AVX2 on a shared Xeon here (the loop is
// memory-bound), and what the B200 boxes' cores were measured with is the AVX2 form.
int isa_level() {
    static const int v = [] {
        const bool avx2 = __builtin_cpu_supports("avx2");
        const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
        int level = avx2 ? 1 : 0;
        if (const char* e = getenv("RNB_HOST_ISA")) {
            if (!strcmp(e, "scalar")) level = 0;
            else if (!strcmp(e, "avx512") && avx512) level = 2;
        }
        return level;
    }();
    return v;
}

int default_threads() {
    if (const char* e = getenv("RNB_HOST_THREADS")) {
        const int t = atoi(e);
        if (t >= 1) return std::min(t, 256);
    }
    int cpus = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
    if (cpus <= 0) cpus = static_cast<int>(std::thread::hardware_concurrency());
    if (cpus <= 0) cpus = 1;
    // one process per GPU under torchrun: the ranks of a node share its cores
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
        const int w = atoi(e);
        if (w > 1) cpus = std::max(1, cpus / w);
    }
    return std::min(cpus, 16);
}

constexpr size_t ITEM = 32768;  // elements per work item: 128 KB read, 64 KB written

// A forked child inherits the pool object but none of its threads (and possibly a locked mutex): it converts on the
// calling thread alone.
std::atomic<bool> g_forked_child{false};

}  // namespace

void f32_to_bf16_rne(const float* src, uint16_t* dst, size_t n) {
    switch (isa_level()) {
        case 2: convert_avx512(src, dst, n); break;
        case 1: convert_avx2(src, dst, n); break;
        default: convert_scalar(src, dst, n);
    }
}

double host_pack_split(double t_convert, double t_copy_f32, double t_copy_bf16) {
    if (!(t_convert > 0) || !(t_copy_f32 > 0) || !(t_copy_bf16 > 0)) return 0.0;
    const double c = t_convert / 0.8, l1 = t_copy_f32, l2 = t_copy_bf16;
    double f = c + l1 - l2 > 0 ? l1 / (c + l1 - l2) : 1.0;
    f = std::min(1.0, std::max(0.0, f));
    if (f > 0.93) f = 1.0;
    const double t_mixed = std::max(f * c, f * l2 + (1.0 - f) * l1);
    if (f < 0.2 || t_mixed > 0.85 * l1) f = 0.0;
    return f;
}

int host_pack_images(double frac, int batch) {
    if (batch <= 0 || !(frac > 0)) return 0;
    if (frac >= 1.0) return batch;
    if (batch < 32) return frac >= 0.5 ? batch : 0;   // too small to split: the larger side takes it all
    int nb = static_cast<int>(frac * batch / 16.0 + 0.5) * 16;
    if (nb < 16) nb = 0;
    if (nb > batch - 16) nb = batch;
    return nb;
}

struct HostPacker::Impl {
    std::vector<std::thread> workers;
    std::mutex run_mutex;              // one run() at a time
    std::mutex m;
    std::condition_variable cv;
    uint64_t gen = 0;
    bool stop = false;
    // the current job
    const float* src = nullptr;
    uint16_t* dst = nullptr;
    size_t n = 0, piece = 0, items_per_piece = 0, nitems = 0;
    std::atomic<size_t> next{0};
    std::atomic<int> active{0};
    std::unique_ptr<std::atomic<uint32_t>[]> done;   // per piece: finished items

    // converts one item, returns false when none is left
    bool take() {
        const size_t g = next.fetch_add(1, std::memory_order_relaxed);
        if (g >= nitems) return false;
        const size_t p = g / items_per_piece, i = g - p * items_per_piece;
        const size_t pend = std::min(n, (p + 1) * piece), a = p * piece + i * ITEM;
        if (a < pend) f32_to_bf16_rne(src + a, dst + a, std::min(ITEM, pend - a));
        done[p].fetch_add(1, std::memory_order_release);
        return true;
    }

    void worker() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
            }
            while (take()) {}
            active.fetch_sub(1, std::memory_order_release);
        }
    }
};

HostPacker::HostPacker() : impl_(new Impl), nthreads_(default_threads()) {
    for (int t = 1; t < nthreads_; ++t) impl_->workers.emplace_back([this] { impl_->worker(); });
    pthread_atfork(nullptr, nullptr, [] { g_forked_child.store(true, std::memory_order_relaxed); });
}

HostPacker::~HostPacker() {
    {
        std::lock_guard<std::mutex> lk(impl_->m);
        impl_->stop = true;
    }
    impl_->cv.notify_all();
    for (auto& t : impl_->workers) t.join();
    delete impl_;
}

HostPacker& HostPacker::instance() {
    static HostPacker* p = new HostPacker();   // never destroyed: worker threads must not be joined from atexit of a dlopen'ed library
    return *p;
}

void HostPacker::run(const float* src, uint16_t* dst, size_t n, size_t piece,
                     const std::function<void(size_t, size_t)>& ready) {
    if (n == 0) return;
    if (piece == 0 || piece > n) piece = n;
    if (g_forked_child.load(std::memory_order_relaxed)) {   // no workers here: piece by piece on this thread
        for (size_t first = 0; first < n; first += piece) {
            const size_t count = std::min(piece, n - first);
            f32_to_bf16_rne(src + first, dst + first, count);
            ready(first, count);
        }
        return;
    }
    Impl& J = *impl_;
    std::lock_guard<std::mutex> run_lock(J.run_mutex);
    const size_t npieces = (n + piece - 1) / piece;
    J.src = src; J.dst = dst; J.n = n; J.piece = piece;
    J.items_per_piece = (piece + ITEM - 1) / ITEM;
    J.nitems = npieces * J.items_per_piece;
    J.done.reset(new std::atomic<uint32_t>[npieces]);
    for (size_t p = 0; p < npieces; ++p) J.done[p].store(0, std::memory_order_relaxed);
    J.next.store(0, std::memory_order_relaxed);
    J.active.store(static_cast<int>(J.workers.size()), std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(J.m);
        ++J.gen;
    }
    J.cv.notify_all();
    size_t p = 0;
    while (p < npieces) {
        if (J.done[p].load(std::memory_order_acquire) == J.items_per_piece) {
            const size_t first = p * piece;
            ready(first, std::min(piece, n - first));
            ++p;
        } else if (!J.take()) {
            _mm_pause();   // every item is taken; the piece is being finished by a worker
        }
    }
    while (J.active.load(std::memory_order_acquire) != 0) _mm_pause();   // no worker still looks at this job
}

}  // namespace rnb
