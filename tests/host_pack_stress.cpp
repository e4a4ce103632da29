// Stress / race check of the conversion pool (csrc/host_pack.cpp), built by tests/test_host_logic.py with
// -fsanitize=thread: several caller threads share the process-wide pool, every piece must be announced once, in
// order, only after all of its elements are converted, and the result must equal the scalar conversion.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../resnet_c_b200/csrc/host_pack.h"

static uint16_t scalar_bf16(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fff;
    return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

int main() {
    std::atomic<int> failures{0};
    const int callers = 3, rounds = 25;
    std::vector<std::thread> th;
    for (int c = 0; c < callers; ++c) {
        th.emplace_back([&, c] {
            const size_t n = 100000 + 37777 * c, piece = 16384 + 4096 * c;   // ragged: last piece and last item short
            std::vector<float> src(n);
            std::vector<uint16_t> dst(n);
            for (int r = 0; r < rounds; ++r) {
                for (size_t i = 0; i < n; ++i) src[i] = static_cast<float>((i * 2654435761u + r * 40503u + c) % 100003) / 777.f - 60.f;
                std::fill(dst.begin(), dst.end(), 0xAAAA);
                size_t expect_first = 0;
                rnb::HostPacker::instance().run(src.data(), dst.data(), n, piece, [&](size_t first, size_t count) {
                    if (first != expect_first || count == 0 || first + count > n) ++failures;
                    for (size_t i = first; i < first + count; ++i)       // the piece is complete when it is announced
                        if (dst[i] != scalar_bf16(src[i])) { ++failures; break; }
                    expect_first = first + count;
                });
                if (expect_first != n) ++failures;
            }
        });
    }
    for (auto& t : th) t.join();
    std::printf("threads %d failures %d\n", rnb::HostPacker::instance().threads(), failures.load());
    return failures.load() ? 1 : 0;
}
