// Host-only check of InRef (csrc/model.h): the sub-batch arithmetic the chunk loop and the two-lane split rely on,
// for every input format and for a mixed BF16 / FP32 batch. Built with nvcc by tests/test_host_logic.py; no GPU needed.
#include <cstdio>
#include <cstdlib>

#include "../resnet_c_b200/csrc/model.h"

#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main() {
    using rnb::InRef;
    const size_t img = 3 * 224 * 224;
    char* base = reinterpret_cast<char*>(0x10000000);
    const float* f32 = reinterpret_cast<const float*>(0x40000000);

    InRef a{base, InRef::F32_NCHW};
    CHECK(a.elem_bytes() == 4 && a.f32() == reinterpret_cast<const float*>(base) && a.u8() == nullptr);
    CHECK(a.at(5, img).p == base + 5 * img * 4 && a.at(5, img).kind == InRef::F32_NCHW && a.at(5, img).p2 == nullptr);

    InRef u{base, InRef::U8_HWC};
    CHECK(u.elem_bytes() == 1 && u.u8() == reinterpret_cast<const uint8_t*>(base) && u.f32() == nullptr);
    CHECK(u.at(7, img).p == base + 7 * img);

    // mixed: images [0, 48) BF16 at p, images [48, 128) FP32 at p2, both indexed by the image number
    InRef m{base, InRef::BF16_NCHW, f32, 48};
    CHECK(m.elem_bytes() == 2 && m.f32() == nullptr && m.u8() == nullptr);
    InRef m16 = m.at(16, img);     // still inside the BF16 part
    CHECK(m16.p == base + 16 * img * 2 && m16.p2 == f32 + 16 * img && m16.nb == 32);
    InRef m64 = m.at(64, img);     // the second lane of a 128-image batch: FP32 images only
    CHECK(m64.p == base + 64 * img * 2 && m64.p2 == f32 + 64 * img && m64.nb == 0);
    InRef m48 = m.at(48, img);
    CHECK(m48.nb == 0);
    // composing offsets (lane split, then chunk loop) equals one offset
    InRef c = m.at(10, img).at(30, img);
    InRef d = m.at(40, img);
    CHECK(c.p == d.p && c.p2 == d.p2 && c.nb == d.nb && c.nb == 8);
    // the device address image b of the ORIGINAL batch is read from is the same through any sub-batch
    for (int off : {0, 16, 40, 48, 100}) {
        InRef s = m.at(off, img);
        for (int b = off; b < 128; b += 13) {
            const int local = b - off;
            const char* got = local < s.nb ? static_cast<const char*>(s.p) + local * img * 2
                                           : reinterpret_cast<const char*>(s.p2 + local * img);
            const char* want = b < 48 ? base + b * img * 2 : reinterpret_cast<const char*>(f32 + b * img);
            CHECK(got == want);
        }
    }
    // a pure BF16 batch (rnb_model_forward_bf16): nb = batch, no FP32 tensor
    InRef p{base, InRef::BF16_NCHW, nullptr, 256};
    CHECK(p.at(128, img).nb == 128 && p.at(128, img).p2 == nullptr);
    std::printf("inref ok\n");
    return 0;
}
