// host_pack.h — CPU side of the host-input paths (rnb_model_forward_host / rnb_model_submit_host).
//
// The reference hands the network FP32 NCHW images from host memory (cuda/inference/main.cu:233-236: Tensor::load of
// the .bin that convert_imgs_to_bin.py wrote, then one cudaMemcpy). On the BF16 / FP8 paths the first thing the stem
// does with such an image is round it to BF16, and 602 KB per image over PCIe is what bounds the end-to-end rate
// (DESIGN.md section 6). So the host paths round on the CPU — the SAME round-to-nearest-even, bit for bit — into a
// pinned staging buffer and upload half the bytes; the stem then reads BF16 NCHW (stem_tc.cu, MODE 2).
// Plain C++ (compiled by g++, no CUDA): model.cu owns the staging buffers, streams and copies.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace rnb {

// dst[i] = BF16(src[i]), round to nearest even; NaN -> 0x7FFF (what cvt.rn.bf16x2.f32 produces). Single thread.
void f32_to_bf16_rne(const float* src, uint16_t* dst, size_t n);

// The split rule of the packed host paths. With c, l1, l2 seconds per FP32 byte of a batch on the cores, on the link as
// FP32 and on the link as BF16 (c = cold-sample conversion time / 0.8: inside a serving loop every core also feeds
// the DMA engine), a fraction f of the images through the cores costs f c of core time and f l2 + (1 - f) l1 of link
// time; both finish together at f = l1 / (c + l1 - l2). Returns that fraction — 1 above 0.93, 0 below a fifth or when
// less than 15 % would be gained over plain copies. Arguments: the three times measured on the same sample.
double host_pack_split(double t_convert, double t_copy_f32, double t_copy_bf16);
// The leading images of a `batch` that go through the cores for fraction `frac`: whole 16-image upload pieces; a
// remainder below one piece joins the other side, and a batch of fewer than 32 images goes to the larger side whole.
int host_pack_images(double frac, int batch);

// A process-wide pool of sleeping worker threads. run() converts src[0, n) into dst with the workers AND the calling
// thread, and calls ready(first_element, count) on the CALLING thread for every consecutive piece of `piece` elements
// as soon as that piece is complete, in order (the caller queues the piece's H2D copy there, so the upload of piece i
// overlaps the conversion of piece i + 1). One run() at a time (callers from several threads are serialised). A
// forked child process has no workers: there run() converts on the calling thread.
class HostPacker {
public:
    static HostPacker& instance();
    int threads() const { return nthreads_; }   // participants of a run(), the caller included
    void run(const float* src, uint16_t* dst, size_t n, size_t piece,
             const std::function<void(size_t first, size_t count)>& ready);
    HostPacker(const HostPacker&) = delete;
    HostPacker& operator=(const HostPacker&) = delete;

private:
    HostPacker();
    ~HostPacker();
    struct Impl;
    Impl* impl_;
    int nthreads_;
};

}  // namespace rnb
