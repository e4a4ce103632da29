// model.h — the planned ResNet inference engine behind rnb_model_* (include/rnb.h).
//
// Mirrors the graph that /root/reference/cuda/inference/main.cu builds by hand (ResnetModel :91-107,
// createLayer :53-89, layerForward :127-166, resnet152Forward :168-226) for five depths, but plans
// it once: weights are BN-folded and repacked at load, every convolution becomes one fused tcgen05
// launch with pre-built TMA descriptors over a liveness-packed activation arena, and each chunk of
// images is replayed from a CUDA graph.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "conv_plan.h"

namespace rnb {

bool stem_fused_enabled();  // stem_tc.cu


struct ConvWeights {
    void* w = nullptr;       // packed [Cout][k][k][Cin] in the activation type, BN folded
    float* bias = nullptr;   // [Cout]
    float* wscale = nullptr; // FP8 path: per-output-channel weight scale [Cout]
    int Cin = 0, Cout = 0, k = 0, stride = 1, pad = 0;  // FP8 path: channel counts padded to multiples of 128
};

struct BlockWeights {
    bool bottleneck = false;
    ConvWeights conv1, conv2, conv3;  // conv3 unused for BasicBlock
    bool has_ds = false;
    ConvWeights ds;
    float* bias3ds = nullptr;  // conv3 shift + downsample shift (fused-tail kernel with folded downsample)
    std::string name;  // "layer{L}.{i}"
};

struct NamedAct {
    void* ptr;
    int C, H, W;
};

// Everything needed to run `n` images through the network once.
struct ChunkPlan {
    int n = 0;
    void* stem_out = nullptr;   // NHWC [n,112,112,64] (CUDA-core stem) or packed NHWC4 input (tensor-core stem)
    void* pool_out = nullptr;   // NHWC [n,56,56,64]
    std::vector<ConvPlan> convs;
    void* last = nullptr;       // NHWC [n,7,7,C_final]
    int last_hw = 0, last_c = 0;
    float* pooled = nullptr;    // [C_final][n] fp32 (transposed, see tail.cu)
    void* pooled_bf16 = nullptr;  // [n][C_final] bf16: A operand of the tensor-core FC (BF16 path)
    float* fc_out = nullptr;    // logits pointer fc_plan was built for
    ConvPlan fc_plan;           // FC as a 1x1 conv on tcgen05 with FP32 output (BF16 path); valid iff pooled_bf16
    std::map<std::string, NamedAct> named;
    // FP8 path: the stem runs in BF16 into pool_raw, pool_out is its E4M3 copy (64 -> 128 channels, zero padded);
    // links[i] = the tensors conv i reads / adds / writes (their per-tensor scales are fixed by calibration)
    void* pool_raw = nullptr;
    struct Link {
        const void* in;
        const void* res;
        void* out;
    };
    std::vector<Link> links;
    std::vector<float*> fp8_vecs;  // per conv: 2 x Cout pre-multiplied epilogue vectors (owned by the model's arena)
    float last_scale = 1.f;  // scale of `last`
};

class Arena {
public:
    void* acquire(size_t bytes);
    void release(void* p);
    void release_all();  // mark every block free (a plan failed half way)
    void free_all();
    size_t total_bytes() const { return total_; }
    bool keep = false;  // debug: never recycle
private:
    struct Block { void* p; size_t bytes; bool busy; };
    std::vector<Block> blocks_;
    size_t total_ = 0;
};

// One batch of input images on the device, in one of the accepted formats (every format has 3 * image * image
// elements per image). BF16_NCHW is the packed host paths' form and may be mixed: images [0, nb) are BF16 at p, images
// [nb, batch) FP32 at p2, both tensors indexed by the image number in the batch.
struct InRef {
    enum Kind { F32_NCHW = 0, U8_HWC = 1, BF16_NCHW = 2 };
    const void* p = nullptr;
    int kind = F32_NCHW;
    const float* p2 = nullptr;
    int nb = 0;
    size_t elem_bytes() const { return kind == F32_NCHW ? 4 : kind == U8_HWC ? 1 : 2; }
    // the sub-batch that starts `images` images further on
    InRef at(size_t images, size_t img_elems) const {
        InRef r = *this;
        if (p) r.p = static_cast<const char*>(p) + images * img_elems * elem_bytes();
        if (p2) r.p2 = p2 + images * img_elems;
        r.nb = nb > static_cast<int>(images) ? nb - static_cast<int>(images) : 0;
        return r;
    }
    const float* f32() const { return kind == F32_NCHW ? static_cast<const float*>(p) : nullptr; }
    const uint8_t* u8() const { return kind == U8_HWC ? static_cast<const uint8_t*>(p) : nullptr; }
};

struct Model {
    std::string arch;
    int esz = 2;            // activation bytes: 1 fp8 (E4M3), 2 bf16, 4 tf32
    // FP8 variant: E4M3 weights (per-output-channel scales) and activations (per-tensor
    // scales), kind::f8f6f4 MMAs, layer-by-layer plan; stem and FC stay BF16. The activation scales are fixed by ONE
    // calibration pass (the first forward, or rnb_model_calibrate): every conv is run once with an amax-recording
    // epilogue on the calibration batch, in network order, on the already quantised inputs.
    bool fp8 = false;
    // Blocks [0, fp8_first_block) run in BF16 with the fused kernels of the BF16 path, the rest in FP8; the hand-over
    // is one re-quantisation launch. Default: layer1 stays BF16 (its 64-channel tensors would have to be padded to
    // 128 FP8 channels: no byte saving and 2-4x the MMA work); RNB_FP8_FROM=0: the whole network in FP8.
    int fp8_first_block = 0;
    // Hand-over of the mixed plan: the first FP8 block's conv1 and downsample read the BF16 block input DIRECTLY
    // (BF16 weights, BF16 MMAs) and write E4M3 — no separate re-quantisation pass over the 411 MB tensor.
    // RNB_FP8_HANDOVER=0: one quantize_pad launch instead (both convs then run in FP8).
    bool fp8_fused_handover = false;
    float* fp8_ones = nullptr;           // [4096] ones: the "weight scale" of a BF16-operand conv with E4M3 output
    bool fp8_calibrated = false;
    float fp8_stem_scale = 1.f;
    std::vector<float> fp8_out_scale;   // per conv launch, in plan order
    float* fp8_amax_dev = nullptr;
    int stem_esz() const { return fp8 ? 2 : esz; }
    int cpad(int c) const { return fp8 ? (c + 127) / 128 * 128 : c; }
    int calibrate_fp8(const float* x, int n);
    int apply_fp8_scales(ChunkPlan& p, cudaStream_t s);
    std::vector<void*> fp8_vec_allocs;  // device scratch of the pre-multiplied vectors of every plan
    int classes = 1000;
    int image = 224;
    int max_batch = 0;
    int chunk = 0;
    int num_sms = 148;
    int device = 0;         // the CUDA device this model lives on (current device at creation); the C ABI entry
                            // points switch to it, so one process can hold replicas on several GPUs
    bool bottleneck = true;
    int final_c = 2048;

    float* stem_w = nullptr;     // folded fp32 [64][3][7][7]
    float* stem_bias = nullptr;  // [64]
    void* stem_wk = nullptr;     // tensor-core stem weights (bf16 path), see stem_tc.cu
    bool stem_tc = false;
    std::vector<BlockWeights> blocks;
    float* fc_w = nullptr;  // [classes][final_c]
    float* fc_b = nullptr;
    void* fc_wq = nullptr;   // [classes_pad][final_c] bf16 (BF16 path: tensor-core FC), rows >= classes zero
    float* fc_bq = nullptr;  // [classes_pad]
    int classes_pad = 0;
    bool fc_tc = false;      // RNB_FC_TC=0 keeps the FP32 CUDA-core FC
    int num_convs = 0;
    double flops_per_image = 0;

    Arena arena;
    std::map<int, ChunkPlan> plans;
    // CUDA graphs: one captured chunk per (n | input kind << 29, x, logits, top1). At most kGraphsPerShape executables
    // are kept per (n, u8) shape; a call with new pointers beyond that RE-TARGETS the least recently used one with
    // cudaGraphExecUpdate (a re-capture, no instantiation), so a caller that allocates fresh outputs every call pays
    // ~0.1 ms of host time, not an instantiate, and nothing is ever flushed wholesale.
    using GraphKey = std::tuple<int64_t, const void*, const void*, void*, void*>;   // shape, x, x2 (mixed input), logits, top1
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        uint64_t used = 0;
    };
    static constexpr int kGraphsPerShape = 4;
    std::map<GraphKey, GraphEntry> graphs;
    uint64_t graph_clock = 0;
    int capture_chunk(ChunkPlan& p, InRef in, float* logits, int32_t* top1, cudaGraph_t* out);
    // All forwards of a model share ONE activation arena: every enqueue waits for the previous one (whatever stream it
    // went to) and leaves an event behind, so forward() on two streams, forward_host() and submit_host() never overlap
    // on the arena. Same-stream back-to-back forwards pay one event record per call.
    cudaEvent_t order_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    int begin_enqueue(cudaStream_t s);
    int end_enqueue(cudaStream_t s);
    cudaStream_t host_compute = nullptr;  // forward_host()'s compute stream
    int create_streams();
    // Planning + tile autotune + graph capture for `batch` images ahead of time, so that forward() never
    // device-synchronises or times trial launches inside a serving loop (rnb_model_warmup).
    int warmup(int batch, bool include_u8);
    cudaStream_t cap_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_events;
    float* host_x_dev = nullptr;    // device staging for forward_host
    float* host_logits_dev = nullptr;
    int32_t* host_top1_dev = nullptr;
    float* scratch_logits = nullptr;  // used when the caller passes logits = NULL
    float* u8_scratch = nullptr;      // normalised FP32 NCHW staging for uint8 input on the non-BF16-stem paths
    // ImageNet mean / std of convert_imgs_to_bin.py:18 (torchvision ImageClassification preset)
    float norm_mean[3] = {0.485f, 0.456f, 0.406f};
    float norm_std[3] = {0.229f, 0.224f, 0.225f};
    int last_chunk_n = 0;
    bool use_graph = true;
    bool alternate_tiles = true;  // consecutive convs walk their tiles in opposite directions (L2 reuse)
    bool autotune = true;         // time the admissible tile families per layer shape at plan time
    // Fused layer1 Bottleneck tail (bneck_l1.cuh). RNB_FUSE: 0 = off, 1 = conv2+conv3+residual,
    // 2 (default) = also fold the downsample conv of block 0; RNB_FUSE_NEXT=0 keeps the next block's
    // conv1 as its own launch.
    // Downsample conv of block 0 (layers 2-4) on a side stream, concurrently with conv1 -> conv2 of the same
    // block: it gets `side_sms` SMs, the main chain the rest (RNB_SIDE_SMS, 0 = off).
    int side_sms = 0;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t fork_ev[4] = {nullptr, nullptr, nullptr, nullptr}, join_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int launch_convs(ChunkPlan& p, cudaStream_t s, cudaEvent_t* per_launch_events);
    int fuse_level = 2;
    bool fuse_next = true;
    std::map<std::tuple<int, int, int, int, int, int, int, int>, int> tuned;  // layer shape (+ SM budget) -> force_bn code
    std::map<std::tuple<int, int, int, int>, int> tuned_c3n1;  // (M, K3, N3, N1) -> 1 fused conv3 + conv1' launch, 0 two plain launches

    // Two lanes (round 2): a batch whose layers leave the persistent grids with short last waves (ResNet-152 at 128
    // images: 98 tiles on 74 CTA pairs in all of layer3) runs as TWO half batches on two streams — a second engine state
    // (arena, plans, graphs, streams) that SHARES the weights — so that the idle SMs of one half's tail run the other
    // half's tiles. Per-image results do not depend on the batch they are in (bit-exact), so the output is unchanged.
    // Decided per batch size at first use by timing both forms (RNB_LANES=1 / 2 forces; needs RNB_AUTOTUNE).
    std::unique_ptr<Model> lane2;
    bool is_lane = false;                 // lane2 of another model: does not own the weights
    std::map<int, int> lane_choice;       // batch -> 1 | 2
    cudaEvent_t lane_fork = nullptr, lane_join = nullptr;
    int make_lane();
    int forward_one(InRef in, int batch, float* logits, int32_t* top1, cudaStream_t s);
    int forward_two(InRef in, int batch, float* logits, int32_t* top1, cudaStream_t s);
    int lanes_for(int batch, InRef in, float* logits, int32_t* top1, cudaStream_t s);

    ~Model();
    int load(const std::string& arch, int dtype, const std::string& dir, int max_batch, int chunk);
    // Pre-packed weight blob: everything load() derives from the 320-932 raw files of a
    // save_weights.py directory (BN-folded, K-major BF16/TF32 conv weights, stem / FC packs, biases) in ONE file
    // with a header and an FNV-1a checksum; load_packed() restores it with one read and one cudaMemcpy.
    int save_packed(const std::string& path);
    int load_packed(const std::string& path, int max_batch, int chunk);
    int configure(const std::string& arch_name, int dtype, int max_batch_, int chunk_);  // shared front half of both loads
    void build_structure();                      // blocks[] with shapes only (no device memory), flops, conv count
    template <class F> void for_each_weight(F&& f);  // every device weight buffer, deterministic order
    void* blob = nullptr;                        // load_packed(): the one device allocation all weights point into
    ChunkPlan* plan_for(int n);
    // uint8 HWC input is normalised with norm_mean / norm_std; BF16 NCHW input needs the BF16 tensor-core stem
    int enqueue_chunk(ChunkPlan& p, InRef in, float* logits, int32_t* top1, cudaStream_t s);
    int enqueue_fc(ChunkPlan& p, float* logits, cudaStream_t s);
    int forward(const float* x, int batch, float* logits, int32_t* top1, cudaStream_t s);
    int forward_u8(const uint8_t* x, int batch, float* logits, int32_t* top1, cudaStream_t s);
    // x = the FP32 image rounded to BF16 (nearest-even), NCHW: bit-identical to forward() on the BF16 / FP8 paths
    int forward_bf16(const uint16_t* x, int batch, float* logits, int32_t* top1, cudaStream_t s);
    bool accepts_bf16_input() const { return stem_tc && stem_esz() == 2 && stem_fused_enabled(); }
    int forward_any(InRef in, int batch, float* logits, int32_t* top1, cudaStream_t s);
    int set_normalization(const float* mean, const float* std);
    int forward_host(const float* x, int batch, float* logits, int32_t* top1);
    // Host packing (host_pack.h): FP32 host input of a BF16-stem model is rounded to BF16 by the host cores and half the
    // bytes cross PCIe — for a FRACTION of every batch: host cores (conversion rate Rc) and the PCIe link work side by
    // side, the leading images of a batch going through the cores as BF16, the rest straight over the link as FP32, in
    // the proportion that lets both finish together. host_pack_mode: 0 / 1 forced, none / all (RNB_HOST_PACK,
    // rnb_model_set_host_pack); -1 = the first host call with pageable input and the first with pinned input each
    // time the conversion and both copies on samples of their own batch and fix the fraction for that kind of
    // memory. A group (rnb_group_*) turns it off: its replicas are fed one after the other by ONE host thread.
    int host_pack_mode = -1;
    double host_pack_frac[2] = {-1, -1};    // auto mode: [pageable, pinned] -> fraction of the images rounded on the host
    double host_pack_last = -1;             // the fraction the most recent host call used
    double host_pack_gbps[2][3] = {{0, 0, 0}, {0, 0, 0}};   // per kind of memory: conversion (FP32 bytes read), FP32 H2D, BF16 H2D
    int host_pack_last_kind = 0;
    uint16_t* host_stage = nullptr;          // pinned staging of forward_host
    // how many leading images of this batch the host cores round to BF16 (0 = plain FP32 copies)
    int host_pack_for(const float* x, int batch, uint16_t** stage, float* x_dev);
    // pipelined host path: two slots, each with its own device input / output buffers
    struct HostSlot {
        uint16_t* stage = nullptr;   // pinned: the batch rounded to BF16 by the host cores (host_pack.h)
        float* x_dev = nullptr;
        float* logits_dev = nullptr;
        int32_t* top1_dev = nullptr;
        cudaEvent_t copied = nullptr, computed = nullptr, done = nullptr;
        bool pending = false;
    } slots[2];
    cudaStream_t pipe_compute = nullptr;
    cudaStream_t pipe_d2h = nullptr;   // the slots' device -> host copies (not in the way of the next forward pass)
    int submit_host(int slot, const float* x, int batch, float* logits, int32_t* top1);
    int submit_host_u8(int slot, const uint8_t* x, int batch, float* logits, int32_t* top1);
    int submit_host_any(int slot, const void* x, bool u8, int batch, float* logits, int32_t* top1);
    int wait_host(int slot);
    int repeat_launch(int batch, int index, int repeat, cudaStream_t s);
    int profile(const float* x, int batch, int iters, int* kind, float* ms, double* flops,
                double* bytes, int max_entries, int* n_entries, cudaStream_t s);
    // stem pre-pass/conv + stem/max-pool + planned conv launches + avg-pool + fc + arg-max
    int launches_per_chunk(int n) {
        ChunkPlan* p = plan_for(n);
        // (float input; the tensor-core stems are one launch — unless RNB_STEM_FUSED=0 — the CUDA-core stem two)
        const int stem = stem_tc && (stem_esz() == 4 || stem_fused_enabled()) ? 1 : 2;
        return p ? static_cast<int>(p->convs.size()) + stem + 3 + (p->pool_raw ? 1 : 0) : 0;
    }
};

}  // namespace rnb

// the opaque handle of include/rnb.h
struct rnb_model {
    rnb::Model impl;
};
