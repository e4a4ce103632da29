/* rnb.h — C ABI of librnb.so, the B200-native (sm_100a) ResNet inference hot path.
 *
 * This is the boundary a host program binds (C/C++ directly, Python through ctypes, anything else
 * through its FFI): plain pointers and sizes, no C++ or torch types. Every entry point cites the
 * reference interface it replaces (paths are into olehskip/resnet.c).
 *
 * Conventions
 *   - All `*_dev` pointers are CUDA device pointers on the device passed to rnb_init().
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - Functions return 0 on success and a non-zero code on failure; rnb_last_error() then returns a
 *     thread-local, human-readable message. Nothing here aborts the process (the C++ shims in
 *     cuda/nn.cu add the reference's abort-on-error behaviour on top).
 *   - Calls are asynchronous on `stream` unless stated otherwise.
 *   - There is no CPU fallback: without a B200-class GPU every compute call fails.
 */
#ifndef RNB_H
#define RNB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNB_OK 0
#define RNB_ERR_INVALID 1   /* bad argument */
#define RNB_ERR_CUDA 2      /* CUDA runtime / driver failure */
#define RNB_ERR_IO 3        /* weight / image file problem */
#define RNB_ERR_UNSUPPORTED 4

/* Arithmetic type of the tensor-core path. Activations and weights are stored in this type,
 * accumulation is always FP32, logits are always FP32. */
#define RNB_DTYPE_BF16 0
#define RNB_DTYPE_TF32 1
/* FP8 variant (whole-model path only): E4M3 weights with one scale per output channel and E4M3 activations with one
 * scale per tensor, tcgen05 kind::f8f6f4 with FP32 accumulation; stem and FC stay BF16, logits FP32. The activation
 * scales are fixed by ONE calibration pass: rnb_model_calibrate(), or implicitly the first chunk of the first
 * forward. The reference has no reduced-precision path; the parity bar of this variant is stated in DESIGN.md. */
#define RNB_DTYPE_FP8 2

typedef struct rnb_model rnb_model_t;

/* ---- library ---------------------------------------------------------------------------- */

/* Prepare CUDA device `device` (check it is sm_100, raise kernel shared-memory limits — per device) and make it the
 * calling thread's current device. Idempotent; may be called for several devices of one process. The first call also
 * runs a one-off self-test of the im2col TMA descriptor path on a tensor smaller than 128 KiB (a driver-dependent
 * work-around is switched on or off according to the result; if neither variant reproduces the FP32 kernel the call
 * fails). A model lives on the device that was current when it was created and every rnb_model_* / rnb_group_* entry
 * point switches to that device for the duration of the call, so one process can drive replicas on several GPUs.
 * The per-op entry points below run on the device that owns their input tensor. */
int rnb_init(int device);
/* Message of the last failing call on this thread ("" if none). */
const char* rnb_last_error(void);
/* "resnet.c_b200 x.y (sm_100a)". */
const char* rnb_version(void);

/* ---- whole-model path (replaces createResnet152 / resnet152Forward / the CPU arg-max,
 *      cuda/inference/main.cu:109-125, 168-226, 243-251) -------------------------------------- */

/* Build a model: `arch` is "resnet18" | "resnet34" | "resnet50" | "resnet101" | "resnet152".
 * `weights_dir` holds one raw little-endian float32 file per state_dict key exactly as written by
 * save_weights.py:8-12 (conv OIHW, bn .weight/.bias/.running_mean/.running_var, fc [out,in]).
 * BatchNorm is folded into the conv weights/bias in FP64 at load time (ops.cu:149-150 evaluates the
 * same expression in double). `max_batch` sizes the activation arena; `chunk` is the number of
 * images pushed through the whole network at a time (0 = library default) so that inter-layer
 * activations stay L2-resident. */
int rnb_model_create(const char* arch, int dtype, const char* weights_dir, int max_batch, int chunk,
                     rnb_model_t** out);
int rnb_model_destroy(rnb_model_t* m);
/* Device the model lives on. */
int rnb_model_device(const rnb_model_t* m);
/* Plan, autotune and capture everything a forward of `batch` images needs (also the uint8 path if include_u8),
 * ahead of time: the first forward of a new batch size otherwise does this inside the call (a device
 * synchronisation plus trial launches — not allowed while the caller's stream is being captured). Blocking. */
int rnb_model_warmup(rnb_model_t* m, int batch, int include_u8);

/* FP8 models: fix the per-tensor activation scales from `batch` images (x_dev [batch,3,224,224] float32 NCHW,
 * batch <= max_batch; only the first `chunk` images are used). Every conv is run on the batch with an epilogue that
 * records max|y|; scale = max / 448 (the largest E4M3 value). Blocking. A second call is a no-op; other dtypes: no-op. */
int rnb_model_calibrate(rnb_model_t* m, const float* x_dev, int batch);

/* Pre-packed weight cache. rnb_model_save_packed() writes everything rnb_model_create() derived from the
 * save_weights.py directory (BN folded into K-major BF16/TF32 conv weights, stem / FC packs, biases) as ONE
 * file: 72-byte header (magic "RNBWGT01", format version — 2 since the stem's weight blocks were re-ordered; files of
 * another version are refused —, arch, dtype, class count, word-wise FNV-1a-64 checksum) + 256-byte-aligned
 * tensors. rnb_model_create_packed() restores a model from it with one read and one host->device copy
 * instead of 320-932 small file reads, copies and fold kernels (tensor.cuh:126-152,184-199); a wrong magic,
 * size or checksum is an error. The packed model computes bit-identical results. */
int rnb_model_save_packed(rnb_model_t* m, const char* path);
int rnb_model_create_packed(const char* path, int max_batch, int chunk, rnb_model_t** out);

/* Forward pass on device buffers: x_dev is [batch,3,224,224] float32 NCHW (the layout of the files
 * written by convert_imgs_to_bin.py:20-23), logits_dev is [batch,num_classes] float32,
 * top1_dev is [batch] int32 (lowest index among ties, as main.cu:243-251); either output may be
 * NULL. The whole pass is replayed from a CUDA graph; up to four executables are kept per batch size and a call with
 * other buffers re-targets the least recently used one (cudaGraphExecUpdate), so reuse your buffers where you can.
 * Concurrency: all forwards of ONE model share its activation arena and are ordered one after the other on the
 * device whatever streams they were enqueued on (an event links consecutive calls); calls on one model must come
 * from one host thread at a time. Use one model per stream / thread for concurrent passes. */
int rnb_model_forward(rnb_model_t* m, const float* x_dev, int batch, float* logits_dev,
                      int32_t* top1_dev, void* stream);

/* Same through HOST buffers (what main.cu does with loadToCuda / fc_out.cpu(), cuda/inference/main.cu:233-251):
 * copies the input host->device chunk by chunk overlapped with compute, runs the pass, copies logits/top-1 back and
 * synchronises. Pinned host memory is used as-is; pageable memory goes through the driver's staging.
 * Host packing (BF16 and FP8 models, 224 x 224): the stem's first act on an FP32 image is to round it to BF16, and
 * 602 KB per image over PCIe bounds the end-to-end rate, so the host paths can round on the HOST cores (a pool of
 * up to 16 threads, RNB_HOST_THREADS; the same round-to-nearest-even, results bit-identical) into a pinned staging
 * buffer and upload half the bytes, piece by piece while the next piece is being rounded. Cores and link share a
 * batch: the leading images go through the cores as BF16 while the rest crosses the link as FP32, in the proportion
 * that lets both finish together (the stem reads either form). Pageable input then needs no driver staging either.
 * The first host call of a model with pageable input, and the first with pinned input, time conversion and copies on
 * samples of their batch and fix that proportion for that kind of memory (none, if the plain copy is not clearly
 * slower; batches of fewer than 32 images are copied plainly and leave the question open); RNB_HOST_PACK=0 / 1 or rnb_model_set_host_pack() force none / all. */
int rnb_model_forward_host(rnb_model_t* m, const float* x_host, int batch, float* logits_host,
                           int32_t* top1_host);

/* Pipelined host path for a stream of batches (a serving loop): rnb_model_submit_host() queues the
 * H2D copy of one batch, its forward pass and the D2H copy of the results into `slot` (0 or 1) and
 * returns without blocking; rnb_model_wait_host() blocks until that slot's logits / top-1 are in the
 * host buffers given at submit time. With two slots the PCIe transfer of batch i+1 overlaps the
 * forward pass of batch i. Host buffers should be pinned; they must stay valid until the wait. */
int rnb_model_submit_host(rnb_model_t* m, int slot, const float* x_host, int batch, float* logits_host,
                          int32_t* top1_host);
int rnb_model_wait_host(rnb_model_t* m, int slot);

/* Host packing control and introspection. mode: -1 decide by timing at the next host call (default), 0 plain FP32
 * copies, 1 round every image to BF16 on the host (RNB_ERR_UNSUPPORTED on a model whose stem does not take BF16
 * input: tf32, or images other than 224 x 224). rnb_model_host_pack() returns what the most recent host call did
 * (0 none / 1 some or all; -1 = no host call yet) and, when info != NULL: info[0..2] = what the decision measured —
 * conversion rate (FP32 bytes read), FP32 H2D, BF16 H2D, in GB/s (zeros when forced) — and info[3] = the fraction
 * of a batch's images rounded on the host. rnb_host_pack_threads(): threads a conversion uses, caller included. */
int rnb_model_set_host_pack(rnb_model_t* m, int mode);
/* Fixes the proportion instead of measuring it: `fraction` of every batch's images (in whole 16-image pieces) is
 * rounded on the host, the rest crosses PCIe as FP32. 0 <= fraction <= 1. */
int rnb_model_set_host_pack_fraction(rnb_model_t* m, double fraction);
int rnb_model_host_pack(const rnb_model_t* m, double info[4]);
int rnb_host_pack_threads(void);
/* The split rule by itself (no GPU needed): from a conversion rate (GB/s of FP32 input read by the pool on a cold
 * sample), the rate of plain FP32 copies and the rate of BF16 copies from the staging buffer (GB/s), the fraction of
 * a batch's images the host paths would round on the host — f = l1 / (c + l1 - l2) in seconds per FP32 byte, c taken
 * at 0.8 of the sampled rate; 1 above 0.93, 0 below 0.2 or with less than 15 % to gain — and, through *images (may be
 * NULL), how many leading images of a `batch` that is (whole 16-image pieces). */
double rnb_host_pack_split(double convert_gbps, double h2d_f32_gbps, double h2d_bf16_gbps, int batch, int* images);
/* The conversion itself on host memory (no GPU involved): dst[i] = BF16(src[i]), round to nearest even, NaN ->
 * 0x7FFF, by the same thread pool. What the packed host paths upload. */
int rnb_f32_to_bf16_host(const float* src, uint16_t* dst, size_t n);
/* Forward on a DEVICE tensor that already holds such BF16 values, [batch,3,224,224] NCHW: bit-identical to
 * rnb_model_forward() on the FP32 tensor it was rounded from (BF16 / FP8 models only). */
int rnb_model_forward_bf16(rnb_model_t* m, const uint16_t* x_dev, int batch, float* logits_dev,
                           int32_t* top1_dev, void* stream);

/* Decoded-image input: x is uint8 HWC [batch][224][224][3] — the output of JPEG decode + resize 256 +
 * centre-crop 224 (convert_imgs_to_bin.py:12). The remaining step of that script (:18, torchvision
 * ToTensor + Normalize: (float(u8) / 255 - mean) / std in FP32) is fused into the stem's layout pre-pass,
 * bit-identically to feeding rnb_model_forward() the float tensor the script writes. A quarter of the
 * bytes cross PCIe. mean / std default to the ImageNet constants of the torchvision preset. */
int rnb_model_set_normalization(rnb_model_t* m, const float mean[3], const float std[3]);
int rnb_model_forward_u8(rnb_model_t* m, const uint8_t* x_dev, int batch, float* logits_dev,
                         int32_t* top1_dev, void* stream);
int rnb_model_submit_host_u8(rnb_model_t* m, int slot, const uint8_t* x_host, int batch, float* logits_host,
                             int32_t* top1_host);

/* Introspection used by the benchmarks. */
int rnb_model_num_classes(const rnb_model_t* m);
int rnb_model_num_convs(const rnb_model_t* m);
/* Kernel launches one forward of `batch` images issues (graph nodes). */
int rnb_model_launches_per_forward(rnb_model_t* m, int batch);
/* Algorithmic FLOPs per image: 2*MAC over convs + fc. */
double rnb_model_flops_per_image(const rnb_model_t* m);
/* Per-launch device timing of one chunk of min(batch, chunk) images, launched WITHOUT the graph and
 * with a CUDA event between consecutive launches on `stream` (average over `iters` passes, after one
 * warm-up pass). For launch i: kind_out[i] (0 stem conv, 1 max-pool, 2 tcgen05 conv, 3 avg-pool,
 * 4 fc, 5 arg-max), ms_out[i], flops_out[i] (algorithmic 2*MAC, 0 for pools/arg-max) and
 * bytes_out[i] (algorithmic HBM bytes: operands read once + result written once). Arrays hold
 * `max_entries`; the count is returned through *n_entries. Synchronises the stream. */
int rnb_model_profile(rnb_model_t* m, const float* x_dev, int batch, int iters, int* kind_out,
                      float* ms_out, double* flops_out, double* bytes_out, int max_entries,
                      int* n_entries, void* stream);
/* Profiling aid (energy per launch, tools/energy_profile.py): enqueue conv launch `index` (0-based among the
 * kind-2 entries of rnb_model_profile) of the plan for `batch` images `repeat` times back to back on `stream`, on the
 * buffers of the last forward / profile pass. No synchronisation. Results of a later forward are unaffected (every
 * launch rewrites its own output from unchanged inputs). */
int rnb_model_repeat_launch(rnb_model_t* m, int batch, int index, int repeat, void* stream);
/* Copy an intermediate activation of the LAST forward (chunk 0) to `out_dev` as float32 NCHW.
 * `name` is "maxpool" | "layer{L}.{i}" (block output) | "avgpool" (and "stem", the 112x112 conv1 output, only on
 * the CUDA-core stem path: the tensor-core stem fuses conv1 + pool and never materialises it). Returns the element
 * count through *numel (call with out_dev = NULL to query). Debug / parity use only: needs RNB_KEEP_ACTIVATIONS=1 in
 * the environment when the model is created (otherwise the arena recycles the buffers and the call fails with
 * RNB_ERR_UNSUPPORTED). */
int rnb_model_get_activation(rnb_model_t* m, const char* name, float* out_dev, int64_t* numel,
                             void* stream);

/* The step before the path, first half of the torchvision preset the reference applies to a PIL image
 * (convert_imgs_to_bin.py:12,18): resize so that the short side is `resize` (256; long side int(resize * long / short)),
 * antialiased bilinear with Pillow's 8-bit fixed-point arithmetic (horizontal pass, uint8 intermediate, vertical
 * pass), then centre crop `crop` x `crop` (224; offsets int(round((size - crop) / 2.0)), half to even). Bit-exact
 * against Pillow / torchvision. img_dev: n_images decoded images of ONE size, uint8 HWC [n][H][W][3]; out_dev
 * [n][crop][crop][3] — the input of rnb_model_forward_u8. Coefficient tables are cached per (H, W) (the first call of
 * a size allocates and uploads them synchronously); runs on the device that owns img_dev. JPEG decode stays on the
 * host. */
int rnb_resize_crop_u8(const uint8_t* img_dev, int n_images, int H, int W, uint8_t* out_dev, int resize, int crop,
                       void* stream);

/* ---- multi-GPU: data-parallel replicas driven by ONE process. The reference has a single
 *      device and B = 1 (cuda/inference/main.cu:230); per-image semantics are those of main.cu:168-251. ------------- */

typedef struct rnb_group rnb_group_t;

/* One full replica of the model on every device of `devices` (own weights, arena, stream, CUDA graphs; rnb_init() is
 * called for each). A batch shards by image into contiguous slices, earlier replicas take the remainder
 * (rnb_group_shard). No communication during the forward. The gather at the end is FUSED INTO THE LAST KERNELS: where
 * devices[r] can map devices[0]'s memory (NVLink peer access, enabled here), replica r's FC epilogue (TMA store) and
 * arg-max write their rows directly into the gathering buffers on devices[0]; otherwise (or with
 * RNB_GROUP_GATHER=copy in the environment) they land in a local buffer and one cudaMemcpyPeerAsync per output moves
 * them. A device may appear more than once (two replicas sharing a GPU). */
int rnb_group_create(const char* arch, int dtype, const char* weights_dir, const int* devices, int n_devices,
                     int max_batch_per_device, rnb_group_t** out);
int rnb_group_destroy(rnb_group_t* g);
int rnb_group_size(const rnb_group_t* g);
/* Replica r (borrowed: destroyed with the group). */
rnb_model_t* rnb_group_model(rnb_group_t* g, int r);
/* 1 if replica r stores its results straight into the root's buffers (peer-mapped), 0 if it goes through a copy. */
int rnb_group_direct_stores(const rnb_group_t* g, int r);
/* The sharding rule by itself (no GPU needed): `total` images over `world` replicas in contiguous slices, the first
 * total % world replicas take one image more. */
int rnb_shard_bounds(int total, int world, int rank, int* first, int* count);
/* Slice [*first, *first + *count) of a `batch`-image step owned by replica r. */
int rnb_group_shard(const rnb_group_t* g, int batch, int r, int* first, int* count);
/* Plan / autotune / capture every replica for its shard of `batch` images ahead of time (blocking). */
int rnb_group_warmup(rnb_group_t* g, int batch);
/* Sharded forward on device buffers. x_dev[r] is replica r's slice, [count_r,3,224,224] float32 NCHW ON devices[r];
 * logits_root_dev [batch,classes] and top1_root_dev [batch] live ON devices[0] (either may be NULL) and receive all
 * rows in image order. Asynchronous: every replica runs on its own stream; `root_stream` (a stream of devices[0],
 * NULL = the group's own) is made to wait for all of them, so work enqueued on it afterwards sees the gathered
 * results. rnb_group_synchronize() blocks until everything enqueued so far has finished. One host thread at a time. */
int rnb_group_forward(rnb_group_t* g, const float* const* x_dev, int batch, float* logits_root_dev,
                      int32_t* top1_root_dev, void* root_stream);
/* Same with decoded uint8 HWC slices [count_r,224,224,3] (rnb_model_forward_u8). */
int rnb_group_forward_u8(rnb_group_t* g, const uint8_t* const* x_dev, int batch, float* logits_root_dev,
                         int32_t* top1_root_dev, void* root_stream);
int rnb_group_synchronize(rnb_group_t* g);
/* Host-buffer paths: x_host [batch,3,224,224] (pinned memory recommended); every replica copies its own slice to its
 * device, forwards it and copies its own rows of logits_host [batch,classes] / top1_host [batch] back — the host
 * buffer is the gather. submit/wait: two slots as rnb_model_submit_host; forward_host = submit + wait (blocking). */
int rnb_group_submit_host(rnb_group_t* g, int slot, const float* x_host, int batch, float* logits_host,
                          int32_t* top1_host);
int rnb_group_submit_host_u8(rnb_group_t* g, int slot, const uint8_t* x_host, int batch, float* logits_host,
                             int32_t* top1_host);
int rnb_group_wait_host(rnb_group_t* g, int slot);
int rnb_group_forward_host(rnb_group_t* g, const float* x_host, int batch, float* logits_host, int32_t* top1_host);

/* ---- fused tensor-core convolution on caller-owned FP32 NCHW tensors
 *      (conv2dForwardKernel + batchNorm2dForwardKernel [+ addForwardKernel] [+ reluForwardKernel],
 *      ops.cu:14-48,139-151,153-160,130-137 as chained by layerForward, main.cu:127-166) -------- */

/* y = act( bn(conv(x, w)) + residual ). x [B,Cin,H,W]; w [Cout,Cin,k,k]; bn_* are [Cout] or all
 * NULL (no BN); residual [B,Cout,OH,OW] or NULL; k in {1,3}; Cin % 64 == 0, Cout % 64 == 0.
 * Internally converts to NHWC `dtype`, runs the tcgen05 implicit-GEMM kernel, converts back. */
int rnb_conv_bn_act_forward(const float* x_dev, const float* w_dev, const float* bn_weight_dev,
                            const float* bn_bias_dev, const float* bn_mean_dev,
                            const float* bn_var_dev, const float* residual_dev, float* out_dev,
                            int B, int Cin, int H, int W, int Cout, int k, int stride, int pad,
                            int relu, int dtype, void* stream);

/* The same as a PLANNED object, for callers that run one block many times (cuda/nn.cu's Conv2d / BasicBlock /
 * Bottleneck modules): BN is folded and the weights are packed once, at creation (the weights are COPIED: later
 * changes to the caller's tensors are not seen); per input shape the NHWC staging tensors and the launch descriptors
 * are built once and cached; intermediates of a block stay NHWC between its launches. A layer1-shaped Bottleneck
 * (64 -> 64 -> 256, stride 1, BF16) runs conv2 + conv3 + shortcut (+ downsample folded into conv3's accumulator) as
 * ONE fused launch, like the whole-model planner. Semantics of layerForward (main.cu:127-166) / torchvision BasicBlock. */
#define RNB_BLOCK_CONV 0        /* convs[0]                                  y = act(bn(conv(x)) [+ residual])      */
#define RNB_BLOCK_BASIC 1       /* convs = conv1, conv2 [, downsample]       BasicBlock                             */
#define RNB_BLOCK_BOTTLENECK 2  /* convs = conv1, conv2, conv3 [, downsample] Bottleneck (stride on conv2)          */
typedef struct rnb_block rnb_block_t;
typedef struct rnb_conv_params {
    const float* w;          /* [Cout][Cin][k][k] float32, device */
    const float* bn_weight;  /* [Cout] each, or all four NULL (no BN) */
    const float* bn_bias;
    const float* bn_mean;
    const float* bn_var;
    int Cin, Cout, k, stride, pad;
} rnb_conv_params_t;
int rnb_block_create(int kind, int dtype, const rnb_conv_params_t* convs, int n_convs, rnb_block_t** out);
int rnb_block_destroy(rnb_block_t* b);
/* x [B,Cin,H,W] -> out [B,Cout,OH,OW], float32 NCHW device tensors. residual_dev / relu: conv blocks only (a residual
 * block adds its own shortcut and always ends in ReLU). Asynchronous on `stream`; one host thread at a time. */
int rnb_block_forward(rnb_block_t* b, const float* x_dev, int B, int H, int W, const float* residual_dev, int relu,
                      float* out_dev, void* stream);
/* Tensor-core launches of the cached plan for this input shape (0 if it has not run yet). */
int rnb_block_num_launches(rnb_block_t* b, int B, int H, int W);

/* One FP8 (E4M3) convolution with explicit scales — the kernel of the RNB_DTYPE_FP8 model on caller-owned FP32 NCHW
 * tensors (tests): x / in_scale and residual / res_scale are rounded to E4M3 (NHWC, channels zero-padded to multiples
 * of 128), the weights are BN-folded and quantised with one scale per output channel (max|w| / 448),
 * y = act(acc * wscale[c] * in_scale + shift[c] (+ residual_q * res_scale)) is rounded to E4M3 as y / out_scale, and
 * out_dev receives the de-quantised result. k in {1,3}. */
int rnb_conv_fp8_forward(const float* x_dev, const float* w_dev, const float* bn_weight_dev, const float* bn_bias_dev,
                         const float* bn_mean_dev, const float* bn_var_dev, const float* residual_dev, float* out_dev,
                         int B, int Cin, int H, int W, int Cout, int k, int stride, int pad, int relu, float in_scale,
                         float res_scale, float out_scale, void* stream);

/* Fused stem: conv 7x7/2 pad 3 (3 -> 64) + BN + ReLU + maxpool 3x3/2 pad 1
 * (main.cu:181-192). x [B,3,H,W] -> out [B,64,OH/2,OW/2] float32 NCHW. */
int rnb_stem_forward(const float* x_dev, const float* w_dev, const float* bn_weight_dev,
                     const float* bn_bias_dev, const float* bn_mean_dev, const float* bn_var_dev,
                     float* out_dev, int B, int H, int W, int dtype, void* stream);

/* Fused tail: global average pool + fc + arg-max (main.cu:213-224, 243-251).
 * x [B,C,HW] float32 NCHW, fc_w [classes,C], fc_b [classes] -> logits [B,classes], top1 [B]. */
int rnb_tail_forward(const float* x_dev, const float* fc_w_dev, const float* fc_b_dev,
                     float* logits_dev, int32_t* top1_dev, int B, int C, int HW, int classes,
                     void* stream);

/* ---- per-op FP32 NCHW entry points with the reference kernels' exact semantics
 *      (what cuda/nn.cu's module forwards call; ops.cuh:15-32) --------------------------------- */

/* conv2dForwardKernel, ops.cu:14-48 — square kernel, zero padding, no bias. */
int rnb_conv2d_forward(const float* x_dev, float* out_dev, const float* w_dev, int B, int Cin, int H,
                       int W, int Cout, int k, int stride, int pad, void* stream);
/* batchNorm2dForwardKernel, ops.cu:139-151 — eval mode, eps 1e-5 evaluated in double; in place OK. */
int rnb_batchnorm2d_forward(const float* x_dev, float* out_dev, const float* weight_dev,
                            const float* bias_dev, const float* mean_dev, const float* var_dev, int B,
                            int C, int HW, void* stream);
/* reluForwardKernel, ops.cu:130-137 — in place OK. */
int rnb_relu_forward(const float* x_dev, float* out_dev, int64_t n, void* stream);
/* addForwardKernel, ops.cu:153-160 — in place OK. */
int rnb_add_forward(const float* a_dev, const float* b_dev, float* out_dev, int64_t n, void* stream);
/* maxPool2dKernel, ops.cu:50-78 — -inf init, out-of-bounds taps skipped. */
int rnb_maxpool2d_forward(const float* x_dev, float* out_dev, int B, int C, int H, int W, int k,
                          int stride, int pad, void* stream);
/* avgPool2dKernel, ops.cu:80-108 — divides by k*k even when taps are skipped. */
int rnb_avgpool2d_forward(const float* x_dev, float* out_dev, int B, int C, int H, int W, int k,
                          int stride, int pad, void* stream);
/* linearForwardKernel, ops.cu:110-128 — bias may be NULL. */
int rnb_linear_forward(const float* x_dev, float* out_dev, const float* w_dev, const float* bias_dev,
                       int B, int in_features, int out_features, void* stream);
/* Row-wise arg-max, lowest index on ties (main.cu:243-251). */
int rnb_argmax_forward(const float* x_dev, int32_t* out_dev, int B, int n, void* stream);

/* The step after the path (main.cu:240-251 and beyond): row softmax of the logits and the k most probable
 * classes (value descending, lowest index first on ties), 1 <= k <= min(n, 32). probs_full_dev [B][n] may be
 * NULL; top_probs_dev [B][k], top_idx_dev [B][k]. */
int rnb_softmax_topk_forward(const float* logits_dev, float* probs_full_dev, float* top_probs_dev,
                             int32_t* top_idx_dev, int B, int n, int k, void* stream);

/* Result writer: numel FP32 values from device memory to a raw little-endian file — the format of
 * Tensor::save (tensor.cuh:154-163) and of the cuda_out.bin / torch_out.bin pair that the reference's check_out()
 * (pytorch_inference.py:8-11) compares. Synchronises the device. */
int rnb_save_f32(const float* dev, int64_t numel, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* RNB_H */
