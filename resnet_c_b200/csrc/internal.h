// internal.h — declarations shared by the translation units of librnb.so (not part of the ABI).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <string>

namespace rnb {

// ---- process state (api.cu). Nothing here is tied to ONE device: rnb_init(d) prepares device d (kernel attributes
// are per device) and every model remembers the device it was created on.
int num_sms();                // SM count of the calling thread's current device (148 if rnb_init() has not seen it)
bool device_ready(int dev);   // rnb_init(dev) has succeeded
void set_error(const std::string& msg);
int fail_cuda(cudaError_t e, const char* what);  // records the message, returns RNB_ERR_CUDA

// Makes `dev` the calling thread's current device for the lifetime of the guard (entry points of a model /
// group run on the model's own device whatever the caller had selected).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (dev >= 0 && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Frees stream-ordered temporaries on EVERY exit path of the per-op entry points (api.cu).
struct AsyncTemps {
    cudaStream_t s;
    void* p[8];
    int n = 0;
    explicit AsyncTemps(cudaStream_t stream) : s(stream) {}
    cudaError_t alloc(void** out, size_t bytes) {
        cudaError_t e = n < 8 ? cudaMallocAsync(out, bytes, s) : cudaErrorMemoryAllocation;
        if (e == cudaSuccess) p[n++] = *out;
        return e;
    }
    ~AsyncTemps() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(p[i], s);
    }
    AsyncTemps(const AsyncTemps&) = delete;
    AsyncTemps& operator=(const AsyncTemps&) = delete;
};

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor in the
// stream drains; it MUST execute griddepcontrol.wait before touching anything an earlier kernel wrote (and every
// kernel launched this way must execute it, so that completion stays transitive). RNB_NO_PDL=1 -> plain launch.
template <class Kernel, class... Args>
inline cudaError_t launch_pdl_small(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args... args) {
    static const bool on = !(getenv("RNB_NO_PDL") && atoi(getenv("RNB_NO_PDL")) != 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- FP32 NCHW per-op kernels (ops_f32.cu)
cudaError_t launch_conv2d_f32(const float* x, float* out, const float* w, int B, int Cin, int H,
                              int W, int Cout, int k, int stride, int pad, cudaStream_t s);
cudaError_t launch_batchnorm2d_f32(const float* x, float* out, const float* weight,
                                   const float* bias, const float* mean, const float* var, int B,
                                   int C, int HW, cudaStream_t s);
cudaError_t launch_relu_f32(const float* x, float* out, int64_t n, cudaStream_t s);
cudaError_t launch_add_f32(const float* a, const float* b, float* out, int64_t n, cudaStream_t s);
cudaError_t launch_pool2d_f32(bool is_max, const float* x, float* out, int B, int C, int H, int W,
                              int k, int stride, int pad, cudaStream_t s);
cudaError_t launch_linear_f32(const float* x, float* out, const float* w, const float* bias, int B,
                              int in_f, int out_f, cudaStream_t s);
cudaError_t launch_argmax_f32(const float* x, int32_t* out, int B, int n, cudaStream_t s);
// row softmax (optional full output) + top-k probabilities / indices (value desc, index asc)
cudaError_t launch_softmax_topk_f32(const float* x, float* probs_full, float* top_p, int32_t* top_i, int B, int n,
                                    int k, cudaStream_t s);

// ---- layout / weight preparation (layout.cu). `esz` = 2 (bf16) or 4 (tf32-rounded fp32).
cudaError_t launch_nchw_to_nhwc(const float* x, void* out, int B, int C, int HW, int esz,
                                cudaStream_t s);
cudaError_t launch_nhwc_to_nchw(const void* x, float* out, int B, int C, int HW, int esz,
                                cudaStream_t s);
// w [Cout][Cin][k][k] fp32 (+ optional BN vectors, all-or-none) -> packed [Cout][k][k][Cin_pad] in
// the activation type and bias[Cout] fp32. The fold is evaluated in double.
cudaError_t launch_fold_pack(const float* w, const float* bn_w, const float* bn_b, const float* bn_m,
                             const float* bn_v, void* packed, float* bias, int Cout, int Cin, int k,
                             int esz, cudaStream_t s);
// Same fold, but keeps fp32 OIHW order (used by the CUDA-core stem).
cudaError_t launch_fold_f32(const float* w, const float* bn_w, const float* bn_b, const float* bn_m,
                            const float* bn_v, float* w_out, float* bias, int Cout, int per_out,
                            cudaStream_t s);

// ---- FP8 (E4M3) quantisation (fp8.cu). Weights: BN folded in double, one scale per output channel, channels padded
// to Cout_pad / Cin_pad (multiples of 128); activations: per-tensor scale.
cudaError_t launch_fold_pack_fp8(const float* w, const float* bn_w, const float* bn_b, const float* bn_m,
                                 const float* bn_v, void* packed, float* wscale, float* bias, int Cout, int Cin, int k,
                                 int Cout_pad, int Cin_pad, cudaStream_t s);
cudaError_t launch_amax_bf16(const void* x, int64_t n, float* amax, cudaStream_t s);
cudaError_t launch_quantize_pad_bf16(const void* x, void* out, int64_t rows, int C, int Cpad, float inv_scale,
                                     cudaStream_t s);
cudaError_t launch_nchw_to_nhwc_fp8(const float* x, void* out, int B, int C, int Cpad, int HW, float inv_scale,
                                    cudaStream_t s);
cudaError_t launch_nhwc_fp8_to_nchw(const void* x, float* out, int B, int C, int Cpad, int HW, float scale,
                                    cudaStream_t s);

// ---- stem + max-pool (stem.cu)
// conv 7x7/2 pad 3, 3 -> 64, + bias + ReLU: x fp32 NCHW [B,3,H,W] -> out NHWC [B,OH,OW,64].
cudaError_t launch_stem_conv(const float* x, const float* w_folded /*[64][3][7][7]*/,
                             const float* bias, void* out, int B, int H, int W, int esz,
                             cudaStream_t s);
// max-pool 3x3/2 pad 1 over NHWC.
cudaError_t launch_maxpool_nhwc(const void* x, void* out, int B, int H, int W, int C, int esz,
                                cudaStream_t s);

// ---- tensor-core stem, BF16 path, 224x224 inputs (stem_tc.cu)
size_t stem_tc_packed_input_bytes(int B);
size_t stem_tc_packed_weight_bytes();
cudaError_t stem_tc_init();
cudaError_t launch_stem_tc_pack_weights(const float* w, const float* bn_w, const float* bn_b,
                                        const float* bn_m, const float* bn_v, void* wk, float* bias,
                                        cudaStream_t s);
// conv7x7/2 + BN + ReLU + maxpool from NCHW fp32. Default: ONE launch (part 1; part 0 is a no-op) whose loader warps
// build the padded NHWC4 bf16 rows in shared memory; RNB_STEM_FUSED=0: part 0 = layout pre-pass into the scratch
// `xp`, part 1 = the same kernel fed from `xp` by bulk copies.
bool stem_fused_enabled();
cudaError_t launch_stem_tc(const float* x, void* xp, const void* wk, const float* bias, void* out, int B,
                           cudaStream_t s);
cudaError_t launch_stem_tc_part(int part, const float* x, void* xp, const void* wk, const float* bias,
                                void* out, int B, cudaStream_t s);
// the kernel fed from an already packed tensor (uint8 input: launch_stem_tc_pack_u8 first)
cudaError_t launch_stem_tc_from_packed(const void* xp, const void* wk, const float* bias, void* out, int B,
                                       cudaStream_t s);
// the kernel fed with images [0, nb) from a BF16 NCHW tensor (the FP32 image rounded to nearest-even on the host:
// host_pack.cpp) and images [nb, B) from an FP32 NCHW tensor, both indexed by the image number in the batch
cudaError_t launch_stem_tc_from_mixed(const void* x_bf16, const float* x_f32, int nb, const void* wk, const float* bias,
                                      void* out, int B, cudaStream_t s);

// Decoded-image input (uint8 HWC [B][224][224][3]) with the /255 + mean/std normalisation of
// convert_imgs_to_bin.py:18 fused in: straight into the packed stem input (BF16 path) ...
cudaError_t launch_stem_tc_pack_u8(const uint8_t* x, void* xp, int B, const float* mean, const float* std,
                                   cudaStream_t s);
// ... or into a normalised FP32 NCHW tensor (every other path)
cudaError_t launch_u8_hwc_to_f32_nchw(const uint8_t* x, float* out, int B, int H, int W, const float* mean,
                                      const float* std, cudaStream_t s);

// ---- tensor-core stem, TF32 path (3-term BF16 split, FP32/TF32 output), 224x224 (stem_tc_split.cu)
size_t stem_tc_split_packed_input_bytes(int B);
size_t stem_tc_split_packed_weight_bytes();
cudaError_t stem_tc_split_init();
cudaError_t launch_stem_tc_split_pack_weights(const float* w, const float* bn_w, const float* bn_b,
                                              const float* bn_m, const float* bn_v, void* wk, float* bias,
                                              cudaStream_t s);
cudaError_t launch_stem_tc_split_part(int part, const float* x, void* xp, const void* wk, const float* bias,
                                      void* out, int B, cudaStream_t s);

// dispatch on the activation type (2 = BF16 kernel, 4 = TF32 split kernel)
inline size_t stem_any_input_bytes(int esz, int B) {
    return esz == 2 ? stem_tc_packed_input_bytes(B) : stem_tc_split_packed_input_bytes(B);
}
inline size_t stem_any_weight_bytes(int esz) {
    return esz == 2 ? stem_tc_packed_weight_bytes() : stem_tc_split_packed_weight_bytes();
}
inline cudaError_t launch_stem_any_pack_weights(int esz, const float* w, const float* bn_w, const float* bn_b,
                                                const float* bn_m, const float* bn_v, void* wk, float* bias,
                                                cudaStream_t s) {
    return esz == 2 ? launch_stem_tc_pack_weights(w, bn_w, bn_b, bn_m, bn_v, wk, bias, s)
                    : launch_stem_tc_split_pack_weights(w, bn_w, bn_b, bn_m, bn_v, wk, bias, s);
}
inline cudaError_t launch_stem_any_part(int esz, int part, const float* x, void* xp, const void* wk,
                                        const float* bias, void* out, int B, cudaStream_t s) {
    return esz == 2 ? launch_stem_tc_part(part, x, xp, wk, bias, out, B, s)
                    : launch_stem_tc_split_part(part, x, xp, wk, bias, out, B, s);
}

// ---- tail (tail.cu)
// global average pool over NHWC [B,HW,C] -> pooledT [C][B] fp32 (transposed, see tail.cu)
// (+ optional row-major BF16 copy pooled_bf16[B][C] for the tensor-core FC; BF16 activations only)
// With pooled_bf16 (tensor-core FC) the FP32 result is written ROW-MAJOR [B][C] instead (nobody reads it on the hot
// path; rnb_model_get_activation hands it out).
// esz == 1: x is E4M3 with per-tensor scale `in_scale` (pooled_bf16 required).
cudaError_t launch_avgpool_nhwc(const void* x, float* pooledT, void* pooled_bf16, int B, int HW, int C, int esz,
                                cudaStream_t s, float in_scale = 1.f);
// fc.weight [classes][C] fp32 -> [cpad][C] bf16 (zero rows beyond classes), fc.bias -> [cpad] fp32
cudaError_t launch_fc_pack(const float* w, const float* b, void* wq, float* bq, int classes, int C, int cpad,
                           cudaStream_t s);
// logits[B,classes] = pooledT[C,B]^T * W[classes,C]^T + bias
cudaError_t launch_fc(const float* pooledT, const float* w, const float* bias, float* logits, int B,
                      int C, int classes, cudaStream_t s);
// out[cols][rows] = in[rows][cols]^T, fp32
cudaError_t launch_transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t s);

}  // namespace rnb
