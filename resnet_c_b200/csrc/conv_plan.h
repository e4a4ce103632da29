// conv_plan.h — one planned (descriptor-complete) launch of the implicit-GEMM conv kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "bneck_c3n1.cuh"
#include "bneck_l1.cuh"
#include "conv3x3_halo2.cuh"
#include "conv3x3_halo.cuh"
#include "conv_igemm.cuh"

namespace rnb {

enum class ActType { FP8 = 1, BF16 = 2, TF32 = 4 };  // value = bytes per element (FP8 = E4M3, scaled)

struct ConvDesc {
    int B, H, W, Cin, Cout, ksize, stride, pad;
    bool relu;
    bool reverse = false;  // walk the tiles last-to-first (see ConvGeom::reverse)
    ActType act;
    const void* in;        // NHWC [B][H][W][Cin]
    const void* weight;    // [Cout][ksize][ksize][Cin], BN-folded, same element type as `in`
    const float* bias;     // [Cout] fp32 (folded BN shift)
    const void* residual;  // NHWC [B][OH][OW][Cout] or nullptr
    void* out;             // NHWC [B][OH][OW][Cout]
    // FC mode (BF16 operands only): the output is UNROUNDED FP32, row-major [M][out_cols] with
    // out_cols <= Cout (weights / bias are padded to Cout rows; the padded columns are clipped by TMA)
    bool out_f32 = false;
    int out_cols = 0;
    // FP8 only (see ConvGeom): per-output-channel weight scales, per-tensor activation scales, and a device scratch
    // of 2 * Cout floats that receives the pre-multiplied epilogue vectors (fp8_premultiply)
    const float* chan_scale = nullptr;  // [Cout] weight scales
    float in_scale = 1.f, res_scale = 1.f, out_scale = 1.f;
    float* fp8_vecs = nullptr;          // [2][Cout]: chan_scale', bias'
    // BF16 operands, E4M3 output (act == BF16): the hand-over conv of a mixed FP8 plan. chan_scale = a vector of ones,
    // in_scale = 1; residual (if any) is E4M3 too.
    bool out_fp8 = false;
};

// Fused layer1 Bottleneck tail (bneck_l1.cuh): conv2 3x3 + conv3 1x1 + shortcut + ReLU, optionally
// with the downsample conv folded into conv3's accumulator and the next block's conv1 appended.
struct BneckDesc {
    int B, H, W;             // t1 is NHWC [B][H][W][64]
    bool reverse = false;
    const void* t1;          // conv1 output of this block
    const void* w2;          // [64][3][3][64] bf16, BN folded
    const float* bias2;
    const void* w3;          // [256][64]
    const float* bias3;      // conv3 shift; with `wds`: conv3 shift + downsample shift
    const void* wds;         // [256][64] downsample weights, or nullptr
    const void* shortcut;    // wds ? block input NHWC [B][H][W][64] : residual NHWC [B][H][W][256]
    void* y;                 // NHWC [B][H][W][256]
    const void* w1n;         // next block's conv1 weights [n1][256], or nullptr
    int n1 = 64;             // its output channels: 64 (same layer) or 128 (first block of the next layer)
    const float* bias1n;
    void* t1n;               // next block's conv1 output NHWC [B][H][W][64]
};

// conv3 (128 -> 512) + shortcut + ReLU fused with the next block's conv1 (512 -> 128) (bneck_c3n1.cuh)
struct C3n1Desc {
    int M;                 // pixel rows
    int K3 = 128, N3 = 512, N1 = 128;  // conv3: K3 -> N3, conv1': N3 -> N1. (128,512,128) or (256,1024,256)
    bool reverse = false;
    const void* t2;        // [M][128]
    const void* w3;        // [512][128]
    const float* bias3;
    const void* residual;  // [M][512]
    void* y;               // [M][512]
    const void* w1n;       // [128][512]
    const float* bias1n;
    void* t1n;             // [M][128]
};

struct ConvPlan {
    CUtensorMap tmA, tmB, tmOut, tmRes;
    CUtensorMap tmBh;  // CTA-pair kernel, tail split: weight boxes of bn/4 rows (ConvGeom::split_from)
    CUtensorMap tmW3, tmWds, tmW1n, tmT1n;  // fused Bottleneck tail only
    int bneck;    // 0 = plain conv, 1 = fused tail with residual tensor, 2 = fused tail with folded downsample,
                  // 3 = conv3 + next conv1 (bneck_c3n1.cuh; geometry in cg / cp), resident weights (layer2 shape),
                  // 4 = the same with streamed weights (layer3 shape),
                  // 5 = fused layer1 tail with residual tensor + the NEXT LAYER's conv1 (256 -> 128)
    int rings;    // bneck == 4: shared-memory split of bneck_c3n1s_kernel (RNB_C3N1S_RINGS at plan time, 0 = default)
    C3n1Geom cg;
    C3n1Params cp;
    BneckGeom bg;
    BneckParams bp;
    ConvGeom g;
    const float* bias;
    int bn;       // tile N
    int ctas;     // 1 = single-CTA 128-pixel tiles, 2 = CTA-pair 256-pixel tiles (cta_group::2)
    int resb;     // 1 = CTA-pair BN = 128 kernel with RESIDENT weights (force_bn code 31128: 3x3, 128 -> 128, bf16)
    int w16;      // 1 = single-CTA BN = 128 kernel with 16 epilogue warps (force_bn code 20128; bf16 / fp8)
    int deep;     // 1 = deeper smem pipeline / fewer staging buffers variant of the tile family (bf16)
    int f32out;   // 1 = FP32-output variant of the single-CTA kernel (FC layer)
    int halo;     // 1 = conv3x3_halo_kernel (3x3/1, 64->64, bf16): tmA/tmOut are 4-D tiled maps, geometry in hg
    HaloGeom hg;
    int halo2;    // 1 = conv3x3_halo2_kernel (3x3/1, 64->64, tf32, CTA pair): tmA/tmRes/tmOut are 4-D maps, geometry in h2g
    Halo2Geom h2g;
    int osz;      // output element bytes (= esz except for the BF16 -> E4M3 hand-over convs)
    int esz;      // element bytes
    int grid;     // persistent CTAs
    int side;     // 1 = launched on the engine's side stream (downsample conv overlapped with conv1/conv2)
    int join;     // 1 = must wait for the side stream before it starts (consumes the downsample output)
    // quant = 1: not a conv but the BF16 -> E4M3 re-quantisation of an activation tensor (the hand-over from the BF16
    // layers to the FP8 layers of a mixed plan): q_dst[rows][q_cpad] = RN_e4m3(q_src[rows][q_c] * q_inv_scale)
    int quant;
    const void* q_src;
    void* q_dst;
    long long q_rows;
    int q_c, q_cpad;
    float q_inv_scale;
    // FP8: the un-multiplied vectors and the scratch the kernel reads (kept so that calibration can re-scale a plan)
    const float* fp8_wscale;
    const float* fp8_shift;
    float* fp8_vecs;
    double flops;  // 2*M*N*K
    double bytes;  // algorithmic HBM bytes: input + weights + bias (+ residual) read once, output written once
};

// force_bn codes: 0 = heuristic, 64 / 128 = single-CTA tiles, 1128 / 1256 = CTA-pair tiles,
// 3064 = halo-resident 3x3 kernel (only where conv_plan_halo_ok()).
bool conv_plan_halo_ok(const ConvDesc& d);
bool conv_plan_halo2_ok(const ConvDesc& d);  // force_bn code 4064
// Builds the tensor maps and tile geometry. Returns 0 on success; on failure writes a message to
// `err` (if non-null, at most errlen bytes).
int conv_plan_init(ConvPlan* plan, const ConvDesc& d, int num_sms, int force_bn, char* err,
                   int errlen);
bool bneck_plan_ok(int H, int W, int esz);
bool c3n1_shape_ok(int K3, int N3, int N1);
int c3n1_plan_init(ConvPlan* plan, const C3n1Desc& d, int num_sms, char* err, int errlen);
int bneck_plan_init(ConvPlan* plan, const BneckDesc& d, int num_sms, char* err, int errlen);
// FP8: (re)compute the pre-multiplied epilogue vectors of a plan for the given per-tensor scales and point the launch
// geometry at them: chan_scale' = wscale * in / out, bias' = shift / out, res_mul = res / out. Enqueued on `stream`.
cudaError_t fp8_premultiply(ConvPlan* plan, float in_scale, float res_scale, float out_scale, cudaStream_t stream);
// Enqueues the kernel on `stream` (no synchronisation).
cudaError_t conv_plan_launch(const ConvPlan& plan, cudaStream_t stream);
// One-time per process: raise the dynamic shared memory limit of every kernel instantiation.
cudaError_t conv_kernels_init();

}  // namespace rnb
