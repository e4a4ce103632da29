// conv_plan.cu — host side of the implicit-GEMM conv: tile selection, TMA descriptors, launch.
#include "conv_plan.h"

#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "conv_igemm2.cuh"
#include "internal.h"
#include "tensormap.h"

namespace rnb {

// Tile configurations (BN, element bytes, smem stages, staging buffers). All use BM = 128.
using CfgBf16N128 = ConvCfg<128, 2, 4, 3>;
using CfgBf16N64 = ConvCfg<64, 2, 6, 3>;
using CfgTf32N128 = ConvCfg<128, 4, 3, 2>;
using CfgTf32N64 = ConvCfg<64, 4, 5, 3>;

// 2-CTA (cta_group::2) configurations: 256-pixel tiles, BN = 256 or 128.
using Cfg2Bf16N256 = Conv2Cfg<256, 2, 4, 3>;
using Cfg2Bf16N128 = Conv2Cfg<128, 2, 5, 3>;
using Cfg2Tf32N256 = Conv2Cfg<256, 4, 4, 3>;
using Cfg2Tf32N128 = Conv2Cfg<128, 4, 5, 3>;
// FP8 (E4M3, kind::f8f6f4): 128 channels per K block, 128 output bytes per staging row
using CfgFp8N128 = ConvCfg<128, 1, 4, 4>;   // (4 stages, 4 staging boxes: 1 % faster than (5, 3) and (3, 6), profiles/ab_r2.txt)
// BF16 operands, E4M3 output (hand-over convs of the mixed FP8 plan)
using CfgBf16N128F8 = ConvCfg<128, 2, 4, 3, 1>;
using Cfg2Bf16N256F8 = Conv2Cfg<256, 2, 4, 3, 1>;
using Cfg2Bf16N128F8 = Conv2Cfg<128, 2, 5, 3, 1>;
static_assert(CfgBf16N128F8::SMEM_BYTES <= 232448 && Cfg2Bf16N256F8::SMEM_BYTES <= 232448 &&
                  Cfg2Bf16N128F8::SMEM_BYTES <= 232448,
              "smem budget");
// deepest rings (bf16): eight / six stages in flight and ONE staging buffer — force codes 12128 / 12256
using Cfg2Fp8N128DD = Conv2Cfg<128, 1, 8, 1>;
using Cfg2Fp8N256DD = Conv2Cfg<256, 1, 6, 1>;
static_assert(Cfg2Fp8N128DD::SMEM_BYTES <= 232448 && Cfg2Fp8N256DD::SMEM_BYTES <= 232448, "smem budget");
using Cfg2Bf16N128DD = Conv2Cfg<128, 2, 8, 1>;
using Cfg2Bf16N256DD = Conv2Cfg<256, 2, 6, 1>;
static_assert(Cfg2Bf16N128DD::SMEM_BYTES <= 232448 && Cfg2Bf16N256DD::SMEM_BYTES <= 232448, "smem budget");
// resident weights (3x3, 128 -> 128: 18 K blocks = 144 KB per CTA), 3-deep A ring, one staging buffer: force_bn code 31128
using Cfg2Bf16N128R = Conv2Cfg<128, 2, 3, 1, 2, 18>;
static_assert(Cfg2Bf16N128R::SMEM_BYTES <= 232448, "smem budget");
// 16 epilogue warps (one 32-column chunk per warp and tile): force_bn code 20128
using CfgBf16N128W16 = ConvCfg<128, 2, 4, 3, 2, 16>;
using CfgFp8N128W16 = ConvCfg<128, 1, 4, 4, 1, 16>;
static_assert(CfgBf16N128W16::SMEM_BYTES <= 232448 && CfgFp8N128W16::SMEM_BYTES <= 232448, "smem budget");
using CfgFp8N128B = ConvCfg<128, 1, 3, 6>;   // RNB_FP8_CFG=1: shallower K ring, six staging boxes (residual prefetch depth)
using CfgFp8N128C = ConvCfg<128, 1, 5, 3>;   // RNB_FP8_CFG=2
static_assert(CfgFp8N128B::SMEM_BYTES <= 232448 && CfgFp8N128C::SMEM_BYTES <= 232448, "smem budget");
using Cfg2Fp8N256 = Conv2Cfg<256, 1, 5, 3>;
using Cfg2Fp8N128 = Conv2Cfg<128, 1, 6, 3>;
static_assert(CfgFp8N128::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Fp8N256::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Fp8N128::SMEM_BYTES <= 232448, "smem budget");
// deep variants (bf16 only): one/two more pipeline stages, two staging buffers
using CfgBf16N128D = ConvCfg<128, 2, 5, 2>;
using CfgBf16N64F32 = ConvCfg<64, 2, 4, 3, 4>;  // BF16 operands, FP32 output (FC)
static_assert(CfgBf16N64F32::SMEM_BYTES <= 232448, "smem budget");
using Cfg2Bf16N256D = Conv2Cfg<256, 2, 5, 2>;
using Cfg2Bf16N128D = Conv2Cfg<128, 2, 6, 2>;
static_assert(CfgBf16N128D::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Bf16N256D::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Bf16N128D::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Bf16N256::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Bf16N128::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Tf32N256::SMEM_BYTES <= 232448, "smem budget");
static_assert(Cfg2Tf32N128::SMEM_BYTES <= 232448, "smem budget");

static_assert(CfgBf16N128::SMEM_BYTES <= 232448, "smem budget");
static_assert(CfgBf16N64::SMEM_BYTES <= 232448, "smem budget");
static_assert(CfgTf32N128::SMEM_BYTES <= 232448, "smem budget");
static_assert(CfgTf32N64::SMEM_BYTES <= 232448, "smem budget");

template <class Cfg>
static cudaError_t set_smem() {
    return cudaFuncSetAttribute(conv_igemm_kernel<Cfg>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
}
template <class Cfg>
static cudaError_t set_smem2() {
    return cudaFuncSetAttribute(conv_igemm2_kernel<Cfg>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
}

cudaError_t conv_kernels_init() {
    cudaError_t e;
    if ((e = set_smem<CfgBf16N128>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgBf16N64>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgTf32N128>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgTf32N64>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N256>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N128>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Tf32N256>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Tf32N128>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgFp8N128>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgFp8N128B>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgBf16N128F8>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N256F8>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N128F8>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N128R>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N128DD>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Fp8N128DD>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Fp8N256DD>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N256DD>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgBf16N128W16>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgFp8N128W16>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgFp8N128C>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Fp8N256>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Fp8N128>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgBf16N128D>()) != cudaSuccess) return e;
    if ((e = set_smem<CfgBf16N64F32>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N256D>()) != cudaSuccess) return e;
    if ((e = set_smem2<Cfg2Bf16N128D>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_halo_kernel<HaloCfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HaloCfg::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_l1_kernel<BneckCfg<false>>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  BneckCfg<false>::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(conv3x3_halo2_kernel<Halo2Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Halo2Cfg::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_c3n1s_kernel<C3n1sL3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C3n1sL3::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_c3n1s_kernel<C3n1sL3P5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C3n1sL3P5::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_c3n1s_kernel<C3n1sL3W3P5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C3n1sL3W3P5::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_c3n1s_kernel<C3n1sL3P6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C3n1sL3P6::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_c3n1_kernel<C3n1Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C3n1Cfg::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_l1_kernel<BneckCfg<false, 128>>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  BneckCfg<false, 128>::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_l1_kernel<BneckCfg<true, 64, 1>>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  BneckCfg<true, 64, 1>::SMEM_BYTES)) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(bneck_l1_kernel<BneckCfg<true>>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  BneckCfg<true>::SMEM_BYTES)) != cudaSuccess)
        return e;
    return cudaSuccess;
}

// Persistent CTA pairs: with RNB_BALANCE (bit mask: 1 fused layer2/3 tails, 2 fused layer1 tails, 4 generic pair convs,
// 8 TF32 halo convs) the grid is shrunk to ceil(tiles / waves) pairs so that every pair runs the same number of tiles
// (196 tiles: 66 pairs x 3 instead of 48 x 3 + 26 x 2) — the same critical path with less HBM contention per wave.
static int pair_count(int tiles, int num_sms, int bit) {
    const int max_pairs = num_sms / 2;
    if (tiles <= max_pairs) return tiles;
    const int mask = getenv("RNB_BALANCE") ? atoi(getenv("RNB_BALANCE")) : 0;
    if (!(mask & bit)) return max_pairs;
    const int waves = (tiles + max_pairs - 1) / max_pairs;
    return (tiles + waves - 1) / waves;
}

static int fail(char* err, int errlen, const char* msg, int code) {
    if (err && errlen > 0) snprintf(err, errlen, "%s (code %d)", msg, code);
    return code ? code : -1;
}

bool conv_plan_halo_ok(const ConvDesc& d) {
    return d.act == ActType::BF16 && d.ksize == 3 && d.stride == 1 && d.pad == 1 && d.Cin == 64 &&
           d.Cout == 64 && !d.residual && d.W <= 62 && d.W >= 8 && d.H % 2 == 0;
}

static int halo_plan_init(ConvPlan* plan, const ConvDesc& d, int num_sms, char* err, int errlen) {
    if (!conv_plan_halo_ok(d)) return fail(err, errlen, "conv_plan: layer is not eligible for the halo kernel", -7);
    plan->halo = 1;
    plan->bn = 64;
    plan->esz = 2;
    plan->ctas = 1;
    plan->bias = d.bias;
    HaloGeom& g = plan->hg;
    g.N = d.B; g.H = d.H; g.W = d.W;
    g.rows_per_img = d.H / 2;
    g.tiles = d.B * g.rows_per_img;
    g.relu = d.relu ? 1 : 0;
    g.reverse = d.reverse ? 1 : 0;
    plan->grid = g.tiles < num_sms ? g.tiles : num_sms;
    const double M = 1.0 * d.B * d.H * d.W;
    plan->flops = 2.0 * M * 64 * 576;
    plan->bytes = 2.0 * M * 64 * 2 + 64.0 * 576 * 2 + 256;
    const uint64_t dims[4] = {64, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
    const uint64_t strides[3] = {128, 128ull * d.W, 128ull * d.W * d.H};
    const uint32_t box_in[4] = {64, 64, 2, 1};
    const uint32_t box_out[4] = {64, static_cast<uint32_t>(d.W), 1, 1};
    int r;
    if ((r = make_tiled_nd(&plan->tmA, TmDtype::BF16, d.in, 4, dims, strides, box_in, true)) != 0)
        return fail(err, errlen, "conv_plan: halo input tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmB, TmDtype::BF16, d.weight, 64, 576, 64)) != 0)
        return fail(err, errlen, "conv_plan: halo weight tensor map failed", r);
    if ((r = make_tiled_nd(&plan->tmOut, TmDtype::BF16, d.out, 4, dims, strides, box_out, true)) != 0)
        return fail(err, errlen, "conv_plan: halo output tensor map failed", r);
    plan->tmRes = plan->tmOut;
    return 0;
}

bool conv_plan_halo2_ok(const ConvDesc& d) {
    return d.act == ActType::TF32 && d.ksize == 3 && d.stride == 1 && d.pad == 1 && d.Cin == 64 && d.Cout == 64 &&
           d.W <= 62 && d.W >= 8 && d.H >= 4 && d.H % 4 == 0 && !d.out_f32;
}

static int halo2_plan_init(ConvPlan* plan, const ConvDesc& d, int num_sms, char* err, int errlen) {
    if (!conv_plan_halo2_ok(d)) return fail(err, errlen, "conv_plan: layer is not eligible for the TF32 halo kernel", -7);
    plan->halo2 = 1;
    plan->bn = 64;
    plan->esz = 4;
    plan->ctas = 2;
    plan->bias = d.bias;
    Halo2Geom& g = plan->h2g;
    g.N = d.B; g.H = d.H; g.W = d.W;
    g.tiles_per_img = d.H / 4;
    g.tiles = d.B * g.tiles_per_img;
    g.relu = d.relu ? 1 : 0;
    g.has_res = d.residual ? 1 : 0;
    g.reverse = d.reverse ? 1 : 0;
    const int pairs = pair_count(g.tiles, num_sms, 8);
    plan->grid = 2 * pairs;
    const double M = 1.0 * d.B * d.H * d.W;
    plan->flops = 2.0 * M * 64 * 576;
    plan->bytes = (d.residual ? 3.0 : 2.0) * M * 64 * 4 + 64.0 * 576 * 4 + 256;
    const uint64_t dims[4] = {64, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
    const uint64_t strides[3] = {256, 256ull * d.W, 256ull * d.W * d.H};
    const uint32_t box_in[4] = {32, 64, 2, 1};
    const uint32_t box_out[4] = {32, static_cast<uint32_t>(d.W), 1, 1};
    int r;
    if ((r = make_tiled_nd(&plan->tmA, TmDtype::F32, d.in, 4, dims, strides, box_in, true)) != 0)
        return fail(err, errlen, "conv_plan: halo2 input tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmB, TmDtype::F32, d.weight, 64, 576, 32)) != 0)
        return fail(err, errlen, "conv_plan: halo2 weight tensor map failed", r);
    if ((r = make_tiled_nd(&plan->tmOut, TmDtype::F32, d.out, 4, dims, strides, box_out, true)) != 0)
        return fail(err, errlen, "conv_plan: halo2 output tensor map failed", r);
    plan->tmRes = plan->tmA;
    if (d.residual && (r = make_tiled_nd(&plan->tmRes, TmDtype::F32, d.residual, 4, dims, strides, box_in, true)) != 0)
        return fail(err, errlen, "conv_plan: halo2 residual tensor map failed", r);
    return 0;
}

bool bneck_plan_ok(int H, int W, int esz) { return esz == 2 && W >= 8 && W <= 62 && H >= 4 && H % 4 == 0; }

int bneck_plan_init(ConvPlan* plan, const BneckDesc& d, int num_sms, char* err, int errlen) {
    memset(plan, 0, sizeof(*plan));
    if (!bneck_plan_ok(d.H, d.W, 2)) return fail(err, errlen, "bneck_plan: geometry not eligible", -7);
    if (d.w1n && d.n1 != 64 && (d.n1 != 128 || d.wds))
        return fail(err, errlen, "bneck_plan: next conv1 must have 64 (or, without folded downsample, 128) outputs", -7);
    const int n1 = d.w1n ? d.n1 : 64;
    plan->bneck = d.wds ? 2 : (n1 == 128 ? 5 : 1);
    plan->bn = 256;
    plan->esz = 2;
    plan->ctas = 2;
    BneckGeom& g = plan->bg;
    g.N = d.B; g.H = d.H; g.W = d.W;
    g.tiles_per_img = d.H / 4;
    g.tiles = d.B * g.tiles_per_img;
    g.has_next = d.w1n ? 1 : 0;
    g.reverse = d.reverse ? 1 : 0;
    g.debug = getenv("RNB_BNECK_DEBUG") ? atoi(getenv("RNB_BNECK_DEBUG")) : 0;
    plan->bp.bias2 = d.bias2;
    plan->bp.bias3 = d.bias3;
    plan->bp.bias1n = d.bias1n ? d.bias1n : d.bias2;
    const int pairs = pair_count(g.tiles, num_sms, 2);
    plan->grid = 2 * pairs;
    const double M = 1.0 * d.B * d.H * d.W;
    const double macs = 64.0 * 576 + 256.0 * 64 + (d.wds ? 256.0 * 64 : 0.0) + (d.w1n ? 1.0 * n1 * 256 : 0.0);
    plan->flops = 2.0 * M * macs;
    plan->bytes = 2.0 * M * (64 + (d.wds ? 64 : 256) + 256 + (d.w1n ? n1 : 0)) + 2.0 * macs + 4.0 * (64 + 256 + 64);
    auto act_map = [&](CUtensorMap* tm, const void* base, int C, uint32_t bw, uint32_t bh) {
        const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                                  static_cast<uint64_t>(d.B)};
        const uint64_t strides[3] = {2ull * C, 2ull * C * d.W, 2ull * C * d.W * d.H};
        const uint32_t box[4] = {64, bw, bh, 1};
        return make_tiled_nd(tm, TmDtype::BF16, base, 4, dims, strides, box, true);
    };
    int r;
    if ((r = act_map(&plan->tmA, d.t1, 64, 64, 4)) != 0) return fail(err, errlen, "bneck_plan: t1 tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmB, TmDtype::BF16, d.w2, 64, 576, 32)) != 0)
        return fail(err, errlen, "bneck_plan: w2 tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmW3, TmDtype::BF16, d.w3, 256, 64, 64)) != 0)
        return fail(err, errlen, "bneck_plan: w3 tensor map failed", r);
    plan->tmWds = plan->tmW3;
    if (d.wds && (r = make_tiled_2d(&plan->tmWds, TmDtype::BF16, d.wds, 256, 64, 64)) != 0)
        return fail(err, errlen, "bneck_plan: wds tensor map failed", r);
    if ((r = act_map(&plan->tmRes, d.shortcut, d.wds ? 64 : 256, 64, 2)) != 0)
        return fail(err, errlen, "bneck_plan: shortcut tensor map failed", r);
    if ((r = act_map(&plan->tmOut, d.y, 256, static_cast<uint32_t>(d.W), 1)) != 0)
        return fail(err, errlen, "bneck_plan: y tensor map failed", r);
    plan->tmW1n = plan->tmB;
    plan->tmT1n = plan->tmOut;
    if (d.w1n) {
        if ((r = make_tiled_2d(&plan->tmW1n, TmDtype::BF16, d.w1n, n1, 256, n1 / 2)) != 0)
            return fail(err, errlen, "bneck_plan: w1n tensor map failed", r);
        if ((r = act_map(&plan->tmT1n, d.t1n, n1, static_cast<uint32_t>(d.W), 1)) != 0)
            return fail(err, errlen, "bneck_plan: t1n tensor map failed", r);
    }
    return 0;
}

bool c3n1_shape_ok(int K3, int N3, int N1) {
    return (K3 == 128 && N3 == 512 && N1 == 128) || (K3 == 256 && N3 == 1024 && N1 == 256);
}

int c3n1_plan_init(ConvPlan* plan, const C3n1Desc& d, int num_sms, char* err, int errlen) {
    memset(plan, 0, sizeof(*plan));
    if (d.M <= 0) return fail(err, errlen, "c3n1_plan: bad M", -5);
    if (!c3n1_shape_ok(d.K3, d.N3, d.N1)) return fail(err, errlen, "c3n1_plan: unsupported channel counts", -7);
    const bool streamed = d.K3 != 128;
    plan->bneck = streamed ? 4 : 3;
    plan->rings = getenv("RNB_C3N1S_RINGS") ? atoi(getenv("RNB_C3N1S_RINGS")) : 0;
    plan->bn = 128;
    plan->esz = 2;
    plan->ctas = 2;
    plan->cg.M = d.M;
    plan->cg.tiles = (d.M + 255) / 256;
    plan->cg.reverse = d.reverse ? 1 : 0;
    plan->cp.bias3 = d.bias3;
    plan->cp.bias1n = d.bias1n;
    const int pairs = pair_count(plan->cg.tiles, num_sms, 1);
    plan->grid = 2 * pairs;
    const double M = d.M;
    const double wel = 1.0 * d.N3 * d.K3 + 1.0 * d.N1 * d.N3;
    plan->flops = 2.0 * M * wel;
    plan->bytes = 2.0 * M * (d.K3 + 2.0 * d.N3 + d.N1) + 2.0 * wel + 4.0 * (d.N3 + d.N1);
    int r;
    if ((r = make_tiled_2d(&plan->tmA, TmDtype::BF16, d.t2, d.M, d.K3, 128)) != 0)
        return fail(err, errlen, "c3n1_plan: t2 tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmB, TmDtype::BF16, d.w3, d.N3, d.K3, 64)) != 0)
        return fail(err, errlen, "c3n1_plan: w3 tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmW1n, TmDtype::BF16, d.w1n, d.N1, d.N3, streamed ? d.N1 / 2 : 64)) != 0)
        return fail(err, errlen, "c3n1_plan: w1n tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmRes, TmDtype::BF16, d.residual, d.M, d.N3, 128)) != 0)
        return fail(err, errlen, "c3n1_plan: residual tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmOut, TmDtype::BF16, d.y, d.M, d.N3, 128)) != 0)
        return fail(err, errlen, "c3n1_plan: y tensor map failed", r);
    if ((r = make_tiled_2d(&plan->tmT1n, TmDtype::BF16, d.t1n, d.M, d.N1, 128)) != 0)
        return fail(err, errlen, "c3n1_plan: t1n tensor map failed", r);
    plan->tmW3 = plan->tmB;
    plan->tmWds = plan->tmB;
    return 0;
}

int conv_plan_init(ConvPlan* plan, const ConvDesc& d, int num_sms, int force_bn, char* err,
                   int errlen) {
    memset(plan, 0, sizeof(*plan));
    if (d.out_fp8 && d.act != ActType::BF16)
        return fail(err, errlen, "conv_plan: E4M3 output with other operands than BF16 is the FP8 path itself", -9);
    if (d.act == ActType::FP8 || d.out_fp8) {
        if (d.out_f32 || !d.chan_scale || !d.fp8_vecs || d.Cout % 128 != 0 || !(d.out_scale > 0.f))
            return fail(err, errlen, "conv_plan: FP8 needs channel scales + scratch, Cout % 128 == 0 and a positive output scale", -9);
        if (force_bn == 64 || force_bn == 3064 || force_bn == 4064 ||
            (force_bn >= 10000 && force_bn < 20000 && !(d.act == ActType::FP8 && (force_bn == 12128 || force_bn == 12256))))
            return fail(err, errlen, "conv_plan: tile family not available in FP8", -9);
    }
    if (d.out_f32) {
        if (d.act != ActType::BF16 || d.residual || d.out_cols <= 0 || d.out_cols > d.Cout || d.out_cols % 4 != 0)
            return fail(err, errlen, "conv_plan: FP32-output mode needs BF16 operands, no residual, out_cols % 4 == 0", -8);
        force_bn = 64;
    }
    if (force_bn == 3064 || (force_bn == 0 && d.act != ActType::FP8 && !d.out_fp8 && conv_plan_halo_ok(d) && !getenv("RNB_NO_HALO")))
        return halo_plan_init(plan, d, num_sms, err, errlen);
    if (force_bn == 4064 || (force_bn == 0 && conv_plan_halo2_ok(d) && !getenv("RNB_NO_HALO")))
        return halo2_plan_init(plan, d, num_sms, err, errlen);
    // +10000: "deep" variant of a tile family — one or two more shared-memory stages in flight paid
    // for with one staging buffer less (for layers whose K loop, not whose epilogue, is the bottleneck)
    // 31128: CTA-pair BN = 128 tiles with the whole weight matrix resident in shared memory (one N tile, 18 K blocks)
    bool resb = false;
    if (force_bn == 31128) {
        if (d.act != ActType::BF16 || d.out_f32 || d.out_fp8 || d.Cout != 128 || d.ksize * d.ksize * d.Cin != 18 * 64)
            return fail(err, errlen, "conv_plan: the resident-weight kernel is the bf16 3x3 128 -> 128 one", -9);
        resb = true;
        force_bn = 1128;
    }
    // +20000: single-CTA BN = 128 tiles with SIXTEEN epilogue warps (one 32-column chunk per warp and tile)
    if (force_bn == 20128) {
        if (d.act == ActType::TF32 || d.out_f32 || d.out_fp8 || d.Cout % 128 != 0)
            return fail(err, errlen, "conv_plan: the 16-warp epilogue exists for bf16 / fp8 tiles of 128 columns", -9);
        plan->w16 = 1;
        force_bn = 128;
    }
    // +10000 / +11000: deep (one / two more stages, two staging buffers) / deepest (ring as deep as shared memory allows,
    // one staging buffer; CTA pairs only) variants of a tile family
    int deep = force_bn >= 10000 ? 1 : 0;
    if (force_bn == 12128 || force_bn == 12256) {
        if (d.act == ActType::TF32 || d.out_fp8)
            return fail(err, errlen, "conv_plan: the deepest-ring variants are bf16 / fp8 kernels", -9);
        deep = 2;
        force_bn -= 1000;
    }
    if (deep) force_bn -= 10000;
    plan->deep = deep;
    const int esz = static_cast<int>(d.act);
    const int osz = d.out_fp8 ? 1 : esz;
    const int bk = 128 / esz;
    if (d.ksize != 1 && d.ksize != 3) return fail(err, errlen, "conv_plan: ksize must be 1 or 3", -2);
    if (d.Cin % bk != 0) return fail(err, errlen, "conv_plan: Cin must be a multiple of the K block", -3);
    if (d.Cout % 64 != 0) return fail(err, errlen, "conv_plan: Cout must be a multiple of 64", -4);
    const int OH = (2 * d.pad + d.H - d.ksize) / d.stride + 1;  // convOutputSize, ops.cuh:9-13
    const int OW = (2 * d.pad + d.W - d.ksize) / d.stride + 1;
    const long long M = 1LL * d.B * OH * OW;
    if (M <= 0 || M > 0x7fffffffLL) return fail(err, errlen, "conv_plan: bad M", -5);

    // Tile selection. force_bn: 64 / 128 = single-CTA tiles, 1128 / 1256 = CTA-pair tiles (testing).
    // Default: the CTA-pair kernel (256-pixel tiles, half the L2 bytes per MAC) whenever the layer
    // offers at least one full wave of pair tiles; single-CTA 128-pixel tiles otherwise.
    int bn = (d.Cout % 128 == 0) ? 128 : 64;
    int ctas = 1;
    const char* kind_env = getenv("RNB_CONV_KIND");  // "1" forces single-CTA tiles everywhere
    const bool allow_pair = !(kind_env && atoi(kind_env) == 1);
    if (force_bn == 64 || force_bn == 128) {
        if (d.Cout % force_bn != 0) return fail(err, errlen, "conv_plan: forced BN does not divide Cout", -6);
        bn = force_bn;
    } else if (force_bn == 1128 || force_bn == 1256) {
        if (d.Cout % (force_bn - 1000) != 0) return fail(err, errlen, "conv_plan: forced BN does not divide Cout", -6);
        bn = force_bn - 1000;
        ctas = 2;
    } else if (allow_pair && d.Cout % 128 == 0 && (!d.residual || d.ksize == 3 || getenv("RNB_PAIR_RES")) &&
               d.ksize * d.ksize * d.Cin > 64) {
        // Measured on B200 (profiles/): pairs win wherever the K loop dominates (3x3 and wide 1x1
        // layers, up to 1.45x); layers whose time is the epilogue's HBM traffic (1x1 with residual add, or a
        // single K block) are no faster with pairs and slightly slower, so they keep 128-pixel tiles. A 3x3
        // conv with residual (conv2 of a BasicBlock) is K-loop-bound: pairs (88-106 us vs 148 us, TF32).
        const int pbn = d.Cout % 256 == 0 ? 256 : 128;
        const long long pair_tiles = ((M + 255) / 256) * (d.Cout / pbn);
        if (pair_tiles >= num_sms / 2) {
            bn = pbn;
            ctas = 2;
        }
    }
    ConvGeom& g = plan->g;
    g.M = static_cast<int>(M);
    g.OH = OH;
    g.OW = OW;
    g.Cout = d.Cout;
    g.stride = d.stride;
    g.lower = -d.pad;
    g.ksize = d.ksize;
    g.kblocks_per_tap = d.Cin / bk;
    g.num_kblocks = d.ksize * d.ksize * g.kblocks_per_tap;
    g.m_tiles = static_cast<int>(ctas == 2 ? (M + 255) / 256 : (M + 127) / 128);
    g.n_tiles = d.Cout / bn;
    g.relu = d.relu ? 1 : 0;
    g.has_res = d.residual ? 1 : 0;
    g.reverse = d.reverse ? 1 : 0;
    plan->bias = d.bias;
    plan->bn = bn;
    plan->esz = esz;
    plan->osz = osz;
    plan->ctas = ctas;
    plan->resb = resb ? 1 : 0;
    const int tiles = g.m_tiles * g.n_tiles;
    g.split_from = tiles;
    g.chan_scale = d.chan_scale;
    g.res_mul = 1.f;
    g.amax = nullptr;
    plan->fp8_wscale = d.chan_scale;
    plan->fp8_shift = d.bias;
    plan->fp8_vecs = d.fp8_vecs;
    if (ctas == 2) {
        int pairs = pair_count(tiles, num_sms, 4);
        // Tail split (conv_igemm2.cuh): when the last wave of the persistent grid fills at most half of the
        // pairs, its tiles run as two N halves on twice as many pairs: 98 tiles on 74 pairs take 1.5 instead
        // of 2 tile times (layer4 at 256 images, layer3 of ResNet-152 at 128). Bit-identical. RNB_NO_SPLIT=1: off.
        const bool no_split = getenv("RNB_NO_SPLIT") && atoi(getenv("RNB_NO_SPLIT")) != 0;
        const int nsub = osz == 1 ? bn / 128 : bn / (2 * (128 / esz));  // Conv2Cfg::NSUB
        if (!no_split && nsub % 2 == 0 && !resb) {
            const int max_pairs = num_sms / 2;
            const int rem = tiles % pairs;
            if (tiles > pairs && rem > 0 && 2 * rem <= pairs) {
                g.split_from = tiles - rem;
            } else if (2 * tiles <= max_pairs) {
                g.split_from = 0;  // less than half a wave: every tile as two halves on 2 x tiles pairs
                pairs = 2 * tiles;
            }
        }
        plan->grid = 2 * pairs;
    } else {
        plan->grid = tiles < num_sms ? tiles : num_sms;
    }
    plan->flops = 2.0 * static_cast<double>(M) * d.Cout * (1.0 * d.ksize * d.ksize * d.Cin);
    plan->bytes = 1.0 * d.B * d.H * d.W * d.Cin * esz + 1.0 * d.Cout * d.ksize * d.ksize * d.Cin * esz +
                  4.0 * d.Cout + (d.residual ? 2.0 : 1.0) * static_cast<double>(M) * d.Cout * osz;

    const TmDtype dt = d.act == ActType::BF16 ? TmDtype::BF16 : (d.act == ActType::FP8 ? TmDtype::U8 : TmDtype::F32);
    int r;
    if ((r = make_im2col_nhwc(&plan->tmA, dt, d.in, d.B, d.H, d.W, d.Cin, d.ksize, d.stride, d.pad,
                              128)) != 0)
        return fail(err, errlen, "conv_plan: im2col tensor map (A) failed", r);
    const uint64_t K = 1ull * d.ksize * d.ksize * d.Cin;
    if ((r = make_tiled_2d(&plan->tmB, dt, d.weight, d.Cout, K, ctas == 2 ? bn / 2 : bn)) != 0)
        return fail(err, errlen, "conv_plan: tiled tensor map (B) failed", r);
    plan->tmBh = plan->tmB;
    if (ctas == 2 && g.split_from < tiles && (r = make_tiled_2d(&plan->tmBh, dt, d.weight, d.Cout, K, bn / 4)) != 0)
        return fail(err, errlen, "conv_plan: tiled tensor map (B, half units) failed", r);
    plan->f32out = d.out_f32 ? 1 : 0;
    if (d.out_f32) {
        if ((r = make_tiled_2d(&plan->tmOut, TmDtype::F32, d.out, M, d.out_cols, 128)) != 0)
            return fail(err, errlen, "conv_plan: tiled tensor map (fp32 out) failed", r);
        plan->bytes = 1.0 * d.B * d.H * d.W * d.Cin * esz + 1.0 * d.out_cols * d.Cin * esz + 4.0 * d.out_cols +
                      4.0 * static_cast<double>(M) * d.out_cols;
        plan->flops = 2.0 * static_cast<double>(M) * d.out_cols * d.Cin;
    } else if ((r = make_tiled_2d(&plan->tmOut, d.out_fp8 ? TmDtype::U8 : dt, d.out, M, d.Cout, 128)) != 0)
        return fail(err, errlen, "conv_plan: tiled tensor map (out) failed", r);
    const void* res = d.residual ? d.residual : d.out;
    if (d.out_f32)
        plan->tmRes = plan->tmOut;
    else if ((r = make_tiled_2d(&plan->tmRes, d.out_fp8 ? TmDtype::U8 : dt, res, M, d.Cout, 128)) != 0)
        return fail(err, errlen, "conv_plan: tiled tensor map (residual) failed", r);
    return 0;
}

namespace {
__global__ void fp8_premultiply_kernel(const float* __restrict__ wscale, const float* __restrict__ shift,
                                       float* __restrict__ vecs, int n, float in_over_out, float inv_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        vecs[i] = wscale[i] * in_over_out;
        vecs[n + i] = shift[i] * inv_out;
    }
}
}  // namespace

cudaError_t fp8_premultiply(ConvPlan* plan, float in_scale, float res_scale, float out_scale, cudaStream_t stream) {
    const int n = plan->g.Cout;
    fp8_premultiply_kernel<<<(n + 255) / 256, 256, 0, stream>>>(plan->fp8_wscale, plan->fp8_shift, plan->fp8_vecs, n,
                                                               in_scale / out_scale, 1.f / out_scale);
    plan->g.chan_scale = plan->fp8_vecs;
    plan->bias = plan->fp8_vecs + n;
    plan->g.res_mul = res_scale / out_scale;
    return cudaGetLastError();
}

// Every conv kernel is launched with programmatic stream serialization (PDL): its prologue may run
// while the previous kernel of the stream drains; the kernels call griddepcontrol.wait before they
// touch global tensors. RNB_NO_PDL=1 turns the attribute off (plain stream order).
static bool pdl_enabled() {
    static const bool on = !(getenv("RNB_NO_PDL") && atoi(getenv("RNB_NO_PDL")) != 0);
    return on;
}

template <class Kernel, class... Args>
static cudaError_t launch_pdl(Kernel kernel, int grid, int threads, int smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <class Cfg>
static cudaError_t launch(const ConvPlan& p, cudaStream_t stream) {
    return launch_pdl(conv_igemm_kernel<Cfg>, p.grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, p.tmA, p.tmB,
                      p.tmOut, p.tmRes, p.bias, p.g);
}

template <class Cfg>
static cudaError_t launch2(const ConvPlan& p, cudaStream_t stream) {
    return launch_pdl(conv_igemm2_kernel<Cfg>, p.grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, p.tmA, p.tmB,
                      p.tmBh, p.tmOut, p.tmRes, p.bias, p.g);
}

#ifdef RNB_TIMELINE
// tools/c3n1s_timeline.py: copies the in-kernel timeline of the last bneck_c3n1s launch to the host
extern "C" int rnb_debug_read_timeline(long long* host, int* counts) {
    if (cudaMemcpyFromSymbol(counts, g_c3n1s_tl_n, 4 * sizeof(int)) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(host, g_c3n1s_tl, 4 * kTlMax * sizeof(long long)) != cudaSuccess) return -1;
    return kTlMax;
}
#endif

cudaError_t conv_plan_launch(const ConvPlan& p, cudaStream_t stream) {
    if (p.quant) return launch_quantize_pad_bf16(p.q_src, p.q_dst, p.q_rows, p.q_c, p.q_cpad, p.q_inv_scale, stream);
    if (p.bneck == 4) {
        // how the 224 KB of shared memory are split between the weight rings and the staging boxes (bit-identical)
        const int rings = p.rings;
        if (rings == 1)
            return launch_pdl(bneck_c3n1s_kernel<C3n1sL3P5>, p.grid, C3n1sL3P5::THREADS, C3n1sL3P5::SMEM_BYTES, stream,
                              p.tmA, p.tmB, p.tmW1n, p.tmRes, p.tmOut, p.tmT1n, p.cp, p.cg);
        if (rings == 2)
            return launch_pdl(bneck_c3n1s_kernel<C3n1sL3W3P5>, p.grid, C3n1sL3W3P5::THREADS, C3n1sL3W3P5::SMEM_BYTES,
                              stream, p.tmA, p.tmB, p.tmW1n, p.tmRes, p.tmOut, p.tmT1n, p.cp, p.cg);
        if (rings == 3)
            return launch_pdl(bneck_c3n1s_kernel<C3n1sL3P6>, p.grid, C3n1sL3P6::THREADS, C3n1sL3P6::SMEM_BYTES, stream,
                              p.tmA, p.tmB, p.tmW1n, p.tmRes, p.tmOut, p.tmT1n, p.cp, p.cg);
        return launch_pdl(bneck_c3n1s_kernel<C3n1sL3>, p.grid, C3n1sL3::THREADS, C3n1sL3::SMEM_BYTES, stream, p.tmA, p.tmB,
                          p.tmW1n, p.tmRes, p.tmOut, p.tmT1n, p.cp, p.cg);
    }
    if (p.bneck == 3)
        return launch_pdl(bneck_c3n1_kernel<C3n1Cfg>, p.grid, C3n1Cfg::THREADS, C3n1Cfg::SMEM_BYTES, stream, p.tmA, p.tmB,
                          p.tmW1n, p.tmRes, p.tmOut, p.tmT1n, p.cp, p.cg);
    if (p.bneck == 1)
        return launch_pdl(bneck_l1_kernel<BneckCfg<false>>, p.grid, BneckCfg<false>::THREADS,
                          BneckCfg<false>::SMEM_BYTES, stream, p.tmA, p.tmB, p.tmW3, p.tmWds, p.tmW1n, p.tmRes,
                          p.tmOut, p.tmT1n, p.bp, p.bg);
    if (p.bneck == 5)
        return launch_pdl(bneck_l1_kernel<BneckCfg<false, 128>>, p.grid, BneckCfg<false, 128>::THREADS,
                          BneckCfg<false, 128>::SMEM_BYTES, stream, p.tmA, p.tmB, p.tmW3, p.tmWds, p.tmW1n, p.tmRes,
                          p.tmOut, p.tmT1n, p.bp, p.bg);
    if (p.bneck == 2) {
        // RNB_DS_NABUF=2: two input tiles in flight and three staging boxes (the first form of this kernel)
        static const bool one_tile = !(getenv("RNB_DS_NABUF") && atoi(getenv("RNB_DS_NABUF")) == 2);
        if (one_tile)
            return launch_pdl(bneck_l1_kernel<BneckCfg<true, 64, 1>>, p.grid, BneckCfg<true, 64, 1>::THREADS,
                              BneckCfg<true, 64, 1>::SMEM_BYTES, stream, p.tmA, p.tmB, p.tmW3, p.tmWds, p.tmW1n,
                              p.tmRes, p.tmOut, p.tmT1n, p.bp, p.bg);
        return launch_pdl(bneck_l1_kernel<BneckCfg<true>>, p.grid, BneckCfg<true>::THREADS,
                          BneckCfg<true>::SMEM_BYTES, stream, p.tmA, p.tmB, p.tmW3, p.tmWds, p.tmW1n, p.tmRes,
                          p.tmOut, p.tmT1n, p.bp, p.bg);
    }
    if (p.halo2)
        return launch_pdl(conv3x3_halo2_kernel<Halo2Cfg>, p.grid, Halo2Cfg::THREADS, Halo2Cfg::SMEM_BYTES, stream, p.tmA,
                          p.tmB, p.tmRes, p.tmOut, p.bias, p.h2g);
    if (p.halo) {
        return launch_pdl(conv3x3_halo_kernel<HaloCfg>, p.grid, HaloCfg::THREADS, HaloCfg::SMEM_BYTES, stream,
                          p.tmA, p.tmB, p.tmOut, p.bias, p.hg);
    }
    if (p.esz == 1) {
        if (p.ctas == 2 && p.deep == 2) return p.bn == 256 ? launch2<Cfg2Fp8N256DD>(p, stream) : launch2<Cfg2Fp8N128DD>(p, stream);
        if (p.ctas == 2) return p.bn == 256 ? launch2<Cfg2Fp8N256>(p, stream) : launch2<Cfg2Fp8N128>(p, stream);
        if (p.w16) return launch<CfgFp8N128W16>(p, stream);
        static const int cfg = getenv("RNB_FP8_CFG") ? atoi(getenv("RNB_FP8_CFG")) : 0;
        if (cfg == 1) return launch<CfgFp8N128B>(p, stream);
        if (cfg == 2) return launch<CfgFp8N128C>(p, stream);
        return launch<CfgFp8N128>(p, stream);
    }
    if (p.f32out) return launch<CfgBf16N64F32>(p, stream);
    if (p.esz == 2 && p.osz == 1) {
        if (p.ctas == 2) return p.bn == 256 ? launch2<Cfg2Bf16N256F8>(p, stream) : launch2<Cfg2Bf16N128F8>(p, stream);
        return launch<CfgBf16N128F8>(p, stream);
    }
    if (p.w16 && p.esz == 2) return launch<CfgBf16N128W16>(p, stream);
    if (p.deep == 2 && p.esz == 2 && p.ctas == 2)
        return p.bn == 256 ? launch2<Cfg2Bf16N256DD>(p, stream) : launch2<Cfg2Bf16N128DD>(p, stream);
    if (p.deep && p.esz == 2) {
        if (p.ctas == 2) return p.bn == 256 ? launch2<Cfg2Bf16N256D>(p, stream) : launch2<Cfg2Bf16N128D>(p, stream);
        if (p.bn == 128) return launch<CfgBf16N128D>(p, stream);
    }
    if (p.resb) return launch2<Cfg2Bf16N128R>(p, stream);
    if (p.ctas == 2) {
        if (p.esz == 2) return p.bn == 256 ? launch2<Cfg2Bf16N256>(p, stream) : launch2<Cfg2Bf16N128>(p, stream);
        return p.bn == 256 ? launch2<Cfg2Tf32N256>(p, stream) : launch2<Cfg2Tf32N128>(p, stream);
    }
    if (p.esz == 2) return p.bn == 128 ? launch<CfgBf16N128>(p, stream) : launch<CfgBf16N64>(p, stream);
    return p.bn == 128 ? launch<CfgTf32N128>(p, stream) : launch<CfgTf32N64>(p, stream);
}

}  // namespace rnb
