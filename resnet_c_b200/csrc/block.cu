// block.cu — rnb_block_*: ONE residual block (or one conv + BN) as a planned object behind the reference's module
// classes (cuda/nn.cuh Bottleneck / BasicBlock / Conv2d, whose forwards /root/reference/cuda/inference/main.cu:127-166
// spells out op by op: conv2dForwardKernel -> batchNorm2dForwardKernel -> reluForwardKernel -> ... -> addForwardKernel).
//
// What the per-call path (rnb_conv_bn_act_forward) re-did on every forward now happens ONCE, at creation: BN folded
// into the weights in FP64 and the result packed K-major in the activation type; per input shape (cached) the NHWC
// staging tensors and one descriptor-complete ConvPlan per launch. A forward is then: NCHW -> NHWC of x, the block's
// 2-4 tensor-core launches with every intermediate kept in NHWC, NHWC -> NCHW of the result. Layer1-shaped Bottlenecks
// (64 -> 64 -> 256, stride 1, BF16) run conv2 + conv3 + shortcut (+ folded downsample) as the fused bneck_l1 launch,
// exactly as the whole-model planner does (model.cu).
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/rnb.h"
#include "conv_plan.h"
#include "internal.h"

using namespace rnb;

namespace {

struct PackedConv {
    void* w = nullptr;
    float* bias = nullptr;
    int Cin = 0, Cout = 0, k = 0, stride = 1, pad = 0;
};

struct ShapePlan {
    std::vector<void*> bufs;       // owned device buffers
    void* x = nullptr;             // NHWC input staging
    void* res = nullptr;           // NHWC residual staging (conv kind with an explicit residual)
    void* y = nullptr;             // NHWC output
    int OH = 0, OW = 0, Cout = 0;
    std::vector<ConvPlan> launches;
};

}  // namespace

struct rnb_block {
    int kind = 0, esz = 2, device = 0;
    std::vector<PackedConv> convs;
    float* bias3ds = nullptr;  // Bottleneck with downsample: conv3 shift + downsample shift (folded-downsample kernel)
    std::map<std::tuple<int, int, int, int, int>, ShapePlan> plans;  // (B, H, W, has residual, relu)
    ~rnb_block() {
        for (auto& kv : plans)
            for (void* p : kv.second.bufs) cudaFree(p);
        for (PackedConv& c : convs) {
            cudaFree(c.w);
            cudaFree(c.bias);
        }
        cudaFree(bias3ds);
    }
};

#define BLK_CUDA(expr)                                        \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return fail_cuda(e__, #expr); \
    } while (0)

static int out_size(int x, int k, int stride, int pad) { return (2 * pad + x - k) / stride + 1; }  // ops.cuh:9-13

static int build_shape_plan(rnb_block* b, int B, int H, int W, bool has_res, bool relu, ShapePlan* sp) {
    const int esz = b->esz;
    const ActType act = esz == 2 ? ActType::BF16 : ActType::TF32;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, b->device);
    auto alloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
        sp->bufs.push_back(p);
        return p;
    };
    auto act_bytes = [&](int C, int h, int w) { return 1ull * B * h * w * C * esz; };
    char err[256];
    auto add = [&](const PackedConv& c, const void* in, int h, int w, const void* res, bool do_relu, void* out) -> int {
        ConvDesc d{};
        d.B = B; d.H = h; d.W = w; d.Cin = c.Cin; d.Cout = c.Cout; d.ksize = c.k; d.stride = c.stride; d.pad = c.pad;
        d.relu = do_relu; d.act = act; d.in = in; d.weight = c.w; d.bias = c.bias; d.residual = res; d.out = out;
        ConvPlan cp;
        if (conv_plan_init(&cp, d, nsm, 0, err, sizeof(err))) {
            set_error(err);
            return RNB_ERR_UNSUPPORTED;
        }
        sp->launches.push_back(cp);
        return RNB_OK;
    };
    const PackedConv& c1 = b->convs[0];
    if (!(sp->x = alloc(act_bytes(c1.Cin, H, W)))) return fail_cuda(cudaGetLastError(), "rnb_block: staging alloc");
    int rc;
    if (b->kind == RNB_BLOCK_CONV) {
        sp->OH = out_size(H, c1.k, c1.stride, c1.pad);
        sp->OW = out_size(W, c1.k, c1.stride, c1.pad);
        sp->Cout = c1.Cout;
        if (!(sp->y = alloc(act_bytes(c1.Cout, sp->OH, sp->OW)))) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
        if (has_res && !(sp->res = alloc(act_bytes(c1.Cout, sp->OH, sp->OW))))
            return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
        return add(c1, sp->x, H, W, sp->res, relu, sp->y);
    }
    const bool bott = b->kind == RNB_BLOCK_BOTTLENECK;
    const PackedConv& c2 = b->convs[1];
    const PackedConv* c3 = bott ? &b->convs[2] : nullptr;
    const PackedConv* ds = b->convs.size() > (bott ? 3u : 2u) ? &b->convs.back() : nullptr;
    const int stride = bott ? c2.stride : c1.stride;
    sp->OH = out_size(H, 3, stride, 1);
    sp->OW = out_size(W, 3, stride, 1);
    sp->Cout = bott ? c3->Cout : c2.Cout;
    if (!(sp->y = alloc(act_bytes(sp->Cout, sp->OH, sp->OW)))) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
    if (bott) {
        // layer1 shape: conv1, then conv2 + conv3 + shortcut (+ downsample folded into conv3's accumulator) in one launch
        const bool fuse = esz == 2 && stride == 1 && c1.Cout == 64 && c2.Cin == 64 && c2.Cout == 64 && c3->Cout == 256 &&
                          H == W && bneck_plan_ok(H, W, esz) && !getenv("RNB_BLOCK_NO_FUSE") &&
                          ((ds && ds->Cin == 64 && ds->stride == 1 && b->bias3ds) || (!ds && c1.Cin == 256));
        void* t1 = alloc(act_bytes(c1.Cout, H, W));
        if (!t1) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
        if ((rc = add(c1, sp->x, H, W, nullptr, true, t1))) return rc;
        if (fuse) {
            BneckDesc bd{};
            bd.B = B; bd.H = H; bd.W = W;
            bd.t1 = t1; bd.w2 = c2.w; bd.bias2 = c2.bias; bd.w3 = c3->w;
            bd.bias3 = ds ? b->bias3ds : c3->bias;
            bd.wds = ds ? ds->w : nullptr;
            bd.shortcut = sp->x;
            bd.y = sp->y;
            ConvPlan cp;
            if (bneck_plan_init(&cp, bd, nsm, err, sizeof(err))) {
                set_error(err);
                return RNB_ERR_UNSUPPORTED;
            }
            sp->launches.push_back(cp);
            return RNB_OK;
        }
        const void* shortcut = sp->x;
        if (ds) {
            void* s = alloc(act_bytes(sp->Cout, sp->OH, sp->OW));
            if (!s) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
            if ((rc = add(*ds, sp->x, H, W, nullptr, false, s))) return rc;
            shortcut = s;
        }
        void* t2 = alloc(act_bytes(c2.Cout, sp->OH, sp->OW));
        if (!t2) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
        if ((rc = add(c2, t1, H, W, nullptr, true, t2))) return rc;
        return add(*c3, t2, sp->OH, sp->OW, shortcut, true, sp->y);
    }
    const void* shortcut = sp->x;
    if (ds) {
        void* s = alloc(act_bytes(sp->Cout, sp->OH, sp->OW));
        if (!s) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
        if ((rc = add(*ds, sp->x, H, W, nullptr, false, s))) return rc;
        shortcut = s;
    }
    void* t1 = alloc(act_bytes(c1.Cout, sp->OH, sp->OW));
    if (!t1) return fail_cuda(cudaGetLastError(), "rnb_block: alloc");
    if ((rc = add(c1, sp->x, H, W, nullptr, true, t1))) return rc;
    return add(c2, t1, sp->OH, sp->OW, shortcut, true, sp->y);
}

extern "C" {

int rnb_block_create(int kind, int dtype, const rnb_conv_params_t* convs, int n_convs, rnb_block_t** out) {
    if (!convs || !out || (dtype != RNB_DTYPE_BF16 && dtype != RNB_DTYPE_TF32)) {
        set_error("rnb_block_create: bad argument");
        return RNB_ERR_INVALID;
    }
    const bool ok_count = (kind == RNB_BLOCK_CONV && n_convs == 1) || (kind == RNB_BLOCK_BASIC && (n_convs == 2 || n_convs == 3)) ||
                          (kind == RNB_BLOCK_BOTTLENECK && (n_convs == 3 || n_convs == 4));
    if (!ok_count) {
        set_error("rnb_block_create: a conv block takes 1 conv, a BasicBlock 2 (+ downsample), a Bottleneck 3 (+ downsample)");
        return RNB_ERR_INVALID;
    }
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || !device_ready(dev)) {
        set_error("rnb_init() has not been called (or failed) for the current CUDA device");
        return RNB_ERR_CUDA;
    }
    const int esz = dtype == RNB_DTYPE_BF16 ? 2 : 4;
    std::unique_ptr<rnb_block> b(new rnb_block());
    b->kind = kind; b->esz = esz; b->device = dev;
    for (int i = 0; i < n_convs; ++i) {
        const rnb_conv_params_t& p = convs[i];
        const int nbn = (p.bn_weight != nullptr) + (p.bn_bias != nullptr) + (p.bn_mean != nullptr) + (p.bn_var != nullptr);
        if (!p.w || (nbn != 0 && nbn != 4) || (p.k != 1 && p.k != 3) || p.Cin % (128 / esz) != 0 || p.Cout % 64 != 0 ||
            p.stride <= 0 || p.pad < 0) {
            set_error("rnb_block_create: unsupported conv (k in {1,3}, Cin % 64 == 0 (32 for tf32), Cout % 64 == 0, BN "
                      "vectors all set or all NULL)");
            return RNB_ERR_UNSUPPORTED;
        }
        PackedConv c;
        c.Cin = p.Cin; c.Cout = p.Cout; c.k = p.k; c.stride = p.stride; c.pad = p.pad;
        b->convs.push_back(c);
        PackedConv& pc = b->convs.back();
        BLK_CUDA(cudaMalloc(&pc.w, 1ull * p.Cout * p.k * p.k * p.Cin * esz));
        BLK_CUDA(cudaMalloc(reinterpret_cast<void**>(&pc.bias), p.Cout * sizeof(float)));
        BLK_CUDA(launch_fold_pack(p.w, p.bn_weight, p.bn_bias, p.bn_mean, p.bn_var, pc.w, pc.bias, p.Cout, p.Cin, p.k,
                                  esz, nullptr));
    }
    BLK_CUDA(cudaStreamSynchronize(nullptr));
    if (kind == RNB_BLOCK_BOTTLENECK && n_convs == 4) {
        const int C = b->convs[2].Cout;
        if (b->convs[3].Cout != C) {
            set_error("rnb_block_create: downsample and conv3 must have the same output channels");
            return RNB_ERR_INVALID;
        }
        std::vector<float> b3(C), bd(C);
        BLK_CUDA(cudaMemcpy(b3.data(), b->convs[2].bias, C * sizeof(float), cudaMemcpyDeviceToHost));
        BLK_CUDA(cudaMemcpy(bd.data(), b->convs[3].bias, C * sizeof(float), cudaMemcpyDeviceToHost));
        for (int c = 0; c < C; ++c) b3[c] += bd[c];
        BLK_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->bias3ds), C * sizeof(float)));
        BLK_CUDA(cudaMemcpy(b->bias3ds, b3.data(), C * sizeof(float), cudaMemcpyHostToDevice));
    }
    *out = b.release();
    return RNB_OK;
}

int rnb_block_destroy(rnb_block_t* b) {
    if (b) {
        DeviceGuard guard(b->device);
        cudaDeviceSynchronize();
        delete b;
    }
    return RNB_OK;
}

int rnb_block_num_launches(rnb_block_t* b, int B, int H, int W) {
    if (!b) return 0;
    for (const auto& kv : b->plans)
        if (std::get<0>(kv.first) == B && std::get<1>(kv.first) == H && std::get<2>(kv.first) == W)
            return static_cast<int>(kv.second.launches.size());
    return 0;
}

int rnb_block_forward(rnb_block_t* b, const float* x_dev, int B, int H, int W, const float* residual_dev, int relu,
                      float* out_dev, void* stream) {
    if (!b || !x_dev || !out_dev || B <= 0 || H <= 0 || W <= 0) {
        set_error("rnb_block_forward: bad argument");
        return RNB_ERR_INVALID;
    }
    if (b->kind != RNB_BLOCK_CONV && residual_dev) {
        set_error("rnb_block_forward: an explicit residual is only meaningful for a conv block");
        return RNB_ERR_INVALID;
    }
    DeviceGuard guard(b->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool do_relu = b->kind != RNB_BLOCK_CONV || relu != 0;
    const auto key = std::make_tuple(B, H, W, residual_dev ? 1 : 0, do_relu ? 1 : 0);
    auto it = b->plans.find(key);
    if (it == b->plans.end()) {
        if (b->plans.size() >= 8) {  // bounded: drop everything once a caller cycles through many shapes
            BLK_CUDA(cudaDeviceSynchronize());
            for (auto& kv : b->plans)
                for (void* p : kv.second.bufs) cudaFree(p);
            b->plans.clear();
        }
        ShapePlan sp;
        int rc = build_shape_plan(b, B, H, W, residual_dev != nullptr, do_relu, &sp);
        if (rc) {
            for (void* p : sp.bufs) cudaFree(p);
            return rc;
        }
        it = b->plans.emplace(key, std::move(sp)).first;
    }
    ShapePlan& sp = it->second;
    const PackedConv& c1 = b->convs[0];
    BLK_CUDA(launch_nchw_to_nhwc(x_dev, sp.x, B, c1.Cin, H * W, b->esz, s));
    if (residual_dev) BLK_CUDA(launch_nchw_to_nhwc(residual_dev, sp.res, B, sp.Cout, sp.OH * sp.OW, b->esz, s));
    for (const ConvPlan& cp : sp.launches) BLK_CUDA(conv_plan_launch(cp, s));
    BLK_CUDA(launch_nhwc_to_nchw(sp.y, out_dev, B, sp.Cout, sp.OH * sp.OW, b->esz, s));
    return RNB_OK;
}

}  // extern "C"
