// model.cu — see model.h
#include "model.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>

#include "../../include/rnb.h"
#include "host_pack.h"
#include "internal.h"

namespace rnb {

#define RNB_CUDA(expr)                                      \
    do {                                                    \
        cudaError_t e__ = (expr);                           \
        if (e__ != cudaSuccess) return fail_cuda(e__, #expr); \
    } while (0)

// ------------------------------------------------------------------------------------ arena
void* Arena::acquire(size_t bytes) {
    bytes = (bytes + 1023) & ~static_cast<size_t>(1023);
    int best = -1;
    for (int i = 0; i < static_cast<int>(blocks_.size()); ++i)
        if (!blocks_[i].busy && blocks_[i].bytes >= bytes &&
            (best < 0 || blocks_[i].bytes < blocks_[best].bytes))
            best = i;
    if (best >= 0) {
        blocks_[best].busy = true;
        return blocks_[best].p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    blocks_.push_back({p, bytes, true});
    total_ += bytes;
    return p;
}
void Arena::release(void* p) {
    if (keep) return;
    for (auto& b : blocks_)
        if (b.p == p) b.busy = false;
}
void Arena::release_all() {
    if (keep) return;
    for (auto& b : blocks_) b.busy = false;
}
void Arena::free_all() {
    for (auto& b : blocks_) cudaFree(b.p);
    blocks_.clear();
    total_ = 0;
}

// ------------------------------------------------------------------------------------ weights
namespace {

// One raw little-endian float32 file per state_dict key (save_weights.py:8-12), read the way
// Tensor::loadToCpu does (tensor.cuh:126-147): size comes from the file, shape from the caller.
int read_f32_file(const std::string& path, size_t expect, std::vector<float>& out) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) {
        set_error("cannot open weight file " + path);
        return RNB_ERR_IO;
    }
    const std::streamsize bytes = f.tellg();
    if (bytes != static_cast<std::streamsize>(expect * sizeof(float))) {
        set_error("weight file " + path + " has " + std::to_string(bytes) + " bytes, expected " +
                  std::to_string(expect * sizeof(float)));
        return RNB_ERR_IO;
    }
    f.seekg(0);
    out.resize(expect);
    f.read(reinterpret_cast<char*>(out.data()), bytes);
    if (f.fail()) {
        set_error("short read on " + path);
        return RNB_ERR_IO;
    }
    return RNB_OK;
}

int upload(const std::vector<float>& h, float** d) {
    RNB_CUDA(cudaMalloc(d, h.size() * sizeof(float)));
    RNB_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return RNB_OK;
}

// Raw (un-folded) tensors live only while one conv is being folded; freed on every exit path of the loaders.
struct BnDev {
    float *w = nullptr, *b = nullptr, *m = nullptr, *v = nullptr;
    void free_all() {
        cudaFree(w); cudaFree(b); cudaFree(m); cudaFree(v);
        w = b = m = v = nullptr;
    }
    ~BnDev() { free_all(); }
};
struct DevTmp {
    float* p = nullptr;
    ~DevTmp() { cudaFree(p); }
};

int load_bn(const std::string& dir, const std::string& name, int C, BnDev& bn) {
    std::vector<float> h;
    int r;
    if ((r = read_f32_file(dir + "/" + name + ".weight", C, h)) || (r = upload(h, &bn.w))) return r;
    if ((r = read_f32_file(dir + "/" + name + ".bias", C, h)) || (r = upload(h, &bn.b))) return r;
    if ((r = read_f32_file(dir + "/" + name + ".running_mean", C, h)) || (r = upload(h, &bn.m))) return r;
    if ((r = read_f32_file(dir + "/" + name + ".running_var", C, h)) || (r = upload(h, &bn.v))) return r;
    return RNB_OK;
}

// conv `cname`.weight + bn `bname`.* -> folded, packed device weights.
int load_conv(const std::string& dir, const std::string& cname, const std::string& bname, int Cin,
              int Cout, int k, int stride, int pad, int esz, ConvWeights& cw) {
    std::vector<float> h;
    int r;
    if ((r = read_f32_file(dir + "/" + cname + ".weight", 1ull * Cout * Cin * k * k, h))) return r;
    DevTmp raw;
    if ((r = upload(h, &raw.p))) return r;
    BnDev bn;
    if ((r = load_bn(dir, bname, Cout, bn))) return r;
    // cw.w / cw.bias belong to the model from here on (Model::~Model frees them, also after a failed load)
    cw.Cin = Cin; cw.Cout = Cout; cw.k = k; cw.stride = stride; cw.pad = pad;
    if (esz == 1) {  // FP8: channels padded to whole 128-byte K blocks / staging boxes, per-channel scales
        const int cin_p = (Cin + 127) / 128 * 128, cout_p = (Cout + 127) / 128 * 128;
        cw.Cin = cin_p; cw.Cout = cout_p;
        RNB_CUDA(cudaMalloc(&cw.w, 1ull * cout_p * cin_p * k * k));
        RNB_CUDA(cudaMalloc(&cw.bias, cout_p * sizeof(float)));
        RNB_CUDA(cudaMalloc(&cw.wscale, cout_p * sizeof(float)));
        RNB_CUDA(launch_fold_pack_fp8(raw.p, bn.w, bn.b, bn.m, bn.v, cw.w, cw.wscale, cw.bias, Cout, Cin, k, cout_p,
                                      cin_p, 0));
        RNB_CUDA(cudaDeviceSynchronize());
        return RNB_OK;
    }
    RNB_CUDA(cudaMalloc(&cw.w, 1ull * Cout * Cin * k * k * esz));
    RNB_CUDA(cudaMalloc(&cw.bias, Cout * sizeof(float)));
    RNB_CUDA(launch_fold_pack(raw.p, bn.w, bn.b, bn.m, bn.v, cw.w, cw.bias, Cout, Cin, k, esz, 0));
    RNB_CUDA(cudaDeviceSynchronize());
    return RNB_OK;
}

struct ArchSpec {
    const char* name;
    bool bottleneck;
    int blocks[4];
};
const ArchSpec kArchs[] = {
    {"resnet18", false, {2, 2, 2, 2}},  {"resnet34", false, {3, 4, 6, 3}},
    {"resnet50", true, {3, 4, 6, 3}},   {"resnet101", true, {3, 4, 23, 3}},
    {"resnet152", true, {3, 8, 36, 3}},  // main.cu:116-119
};

}  // namespace

Model::~Model() {
    DeviceGuard guard(device);
    for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
    if (order_ev) cudaEventDestroy(order_ev);
    if (host_compute) cudaStreamDestroy(host_compute);
    for (auto e : copy_events) cudaEventDestroy(e);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (side_stream) cudaStreamDestroy(side_stream);
    for (auto e : fork_ev) if (e) cudaEventDestroy(e);
    for (auto e : join_ev) if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (pipe_compute) cudaStreamDestroy(pipe_compute);
    if (pipe_d2h) cudaStreamDestroy(pipe_d2h);
    for (HostSlot& hs : slots) {
        cudaFree(hs.x_dev); cudaFree(hs.logits_dev); cudaFree(hs.top1_dev);
        if (hs.stage) cudaFreeHost(hs.stage);
        if (hs.copied) cudaEventDestroy(hs.copied);
        if (hs.done) cudaEventDestroy(hs.done);
        if (hs.computed) cudaEventDestroy(hs.computed);
    }
    arena.free_all();
    cudaFree(u8_scratch);
    cudaFree(fp8_amax_dev);
    if (!is_lane) cudaFree(fp8_ones);
    for (void* v : fp8_vec_allocs) cudaFree(v);
    cudaFree(host_x_dev); cudaFree(host_logits_dev); cudaFree(host_top1_dev); cudaFree(scratch_logits);
    if (host_stage) cudaFreeHost(host_stage);
    if (lane_fork) cudaEventDestroy(lane_fork);
    if (lane_join) cudaEventDestroy(lane_join);
    lane2.reset();
    if (is_lane) return;  // the weights belong to the model this lane was made from
    if (blob) {  // load_packed(): every weight pointer points into this one allocation
        cudaFree(blob);
        return;
    }
    cudaFree(stem_w); cudaFree(stem_bias); cudaFree(stem_wk); cudaFree(fc_w); cudaFree(fc_b);
    cudaFree(fc_wq); cudaFree(fc_bq);
    for (auto& b : blocks) {
        for (ConvWeights* c : {&b.conv1, &b.conv2, &b.conv3, &b.ds}) {
            cudaFree(c->w);
            cudaFree(c->bias);
            cudaFree(c->wscale);
        }
        cudaFree(b.bias3ds);
    }
}

int Model::configure(const std::string& arch_name, int dtype, int max_batch_, int chunk_) {
    const ArchSpec* spec = nullptr;
    for (const auto& a : kArchs)
        if (arch_name == a.name) spec = &a;
    if (!spec) {
        set_error("unknown arch '" + arch_name + "' (resnet18|34|50|101|152)");
        return RNB_ERR_INVALID;
    }
    if (dtype != RNB_DTYPE_BF16 && dtype != RNB_DTYPE_TF32 && dtype != RNB_DTYPE_FP8) {
        set_error("dtype must be RNB_DTYPE_BF16, RNB_DTYPE_TF32 or RNB_DTYPE_FP8");
        return RNB_ERR_INVALID;
    }
    if (max_batch_ <= 0) {
        set_error("max_batch must be positive");
        return RNB_ERR_INVALID;
    }
    arch = arch_name;
    esz = dtype == RNB_DTYPE_BF16 ? 2 : (dtype == RNB_DTYPE_FP8 ? 1 : 4);
    fp8 = dtype == RNB_DTYPE_FP8;
    bottleneck = spec->bottleneck;
    max_batch = max_batch_;
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    num_sms = rnb::num_sms();
    if (chunk_ <= 0) {
        const char* env = getenv("RNB_CHUNK");
        chunk_ = env ? atoi(env) : 0;
        if (chunk_ <= 0) chunk_ = max_batch_;
    }
    chunk = std::min(chunk_, max_batch_);
    const char* ng = getenv("RNB_NO_GRAPH");
    use_graph = !(ng && atoi(ng) != 0);
    const char* na = getenv("RNB_NO_ALTERNATE");
    alternate_tiles = !(na && atoi(na) != 0);
    const char* at = getenv("RNB_AUTOTUNE");
    autotune = !(at && atoi(at) == 0);
    const char* ka = getenv("RNB_KEEP_ACTIVATIONS");  // parity debugging: never recycle arena blocks
    arena.keep = ka && atoi(ka) != 0;
    const char* fu = getenv("RNB_FUSE");
    fuse_level = fu ? atoi(fu) : 2;
    const char* ss = getenv("RNB_SIDE_SMS");
    side_sms = ss ? atoi(ss) : 0;
    if (side_sms < 0 || side_sms > num_sms - 16) side_sms = 0;
    side_sms &= ~1;  // CTA pairs
    const char* fn = getenv("RNB_FUSE_NEXT");
    fuse_next = !(fn && atoi(fn) == 0);
    const char* nostc0 = getenv("RNB_NO_STEM_TC");
    stem_tc = image == 224 && !(nostc0 && atoi(nostc0) != 0);
    if (fp8) {
        if (!stem_tc) {
            set_error("the FP8 variant needs the tensor-core stem (224 x 224 inputs, RNB_NO_STEM_TC unset)");
            return RNB_ERR_UNSUPPORTED;
        }
        side_sms = 0;
        const char* ff = getenv("RNB_FP8_FROM");
        fp8_first_block = ff && atoi(ff) == 0 ? 0 : spec->blocks[0];   // default: layer1 in BF16
        const char* ho = getenv("RNB_FP8_HANDOVER");
        fp8_fused_handover = fp8_first_block > 0 && !(ho && atoi(ho) == 0);
    }
    return RNB_OK;
}

int Model::load(const std::string& arch_name, int dtype, const std::string& dir, int max_batch_,
                int chunk_) {
    int r = configure(arch_name, dtype, max_batch_, chunk_);
    if (r) return r;
    const ArchSpec* spec = nullptr;
    for (const auto& a : kArchs)
        if (arch_name == a.name) spec = &a;
    // ---- stem: conv1 + bn1 (main.cu:110-111), folded in fp32 OIHW for the CUDA-core stem
    {
        std::vector<float> h;
        if ((r = read_f32_file(dir + "/conv1.weight", 64 * 147, h))) return r;
        DevTmp raw;
        if ((r = upload(h, &raw.p))) return r;
        BnDev bn;
        if ((r = load_bn(dir, "bn1", 64, bn))) return r;
        RNB_CUDA(cudaMalloc(&stem_w, 64 * 147 * sizeof(float)));
        RNB_CUDA(cudaMalloc(&stem_bias, 64 * sizeof(float)));
        RNB_CUDA(launch_fold_f32(raw.p, bn.w, bn.b, bn.m, bn.v, stem_w, stem_bias, 64, 147, 0));
        if (stem_tc) {
            RNB_CUDA(cudaMalloc(&stem_wk, stem_any_weight_bytes(stem_esz())));
            RNB_CUDA(launch_stem_any_pack_weights(stem_esz(), raw.p, bn.w, bn.b, bn.m, bn.v, stem_wk, stem_bias, 0));
        }
        RNB_CUDA(cudaDeviceSynchronize());
    }
    double macs = 64.0 * 147 * 112 * 112;
    // ---- residual layers (createLayer, main.cu:53-89; BasicBlock per torchvision)
    int in_c = 64;
    int hw = 56;
    num_convs = 1;
    for (int L = 0; L < 4; ++L) {
        const int mid = 64 << L;
        const int out_c = bottleneck ? mid * 4 : mid;
        const int layer_stride = L == 0 ? 1 : 2;
        for (int i = 0; i < spec->blocks[L]; ++i) {
            blocks.emplace_back();  // owned by the model at once: a failed load frees what was already folded
            BlockWeights& bw = blocks.back();
            bw.bottleneck = bottleneck;
            bw.name = "layer" + std::to_string(L + 1) + "." + std::to_string(i);
            const std::string p = bw.name + ".";
            const int stride = i == 0 ? layer_stride : 1;
            const int out_hw = hw / stride;
            // mixed FP8 plan: the blocks before fp8_first_block keep BF16 weights
            const int bidx = static_cast<int>(blocks.size()) - 1;
            const int esz = fp8 && bidx < fp8_first_block ? 2 : this->esz;
            // fused hand-over: conv1 and the downsample of the first FP8 block keep BF16 weights (E4M3 output)
            const bool will_ds = i == 0 && (stride != 1 || in_c != out_c);
            const int esz_in = fp8 && fp8_fused_handover && bidx == fp8_first_block && will_ds ? 2 : esz;
            if (bottleneck) {
                if ((r = load_conv(dir, p + "conv1", p + "bn1", in_c, mid, 1, 1, 0, esz_in, bw.conv1))) return r;
                if ((r = load_conv(dir, p + "conv2", p + "bn2", mid, mid, 3, stride, 1, esz, bw.conv2))) return r;
                if ((r = load_conv(dir, p + "conv3", p + "bn3", mid, out_c, 1, 1, 0, esz, bw.conv3))) return r;
                macs += 1.0 * hw * hw * in_c * mid + 1.0 * out_hw * out_hw * mid * mid * 9 +
                        1.0 * out_hw * out_hw * mid * out_c;
                num_convs += 3;
            } else {
                if ((r = load_conv(dir, p + "conv1", p + "bn1", in_c, mid, 3, stride, 1, esz_in, bw.conv1))) return r;
                if ((r = load_conv(dir, p + "conv2", p + "bn2", mid, mid, 3, 1, 1, esz, bw.conv2))) return r;
                macs += 1.0 * out_hw * out_hw * in_c * mid * 9 + 1.0 * out_hw * out_hw * mid * mid * 9;
                num_convs += 2;
            }
            if (i == 0 && (stride != 1 || in_c != out_c)) {  // main.cu:71
                bw.has_ds = true;
                if ((r = load_conv(dir, p + "downsample.0", p + "downsample.1", in_c, out_c, 1, stride, 0,
                                   esz_in, bw.ds)))
                    return r;
                macs += 1.0 * out_hw * out_hw * in_c * out_c;
                num_convs += 1;
                if (bottleneck && esz != 1 && esz_in == esz) {
                    std::vector<float> b3(out_c), bd(out_c);
                    RNB_CUDA(cudaMemcpy(b3.data(), bw.conv3.bias, out_c * sizeof(float), cudaMemcpyDeviceToHost));
                    RNB_CUDA(cudaMemcpy(bd.data(), bw.ds.bias, out_c * sizeof(float), cudaMemcpyDeviceToHost));
                    for (int c = 0; c < out_c; ++c) b3[c] += bd[c];
                    if ((r = upload(b3, &bw.bias3ds))) return r;
                }
            }
            in_c = out_c;
            hw = out_hw;
        }
    }
    final_c = in_c;
    // ---- fc (main.cu:122)
    {
        std::vector<float> h;
        // class count comes from the bias file size
        std::ifstream f(dir + "/fc.bias", std::ios::binary | std::ios::ate);
        if (!f.is_open()) {
            set_error("cannot open weight file " + dir + "/fc.bias");
            return RNB_ERR_IO;
        }
        classes = static_cast<int>(f.tellg() / sizeof(float));
        if (classes <= 0) {
            set_error("fc.bias is empty");
            return RNB_ERR_IO;
        }
        if ((r = read_f32_file(dir + "/fc.bias", classes, h)) || (r = upload(h, &fc_b))) return r;
        if ((r = read_f32_file(dir + "/fc.weight", 1ull * classes * final_c, h)) || (r = upload(h, &fc_w)))
            return r;
        macs += 1.0 * classes * final_c;
        // BF16 path: FC on tensor cores (BF16 operands, FP32 accumulate and FP32 logits)
        const char* ft = getenv("RNB_FC_TC");
        fc_tc = esz <= 2 && final_c % 64 == 0 && classes % 4 == 0 && (fp8 || !(ft && atoi(ft) == 0));
        if (fc_tc) {
            classes_pad = (classes + 63) / 64 * 64;
            RNB_CUDA(cudaMalloc(&fc_wq, 1ull * classes_pad * final_c * 2));
            RNB_CUDA(cudaMalloc(&fc_bq, classes_pad * sizeof(float)));
            RNB_CUDA(launch_fc_pack(fc_w, fc_b, fc_wq, fc_bq, classes, final_c, classes_pad, 0));
            RNB_CUDA(cudaDeviceSynchronize());
        }
    }
    flops_per_image = 2.0 * macs;
    if (fp8) {
        std::vector<float> ones(4096, 1.f);
        if ((r = upload(ones, &fp8_ones))) return r;
    }
    int rs = create_streams();
    if (rs) return rs;
    set_error("");
    return RNB_OK;
}

int Model::create_streams() {
    RNB_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
    RNB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    RNB_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    RNB_CUDA(cudaStreamCreateWithFlags(&host_compute, cudaStreamNonBlocking));
    RNB_CUDA(cudaEventCreateWithFlags(&order_ev, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) {
        RNB_CUDA(cudaEventCreateWithFlags(&fork_ev[i], cudaEventDisableTiming));
        RNB_CUDA(cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming));
    }
    return RNB_OK;
}

// Serialise every use of the shared activation arena across streams (see model.h).
int Model::begin_enqueue(cudaStream_t s) {
    if (has_last && last_stream != s) RNB_CUDA(cudaStreamWaitEvent(s, order_ev, 0));
    return RNB_OK;
}
int Model::end_enqueue(cudaStream_t s) {
    RNB_CUDA(cudaEventRecord(order_ev, s));
    last_stream = s;
    has_last = true;
    return RNB_OK;
}

// ------------------------------------------------------------------------------------ packed weight blob
// blocks[] with shapes only, exactly as load() lays them out (createLayer, main.cu:53-89)
void Model::build_structure() {
    const ArchSpec* spec = nullptr;
    for (const auto& a : kArchs)
        if (arch == a.name) spec = &a;
    blocks.clear();
    double macs = 64.0 * 147 * 112 * 112;
    int in_c = 64, hw = 56;
    num_convs = 1;
    auto shape = [](ConvWeights& c, int Cin, int Cout, int k, int stride, int pad) {
        c.Cin = Cin; c.Cout = Cout; c.k = k; c.stride = stride; c.pad = pad;
    };
    for (int L = 0; L < 4; ++L) {
        const int mid = 64 << L;
        const int out_c = bottleneck ? mid * 4 : mid;
        for (int i = 0; i < spec->blocks[L]; ++i) {
            BlockWeights bw;
            bw.bottleneck = bottleneck;
            bw.name = "layer" + std::to_string(L + 1) + "." + std::to_string(i);
            const int stride = i == 0 ? (L == 0 ? 1 : 2) : 1;
            const int out_hw = hw / stride;
            if (bottleneck) {
                shape(bw.conv1, in_c, mid, 1, 1, 0);
                shape(bw.conv2, mid, mid, 3, stride, 1);
                shape(bw.conv3, mid, out_c, 1, 1, 0);
                macs += 1.0 * hw * hw * in_c * mid + 1.0 * out_hw * out_hw * mid * mid * 9 + 1.0 * out_hw * out_hw * mid * out_c;
                num_convs += 3;
            } else {
                shape(bw.conv1, in_c, mid, 3, stride, 1);
                shape(bw.conv2, mid, mid, 3, 1, 1);
                macs += 1.0 * out_hw * out_hw * in_c * mid * 9 + 1.0 * out_hw * out_hw * mid * mid * 9;
                num_convs += 2;
            }
            if (i == 0 && (stride != 1 || in_c != out_c)) {
                bw.has_ds = true;
                shape(bw.ds, in_c, out_c, 1, stride, 0);
                macs += 1.0 * out_hw * out_hw * in_c * out_c;
                num_convs += 1;
            }
            blocks.push_back(bw);
            in_c = out_c;
            hw = out_hw;
        }
    }
    final_c = in_c;
    macs += 1.0 * classes * final_c;
    flops_per_image = 2.0 * macs;
}

template <class F>
void Model::for_each_weight(F&& f) {
    f(reinterpret_cast<void**>(&stem_w), 64ull * 147 * sizeof(float));
    f(reinterpret_cast<void**>(&stem_bias), 64ull * sizeof(float));
    if (stem_tc) f(&stem_wk, stem_any_weight_bytes(stem_esz()));
    for (auto& b : blocks) {
        for (ConvWeights* c : {&b.conv1, &b.conv2, &b.conv3, &b.ds}) {
            if (c->Cout == 0) continue;  // conv3 of a BasicBlock, ds of a block without downsample
            f(&c->w, 1ull * c->Cout * c->Cin * c->k * c->k * esz);
            f(reinterpret_cast<void**>(&c->bias), 1ull * c->Cout * sizeof(float));
        }
        if (b.has_ds && b.bottleneck) f(reinterpret_cast<void**>(&b.bias3ds), 1ull * b.conv3.Cout * sizeof(float));
    }
    f(reinterpret_cast<void**>(&fc_w), 1ull * classes * final_c * sizeof(float));
    f(reinterpret_cast<void**>(&fc_b), 1ull * classes * sizeof(float));
    if (fc_tc) {
        f(&fc_wq, 1ull * classes_pad * final_c * 2);
        f(reinterpret_cast<void**>(&fc_bq), 1ull * classes_pad * sizeof(float));
    }
}

namespace {
struct PackedHeader {
    char magic[8];        // "RNBWGT01"
    uint32_t version;     // 2 (1: the first stem weight layout [kh][j][oc][e], window starting at tap 0)
    uint32_t esz;         // 2 = BF16 operands, 4 = TF32
    char arch[16];
    uint32_t classes, classes_pad;
    uint32_t stem_tc, fc_tc;
    uint32_t num_tensors;
    uint32_t reserved;
    uint64_t payload_bytes;
    uint64_t checksum;    // FNV-1a over the payload taken as 64-bit words, four interleaved lanes
};
// The payload is a multiple of 256 bytes. Word-wise FNV-1a in four independent lanes (word i goes to lane
// i & 3), lanes folded at the end: ~10 GB/s on one core instead of ~1 GB/s byte by byte — the checksum must not
// cost more than the read it protects.
uint64_t fnv1a(const uint8_t* p, size_t n) {
    const uint64_t prime = 1099511628211ull;
    uint64_t h[4] = {1469598103934665603ull, 1469598103934665603ull ^ 1, 1469598103934665603ull ^ 2,
                     1469598103934665603ull ^ 3};
    const size_t words = n / 8;
    size_t i = 0;
    for (; i + 4 <= words; i += 4) {
        uint64_t w[4];
        memcpy(w, p + 8 * i, 32);
        h[0] = (h[0] ^ w[0]) * prime;
        h[1] = (h[1] ^ w[1]) * prime;
        h[2] = (h[2] ^ w[2]) * prime;
        h[3] = (h[3] ^ w[3]) * prime;
    }
    uint64_t r = 1469598103934665603ull;
    for (int k = 0; k < 4; ++k) r = (r ^ h[k]) * prime;
    for (size_t b = 8 * i; b < n; ++b) r = (r ^ p[b]) * prime;
    return r;
}
constexpr size_t kAlign = 256;
size_t aligned(size_t b) { return (b + kAlign - 1) / kAlign * kAlign; }
}  // namespace

int Model::save_packed(const std::string& path) {
    if (fp8) {
        set_error("save_packed: the FP8 variant has no packed format (its scales depend on the calibration batch)");
        return RNB_ERR_UNSUPPORTED;
    }
    size_t total = 0;
    uint32_t count = 0;
    for_each_weight([&](void**, size_t bytes) { total += aligned(bytes); ++count; });
    std::vector<uint8_t> host(total, 0);
    size_t off = 0;
    cudaError_t ce = cudaSuccess;
    for_each_weight([&](void** p, size_t bytes) {
        if (ce == cudaSuccess) ce = cudaMemcpy(host.data() + off, *p, bytes, cudaMemcpyDeviceToHost);
        off += aligned(bytes);
    });
    if (ce != cudaSuccess) return fail_cuda(ce, "save_packed: cudaMemcpy");
    PackedHeader h{};
    memcpy(h.magic, "RNBWGT01", 8);
    h.version = 2;
    h.esz = static_cast<uint32_t>(esz);
    snprintf(h.arch, sizeof(h.arch), "%s", arch.c_str());
    h.classes = classes; h.classes_pad = classes_pad;
    h.stem_tc = stem_tc ? 1 : 0; h.fc_tc = fc_tc ? 1 : 0;
    h.num_tensors = count;
    h.payload_bytes = total;
    h.checksum = fnv1a(host.data(), total);
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f.is_open()) {
        set_error("save_packed: cannot open " + path);
        return RNB_ERR_IO;
    }
    f.write(reinterpret_cast<const char*>(&h), sizeof(h));
    f.write(reinterpret_cast<const char*>(host.data()), static_cast<std::streamsize>(total));
    if (f.fail()) {
        set_error("save_packed: short write on " + path);
        return RNB_ERR_IO;
    }
    return RNB_OK;
}

int Model::load_packed(const std::string& path, int max_batch_, int chunk_) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) {
        set_error("cannot open packed weight file " + path);
        return RNB_ERR_IO;
    }
    const std::streamsize fsize = f.tellg();
    PackedHeader h{};
    if (fsize < static_cast<std::streamsize>(sizeof(h))) {
        set_error("packed weight file " + path + " is truncated");
        return RNB_ERR_IO;
    }
    f.seekg(0);
    f.read(reinterpret_cast<char*>(&h), sizeof(h));
    if (memcmp(h.magic, "RNBWGT01", 8) != 0 || h.version != 2 || (h.esz != 2 && h.esz != 4)) {
        set_error("packed weight file " + path + ": bad magic / version");
        return RNB_ERR_IO;
    }
    if (fsize != static_cast<std::streamsize>(sizeof(h) + h.payload_bytes)) {
        set_error("packed weight file " + path + ": size does not match its header");
        return RNB_ERR_IO;
    }
    h.arch[sizeof(h.arch) - 1] = 0;
    int r = configure(h.arch, h.esz == 2 ? RNB_DTYPE_BF16 : RNB_DTYPE_TF32, max_batch_, chunk_);
    if (r) return r;
    if (static_cast<uint32_t>(stem_tc ? 1 : 0) != h.stem_tc) {
        set_error("packed weight file was written with a different RNB_NO_STEM_TC setting");
        return RNB_ERR_INVALID;
    }
    classes = static_cast<int>(h.classes);
    classes_pad = static_cast<int>(h.classes_pad);
    fc_tc = h.fc_tc != 0;
    build_structure();
    size_t total = 0;
    uint32_t count = 0;
    for_each_weight([&](void**, size_t bytes) { total += aligned(bytes); ++count; });
    if (total != h.payload_bytes || count != h.num_tensors) {
        set_error("packed weight file " + path + ": tensor table does not match architecture " + arch);
        return RNB_ERR_IO;
    }
    std::unique_ptr<uint8_t[]> host(new uint8_t[total]);  // (uninitialised on purpose)
    f.read(reinterpret_cast<char*>(host.get()), static_cast<std::streamsize>(total));
    if (f.fail()) {
        set_error("short read on " + path);
        return RNB_ERR_IO;
    }
    RNB_CUDA(cudaMalloc(&blob, total));
    RNB_CUDA(cudaMemcpyAsync(blob, host.get(), total, cudaMemcpyHostToDevice, 0));  // overlaps the checksum
    const bool sum_ok = fnv1a(host.get(), total) == h.checksum;
    RNB_CUDA(cudaStreamSynchronize(0));
    if (!sum_ok) {
        set_error("packed weight file " + path + ": checksum mismatch (corrupt file)");
        return RNB_ERR_IO;
    }
    size_t off = 0;
    for_each_weight([&](void** p, size_t bytes) {
        *p = static_cast<uint8_t*>(blob) + off;
        off += aligned(bytes);
    });
    int rs = create_streams();
    if (rs) return rs;
    set_error("");
    return RNB_OK;
}

// ------------------------------------------------------------------------------------ planning
ChunkPlan* Model::plan_for(int n) {
    auto it = plans.find(n);
    if (it != plans.end()) return &it->second;
    // A new plan is rare (first use of a batch size). Tuning launches kernels on the arena, so nothing
    // else may be in flight on it.
    if (autotune) cudaDeviceSynchronize();
    ChunkPlan p;
    p.n = n;
    // element size of the tensors being planned: esz, except in the BF16 head of a mixed FP8 plan
    int cur_esz = fp8 && fp8_first_block > 0 ? 2 : esz;
    auto bytes = [&](int c, int h) { return 1ull * n * h * h * (cur_esz == 1 ? cpad(c) : c) * cur_esz; };
    // A plan that fails half way must not leave its blocks busy: between plans nothing is (every plan releases all
    // of its blocks at the end, chunks run back to back), so a failure simply frees the lot.
    auto fail_alloc = [&]() -> ChunkPlan* {
        arena.release_all();
        set_error("activation arena allocation failed");
        return nullptr;
    };
    struct PlanGuard {
        Arena& a;
        bool ok = false;
        ~PlanGuard() { if (!ok) a.release_all(); }
    } plan_guard{arena};
    const int s_hw = (6 + image - 7) / 2 + 1;     // 112
    const int p_hw = (2 + s_hw - 3) / 2 + 1;      // 56
    if (!(p.stem_out = arena.acquire(stem_tc ? stem_any_input_bytes(stem_esz(), n) : bytes(64, s_hw))))
        return fail_alloc();
    if (fp8 && fp8_first_block == 0 && !(p.pool_raw = arena.acquire(1ull * n * p_hw * p_hw * 64 * 2))) return fail_alloc();
    if (!(p.pool_out = arena.acquire(bytes(64, p_hw)))) return fail_alloc();
    if (!stem_tc) p.named["stem"] = {p.stem_out, 64, s_hw, s_hw};
    p.named["maxpool"] = {p.pool_out, 64, p_hw, p_hw};
    arena.release(p.stem_out);  // dead once the pool has run
    if (p.pool_raw) arena.release(p.pool_raw);  // dead once its E4M3 copy exists (launches are stream-ordered)

    void* x = p.pool_out;
    int hw = p_hw;
    char err[256];
    // FP8: per-tensor scale of every live activation buffer (provisional 1.0 until calibrated)
    std::map<const void*, float> scale_of;
    scale_of[p.pool_out] = fp8_calibrated ? fp8_stem_scale : 1.f;
    const void* handover_src = nullptr;  // BF16 input of the first FP8 block while that block is being planned
    int sm_budget = num_sms;  // SMs the next planned conv may use (reduced while a downsample conv runs beside it)
    // rows > 0: a 1x1 / stride-1 conv over `rows` pixel rows of its [M][C] operands (pointers already offset) instead of
    // the whole n x in_hw x in_hw tensor — the un-fused remainder of a partially fused launch
    auto add_conv = [&](const ConvWeights& cw, const void* in, int in_hw, const void* res, bool relu,
                        void* out, int rows = 0) -> int {
        ConvDesc d{};
        d.B = n; d.H = in_hw; d.W = in_hw; d.Cin = cw.Cin; d.Cout = cw.Cout;
        if (rows > 0) {
            d.B = 1; d.H = 1; d.W = rows;
            in_hw = -rows;  // autotune key of the row-range form
        }
        d.ksize = cw.k; d.stride = cw.stride; d.pad = cw.pad; d.relu = relu;
        // hand-over conv of a mixed FP8 plan: BF16 input tensor + BF16 weights, E4M3 output
        const bool ho = handover_src != nullptr && in == handover_src;
        d.act = ho ? ActType::BF16 : (cur_esz == 2 ? ActType::BF16 : (cur_esz == 1 ? ActType::FP8 : ActType::TF32));
        d.out_fp8 = ho;
        const bool fp8 = cur_esz == 1;  // this launch (shadows the model-wide flag inside add_conv)
        // Boustrophedon over the launch sequence: every conv walks its tiles in the opposite
        // direction of the previous one, so it begins where its producer just finished (L2-hot).
        d.reverse = alternate_tiles && (p.convs.size() & 1) != 0;
        d.in = in; d.weight = cw.w; d.bias = cw.bias; d.residual = res; d.out = out;
        if (!fp8 && this->fp8) p.links.push_back({in, res, out});  // BF16 launch of a mixed plan: no scales
        if (fp8) {
            const size_t idx = p.convs.size();
            d.chan_scale = ho ? fp8_ones : cw.wscale;
            d.in_scale = ho ? 1.f : scale_of[in];
            d.res_scale = res ? scale_of[res] : 1.f;
            d.out_scale = fp8_calibrated && idx < fp8_out_scale.size() ? fp8_out_scale[idx] : 1.f;
            scale_of[out] = d.out_scale;
            p.links.push_back({in, res, out});
            float* vecs = nullptr;
            if (cudaMalloc(&vecs, 2ull * cw.Cout * sizeof(float)) != cudaSuccess) {
                set_error("FP8 epilogue vector allocation failed");
                return RNB_ERR_CUDA;
            }
            fp8_vec_allocs.push_back(vecs);
            d.fp8_vecs = vecs;
        }
        ConvPlan cp;
        int rc = conv_plan_init(&cp, d, sm_budget, 0, err, sizeof(err));
        if (rc) {
            set_error(err);
            return rc;
        }
        if (autotune) {
            // Measure the tile families this layer admits on the buffers it will really use and keep
            // the fastest (results are cached per layer shape). The heuristic choice above is the
            // fallback and the first candidate.
            const std::tuple<int, int, int, int, int, int, int, int> key{n, in_hw, cw.Cin, cw.Cout, cw.k,
                                                                         cw.stride, res ? 1 : 0, sm_budget};
            auto hit = tuned.find(key);
            int best_force = hit != tuned.end() ? hit->second : -1;
            if (best_force < 0) {
                // (20128, the 16-epilogue-warp variant, was a candidate for one session: never selected on any layer of any config)
                const int cands[12] = {64, 128, 1128, 1256, 3064, 4064, 10128, 11128, 11256, 12128, 12256, 31128};
                float best_ms = 1e30f;
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                for (int force : cands) {
                    if (fp8 && force != 128 && force != 1128 && force != 1256 && !((force == 12128 || force == 12256) && !ho))
                        continue;
                    if (force == 20128 && (cur_esz == 4 || cw.Cout % 128 != 0)) continue;
                    if (force == 31128 && (cur_esz != 2 || ho || cw.Cout != 128 || cw.k != 3 || cw.Cin != 128)) continue;
                    if (force == 3064 && !conv_plan_halo_ok(d)) continue;
                    if (force == 4064 && !conv_plan_halo2_ok(d)) continue;
                    // deep-pipeline variants trade a staging buffer for pipeline stages: only for layers
                    // whose epilogue is light (no residual prefetch) and whose K loop is long; timing a
                    // residual layer alone flatters them (measured in the full network: slower)
                    if (force >= 10000 && ((cur_esz != 2 && !(cur_esz == 1 && (force == 12128 || force == 12256))) || res ||
                                           cw.k * cw.k * cw.Cin < 256))
                        continue;
                    if (cw.Cout % (force % 1000) != 0) continue;
                    if (force == 64 && cw.Cout % 128 == 0 && 1LL * n * in_hw * in_hw > 4096) continue;
                    ConvPlan trial;
                    if (conv_plan_init(&trial, d, sm_budget, force, err, sizeof(err))) continue;
                    if (fp8 && fp8_premultiply(&trial, d.in_scale, d.res_scale, d.out_scale, cap_stream) != cudaSuccess) continue;
                    bool ok = true;
                    for (int i = 0; i < 2 && ok; ++i) ok = conv_plan_launch(trial, cap_stream) == cudaSuccess;
                    cudaEventRecord(e0, cap_stream);
                    for (int i = 0; i < 5 && ok; ++i) ok = conv_plan_launch(trial, cap_stream) == cudaSuccess;
                    cudaEventRecord(e1, cap_stream);
                    if (cudaStreamSynchronize(cap_stream) != cudaSuccess || !ok) continue;
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (getenv("RNB_VERBOSE") && atoi(getenv("RNB_VERBOSE")) >= 2)
                        fprintf(stderr, "rnb autotune: n=%d %dx%d %d->%d k%d s%d res=%d : tile code %d %.1f us\n", n, in_hw,
                                in_hw, cw.Cin, cw.Cout, cw.k, cw.stride, res ? 1 : 0, force, ms * 200.f);
                    if (ms < best_ms) {
                        best_ms = ms;
                        best_force = force;
                    }
                }
                cudaEventDestroy(e0);
                cudaEventDestroy(e1);
                if (best_force < 0) best_force = 0;
                tuned[key] = best_force;
            }
            if (best_force > 0) {
                rc = conv_plan_init(&cp, d, sm_budget, best_force, err, sizeof(err));
                if (rc) {
                    set_error(err);
                    return rc;
                }
            }
        }
        if (getenv("RNB_VERBOSE"))
            fprintf(stderr, "rnb plan: conv#%zu n=%d %dx%d %d->%d k%d s%d res=%d : %s tile %dx%d grid %d\n",
                    p.convs.size(), n, in_hw, in_hw, cw.Cin, cw.Cout, cw.k, cw.stride, res ? 1 : 0,
                    cp.halo2 ? "halo-pair" : cp.halo ? "halo" : (cp.ctas == 2 ?  (cp.deep == 2 ? "pair-deepest" : cp.deep ? "pair-deep" : (cp.resb ? "pair-resident-B" : (cp.g.split_from < cp.g.m_tiles * cp.g.n_tiles ? "pair-split" : "pair"))) : (cp.deep ? "single-deep" : (cp.w16 ? "single-16w" : "single"))),
                    cp.ctas == 2 ? 256 : 128, cp.bn, cp.grid);
        if (fp8 && fp8_premultiply(&cp, d.in_scale, d.res_scale, d.out_scale, cap_stream) != cudaSuccess) {
            set_error("FP8 epilogue vector setup failed");
            return RNB_ERR_CUDA;
        }
        p.convs.push_back(cp);
        return 0;
    };
    const bool c3n1_auto = !(getenv("RNB_C3N1_AUTO") && atoi(getenv("RNB_C3N1_AUTO")) == 0);
    void* pre_t1 = nullptr;  // this block's conv1 output, already produced by the previous fused launch
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
        const BlockWeights& bw = blocks[bi];
        handover_src = nullptr;
        if (fp8 && static_cast<int>(bi) == fp8_first_block && fp8_first_block > 0 && fp8_fused_handover && bw.has_ds &&
            bw.conv1.wscale == nullptr) {
            // fused hand-over: this block's conv1 and downsample read the BF16 x themselves and write E4M3
            handover_src = x;
        } else if (fp8 && static_cast<int>(bi) == fp8_first_block && fp8_first_block > 0) {
            // hand-over of a mixed plan: the BF16 activation x becomes an E4M3 tensor (scale fixed by calibration)
            const int c = bw.bottleneck ? bw.conv1.Cin : bw.conv1.Cin;   // already padded for the FP8 block
            const int c_real = blocks[bi - 1].bottleneck ? blocks[bi - 1].conv3.Cout : blocks[bi - 1].conv2.Cout;
            cur_esz = 1;
            void* xq = arena.acquire(bytes(c, hw));
            if (!xq) return fail_alloc();
            ConvPlan q;
            memset(&q, 0, sizeof(q));
            q.quant = 1;
            q.q_src = x; q.q_dst = xq; q.q_rows = 1LL * n * hw * hw; q.q_c = c_real; q.q_cpad = cpad(c_real);
            const size_t idx = p.convs.size();
            const float sc = fp8_calibrated && idx < fp8_out_scale.size() ? fp8_out_scale[idx] : 1.f;
            q.q_inv_scale = 1.f / sc;
            q.esz = 1;
            q.bytes = 1.0 * q.q_rows * (2.0 * q.q_c + q.q_cpad);
            scale_of[xq] = sc;
            p.links.push_back({x, nullptr, xq});
            p.convs.push_back(q);
            arena.release(x);
            x = xq;
        }
        const int besz = fp8 && static_cast<int>(bi) < fp8_first_block ? 2 : esz;
        cur_esz = besz;
        const int stride = bw.bottleneck ? bw.conv2.stride : bw.conv1.stride;
        const int out_hw = hw / stride;
        const int out_c = bw.bottleneck ? bw.conv3.Cout : bw.conv2.Cout;
        // layer1-shaped Bottleneck (64 -> 64 -> 256, stride 1): conv2 + conv3 + shortcut in one launch
        const bool fuse = fuse_level >= 1 && bw.bottleneck && stride == 1 && bw.conv2.Cin == 64 &&
                          bw.conv2.Cout == 64 && out_c == 256 && bneck_plan_ok(hw, hw, besz);
        const bool fuse_ds = fuse && fuse_level >= 2 && bw.has_ds && bw.ds.stride == 1 && bw.ds.Cin == 64 &&
                             bw.bias3ds;
        void* shortcut = x;
        void* ds = nullptr;
        // Bottleneck block 0 of layers 2-4: the downsample conv is independent of conv1 -> conv2 and (stride 2,
        // write-heavy) HBM-bound while conv2 is tensor-bound: run it beside them on its own slice of the SMs
        const bool overlap_ds = side_sms > 0 && bw.has_ds && !fuse_ds && bw.bottleneck && !fuse;
        if (bw.has_ds && !fuse_ds) {
            if (!(ds = arena.acquire(bytes(out_c, out_hw)))) return fail_alloc();
            sm_budget = overlap_ds ? side_sms : num_sms;
            if (add_conv(bw.ds, x, hw, nullptr, false, ds)) return nullptr;
            if (overlap_ds) p.convs.back().side = 1;
            shortcut = ds;
        }
        sm_budget = overlap_ds ? num_sms - side_sms : num_sms;
        void* y = nullptr;
        if (fuse) {
            void* t1 = pre_t1;
            pre_t1 = nullptr;
            if (!t1) {
                if (!(t1 = arena.acquire(bytes(64, hw)))) return fail_alloc();
                if (add_conv(bw.conv1, x, hw, nullptr, true, t1)) return nullptr;
            }
            const BlockWeights* nb = bi + 1 < blocks.size() ? &blocks[bi + 1] : nullptr;
            bool next = fuse_next && nb && nb->bottleneck && !nb->has_ds && nb->conv1.Cin == 256 &&
                        nb->conv1.Cout == 64 && nb->conv2.stride == 1 && nb->conv2.Cin == 64 &&
                        nb->conv2.Cout == 64 && nb->conv3.Cout == 256;
            int n1 = 64;
            // last block of the layer: the NEXT LAYER's conv1 (256 -> 128, same resolution; the stride sits on
            // its conv2) can ride along instead (BneckCfg<false, 128>). Bit-identical, but measured no faster
            // (the fused launch grows 165 -> 252 us, the separate conv1 it replaces took 101 us: a seventh
            // epilogue step per tile and one staging box fewer), so it is opt-in: RNB_L1L2=1.
            const char* l1l2 = getenv("RNB_L1L2");
            if (!next && l1l2 && atoi(l1l2) != 0 && fuse_next && !fuse_ds && nb && nb->bottleneck && nb->has_ds &&
                nb->conv1.Cin == 256 && nb->conv1.Cout == 128 && nb->conv1.k == 1 && nb->conv1.stride == 1) {
                next = true;
                n1 = 128;
            }
            void* t1n = nullptr;
            if (next && !(t1n = arena.acquire(bytes(n1, hw)))) return fail_alloc();
            if (!(y = arena.acquire(bytes(out_c, hw)))) return fail_alloc();
            BneckDesc bd{};
            bd.B = n; bd.H = hw; bd.W = hw;
            bd.reverse = alternate_tiles && (p.convs.size() & 1) != 0;
            bd.t1 = t1; bd.w2 = bw.conv2.w; bd.bias2 = bw.conv2.bias; bd.w3 = bw.conv3.w;
            bd.bias3 = fuse_ds ? bw.bias3ds : bw.conv3.bias;
            bd.wds = fuse_ds ? bw.ds.w : nullptr;
            bd.shortcut = shortcut;
            bd.y = y;
            bd.w1n = next ? nb->conv1.w : nullptr;
            bd.bias1n = next ? nb->conv1.bias : nullptr;
            bd.n1 = n1;
            bd.t1n = t1n;
            ConvPlan cp;
            if (bneck_plan_init(&cp, bd, num_sms, err, sizeof(err))) {
                set_error(err);
                return nullptr;
            }
            if (getenv("RNB_VERBOSE"))
                fprintf(stderr, "rnb plan: conv#%zu n=%d %dx%d fused bottleneck tail%s%s grid %d\n", p.convs.size(), n,
                        hw, hw, fuse_ds ? " +downsample" : "", next ? " +next conv1" : "", cp.grid);
            if (fp8) p.links.push_back({nullptr, nullptr, nullptr});  // fused BF16 launch of a mixed plan
            p.convs.push_back(cp);
            arena.release(t1);
            pre_t1 = t1n;
        } else if (bw.bottleneck) {
            // conv1 -> bn1 -> relu -> conv2(stride) -> bn2 -> relu -> conv3 -> bn3 -> +shortcut -> relu
            // (layerForward, main.cu:138-163)
            void* t1 = pre_t1;  // already produced by the previous block's fused conv3 + conv1' launch
            pre_t1 = nullptr;
            if (!t1) {
                if (!(t1 = arena.acquire(bytes(bw.conv1.Cout, hw)))) return fail_alloc();
                if (add_conv(bw.conv1, x, hw, nullptr, true, t1)) return nullptr;
            }
            void* t2 = arena.acquire(bytes(bw.conv2.Cout, out_hw));
            if (!t2) return fail_alloc();
            if (add_conv(bw.conv2, t1, hw, nullptr, true, t2)) return nullptr;
            arena.release(t1);
            if (!(y = arena.acquire(bytes(out_c, out_hw)))) return fail_alloc();
            sm_budget = num_sms;
            // layer2-shaped blocks (128 -> 512): conv3 + shortcut + ReLU and the NEXT block's conv1 in one launch
            const BlockWeights* nb = bi + 1 < blocks.size() ? &blocks[bi + 1] : nullptr;
            // (layer2 shape 128 -> 512 -> 128: resident weights; layer3 shape 256 -> 1024 -> 256: streamed weights;
            // RNB_C3N1=0 off, 1 = layer2 only, default both)
            const char* c3env = getenv("RNB_C3N1");
            const int c3level = c3env ? atoi(c3env) : 2;
            const bool c3n1 = fuse_level >= 1 && fuse_next && besz == 2 && nb && nb->bottleneck && !nb->has_ds &&
                              nb->conv1.Cin == out_c && nb->conv2.stride == 1 &&
                              c3n1_shape_ok(bw.conv3.Cin, out_c, nb->conv1.Cout) &&
                              c3level >= (bw.conv3.Cin == 128 ? 1 : 2);
            // fused (one launch: conv3 + shortcut + next conv1) or plain (two launches with tile autotune and tail
            // split)? Which wins depends on how the tile count of THIS batch size falls on the 74 CTA pairs (fused
            // layer3 at 98 tiles: a full wave + a third; at 149: two waves + one tile), so with the autotuner on the
            // first such block of a shape is timed both ways on its real buffers (RNB_C3N1_AUTO=0: always fused).
            bool c3n1_use = c3n1;
            if (c3n1 && autotune && c3n1_auto) {
                const std::tuple<int, int, int, int> key{n * out_hw * out_hw, bw.conv3.Cin, out_c, nb->conv1.Cout};
                auto hit = tuned_c3n1.find(key);
                if (hit != tuned_c3n1.end()) {
                    c3n1_use = hit->second != 0;
                } else {
                    void* t1n_trial = arena.acquire(bytes(nb->conv1.Cout, out_hw));
                    if (!t1n_trial) return fail_alloc();
                    const size_t first = p.convs.size();
                    float ms_plain = 1e30f, ms_fused = 1e30f;
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0);
                    cudaEventCreate(&e1);
                    if (!add_conv(bw.conv3, t2, out_hw, shortcut, true, y) &&
                        !add_conv(nb->conv1, y, out_hw, nullptr, true, t1n_trial)) {
                        bool ok = true;
                        for (int i = 0; i < 7 && ok; ++i) {
                            if (i == 2) cudaEventRecord(e0, cap_stream);
                            ok = conv_plan_launch(p.convs[first], cap_stream) == cudaSuccess &&
                                 conv_plan_launch(p.convs[first + 1], cap_stream) == cudaSuccess;
                        }
                        cudaEventRecord(e1, cap_stream);
                        if (cudaStreamSynchronize(cap_stream) == cudaSuccess && ok) cudaEventElapsedTime(&ms_plain, e0, e1);
                    }
                    p.convs.resize(first);
                    if (fp8) p.links.resize(first);
                    C3n1Desc td{};
                    td.M = n * out_hw * out_hw;
                    td.K3 = bw.conv3.Cin; td.N3 = out_c; td.N1 = nb->conv1.Cout;
                    td.t2 = t2; td.w3 = bw.conv3.w; td.bias3 = bw.conv3.bias; td.residual = shortcut; td.y = y;
                    td.w1n = nb->conv1.w; td.bias1n = nb->conv1.bias; td.t1n = t1n_trial;
                    ConvPlan tf;
                    if (!c3n1_plan_init(&tf, td, num_sms, err, sizeof(err))) {
                        bool ok = true;
                        for (int i = 0; i < 7 && ok; ++i) {
                            if (i == 2) cudaEventRecord(e0, cap_stream);
                            ok = conv_plan_launch(tf, cap_stream) == cudaSuccess;
                        }
                        cudaEventRecord(e1, cap_stream);
                        if (cudaStreamSynchronize(cap_stream) == cudaSuccess && ok) cudaEventElapsedTime(&ms_fused, e0, e1);
                    }
                    cudaEventDestroy(e0);
                    cudaEventDestroy(e1);
                    arena.release(t1n_trial);
                    c3n1_use = ms_fused <= 1.03f * ms_plain;  // the fused form unless the plain one is clearly faster
                    tuned_c3n1[key] = c3n1_use ? 1 : 0;
                    if (getenv("RNB_VERBOSE"))
                        fprintf(stderr, "rnb plan: conv3 + next conv1 at M=%d (%d -> %d -> %d): fused %.1f us, plain %.1f us\n",
                                std::get<0>(key), bw.conv3.Cin, out_c, nb->conv1.Cout, ms_fused * 200.f, ms_plain * 200.f);
                }
            }
            if (c3n1_use) {
                void* t1n = arena.acquire(bytes(nb->conv1.Cout, out_hw));
                if (!t1n) return fail_alloc();
                const int M = n * out_hw * out_hw;
                // Wave tail of the streamed (layer3-shaped) fused launch: it cannot be split in N (conv1' consumes whole
                // rows of y) nor in M (cta_group::2 with 64 rows per CTA costs the same cycles). RNB_C3N1_HYBRID=1
                // (opt-in, bit-identical, measured SLOWER: ResNet-152 B=128 3.76 vs 3.60 ms — two more launches per block
                // cost more than the saved half wave): the fused launch takes the whole waves only and the remaining
                // rows run as the plain conv3 (+ shortcut) and conv1' launches.
                int fused_rows = M;
                {
                    const int tiles = (M + 255) / 256, pairs = num_sms / 2, rem = tiles % pairs;
                    const char* hy = getenv("RNB_C3N1_HYBRID");
                    if (bw.conv3.Cin != 128 && tiles > pairs && rem > 0 && 2 * rem <= pairs && hy && atoi(hy) != 0)
                        fused_rows = (tiles - rem) * 256;
                }
                C3n1Desc cd{};
                cd.M = fused_rows;
                cd.K3 = bw.conv3.Cin; cd.N3 = out_c; cd.N1 = nb->conv1.Cout;
                cd.reverse = alternate_tiles && (p.convs.size() & 1) != 0;
                cd.t2 = t2; cd.w3 = bw.conv3.w; cd.bias3 = bw.conv3.bias; cd.residual = shortcut; cd.y = y;
                cd.w1n = nb->conv1.w; cd.bias1n = nb->conv1.bias; cd.t1n = t1n;
                ConvPlan cp;
                if (c3n1_plan_init(&cp, cd, num_sms, err, sizeof(err))) {
                    set_error(err);
                    return nullptr;
                }
                if (getenv("RNB_VERBOSE"))
                    fprintf(stderr, "rnb plan: conv#%zu n=%d %dx%d fused conv3 + next conv1 grid %d rows %d of %d\n",
                            p.convs.size(), n, out_hw, out_hw, cp.grid, fused_rows, M);
                if (fp8) p.links.push_back({nullptr, nullptr, nullptr});
                p.convs.push_back(cp);
                if (fused_rows < M) {
                    const size_t e2 = 2;  // BF16
                    auto off = [&](const void* base, int C) {
                        return static_cast<const uint8_t*>(base) + 1ull * fused_rows * C * e2;
                    };
                    uint8_t* y_rem = const_cast<uint8_t*>(off(y, out_c));
                    if (add_conv(bw.conv3, off(t2, bw.conv3.Cin), out_hw, off(shortcut, out_c), true, y_rem, M - fused_rows))
                        return nullptr;
                    if (add_conv(nb->conv1, y_rem, out_hw, nullptr, true, const_cast<uint8_t*>(off(t1n, nb->conv1.Cout)),
                                 M - fused_rows))
                        return nullptr;
                }
                pre_t1 = t1n;
            } else if (add_conv(bw.conv3, t2, out_hw, shortcut, true, y)) {
                return nullptr;
            }
            if (overlap_ds) p.convs.back().join = 1;
            arena.release(t2);
        } else {
            void* t1 = arena.acquire(bytes(bw.conv1.Cout, out_hw));
            if (!t1) return fail_alloc();
            if (add_conv(bw.conv1, x, hw, nullptr, true, t1)) return nullptr;
            if (!(y = arena.acquire(bytes(out_c, out_hw)))) return fail_alloc();
            if (add_conv(bw.conv2, t1, out_hw, shortcut, true, y)) return nullptr;
            arena.release(t1);
        }
        if (ds) arena.release(ds);
        arena.release(x);
        x = y;
        hw = out_hw;
        p.named[bw.name] = {y, out_c, hw, hw};
    }
    p.last = x;
    p.last_hw = hw * hw;
    p.last_c = final_c;
    p.last_scale = fp8 ? scale_of[x] : 1.f;
    p.pooled = static_cast<float*>(arena.acquire(1ull * n * final_c * sizeof(float)));
    if (!p.pooled) return fail_alloc();
    p.named["avgpool"] = {p.pooled, final_c, 1, 1};
    if (fc_tc) {
        if (!(p.pooled_bf16 = arena.acquire(1ull * n * final_c * 2))) return fail_alloc();
        // the logits buffer is the caller's: the output tensor map is patched per launch target in
        // enqueue_chunk (fc_plan_for), here only the shape-dependent part is fixed
    }
    // Everything is released again: the next plan (another chunk size) may share the blocks, since
    // chunks run back to back on one stream.
    arena.release(x);
    arena.release(p.pooled);
    if (p.pooled_bf16) arena.release(p.pooled_bf16);
    if (fp8) cudaStreamSynchronize(cap_stream);  // the pre-multiplied epilogue vectors were written on cap_stream
    plan_guard.ok = true;
    auto ins = plans.emplace(n, std::move(p));
    return &ins.first->second;
}

int Model::enqueue_chunk(ChunkPlan& p, InRef in, float* logits, int32_t* top1, cudaStream_t s) {
    const int n = p.n;
    const int s_hw = (6 + image - 7) / 2 + 1;
    const float* x = in.f32();
    const uint8_t* x_u8 = in.u8();
    if (in.kind == InRef::BF16_NCHW && !accepts_bf16_input()) {
        set_error("BF16 NCHW input needs the one-launch BF16 tensor-core stem (bf16 / fp8 model, 224 x 224)");
        return RNB_ERR_UNSUPPORTED;
    }
    if (x_u8 && !(stem_tc && stem_esz() == 2)) {
        // generic path: normalise into an FP32 NCHW staging tensor, then proceed as with float input
        RNB_CUDA(launch_u8_hwc_to_f32_nchw(x_u8, u8_scratch, n, image, image, norm_mean, norm_std, s));
        x = u8_scratch;
        x_u8 = nullptr;
    }
    if (stem_tc) {
        void* pool = p.pool_raw ? p.pool_raw : p.pool_out;
        if (x_u8) {
            RNB_CUDA(launch_stem_tc_pack_u8(x_u8, p.stem_out, n, norm_mean, norm_std, s));
            RNB_CUDA(launch_stem_tc_from_packed(p.stem_out, stem_wk, stem_bias, pool, n, s));
        } else if (in.kind == InRef::BF16_NCHW) {
            RNB_CUDA(launch_stem_tc_from_mixed(in.p, in.p2, std::min(in.nb, n), stem_wk, stem_bias, pool, n, s));
        } else {
            RNB_CUDA(launch_stem_any_part(stem_esz(), 0, x, p.stem_out, stem_wk, stem_bias, pool, n, s));
            RNB_CUDA(launch_stem_any_part(stem_esz(), 1, x, p.stem_out, stem_wk, stem_bias, pool, n, s));
        }
        if (p.pool_raw) {
            const int p_hw = (2 + s_hw - 3) / 2 + 1;
            RNB_CUDA(launch_quantize_pad_bf16(p.pool_raw, p.pool_out, 1LL * n * p_hw * p_hw, 64, cpad(64),
                                              1.f / fp8_stem_scale, s));
        }
    } else {
        RNB_CUDA(launch_stem_conv(x, stem_w, stem_bias, p.stem_out, n, image, image, esz, s));
        RNB_CUDA(launch_maxpool_nhwc(p.stem_out, p.pool_out, n, s_hw, s_hw, 64, esz, s));
    }
    {
        int rc = launch_convs(p, s, nullptr);
        if (rc) return rc;
    }
    RNB_CUDA(launch_avgpool_nhwc(p.last, p.pooled, p.pooled_bf16, n, p.last_hw, p.last_c, esz, s, p.last_scale));
    int r = enqueue_fc(p, logits, s);
    if (r) return r;
    if (top1) RNB_CUDA(launch_argmax_f32(logits, top1, n, classes, s));
    return RNB_OK;
}

// The planned conv launches of one chunk, in order. A launch marked `side` goes to the side stream (forked
// from `s` by an event, so it also works under stream capture); the launch marked `join` waits for it.
// `evs` (profiling): an event is recorded on `s` after every launch.
int Model::launch_convs(ChunkPlan& p, cudaStream_t s, cudaEvent_t* evs) {
    static const bool sync_each = getenv("RNB_SYNC_EACH") != nullptr;  // bring-up: localise a faulting launch
    int nfork = 0;
    for (size_t i = 0; i < p.convs.size(); ++i) {
        const ConvPlan& cp = p.convs[i];
        if (cp.side) {
            const int k = nfork++ & 3;
            RNB_CUDA(cudaEventRecord(fork_ev[k], s));
            RNB_CUDA(cudaStreamWaitEvent(side_stream, fork_ev[k], 0));
            RNB_CUDA(conv_plan_launch(cp, side_stream));
            RNB_CUDA(cudaEventRecord(join_ev[k], side_stream));
        } else {
            if (cp.join && nfork > 0) RNB_CUDA(cudaStreamWaitEvent(s, join_ev[(nfork - 1) & 3], 0));
            RNB_CUDA(conv_plan_launch(cp, s));
        }
        if (evs) RNB_CUDA(cudaEventRecord(evs[i], s));
        if (sync_each) {
            cudaError_t e = cudaStreamSynchronize(s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(side_stream);
            if (e != cudaSuccess) {
                set_error("conv launch #" + std::to_string(i) + " failed: " + cudaGetErrorString(e));
                return RNB_ERR_CUDA;
            }
        }
    }
    return RNB_OK;
}

// logits = pooled x fc.weight^T + fc.bias (Linear::forward, nn.cu:55-64; main.cu:222-224). BF16 path:
// the single-CTA tcgen05 conv kernel as a 1x1 "conv" over [n,1,1,C] with unrounded FP32 output straight
// into the caller's logits buffer (plan cached per output pointer); otherwise the FP32 CUDA-core kernel.
int Model::enqueue_fc(ChunkPlan& p, float* logits, cudaStream_t s) {
    if (!p.pooled_bf16) {
        RNB_CUDA(launch_fc(p.pooled, fc_w, fc_b, logits, p.n, p.last_c, classes, s));
        return RNB_OK;
    }
    if (p.fc_out != logits) {
        ConvDesc d{};
        d.B = p.n; d.H = 1; d.W = 1; d.Cin = p.last_c; d.Cout = classes_pad;
        d.ksize = 1; d.stride = 1; d.pad = 0; d.relu = false; d.act = ActType::BF16;
        d.in = p.pooled_bf16; d.weight = fc_wq; d.bias = fc_bq; d.residual = nullptr; d.out = logits;
        d.out_f32 = true; d.out_cols = classes;
        char err[256];
        if (conv_plan_init(&p.fc_plan, d, num_sms, 0, err, sizeof(err))) {
            set_error(err);
            return RNB_ERR_CUDA;
        }
        p.fc_out = logits;
    }
    RNB_CUDA(conv_plan_launch(p.fc_plan, s));
    return RNB_OK;
}

// ------------------------------------------------------------------------------------ FP8 calibration
// Writes the calibrated per-tensor scales into the launch geometry of a plan (plans made before / after calibration).
int Model::apply_fp8_scales(ChunkPlan& p, cudaStream_t s) {
    std::map<const void*, float> scale_of;
    scale_of[p.pool_out] = fp8_stem_scale;
    for (size_t i = 0; i < p.convs.size(); ++i) {
        const ChunkPlan::Link& l = p.links[i];
        ConvPlan& cp = p.convs[i];
        if (cp.quant) {
            cp.q_inv_scale = 1.f / fp8_out_scale[i];
            scale_of[l.out] = fp8_out_scale[i];
            continue;
        }
        if (!cp.fp8_vecs) continue;  // BF16 head of a mixed plan
        // (a BF16 input — the hand-over convs — has no entry: scale 1)
        const float s_in = scale_of.count(l.in) ? scale_of[l.in] : 1.f;
        RNB_CUDA(fp8_premultiply(&cp, s_in, l.res ? scale_of[l.res] : 1.f, fp8_out_scale[i], s));
        cp.g.amax = nullptr;
        scale_of[l.out] = fp8_out_scale[i];
    }
    p.last_scale = scale_of[p.last];
    return RNB_OK;
}

// One pass over `n` calibration images, in network order: the stem runs in BF16 and its output maximum fixes the first
// scale; every conv is then launched twice on its already quantised inputs — once with an amax-recording epilogue
// (output scale 1), once for real with scale = amax / 448. Blocking; runs on cap_stream.
int Model::calibrate_fp8(const float* x, int n) {
    if (!fp8 || fp8_calibrated) return RNB_OK;
    cudaStream_t s = cap_stream;
    RNB_CUDA(cudaDeviceSynchronize());
    ChunkPlan* pp = plan_for(n);
    if (!pp) return RNB_ERR_CUDA;
    ChunkPlan& p = *pp;
    if (!fp8_amax_dev) RNB_CUDA(cudaMalloc(&fp8_amax_dev, sizeof(float)));
    const int s_hw = (6 + image - 7) / 2 + 1, p_hw = (2 + s_hw - 3) / 2 + 1;
    auto read_amax = [&](float* out) -> int {
        RNB_CUDA(cudaMemcpyAsync(out, fp8_amax_dev, sizeof(float), cudaMemcpyDeviceToHost, s));
        RNB_CUDA(cudaStreamSynchronize(s));
        return RNB_OK;
    };
    void* pool = p.pool_raw ? p.pool_raw : p.pool_out;
    RNB_CUDA(launch_stem_any_part(2, 0, x, p.stem_out, stem_wk, stem_bias, pool, n, s));
    RNB_CUDA(launch_stem_any_part(2, 1, x, p.stem_out, stem_wk, stem_bias, pool, n, s));
    float amax = 0.f;
    int r = RNB_OK;
    if (p.pool_raw) {  // whole network in FP8: the stem output is the first quantised tensor
        RNB_CUDA(cudaMemsetAsync(fp8_amax_dev, 0, sizeof(float), s));
        RNB_CUDA(launch_amax_bf16(p.pool_raw, 1LL * n * p_hw * p_hw * 64, fp8_amax_dev, s));
        if ((r = read_amax(&amax))) return r;
        fp8_stem_scale = amax > 0.f ? amax / 448.f : 1.f;
        RNB_CUDA(launch_quantize_pad_bf16(p.pool_raw, p.pool_out, 1LL * n * p_hw * p_hw, 64, cpad(64), 1.f / fp8_stem_scale, s));
    }
    fp8_out_scale.assign(p.convs.size(), 1.f);
    std::map<const void*, float> scale_of;
    scale_of[p.pool_out] = fp8_stem_scale;
    for (size_t i = 0; i < p.convs.size(); ++i) {
        ConvPlan& cp = p.convs[i];
        const ChunkPlan::Link& l = p.links[i];
        if (cp.quant) {  // hand-over of a mixed plan: scale from the maximum of the BF16 tensor
            RNB_CUDA(cudaMemsetAsync(fp8_amax_dev, 0, sizeof(float), s));
            RNB_CUDA(launch_amax_bf16(cp.q_src, cp.q_rows * cp.q_c, fp8_amax_dev, s));
            if ((r = read_amax(&amax))) return r;
            const float sc = amax > 0.f ? amax / 448.f : 1.f;
            fp8_out_scale[i] = sc;
            cp.q_inv_scale = 1.f / sc;
            RNB_CUDA(conv_plan_launch(cp, s));
            scale_of[l.out] = sc;
            continue;
        }
        if (!cp.fp8_vecs) {  // BF16 launch
            RNB_CUDA(conv_plan_launch(cp, s));
            continue;
        }
        const float s_in = scale_of.count(l.in) ? scale_of[l.in] : 1.f, s_res = l.res ? scale_of[l.res] : 1.f;
        RNB_CUDA(fp8_premultiply(&cp, s_in, s_res, 1.f, s));
        cp.g.amax = fp8_amax_dev;
        RNB_CUDA(cudaMemsetAsync(fp8_amax_dev, 0, sizeof(float), s));
        RNB_CUDA(conv_plan_launch(cp, s));
        if ((r = read_amax(&amax))) return r;
        if (!(amax == amax) || amax > 3.0e38f) {
            set_error("FP8 calibration: conv #" + std::to_string(i) + " produced a non-finite maximum");
            return RNB_ERR_CUDA;
        }
        const float sc = amax > 0.f ? amax / 448.f : 1.f;
        fp8_out_scale[i] = sc;
        RNB_CUDA(fp8_premultiply(&cp, s_in, s_res, sc, s));
        cp.g.amax = nullptr;
        RNB_CUDA(conv_plan_launch(cp, s));
        scale_of[l.out] = sc;
    }
    RNB_CUDA(cudaStreamSynchronize(s));
    fp8_calibrated = true;
    for (auto& kv : plans) {
        if ((r = apply_fp8_scales(kv.second, s))) return r;
    }
    RNB_CUDA(cudaStreamSynchronize(s));
    // graphs captured with provisional scales (none in practice: calibration precedes the first capture)
    for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
    graphs.clear();
    if (getenv("RNB_VERBOSE")) {
        fprintf(stderr, "rnb fp8 calibration: stem scale %.4g;", fp8_stem_scale);
        for (size_t i = 0; i < fp8_out_scale.size(); ++i) fprintf(stderr, " %.3g", fp8_out_scale[i]);
        fprintf(stderr, "\n");
    }
    return RNB_OK;
}

int Model::forward(const float* x, int batch, float* logits, int32_t* top1, cudaStream_t s) {
    return forward_any(InRef{x, InRef::F32_NCHW}, batch, logits, top1, s);
}

int Model::forward_u8(const uint8_t* x, int batch, float* logits, int32_t* top1, cudaStream_t s) {
    return forward_any(InRef{x, InRef::U8_HWC}, batch, logits, top1, s);
}

int Model::forward_bf16(const uint16_t* x, int batch, float* logits, int32_t* top1, cudaStream_t s) {
    if (!accepts_bf16_input()) {
        set_error("BF16 NCHW input needs the one-launch BF16 tensor-core stem (bf16 / fp8 model, 224 x 224)");
        return RNB_ERR_UNSUPPORTED;
    }
    return forward_any(InRef{x, InRef::BF16_NCHW, nullptr, batch}, batch, logits, top1, s);
}

int Model::set_normalization(const float* mean, const float* std) {
    for (int c = 0; c < 3; ++c) {
        if (!(std[c] > 0.f)) {
            set_error("set_normalization: std must be positive");
            return RNB_ERR_INVALID;
        }
    }
    cudaDeviceSynchronize();
    for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);  // the constants are baked into captured launches
    graphs.clear();
    for (int c = 0; c < 3; ++c) {
        norm_mean[c] = mean[c];
        norm_std[c] = std[c];
    }
    if (lane2) return lane2->set_normalization(mean, std);
    return RNB_OK;
}

// ------------------------------------------------------------------------------------ two lanes
int Model::make_lane() {
    std::unique_ptr<Model> l(new Model());
    l->is_lane = true;
    l->arch = arch; l->esz = esz; l->classes = classes; l->image = image; l->max_batch = max_batch; l->chunk = chunk;
    l->num_sms = num_sms; l->device = device; l->bottleneck = bottleneck; l->final_c = final_c;
    l->stem_w = stem_w; l->stem_bias = stem_bias; l->stem_wk = stem_wk; l->stem_tc = stem_tc;
    l->blocks = blocks;
    l->fc_w = fc_w; l->fc_b = fc_b; l->fc_wq = fc_wq; l->fc_bq = fc_bq; l->classes_pad = classes_pad; l->fc_tc = fc_tc;
    l->num_convs = num_convs; l->flops_per_image = flops_per_image;
    for (int c = 0; c < 3; ++c) {
        l->norm_mean[c] = norm_mean[c];
        l->norm_std[c] = norm_std[c];
    }
    l->use_graph = use_graph; l->alternate_tiles = alternate_tiles; l->autotune = autotune; l->side_sms = 0;
    l->fuse_level = fuse_level; l->fuse_next = fuse_next; l->arena.keep = arena.keep;
    // FP8: the lane runs with the SAME calibrated scales (a lane is only made after calibration)
    l->fp8 = fp8; l->fp8_first_block = fp8_first_block; l->fp8_fused_handover = fp8_fused_handover;
    l->fp8_calibrated = fp8_calibrated; l->fp8_stem_scale = fp8_stem_scale; l->fp8_out_scale = fp8_out_scale;
    l->fp8_ones = fp8_ones;
    int r = l->create_streams();
    if (r) return r;
    RNB_CUDA(cudaEventCreateWithFlags(&lane_fork, cudaEventDisableTiming));
    RNB_CUDA(cudaEventCreateWithFlags(&lane_join, cudaEventDisableTiming));
    lane2 = std::move(l);
    return RNB_OK;
}

int Model::forward_two(InRef in, int batch, float* logits, int32_t* top1, cudaStream_t s) {
    const int n0 = (batch + 1) / 2, n1 = batch - n0;
    const size_t img = 3ull * image * image;
    cudaStream_t s2 = lane2->host_compute;
    RNB_CUDA(cudaEventRecord(lane_fork, s));
    RNB_CUDA(cudaStreamWaitEvent(s2, lane_fork, 0));
    int r = forward_one(in, n0, logits, top1, s);
    if (r) return r;
    r = lane2->forward_one(in.at(n0, img), n1,
                           logits ? logits + 1ull * n0 * classes : nullptr, top1 ? top1 + n0 : nullptr, s2);
    if (r) return r;
    RNB_CUDA(cudaEventRecord(lane_join, s2));
    RNB_CUDA(cudaStreamWaitEvent(s, lane_join, 0));
    return RNB_OK;
}

// 1 or 2 lanes for this batch size; the first use of a size times both forms on the caller's buffers (blocking)
int Model::lanes_for(int batch, InRef in, float* logits, int32_t* top1, cudaStream_t s) {
    if (is_lane || (fp8 && !fp8_calibrated) || batch < 2 || batch > chunk || arena.keep) return 1;
    auto it = lane_choice.find(batch);
    if (it != lane_choice.end()) return it->second;
    const int forced = getenv("RNB_LANES") ? atoi(getenv("RNB_LANES")) : 0;
    int choice = 1;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
    if (forced == 1 || forced == 2) {
        choice = forced;
    } else if (autotune && !capturing && logits) {
        if (!lane2 && make_lane()) return 1;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float ms[2] = {1e30f, 1e30f};
        for (int form = 0; form < 2; ++form) {
            bool ok = true;
            for (int i = 0; i < 2 && ok; ++i)
                ok = (form ? forward_two(in, batch, logits, top1, s) : forward_one(in, batch, logits, top1, s)) == RNB_OK;
            cudaEventRecord(e0, s);
            for (int i = 0; i < 4 && ok; ++i)
                ok = (form ? forward_two(in, batch, logits, top1, s) : forward_one(in, batch, logits, top1, s)) == RNB_OK;
            cudaEventRecord(e1, s);
            if (cudaStreamSynchronize(s) == cudaSuccess && ok) cudaEventElapsedTime(&ms[form], e0, e1);
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        choice = ms[1] < 0.98f * ms[0] ? 2 : 1;
        if (getenv("RNB_VERBOSE"))
            fprintf(stderr, "rnb lanes: batch %d one lane %.3f ms, two lanes %.3f ms per step -> %d\n", batch, ms[0] / 4,
                    ms[1] / 4, choice);
    }
    else {
        return 1;  // no decision possible now (no output buffer to time on, or the stream is capturing): ask again later
    }
    if (choice == 2 && !lane2 && make_lane()) choice = 1;
    lane_choice[batch] = choice;
    return choice;
}

int Model::forward_any(InRef in, int batch, float* logits, int32_t* top1, cudaStream_t s) {
    if (batch <= 0 || batch > max_batch) {
        set_error("batch must be in [1, max_batch]");
        return RNB_ERR_INVALID;
    }
    if (!in.p) {
        set_error("x_dev is NULL");
        return RNB_ERR_INVALID;
    }
    if (lanes_for(batch, in, logits, top1, s) == 2) return forward_two(in, batch, logits, top1, s);
    return forward_one(in, batch, logits, top1, s);
}

int Model::forward_one(InRef in, int batch, float* logits, int32_t* top1, cudaStream_t s) {
    const float* x = in.f32();
    const uint8_t* x_u8 = in.u8();
    if (in.kind == InRef::BF16_NCHW && fp8 && !fp8_calibrated) {
        set_error("an FP8 model is calibrated on FP32 or uint8 input: calibrate before feeding BF16 input");
        return RNB_ERR_INVALID;
    }
    if (x_u8 && (!(stem_tc && stem_esz() == 2) || (fp8 && !fp8_calibrated)) && !u8_scratch)
        RNB_CUDA(cudaMalloc(&u8_scratch, 1ull * chunk * 3 * image * image * sizeof(float)));
    if (!logits) {
        if (!scratch_logits)
            RNB_CUDA(cudaMalloc(&scratch_logits, 1ull * max_batch * classes * sizeof(float)));
        logits = scratch_logits;
    }
    const size_t img_elems = 3ull * image * image;
    // Planning a new chunk size device-synchronises and times trial launches: not inside somebody's stream capture.
    for (int off = 0; off < batch; off += chunk) {
        if (plans.count(std::min(chunk, batch - off))) continue;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
            set_error("forward: this batch size has not been planned yet and the stream is capturing; call "
                      "rnb_model_warmup(batch) first");
            return RNB_ERR_INVALID;
        }
    }
    if (fp8 && !fp8_calibrated) {
        // first forward of an FP8 model: its first chunk is the calibration batch (blocking; not under capture —
        // the planning check above has already refused a capturing stream, since nothing is planned yet)
        const int n = std::min(chunk, batch);
        RNB_CUDA(cudaStreamSynchronize(s));
        const float* xcal = x;
        if (!xcal) {
            RNB_CUDA(launch_u8_hwc_to_f32_nchw(x_u8, u8_scratch, n, image, image, norm_mean, norm_std, cap_stream));
            xcal = u8_scratch;
        }
        int r = calibrate_fp8(xcal, n);
        if (r) return r;
    }
    {
        int r = begin_enqueue(s);
        if (r) return r;
    }
    for (int off = 0; off < batch; off += chunk) {
        const int n = std::min(chunk, batch - off);
        ChunkPlan* p = plan_for(n);
        if (!p) return RNB_ERR_CUDA;
        const InRef inc = in.at(off, img_elems);
        float* lc = logits + 1ull * off * classes;
        int32_t* tc = top1 ? top1 + off : nullptr;
        if (!use_graph) {
            int r = enqueue_chunk(*p, inc, lc, tc, s);
            if (r) return r;
            continue;
        }
        // (U8_HWC = bit 29, BF16_NCHW = bit 30; the count of BF16 images of a mixed batch above bit 32)
        const int64_t shape = n | (inc.kind << 29) |
                              (inc.kind == InRef::BF16_NCHW ? static_cast<int64_t>(std::min(inc.nb, n)) << 32 : 0);
        const GraphKey key{shape, inc.p, inc.p2, lc, tc};
        auto g = graphs.find(key);
        if (g == graphs.end()) {
            // executables already held for this shape, least recently used first
            int held = 0;
            auto lru = graphs.end();
            for (auto it = graphs.begin(); it != graphs.end(); ++it) {
                if (std::get<0>(it->first) != shape) continue;
                ++held;
                if (lru == graphs.end() || it->second.used < lru->second.used) lru = it;
            }
            cudaGraph_t graph = nullptr;
            int r = capture_chunk(*p, inc, lc, tc, &graph);
            if (r) return r;
            cudaGraphExec_t exec = nullptr;
            if (held >= kGraphsPerShape) {
                // re-target the least recently used executable: same topology, new pointers
                cudaGraphExecUpdateResultInfo info{};
                if (cudaGraphExecUpdate(lru->second.exec, graph, &info) == cudaSuccess) {
                    exec = lru->second.exec;
                } else {
                    cudaGetLastError();
                    cudaGraphExecDestroy(lru->second.exec);
                }
                graphs.erase(lru);
            }
            if (!exec) {
                cudaError_t ce = cudaGraphInstantiate(&exec, graph, 0);
                if (ce != cudaSuccess) {
                    cudaGraphDestroy(graph);
                    return fail_cuda(ce, "cudaGraphInstantiate");
                }
            }
            cudaGraphDestroy(graph);
            g = graphs.emplace(key, GraphEntry{exec, 0}).first;
        }
        g->second.used = ++graph_clock;
        RNB_CUDA(cudaGraphLaunch(g->second.exec, s));
    }
    {
        int r = end_enqueue(s);
        if (r) return r;
    }
    last_chunk_n = std::min(chunk, batch);
    return RNB_OK;
}

// One chunk captured into a graph on cap_stream; the graph is destroyed on every failure path.
int Model::capture_chunk(ChunkPlan& p, InRef in, float* logits, int32_t* top1, cudaGraph_t* out) {
    *out = nullptr;
    RNB_CUDA(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
    const int r = enqueue_chunk(p, in, logits, top1, cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
    if (r || ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        if (r) return r;
        return fail_cuda(ce, "cudaStreamEndCapture");
    }
    *out = graph;
    return RNB_OK;
}

int Model::warmup(int batch, bool include_u8) {
    if (batch <= 0 || batch > max_batch) {
        set_error("warmup: batch must be in [1, max_batch]");
        return RNB_ERR_INVALID;
    }
    if (fp8 && !fp8_calibrated) {
        set_error("warmup: an FP8 model must be calibrated first (rnb_model_calibrate, or a first forward on real data)");
        return RNB_ERR_INVALID;
    }
    // a real forward on scratch buffers: plans every chunk size of this batch, autotunes, allocates lazily
    // created buffers and instantiates one graph per chunk shape (later calls with other pointers re-target it)
    const size_t img_elems = 3ull * image * image;
    float* x = nullptr;
    RNB_CUDA(cudaMalloc(&x, 1ull * batch * img_elems * sizeof(float)));
    cudaMemset(x, 0, 1ull * batch * img_elems * sizeof(float));
    if (!host_top1_dev && cudaMalloc(&host_top1_dev, 1ull * max_batch * sizeof(int32_t)) != cudaSuccess) {
        cudaFree(x);
        return fail_cuda(cudaGetLastError(), "warmup: cudaMalloc");
    }
    // (a real logits buffer, so that the one-lane / two-lane decision for this batch size is timed here and not in
    // the first serving call)
    if (!scratch_logits && cudaMalloc(&scratch_logits, 1ull * max_batch * classes * sizeof(float)) != cudaSuccess) {
        cudaFree(x);
        return fail_cuda(cudaGetLastError(), "warmup: cudaMalloc");
    }
    int r = forward(x, batch, scratch_logits, host_top1_dev, host_compute);
    if (!r && include_u8)
        r = forward_u8(reinterpret_cast<const uint8_t*>(x), batch, scratch_logits, host_top1_dev, host_compute);
    // the graph of the host paths' BF16 input form (host_pack.h), unless packing is ruled out already
    if (!r && accepts_bf16_input() && host_pack_mode != 0)   // (a mixed batch has its own graph: captured at first use)
        r = forward_bf16(reinterpret_cast<const uint16_t*>(x), batch, scratch_logits, host_top1_dev, host_compute);
    cudaError_t ce = cudaStreamSynchronize(host_compute);
    cudaFree(x);
    if (r) return r;
    if (ce != cudaSuccess) return fail_cuda(ce, "warmup: cudaStreamSynchronize");
    return RNB_OK;
}

int Model::repeat_launch(int batch, int index, int repeat, cudaStream_t s) {
    if (batch <= 0 || batch > max_batch || repeat <= 0) {
        set_error("repeat_launch: bad argument");
        return RNB_ERR_INVALID;
    }
    ChunkPlan* pp = plan_for(std::min(chunk, batch));
    if (!pp) return RNB_ERR_CUDA;
    if (index < 0 || index >= static_cast<int>(pp->convs.size())) {
        set_error("repeat_launch: launch index out of range");
        return RNB_ERR_INVALID;
    }
    int ro = begin_enqueue(s);
    if (ro) return ro;
    for (int i = 0; i < repeat; ++i) RNB_CUDA(conv_plan_launch(pp->convs[index], s));
    return end_enqueue(s);
}

int Model::profile(const float* x, int batch, int iters, int* kind, float* ms, double* flops,
                   double* bytes, int max_entries, int* n_entries, cudaStream_t s) {
    if (batch <= 0 || batch > max_batch || iters <= 0 || !x || !n_entries) {
        set_error("profile: bad argument");
        return RNB_ERR_INVALID;
    }
    const int n = std::min(chunk, batch);
    if (fp8 && !fp8_calibrated) {
        int rc = calibrate_fp8(x, n);
        if (rc) return rc;
    }
    ChunkPlan* pp = plan_for(n);
    if (!pp) return RNB_ERR_CUDA;
    ChunkPlan& p = *pp;
    const int L = static_cast<int>(p.convs.size()) + 5;
    *n_entries = L;
    if (max_entries < L) {
        set_error("profile: output arrays too small");
        return RNB_ERR_INVALID;
    }
    if (!scratch_logits)
        RNB_CUDA(cudaMalloc(&scratch_logits, 1ull * max_batch * classes * sizeof(float)));
    if (!host_top1_dev) RNB_CUDA(cudaMalloc(&host_top1_dev, 1ull * max_batch * sizeof(int32_t)));
    {
        int ro = begin_enqueue(s);
        if (ro) return ro;
    }
    std::vector<cudaEvent_t> ev(L + 1);
    for (auto& e : ev) RNB_CUDA(cudaEventCreate(&e));
    std::vector<double> acc(L, 0.0);
    const int s_hw = (6 + image - 7) / 2 + 1, p_hw = (2 + s_hw - 3) / 2 + 1;
    for (int it = -1; it < iters; ++it) {
        int i = 0;
        RNB_CUDA(cudaEventRecord(ev[i], s));
        void* pool = p.pool_raw ? p.pool_raw : p.pool_out;
        if (stem_tc)
            RNB_CUDA(launch_stem_any_part(stem_esz(), 0, x, p.stem_out, stem_wk, stem_bias, pool, n, s));
        else
            RNB_CUDA(launch_stem_conv(x, stem_w, stem_bias, p.stem_out, n, image, image, esz, s));
        RNB_CUDA(cudaEventRecord(ev[++i], s));
        if (stem_tc)
            RNB_CUDA(launch_stem_any_part(stem_esz(), 1, x, p.stem_out, stem_wk, stem_bias, pool, n, s));
        else
            RNB_CUDA(launch_maxpool_nhwc(p.stem_out, p.pool_out, n, s_hw, s_hw, 64, esz, s));
        if (p.pool_raw)  // the E4M3 copy of the stem output is booked with the stem's second launch
            RNB_CUDA(launch_quantize_pad_bf16(p.pool_raw, p.pool_out, 1LL * n * p_hw * p_hw, 64, cpad(64),
                                              1.f / fp8_stem_scale, s));
        RNB_CUDA(cudaEventRecord(ev[++i], s));
        {
            int rc = launch_convs(p, s, &ev[i + 1]);
            if (rc) return rc;
            i += static_cast<int>(p.convs.size());
        }
        RNB_CUDA(launch_avgpool_nhwc(p.last, p.pooled, p.pooled_bf16, n, p.last_hw, p.last_c, esz, s, p.last_scale));
        RNB_CUDA(cudaEventRecord(ev[++i], s));
        {
            int rr = enqueue_fc(p, scratch_logits, s);
            if (rr) return rr;
        }
        RNB_CUDA(cudaEventRecord(ev[++i], s));
        RNB_CUDA(launch_argmax_f32(scratch_logits, host_top1_dev, n, classes, s));
        RNB_CUDA(cudaEventRecord(ev[++i], s));
        RNB_CUDA(cudaStreamSynchronize(s));
        if (it < 0) continue;  // warm-up pass
        for (int k = 0; k < L; ++k) {
            float t = 0.f;
            RNB_CUDA(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
            acc[k] += t;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    {
        int ro = end_enqueue(s);
        if (ro) return ro;
    }
    int i = 0;
    auto put = [&](int k, double f, double b) {
        kind[i] = k;
        ms[i] = static_cast<float>(acc[i] / iters);
        flops[i] = f;
        bytes[i] = b;
        ++i;
    };
    const double img_px = 1.0 * image * image;
    if (stem_tc) {
        // launch 0 = layout pre-pass (fp32 NCHW -> padded NHWC4 bf16), launch 1 = fused conv+BN+ReLU+pool
        // (fused form: launch 0 does not exist — its entry times an empty interval — and launch 1 reads the FP32 image
        // itself)
        const bool fused = stem_esz() == 4 || stem_fused_enabled();  // (the TF32 split stem has no two-launch form)
        const double packed = static_cast<double>(stem_any_input_bytes(stem_esz(), 1));
        put(0, 0.0, fused ? 0.0 : n * (3.0 * img_px * 4 + packed));
        put(1, 2.0 * n * 64 * 147 * s_hw * s_hw,
            n * ((fused ? 3.0 * img_px * 4 : packed) + 64.0 * p_hw * p_hw * stem_esz() +
                 (p.pool_raw ? (64.0 * 2 + 128.0) * p_hw * p_hw : 0.0)) + 28672.0);
    } else {
        put(0, 2.0 * n * 64 * 147 * s_hw * s_hw, n * (3.0 * img_px * 4 + 64.0 * s_hw * s_hw * esz) + 64 * 148 * 4.0);
        put(1, 0.0, n * 64.0 * esz * (1.0 * s_hw * s_hw + 1.0 * p_hw * p_hw));
    }
    for (const ConvPlan& cp : p.convs) put(2, cp.flops, cp.bytes);
    put(3, 0.0, 1.0 * n * p.last_c * (1.0 * p.last_hw * esz + 4.0));  // (FP8: last_c is a multiple of 128 already)
    put(4, 2.0 * n * p.last_c * classes, 4.0 * (1.0 * n * p.last_c + 1.0 * classes * p.last_c + 1.0 * n * classes));
    put(5, 0.0, 4.0 * n * (classes + 1.0));
    return RNB_OK;
}

// FP32 host input of a BF16-stem model: how many leading images of the batch do the host cores round to BF16 (half the
// bytes on PCIe) while the rest crosses as FP32? Fixed once per model and kind of host memory (see model.h).
int Model::host_pack_for(const float* x, int batch, uint16_t** stage, float* x_dev) {
    host_pack_last = 0;
    if (!accepts_bf16_input() || (fp8 && !fp8_calibrated)) return 0;   // (calibration reads the FP32 image)
    if (host_pack_mode < 0) {
        const char* e = getenv("RNB_HOST_PACK");
        if (e && (atoi(e) == 0 || atoi(e) == 1)) host_pack_mode = atoi(e);
    }
    if (host_pack_mode == 0) return 0;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, x) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const int kind = pinned ? 1 : 0;
    host_pack_last_kind = kind;
    if (host_pack_mode < 0 && host_pack_frac[kind] == 0) return 0;
    // a small batch says nothing about a serving loop (thread wake-ups and launch latencies, not rates): it goes the
    // plain way, without staging memory, and the question stays open for the first batch of at least 32 images
    if (host_pack_mode < 0 && host_pack_frac[kind] < 0 && batch < 32) return 0;
    const size_t img_elems = 3ull * image * image;
    if (!*stage && cudaHostAlloc(reinterpret_cast<void**>(stage), 1ull * max_batch * img_elems * sizeof(uint16_t),
                                 cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        *stage = nullptr;
        host_pack_mode = 0;   // no pinned memory to stage in: the plain copy still works
        return 0;
    }
    // the share of a batch, in whole 16-image upload pieces; a remainder below one piece joins the other side
    auto images_of = [&](double frac) {
        const int nb = host_pack_images(frac, batch);
        host_pack_last = static_cast<double>(nb) / batch;
        return nb;
    };
    if (host_pack_mode == 1) return images_of(1.0);
    if (host_pack_frac[kind] > 0) return images_of(host_pack_frac[kind]);
    // Time the parts on samples of this very batch (its memory: pinned or pageable, its NUMA placement): a plain FP32
    // copy, the conversion with the pool, a BF16 copy of the result. Two rounds, the second one counts — on OTHER
    // images where the batch has them, so that the conversion reads memory, not cache; a copy is timed as the faster
    // of two in a row (the first one after an idle stretch has been seen 8x slower).
    using clk = std::chrono::steady_clock;
    const int ns = std::min(batch, 32);
    const size_t n = static_cast<size_t>(ns) * img_elems;
    HostPacker& hp = HostPacker::instance();
    double t[3] = {0, 0, 0};
    auto time_copy = [&](const void* src, size_t bytes) {
        double best = 1e30;
        for (int again = 0; again < 2; ++again) {
            cudaStreamSynchronize(copy_stream);
            const auto t0 = clk::now();
            cudaMemcpyAsync(x_dev, src, bytes, cudaMemcpyHostToDevice, copy_stream);
            cudaStreamSynchronize(copy_stream);
            best = std::min(best, std::chrono::duration<double>(clk::now() - t0).count());
        }
        return best;
    };
    for (int rep = 0; rep < 2; ++rep) {
        const float* xs = x + (rep == 1 && batch >= 2 * ns ? n : 0);
        // (the FP32 copy first: once the cores have read the sample, the DMA engine finds it in their caches and the
        // copy runs at 60 % of its rate from memory — which is where a serving loop's batches are)
        t[1] = time_copy(xs, n * sizeof(float));
        const auto t0 = clk::now();
        hp.run(xs, *stage, n, n, [](size_t, size_t) {});
        t[0] = std::chrono::duration<double>(clk::now() - t0).count();
        t[2] = time_copy(*stage, n * sizeof(uint16_t));
    }
    if (cudaGetLastError() != cudaSuccess) {
        host_pack_frac[kind] = 0;
        return 0;
    }
    for (int i = 0; i < 3; ++i) host_pack_gbps[kind][i] = (i == 2 ? 2.0 : 4.0) * n / std::max(t[i], 1e-9) * 1e-9;
    // cores and link share a batch in the proportion that lets both finish together (host_pack.h)
    const double f = host_pack_split(t[0], t[1], t[2]);
    host_pack_frac[kind] = f;
    if (getenv("RNB_VERBOSE"))
        fprintf(stderr, "rnb host pack (%s input): %d threads convert %.1f GB/s (FP32 read), H2D FP32 %.1f GB/s, H2D BF16 "
                        "%.1f GB/s -> %.0f %% of every batch rounded to BF16 on the host\n", pinned ? "pinned" : "pageable",
                hp.threads(), host_pack_gbps[kind][0], host_pack_gbps[kind][1], host_pack_gbps[kind][2], 100.0 * f);
    return f > 0 ? images_of(f) : 0;
}

int Model::forward_host(const float* x, int batch, float* logits, int32_t* top1) {
    if (batch <= 0 || batch > max_batch) {
        set_error("batch must be in [1, max_batch]");
        return RNB_ERR_INVALID;
    }
    const size_t img_elems = 3ull * image * image;
    if (!host_x_dev) {
        RNB_CUDA(cudaMalloc(&host_x_dev, 1ull * max_batch * img_elems * sizeof(float)));
        RNB_CUDA(cudaMalloc(&host_logits_dev, 1ull * max_batch * classes * sizeof(float)));
        RNB_CUDA(cudaMalloc(&host_top1_dev, 1ull * max_batch * sizeof(int32_t)));
    }
    // Pipeline: every piece's H2D copy is queued on copy_stream up front; the compute stream waits
    // per piece, so piece i+1 crosses PCIe while piece i runs. Pieces are at most 64 images (and at
    // most the engine's chunk): a 256-batch of FP32 input is 154 MB, i.e. as long on PCIe as the whole
    // forward pass, so overlapping the two matters more than the per-launch efficiency of big pieces.
    const char* hc_env = getenv("RNB_HOST_CHUNK");
    int hchunk = hc_env ? atoi(hc_env) : 64;
    if (hchunk <= 0 || hchunk > chunk) hchunk = chunk;
    const int nchunks = (batch + hchunk - 1) / hchunk;
    while (static_cast<int>(copy_events.size()) < nchunks) {
        cudaEvent_t e;
        RNB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        copy_events.push_back(e);
    }
    cudaStream_t compute = host_compute;
    if (2 * host_pack_for(x, batch, &host_stage, host_x_dev) >= batch) {
        // (all or nothing here: the pieces of this path are separate forward passes)
        // the host cores round piece c to BF16 while piece c-1 crosses PCIe and piece c-2 runs
        host_pack_last = 1;
        uint16_t* xd = reinterpret_cast<uint16_t*>(host_x_dev);
        int rc = RNB_OK;
        cudaError_t ce = cudaSuccess;
        HostPacker::instance().run(x, host_stage, batch * img_elems, hchunk * img_elems, [&](size_t first, size_t count) {
            if (rc != RNB_OK || ce != cudaSuccess) return;
            const int off = static_cast<int>(first / img_elems), n = static_cast<int>(count / img_elems);
            ce = cudaMemcpyAsync(xd + first, host_stage + first, count * sizeof(uint16_t), cudaMemcpyHostToDevice, copy_stream);
            if (ce == cudaSuccess) ce = cudaEventRecord(copy_events[off / hchunk], copy_stream);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(compute, copy_events[off / hchunk], 0);
            if (ce == cudaSuccess)
                rc = forward_bf16(xd + first, n, host_logits_dev + 1ull * off * classes, host_top1_dev + off, compute);
        });
        RNB_CUDA(ce);
        if (rc) return rc;
    } else {
        for (int c = 0; c < nchunks; ++c) {
            const int off = c * hchunk;
            const int n = std::min(hchunk, batch - off);
            RNB_CUDA(cudaMemcpyAsync(host_x_dev + off * img_elems, x + off * img_elems,
                                     n * img_elems * sizeof(float), cudaMemcpyHostToDevice, copy_stream));
            RNB_CUDA(cudaEventRecord(copy_events[c], copy_stream));
        }
        for (int c = 0; c < nchunks; ++c) {
            const int off = c * hchunk;
            const int n = std::min(hchunk, batch - off);
            RNB_CUDA(cudaStreamWaitEvent(compute, copy_events[c], 0));
            int r = forward(host_x_dev + off * img_elems, n, host_logits_dev + 1ull * off * classes,
                            host_top1_dev + off, compute);
            if (r) return r;
        }
    }
    if (logits)
        RNB_CUDA(cudaMemcpyAsync(logits, host_logits_dev, 1ull * batch * classes * sizeof(float),
                                 cudaMemcpyDeviceToHost, compute));
    if (top1)
        RNB_CUDA(cudaMemcpyAsync(top1, host_top1_dev, 1ull * batch * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, compute));
    RNB_CUDA(cudaStreamSynchronize(compute));
    return RNB_OK;
}

// Pipelined host path. submit_host() queues, without blocking, the H2D copy of one batch (copy
// stream), its forward pass (compute stream, after the copy) and the D2H copy of logits / top-1;
// wait_host() blocks until that slot's results are in the caller's host buffers. With two slots the
// PCIe transfer of batch i+1 (154 MB for 256 FP32 images — as long as the forward pass itself)
// overlaps the forward pass of batch i, and the full batch still runs as ONE chunk.
int Model::submit_host(int slot, const float* x, int batch, float* logits, int32_t* top1) {
    return submit_host_any(slot, x, false, batch, logits, top1);
}

int Model::submit_host_u8(int slot, const uint8_t* x, int batch, float* logits, int32_t* top1) {
    return submit_host_any(slot, x, true, batch, logits, top1);
}

int Model::submit_host_any(int slot, const void* x, bool u8, int batch, float* logits, int32_t* top1) {
    if (slot < 0 || slot > 1) {
        set_error("slot must be 0 or 1");
        return RNB_ERR_INVALID;
    }
    if (batch <= 0 || batch > max_batch || !x) {
        set_error("submit_host: bad batch or NULL input");
        return RNB_ERR_INVALID;
    }
    HostSlot& hs = slots[slot];
    if (hs.pending) {
        int r = wait_host(slot);
        if (r) return r;
    }
    const size_t img_elems = 3ull * image * image;
    if (!hs.x_dev) {
        RNB_CUDA(cudaMalloc(&hs.x_dev, 1ull * max_batch * img_elems * sizeof(float)));
        RNB_CUDA(cudaMalloc(&hs.logits_dev, 1ull * max_batch * classes * sizeof(float)));
        RNB_CUDA(cudaMalloc(&hs.top1_dev, 1ull * max_batch * sizeof(int32_t)));
        RNB_CUDA(cudaEventCreateWithFlags(&hs.copied, cudaEventDisableTiming));
        RNB_CUDA(cudaEventCreateWithFlags(&hs.done, cudaEventDisableTiming));
        RNB_CUDA(cudaEventCreateWithFlags(&hs.computed, cudaEventDisableTiming));
    }
    if (!pipe_compute) RNB_CUDA(cudaStreamCreateWithFlags(&pipe_compute, cudaStreamNonBlocking));
    if (!pipe_d2h) RNB_CUDA(cudaStreamCreateWithFlags(&pipe_d2h, cudaStreamNonBlocking));
    int r;
    const int nb = u8 ? 0 : host_pack_for(static_cast<const float*>(x), batch, &hs.stage, hs.x_dev);
    if (nb > 0) {
        // FP32 input of a BF16-stem model: the host cores round the first nb images to BF16 (bit for bit what the stem
        // would do) in 16-image pieces, each uploaded as soon as it is complete — half the bytes on PCIe — while the
        // link is busy with the other images as FP32 (queued first: they need no core). This call returns when the
        // last piece is queued; the GPU is busy with the other slot's batch meanwhile.
        const float* xf = static_cast<const float*>(x);
        uint16_t* xd = reinterpret_cast<uint16_t*>(hs.x_dev);   // BF16 images [0, nb) at the front of the buffer; the FP32
        uint16_t* stage = hs.stage;                             // images behind them keep their FP32 positions
        if (nb < batch)
            RNB_CUDA(cudaMemcpyAsync(hs.x_dev + nb * img_elems, xf + nb * img_elems, (batch - nb) * img_elems * sizeof(float),
                                     cudaMemcpyHostToDevice, copy_stream));
        cudaError_t ce = cudaSuccess;
        HostPacker::instance().run(xf, stage, nb * img_elems, 16 * img_elems, [&](size_t first, size_t count) {
            if (ce == cudaSuccess)
                ce = cudaMemcpyAsync(xd + first, stage + first, count * sizeof(uint16_t), cudaMemcpyHostToDevice, copy_stream);
        });
        RNB_CUDA(ce);
        RNB_CUDA(cudaEventRecord(hs.copied, copy_stream));
        RNB_CUDA(cudaStreamWaitEvent(pipe_compute, hs.copied, 0));
        r = forward_any(InRef{xd, InRef::BF16_NCHW, hs.x_dev, nb}, batch, hs.logits_dev, hs.top1_dev, pipe_compute);
    } else {
        // uint8 input: a quarter of the bytes cross PCIe (the device buffer is reused as a byte buffer)
        RNB_CUDA(cudaMemcpyAsync(hs.x_dev, x, batch * img_elems * (u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice,
                                 copy_stream));
        RNB_CUDA(cudaEventRecord(hs.copied, copy_stream));
        RNB_CUDA(cudaStreamWaitEvent(pipe_compute, hs.copied, 0));
        r = u8 ? forward_u8(reinterpret_cast<const uint8_t*>(hs.x_dev), batch, hs.logits_dev, hs.top1_dev, pipe_compute)
               : forward(hs.x_dev, batch, hs.logits_dev, hs.top1_dev, pipe_compute);
    }
    if (r) return r;
    // results go back on their own stream: the other slot's forward pass, queued behind this one on pipe_compute,
    // starts as soon as this one ends instead of after the 1 MB D2H copy
    RNB_CUDA(cudaEventRecord(hs.computed, pipe_compute));
    RNB_CUDA(cudaStreamWaitEvent(pipe_d2h, hs.computed, 0));
    if (logits)
        RNB_CUDA(cudaMemcpyAsync(logits, hs.logits_dev, 1ull * batch * classes * sizeof(float),
                                 cudaMemcpyDeviceToHost, pipe_d2h));
    if (top1)
        RNB_CUDA(cudaMemcpyAsync(top1, hs.top1_dev, 1ull * batch * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                 pipe_d2h));
    RNB_CUDA(cudaEventRecord(hs.done, pipe_d2h));
    hs.pending = true;
    return RNB_OK;
}

int Model::wait_host(int slot) {
    if (slot < 0 || slot > 1) {
        set_error("slot must be 0 or 1");
        return RNB_ERR_INVALID;
    }
    HostSlot& hs = slots[slot];
    if (!hs.pending) return RNB_OK;
    RNB_CUDA(cudaEventSynchronize(hs.done));
    hs.pending = false;
    return RNB_OK;
}

}  // namespace rnb
