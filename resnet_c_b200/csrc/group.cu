// group.cu — rnb_group_*: data-parallel replicas on several GPUs driven by ONE process through the C ABI
// (BASELINE.json north_star: "each rank runs a full replica on its slice, with a single
// gather of logits/top-1 over NVLink only at the end"). The reference has one device and B = 1
// (/root/reference/cuda/inference/main.cu:230); the sharding is new, the per-image semantics are main.cu:168-251.
//
// Replica r is a complete Model on devices[r] (own weights, arena, streams, CUDA graphs). The batch shards by image
// (contiguous slices, earlier replicas take the remainder — the rule of resnet_c_b200/dist.py::shard_bounds). There is
// no communication during the forward. The exchange at the end is fused into the last two kernels of every replica:
// with peer access enabled, the FC kernel's TMA store and the arg-max kernel's stores write their rows of logits /
// top-1 STRAIGHT INTO the gathering buffers on devices[0] over NVLink (a peer-mapped global address in the output
// tensor map) — no collective launch, no staging copy. RNB_GROUP_GATHER=copy (or no peer access between two devices)
// selects the fallback: results land in a local buffer and one cudaMemcpyPeerAsync per output moves them.
#include <algorithm>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/rnb.h"
#include "internal.h"
#include "model.h"

struct rnb_group {
    std::vector<rnb_model*> models;
    std::vector<int> devices;
    std::vector<cudaStream_t> streams;    // replica r's compute stream (on devices[r])
    std::vector<cudaEvent_t> done;        // replica r's forward (and fallback copy) has been enqueued up to here
    std::vector<char> direct;             // replica r stores straight into the root's buffers
    std::vector<float*> stage_logits;     // fallback staging on devices[r]
    std::vector<int32_t*> stage_top1;
    cudaStream_t root_stream = nullptr;   // on devices[0]: made to wait for every replica
    int max_batch = 0;                    // per device
    int classes = 0;
};

using namespace rnb;

#define GRP_CUDA(expr)                                        \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return fail_cuda(e__, #expr); \
    } while (0)

static void shard(int total, int world, int r, int* first, int* count) {
    const int base = total / world, extra = total % world;
    *first = r * base + std::min(r, extra);
    *count = base + (r < extra ? 1 : 0);
}

extern "C" {

int rnb_group_create(const char* arch, int dtype, const char* weights_dir, const int* devices, int n_devices,
                     int max_batch_per_device, rnb_group_t** out) {
    if (!arch || !weights_dir || !devices || !out || n_devices <= 0 || max_batch_per_device <= 0) {
        set_error("rnb_group_create: bad argument");
        return RNB_ERR_INVALID;
    }
    if (dtype == RNB_DTYPE_FP8) {
        set_error("rnb_group_create: the FP8 variant is single-model (its scales come from one calibration batch)");
        return RNB_ERR_UNSUPPORTED;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    std::unique_ptr<rnb_group> g(new rnb_group());
    g->max_batch = max_batch_per_device;
    const char* mode = getenv("RNB_GROUP_GATHER");
    const bool force_copy = mode && std::string(mode) == "copy";
    int rc = RNB_OK;
    for (int r = 0; r < n_devices && rc == RNB_OK; ++r) {
        const int dev = devices[r];
        if ((rc = rnb_init(dev)) != RNB_OK) break;  // also makes `dev` current
        rnb_model_t* m = nullptr;
        if ((rc = rnb_model_create(arch, dtype, weights_dir, max_batch_per_device, 0, &m)) != RNB_OK) break;
        // the replicas' host paths are fed one after the other by the calling thread: no host-side BF16 rounding
        // (it would serialise the replicas; RNB_HOST_PACK=1 still forces it)
        if (!(getenv("RNB_HOST_PACK") && atoi(getenv("RNB_HOST_PACK")) == 1)) m->impl.host_pack_mode = 0;
        g->models.push_back(m);
        g->devices.push_back(dev);
        cudaStream_t s = nullptr;
        cudaEvent_t ev = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            rc = fail_cuda(cudaGetLastError(), "rnb_group_create: stream / event");
            break;
        }
        g->streams.push_back(s);
        g->done.push_back(ev);
        bool direct = dev == devices[0];
        if (!direct && !force_copy) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev, devices[0]) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) direct = true;
                cudaGetLastError();
            }
        }
        g->direct.push_back(direct ? 1 : 0);
        float* sl = nullptr;
        int32_t* st = nullptr;
        g->classes = rnb_model_num_classes(m);
        if (!direct) {
            if (cudaMalloc(&sl, 1ull * max_batch_per_device * g->classes * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&st, 1ull * max_batch_per_device * sizeof(int32_t)) != cudaSuccess) {
                rc = fail_cuda(cudaGetLastError(), "rnb_group_create: staging buffers");
            }
        }
        g->stage_logits.push_back(sl);
        g->stage_top1.push_back(st);
    }
    if (rc == RNB_OK) {
        cudaSetDevice(devices[0]);
        if (cudaStreamCreateWithFlags(&g->root_stream, cudaStreamNonBlocking) != cudaSuccess)
            rc = fail_cuda(cudaGetLastError(), "rnb_group_create: root stream");
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != RNB_OK) {
        const std::string keep = rnb_last_error();
        rnb_group_destroy(g.release());
        set_error(keep);
        return rc;
    }
    *out = g.release();
    return RNB_OK;
}

int rnb_group_destroy(rnb_group_t* g) {
    if (!g) return RNB_OK;
    for (size_t r = 0; r < g->models.size(); ++r) {
        DeviceGuard guard(g->devices[r]);
        cudaDeviceSynchronize();
        if (r < g->streams.size() && g->streams[r]) cudaStreamDestroy(g->streams[r]);
        if (r < g->done.size() && g->done[r]) cudaEventDestroy(g->done[r]);
        if (r < g->stage_logits.size() && g->stage_logits[r]) cudaFree(g->stage_logits[r]);
        if (r < g->stage_top1.size() && g->stage_top1[r]) cudaFree(g->stage_top1[r]);
        rnb_model_destroy(g->models[r]);
    }
    if (g->root_stream && !g->devices.empty()) {
        DeviceGuard guard(g->devices[0]);
        cudaStreamDestroy(g->root_stream);
    }
    delete g;
    return RNB_OK;
}

int rnb_group_size(const rnb_group_t* g) { return g ? static_cast<int>(g->models.size()) : 0; }

rnb_model_t* rnb_group_model(rnb_group_t* g, int r) {
    return g && r >= 0 && r < static_cast<int>(g->models.size()) ? g->models[r] : nullptr;
}

int rnb_group_direct_stores(const rnb_group_t* g, int r) {
    return g && r >= 0 && r < static_cast<int>(g->direct.size()) ? g->direct[r] : 0;
}

int rnb_shard_bounds(int total, int world, int rank, int* first, int* count) {
    if (total < 0 || world <= 0 || rank < 0 || rank >= world || !first || !count) {
        set_error("rnb_shard_bounds: bad argument");
        return RNB_ERR_INVALID;
    }
    shard(total, world, rank, first, count);
    return RNB_OK;
}

int rnb_group_shard(const rnb_group_t* g, int batch, int r, int* first, int* count) {
    if (!g || batch < 0 || r < 0 || r >= static_cast<int>(g->models.size()) || !first || !count) {
        set_error("rnb_group_shard: bad argument");
        return RNB_ERR_INVALID;
    }
    shard(batch, static_cast<int>(g->models.size()), r, first, count);
    return RNB_OK;
}

int rnb_group_warmup(rnb_group_t* g, int batch) {
    if (!g) {
        set_error("rnb_group_warmup: NULL group");
        return RNB_ERR_INVALID;
    }
    const int world = static_cast<int>(g->models.size());
    for (int r = 0; r < world; ++r) {
        int first, count;
        shard(batch, world, r, &first, &count);
        if (count > 0) {
            int rc = rnb_model_warmup(g->models[r], count, 0);
            if (rc) return rc;
        }
    }
    return RNB_OK;
}

static int group_forward_any(rnb_group_t* g, const void* const* x_dev, bool u8, int batch, float* logits_root_dev,
                             int32_t* top1_root_dev, void* root_stream) {
    if (!g || !x_dev || batch <= 0) {
        set_error("rnb_group_forward: bad argument");
        return RNB_ERR_INVALID;
    }
    const int world = static_cast<int>(g->models.size());
    cudaStream_t root = root_stream ? static_cast<cudaStream_t>(root_stream) : g->root_stream;
    for (int r = 0; r < world; ++r) {
        int first, count;
        shard(batch, world, r, &first, &count);
        if (count == 0) continue;
        if (count > g->max_batch || !x_dev[r]) {
            set_error("rnb_group_forward: shard larger than max_batch_per_device, or NULL shard pointer");
            return RNB_ERR_INVALID;
        }
        DeviceGuard guard(g->devices[r]);
        float* lo = logits_root_dev ? logits_root_dev + 1ull * first * g->classes : nullptr;
        int32_t* to = top1_root_dev ? top1_root_dev + first : nullptr;
        float* l_dst = g->direct[r] ? lo : (lo ? g->stage_logits[r] : nullptr);
        int32_t* t_dst = g->direct[r] ? to : (to ? g->stage_top1[r] : nullptr);
        int rc = u8 ? rnb_model_forward_u8(g->models[r], static_cast<const uint8_t*>(x_dev[r]), count, l_dst, t_dst,
                                           g->streams[r])
                    : rnb_model_forward(g->models[r], static_cast<const float*>(x_dev[r]), count, l_dst, t_dst,
                                        g->streams[r]);
        if (rc) return rc;
        if (!g->direct[r]) {
            if (lo)
                GRP_CUDA(cudaMemcpyPeerAsync(lo, g->devices[0], l_dst, g->devices[r],
                                             1ull * count * g->classes * sizeof(float), g->streams[r]));
            if (to)
                GRP_CUDA(cudaMemcpyPeerAsync(to, g->devices[0], t_dst, g->devices[r], 1ull * count * sizeof(int32_t),
                                             g->streams[r]));
        }
        GRP_CUDA(cudaEventRecord(g->done[r], g->streams[r]));
    }
    // the root's stream continues once every replica's rows have been written (events work across devices)
    DeviceGuard guard(g->devices[0]);
    for (int r = 0; r < world; ++r) {
        int first, count;
        shard(batch, world, r, &first, &count);
        if (count > 0) GRP_CUDA(cudaStreamWaitEvent(root, g->done[r], 0));
    }
    return RNB_OK;
}

int rnb_group_forward(rnb_group_t* g, const float* const* x_dev, int batch, float* logits_root_dev,
                      int32_t* top1_root_dev, void* root_stream) {
    return group_forward_any(g, reinterpret_cast<const void* const*>(x_dev), false, batch, logits_root_dev,
                             top1_root_dev, root_stream);
}

int rnb_group_forward_u8(rnb_group_t* g, const uint8_t* const* x_dev, int batch, float* logits_root_dev,
                         int32_t* top1_root_dev, void* root_stream) {
    return group_forward_any(g, reinterpret_cast<const void* const*>(x_dev), true, batch, logits_root_dev,
                             top1_root_dev, root_stream);
}

int rnb_group_synchronize(rnb_group_t* g) {
    if (!g) return RNB_OK;
    for (size_t r = 0; r < g->models.size(); ++r) {
        DeviceGuard guard(g->devices[r]);
        GRP_CUDA(cudaStreamSynchronize(g->streams[r]));
    }
    DeviceGuard guard(g->devices[0]);
    GRP_CUDA(cudaStreamSynchronize(g->root_stream));
    return RNB_OK;
}

static int group_submit_any(rnb_group_t* g, int slot, const void* x_host, bool u8, int batch, float* logits_host,
                            int32_t* top1_host) {
    if (!g || !x_host || batch <= 0) {
        set_error("rnb_group_submit_host: bad argument");
        return RNB_ERR_INVALID;
    }
    const int world = static_cast<int>(g->models.size());
    const size_t img = 3ull * 224 * 224;
    for (int r = 0; r < world; ++r) {
        int first, count;
        shard(batch, world, r, &first, &count);
        if (count == 0) continue;
        // every replica copies ITS rows host->device, forwards, and copies ITS rows of the results back into the
        // caller's host buffers: the host buffer is the gather, no device-side exchange at all
        float* lo = logits_host ? logits_host + 1ull * first * g->classes : nullptr;
        int32_t* to = top1_host ? top1_host + first : nullptr;
        int rc = u8 ? rnb_model_submit_host_u8(g->models[r], slot, static_cast<const uint8_t*>(x_host) + first * img,
                                               count, lo, to)
                    : rnb_model_submit_host(g->models[r], slot, static_cast<const float*>(x_host) + first * img, count,
                                            lo, to);
        if (rc) return rc;
    }
    return RNB_OK;
}

int rnb_group_submit_host(rnb_group_t* g, int slot, const float* x_host, int batch, float* logits_host,
                          int32_t* top1_host) {
    return group_submit_any(g, slot, x_host, false, batch, logits_host, top1_host);
}

int rnb_group_submit_host_u8(rnb_group_t* g, int slot, const uint8_t* x_host, int batch, float* logits_host,
                             int32_t* top1_host) {
    return group_submit_any(g, slot, x_host, true, batch, logits_host, top1_host);
}

int rnb_group_wait_host(rnb_group_t* g, int slot) {
    if (!g) {
        set_error("rnb_group_wait_host: NULL group");
        return RNB_ERR_INVALID;
    }
    for (rnb_model* m : g->models) {
        int rc = rnb_model_wait_host(m, slot);
        if (rc) return rc;
    }
    return RNB_OK;
}

int rnb_group_forward_host(rnb_group_t* g, const float* x_host, int batch, float* logits_host, int32_t* top1_host) {
    int rc = rnb_group_submit_host(g, 0, x_host, batch, logits_host, top1_host);
    if (rc) return rc;
    return rnb_group_wait_host(g, 0);
}

}  // extern "C"
