// resnet_infer_mgpu.cu — the multi-GPU driver: ONE C++ process, one replica per GPU through the C ABI
// (rnb_group_*, include/rnb.h). The reference's driver runs one device and B = 1
// (/root/reference/cuda/inference/main.cu:228-254); this is the same program shape (load weights_bin/, run, print the
// arg-max) for a sharded batch, and it times what bench.py times so that a C/C++ host reproduces the Python numbers:
//
//   resnet_infer_mgpu [arch=resnet50] [bf16|tf32] [gpus=all] [batch_per_gpu=256] [steps=20] [weights_dir=weights_bin]
//                     [image.bin]
//
// Step = every replica forwards its slice; the FC epilogue / arg-max of each replica store their rows straight into
// the gathering buffers on GPU 0 over NVLink (peer-mapped memory), so a step has no collective launch. Timing: CUDA
// events on GPU 0's stream around `steps` steps after 5 warm-up steps (device-resident inputs), then the same through
// host buffers (pinned; H2D + forward + D2H per step, two slots in flight). Prints one JSON line and "max index is N"
// for the first image of every GPU's slice.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "rnb.h"

#define CK(expr)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            fprintf(stderr, "%s failed: %s\n", #expr, cudaGetErrorString(e__));               \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)
#define RNB(expr)                                                                             \
    do {                                                                                      \
        int r__ = (expr);                                                                     \
        if (r__ != RNB_OK) {                                                                  \
            fprintf(stderr, "%s failed (%d): %s\n", #expr, r__, rnb_last_error());            \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)

int main(int argc, char** argv) {
    const std::string arch = argc > 1 ? argv[1] : "resnet50";
    const int dtype = (argc > 2 && !strcmp(argv[2], "tf32")) ? RNB_DTYPE_TF32 : RNB_DTYPE_BF16;
    int ngpu = 0;
    CK(cudaGetDeviceCount(&ngpu));
    if (argc > 3 && atoi(argv[3]) > 0) ngpu = atoi(argv[3]);   // may exceed the device count: replicas then share GPUs
    const int per_gpu = argc > 4 ? atoi(argv[4]) : 256;
    const int steps = argc > 5 ? atoi(argv[5]) : 20;
    const std::string weights_dir = argc > 6 ? argv[6] : "weights_bin";
    const char* image = argc > 7 ? argv[7] : nullptr;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    std::vector<int> devices(ngpu);
    for (int r = 0; r < ngpu; ++r) devices[r] = r % ndev;
    const int batch = per_gpu * ngpu;
    const size_t img = 3ull * 224 * 224;

    rnb_group_t* g = nullptr;
    RNB(rnb_group_create(arch.c_str(), dtype, weights_dir.c_str(), devices.data(), ngpu, per_gpu, &g));
    const int classes = rnb_model_num_classes(rnb_group_model(g, 0));

    // host input: the image file replicated, or a deterministic synthetic batch (LCG, roughly N(0,1) like the
    // normalised images); pinned
    float* x_host = nullptr;
    float *logits_host[2] = {nullptr, nullptr};
    int32_t* top1_host[2] = {nullptr, nullptr};
    CK(cudaMallocHost(&x_host, batch * img * sizeof(float)));
    for (int s = 0; s < 2; ++s) {
        CK(cudaMallocHost(&logits_host[s], 1ull * batch * classes * sizeof(float)));
        CK(cudaMallocHost(&top1_host[s], 1ull * batch * sizeof(int32_t)));
    }
    if (image) {
        FILE* f = fopen(image, "rb");
        if (!f || fread(x_host, sizeof(float), img, f) != img) {
            fprintf(stderr, "cannot read %s\n", image);
            return 4;
        }
        fclose(f);
        for (int b = 1; b < batch; ++b) memcpy(x_host + b * img, x_host, img * sizeof(float));
    } else {
        uint32_t s = 1234567u;
        for (size_t i = 0; i < batch * img; ++i) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) {
                s = s * 1664525u + 1013904223u;
                acc += static_cast<float>(s >> 8) * (1.0f / 16777216.0f);
            }
            x_host[i] = (acc - 2.0f) * 1.7320508f;  // sum of 4 uniforms: variance 1/3 -> scaled to 1
        }
    }

    // device-resident slices, gathering buffers on GPU 0
    std::vector<float*> x_dev(ngpu, nullptr);
    for (int r = 0; r < ngpu; ++r) {
        int first, count;
        RNB(rnb_group_shard(g, batch, r, &first, &count));
        CK(cudaSetDevice(devices[r]));
        CK(cudaMalloc(&x_dev[r], count * img * sizeof(float)));
        CK(cudaMemcpy(x_dev[r], x_host + first * img, count * img * sizeof(float), cudaMemcpyHostToDevice));
    }
    CK(cudaSetDevice(devices[0]));
    float* logits_dev = nullptr;
    int32_t* top1_dev = nullptr;
    CK(cudaMalloc(&logits_dev, 1ull * batch * classes * sizeof(float)));
    CK(cudaMalloc(&top1_dev, 1ull * batch * sizeof(int32_t)));
    cudaStream_t root;
    CK(cudaStreamCreateWithFlags(&root, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    RNB(rnb_group_warmup(g, batch));
    for (int i = 0; i < 5; ++i) RNB(rnb_group_forward(g, x_dev.data(), batch, logits_dev, top1_dev, root));
    RNB(rnb_group_synchronize(g));
    CK(cudaStreamSynchronize(root));
    // the timed region starts when every GPU is idle and ends when GPU 0 has seen every replica's last rows
    CK(cudaEventRecord(e0, root));
    for (int i = 0; i < steps; ++i) RNB(rnb_group_forward(g, x_dev.data(), batch, logits_dev, top1_dev, root));
    CK(cudaEventRecord(e1, root));
    CK(cudaEventSynchronize(e1));
    RNB(rnb_group_synchronize(g));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dev_ips = 1e3 * batch * steps / ms;
    std::vector<int32_t> top1_ref(batch);
    std::vector<float> logits_ref(1ull * batch * classes);
    CK(cudaMemcpy(top1_ref.data(), top1_dev, batch * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(logits_ref.data(), logits_dev, logits_ref.size() * sizeof(float), cudaMemcpyDeviceToHost));

    // end to end through host buffers, two slots in flight
    const int e2e_steps = steps < 10 ? steps : 10;
    for (int s = 0; s < 2; ++s) RNB(rnb_group_submit_host(g, s, x_host, batch, logits_host[s], top1_host[s]));
    for (int s = 0; s < 2; ++s) RNB(rnb_group_wait_host(g, s));
    const auto t0 = std::chrono::steady_clock::now();
    RNB(rnb_group_submit_host(g, 0, x_host, batch, logits_host[0], top1_host[0]));
    for (int i = 1; i < e2e_steps; ++i) {
        RNB(rnb_group_submit_host(g, i & 1, x_host, batch, logits_host[i & 1], top1_host[i & 1]));
        RNB(rnb_group_wait_host(g, (i - 1) & 1));
    }
    RNB(rnb_group_wait_host(g, (e2e_steps - 1) & 1));
    const double e2e_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double e2e_ips = 1.0 * batch * e2e_steps / e2e_s;

    // the gathered device result and the host-path result must agree bit for bit (same replicas, same kernels)
    size_t bad = 0;
    for (int b = 0; b < batch; ++b) bad += top1_ref[b] != top1_host[0][b];
    bad += memcmp(logits_ref.data(), logits_host[0], logits_ref.size() * sizeof(float)) != 0;
    int direct = 0;
    for (int r = 0; r < ngpu; ++r) direct += rnb_group_direct_stores(g, r);

    printf("{\"program\": \"resnet_infer_mgpu\", \"arch\": \"%s\", \"dtype\": \"%s\", \"n_gpus\": %d, \"per_gpu_batch\": %d, "
           "\"steps\": %d, \"ms_per_step\": %.4f, \"images_per_s\": %.1f, \"e2e_images_per_s\": %.1f, "
           "\"replicas_storing_directly_into_gpu0\": %d, \"device_path_equals_host_path\": %s}\n",
           arch.c_str(), dtype == RNB_DTYPE_TF32 ? "tf32" : "bf16", ngpu, per_gpu, steps, ms / steps, dev_ips, e2e_ips,
           direct, bad == 0 ? "true" : "false");
    for (int r = 0; r < ngpu; ++r) {
        int first, count;
        RNB(rnb_group_shard(g, batch, r, &first, &count));
        printf("gpu %d: max index is %d\n", devices[r], top1_ref[first]);
    }
    RNB(rnb_group_destroy(g));
    return bad == 0 ? 0 : 5;
}
