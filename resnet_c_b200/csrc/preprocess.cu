// preprocess.cu — resize 256 + centre-crop 224 of DECODED uint8 images on the GPU: the step before the hot path
// (/root/reference/convert_imgs_to_bin.py:12,18 — torchvision's ImageClassification preset applied to a PIL image,
// i.e. Pillow's antialiased bilinear ImagingResample followed by F.center_crop; the /255 + mean/std half of the preset
// is already fused into the stem pre-pass behind rnb_model_forward_u8). SURVEY.md section 8 f1.
//
// Integer work, bit-exact by construction. The arithmetic restated here is Pillow's src/libImaging/Resample.c
// (third-party, not vendored in the reference; oracle/preprocess.py::resize_crop_u8 is the CPU restatement and is
// pinned against Pillow itself):
//   host   per axis, per output index of the CROP WINDOW only: tap window and double-precision bilinear weights
//          (precompute_coeffs), converted to 22-bit fixed point with round-half-away (normalize_coeffs_8bpc)
//   device horizontal pass  h[y][x] = clamp((2^21 + sum_k in[y][x0 + k] * kx[k]) >> 22)   (uint8, like Pillow's
//          intermediate image), then vertical pass out[y][x] = clamp((2^21 + sum_k h[y0 + k][x] * ky[k]) >> 22)
// One CTA per (image, band of 8 output rows): the horizontal pass of the input rows the band needs goes to shared
// memory, the vertical pass reads it — the intermediate never touches HBM. HBM traffic per image = the decoded
// input read once (+ ~25 % re-read of band halos, L2 hits) + 150 KB written: an HBM-bound byte kernel.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/rnb.h"
#include "internal.h"

namespace rnb {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kBandRows = 8;

struct AxisCoeffs {
    int ksize = 0;
    std::vector<int> lo, n, kk;  // per output index of the crop window: first tap, tap count, ksize coefficients
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc (bilinear, box = whole axis), output indices [first, first + count)
AxisCoeffs axis_coeffs(int in_size, int out_size, int first, int count) {
    AxisCoeffs c;
    const double scale = static_cast<double>(in_size) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    c.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    c.lo.resize(count);
    c.n.resize(count);
    c.kk.assign(static_cast<size_t>(count) * c.ksize, 0);
    const double ss = 1.0 / filterscale;
    std::vector<double> w(c.ksize);
    for (int i = 0; i < count; ++i) {
        const int xx = first + i;
        if (in_size == out_size) {  // ImagingResample skips the pass: identity
            c.lo[i] = xx;
            c.n[i] = 1;
            c.kk[static_cast<size_t>(i) * c.ksize] = 1 << kPrecisionBits;
            continue;
        }
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w[x] = a < 1.0 ? 1.0 - a : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) w[x] /= ww;
            const double v = w[x] * (1 << kPrecisionBits);
            c.kk[static_cast<size_t>(i) * c.ksize + x] = w[x] < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
        }
        c.lo[i] = xmin;
        c.n[i] = xmax;
    }
    return c;
}

// Device-resident plan for one (H, W, resize, crop): coefficient tables of both axes.
struct ResizePlan {
    int H, W, crop, kx, ky, max_rows;  // max_rows: input rows a band of kBandRows output rows can span
    int* x_lo = nullptr;   // [crop]
    int* x_n = nullptr;
    int* x_k = nullptr;    // [crop][kx]
    int* y_lo = nullptr;
    int* y_n = nullptr;
    int* y_k = nullptr;    // [crop][ky]
};

std::mutex g_plan_mutex;
std::map<std::tuple<int, int, int, int, int>, ResizePlan> g_plans;  // (device, H, W, resize, crop)

int round_half_even_div2(int v) {  // Python: int(round(v / 2.0))
    return static_cast<int>(std::nearbyint(v / 2.0));
}

cudaError_t upload(int** dst, const std::vector<int>& src) {
    cudaError_t e = cudaMalloc(dst, src.size() * sizeof(int));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice);
}

// (the plan is returned BY VALUE: the cache may be flushed by another thread once the lock is released)
bool plan_for(int H, int W, int resize, int crop, ResizePlan* out, std::string* err) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    const auto key = std::make_tuple(dev, H, W, resize, crop);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        *out = it->second;
        return true;
    }
    if (g_plans.size() >= 256) {  // a caller cycling through arbitrary image sizes: start over rather than grow for ever
        cudaDeviceSynchronize();  // no launch may still read the tables
        for (auto& kv : g_plans) {
            ResizePlan& q = kv.second;
            cudaFree(q.x_lo); cudaFree(q.x_n); cudaFree(q.x_k); cudaFree(q.y_lo); cudaFree(q.y_n); cudaFree(q.y_k);
        }
        g_plans.clear();
    }
    // torchvision _compute_resized_output_size: short side -> resize, long side int(resize * long / short)
    const int shrt = W <= H ? W : H, lng = W <= H ? H : W;
    const int new_long = static_cast<int>(static_cast<double>(resize) * lng / shrt);
    const int nh = W <= H ? new_long : resize, nw = W <= H ? resize : new_long;
    if (nh < crop || nw < crop) {
        *err = "rnb_resize_crop_u8: the resized image is smaller than the crop";
        return false;
    }
    const int top = round_half_even_div2(nh - crop), left = round_half_even_div2(nw - crop);
    const AxisCoeffs cx = axis_coeffs(W, nw, left, crop), cy = axis_coeffs(H, nh, top, crop);
    ResizePlan p{};
    p.H = H; p.W = W; p.crop = crop; p.kx = cx.ksize; p.ky = cy.ksize;
    p.max_rows = 0;
    for (int b = 0; b < crop; b += kBandRows) {
        const int last = (b + kBandRows < crop ? b + kBandRows : crop) - 1;
        const int rows = cy.lo[last] + cy.n[last] - cy.lo[b];
        if (rows > p.max_rows) p.max_rows = rows;
    }
    if (upload(&p.x_lo, cx.lo) != cudaSuccess || upload(&p.x_n, cx.n) != cudaSuccess ||
        upload(&p.x_k, cx.kk) != cudaSuccess || upload(&p.y_lo, cy.lo) != cudaSuccess ||
        upload(&p.y_n, cy.n) != cudaSuccess || upload(&p.y_k, cy.kk) != cudaSuccess) {
        *err = std::string("rnb_resize_crop_u8: coefficient upload failed: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    g_plans.emplace(key, p);
    *out = p;
    return true;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
    const int v = acc >> kPrecisionBits;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// grid (bands, images); block 256. Shared: hbuf[max_rows][crop * 3] uint8.
__global__ void __launch_bounds__(256) resize_crop_u8_kernel(const uint8_t* __restrict__ in, size_t in_stride,
                                                             uint8_t* __restrict__ out, const ResizePlan p) {
    extern __shared__ uint8_t hbuf[];
    const int band0 = blockIdx.x * kBandRows;
    const int band_rows = min(kBandRows, p.crop - band0);
    const uint8_t* img = in + blockIdx.y * in_stride;
    uint8_t* dst = out + static_cast<size_t>(blockIdx.y) * p.crop * p.crop * 3;
    const int row0 = p.y_lo[band0];
    const int last = band0 + band_rows - 1;
    const int nrows = p.y_lo[last] + p.y_n[last] - row0;
    const int rowlen = p.crop * 3;
    // horizontal pass of input rows [row0, row0 + nrows) for the crop's columns
    for (int e = threadIdx.x; e < nrows * rowlen; e += blockDim.x) {
        const int r = e / rowlen, rem = e - r * rowlen;
        const int x = rem / 3, ch = rem - x * 3;
        const uint8_t* src = img + (static_cast<size_t>(row0 + r) * p.W + p.x_lo[x]) * 3 + ch;
        const int* k = p.x_k + x * p.kx;
        const int n = p.x_n[x];
        int acc = 1 << (kPrecisionBits - 1);
        for (int t = 0; t < n; ++t) acc += static_cast<int>(src[t * 3]) * k[t];
        hbuf[e] = clip8(acc);
    }
    __syncthreads();
    // vertical pass
    for (int e = threadIdx.x; e < band_rows * rowlen; e += blockDim.x) {
        const int r = e / rowlen, rem = e - r * rowlen;
        const int y = band0 + r;
        const uint8_t* src = hbuf + (p.y_lo[y] - row0) * rowlen + rem;
        const int* k = p.y_k + y * p.ky;
        const int n = p.y_n[y];
        int acc = 1 << (kPrecisionBits - 1);
        for (int t = 0; t < n; ++t) acc += static_cast<int>(src[t * rowlen]) * k[t];
        dst[static_cast<size_t>(y) * rowlen + rem] = clip8(acc);
    }
}

}  // namespace
}  // namespace rnb

using namespace rnb;

extern "C" int rnb_resize_crop_u8(const uint8_t* img_dev, int n_images, int H, int W, uint8_t* out_dev, int resize,
                                  int crop, void* stream) {
    if (!img_dev || !out_dev || n_images <= 0 || H <= 0 || W <= 0 || resize <= 0 || crop <= 0 || crop > resize) {
        set_error("rnb_resize_crop_u8: bad argument");
        return RNB_ERR_INVALID;
    }
    cudaPointerAttributes attr{};
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaPointerGetAttributes(&attr, img_dev) == cudaSuccess && attr.type == cudaMemoryTypeDevice && attr.device != prev)
        cudaSetDevice(attr.device);
    else
        cudaGetLastError();
    std::string err;
    ResizePlan plan{};
    const ResizePlan* p = plan_for(H, W, resize, crop, &plan, &err) ? &plan : nullptr;
    int rc = RNB_OK;
    if (!p) {
        set_error(err);
        rc = RNB_ERR_INVALID;
    } else {
        const size_t smem = static_cast<size_t>(p->max_rows) * crop * 3;
        cudaError_t e = cudaSuccess;
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(resize_crop_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e == cudaSuccess && smem > 227 * 1024) {
            set_error("rnb_resize_crop_u8: down-scaling factor too large for the shared-memory band");
            rc = RNB_ERR_UNSUPPORTED;
        } else if (e == cudaSuccess) {
            const dim3 grid((crop + kBandRows - 1) / kBandRows, n_images);
            resize_crop_u8_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
                img_dev, static_cast<size_t>(H) * W * 3, out_dev, *p);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) rc = fail_cuda(e, "rnb_resize_crop_u8");
    }
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    return rc;
}
