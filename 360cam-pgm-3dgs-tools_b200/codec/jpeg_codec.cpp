// libr360codec.so: the C ABI of include/remap360_codec.h over NVIDIA nvJPEG (library code).
// Host-side glue only -- there is no kernel of ours in this file.
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <unordered_map>

#include "remap360_codec.h"

namespace {

thread_local char tl_error[256] = "";

int fail(const char* what, int status) {
    std::snprintf(tl_error, sizeof(tl_error), "%s failed with status %d", what, status);
    return R360_E_CODEC;
}

#define R360_NVJ(call)                                                   \
    do {                                                                 \
        const nvjpegStatus_t st_ = (call);                               \
        if (st_ != NVJPEG_STATUS_SUCCESS) return fail(#call, (int)st_);  \
    } while (0)

// nvJPEG allocates and frees its device and pinned work buffers around every image through the allocator it was
// created with; the default one is cudaMalloc / cudaFree (and cudaHostAlloc), both of which synchronise the whole
// device -- with several host threads encoding at once every thread then waits for every other thread's kernels
// (measured on B200: 12 views of one panorama encode in 0.02 s on one thread, the same work takes 1.4 s per
// panorama with eight threads).  These caches hand freed blocks back out by size instead; blocks stay with the
// process (a panorama job runner works on same-sized images all day).
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> sizes;
    bool pinned = false;

    int take(void** out, size_t n) {
        if (!out) return 1;
        if (n == 0) n = 1;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = free_blocks.lower_bound(n);
            if (it != free_blocks.end() && it->first <= 2 * n + (1u << 20)) {
                *out = it->second;
                free_blocks.erase(it);
                return 0;
            }
        }
        const size_t rounded = (n + ((1u << 16) - 1)) & ~(size_t)((1u << 16) - 1);
        void* p = nullptr;
        const cudaError_t e = pinned ? cudaHostAlloc(&p, rounded, cudaHostAllocDefault) : cudaMalloc(&p, rounded);
        if (e != cudaSuccess) {
            // out of memory with blocks parked in the cache: release them and try once more
            drop_all();
            if ((pinned ? cudaHostAlloc(&p, rounded, cudaHostAllocDefault) : cudaMalloc(&p, rounded)) != cudaSuccess) return 1;
        }
        std::lock_guard<std::mutex> lock(mu);
        sizes[p] = rounded;
        *out = p;
        return 0;
    }
    int give(void* p) {
        if (!p) return 0;
        std::lock_guard<std::mutex> lock(mu);
        auto it = sizes.find(p);
        if (it == sizes.end()) return pinned ? (int)cudaFreeHost(p) : (int)cudaFree(p);
        free_blocks.emplace(it->second, p);
        return 0;
    }
    void drop_all() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto& kv : free_blocks) {
            if (pinned) cudaFreeHost(kv.second); else cudaFree(kv.second);
            sizes.erase(kv.second);
        }
        free_blocks.clear();
    }
};
// one cache per device would be needed if a process drove several devices through this library; the job runners use
// one process per device (remap360/multigpu.py), and blocks are only ever handed back to the device they came from
// because cudaMalloc'd pointers are device-specific and the cache is keyed by the current device below.
BlockCache& dev_cache() {
    static BlockCache caches[16];
    int dev = 0;
    cudaGetDevice(&dev);
    return caches[dev & 15];
}
BlockCache& pinned_cache() {
    static BlockCache c;
    c.pinned = true;
    return c;
}
int dev_malloc(void** p, size_t n) { return dev_cache().take(p, n); }
int dev_free(void* p) { return dev_cache().give(p); }
int pinned_malloc(void** p, size_t n, unsigned int) { return pinned_cache().take(p, n); }
int pinned_free(void* p) { return pinned_cache().give(p); }

bool usable(const r360_images* im, int32_t index) {
    return im && im->data && im->dtype == R360_U8 && (im->channels == 3 || im->channels == 1) &&
           im->width > 0 && im->height > 0 && index >= 0 && index < im->count &&
           im->pitch_bytes >= (int64_t)im->width * im->channels;
}

}  // namespace

struct r360_jpeg {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t decoder = nullptr;
    nvjpegEncoderState_t encoder = nullptr;
    nvjpegEncoderParams_t params = nullptr;
    int quality = -1;
    bool grey = false;
};

extern "C" {

const char* r360_codec_last_error(void) { return tl_error; }

int r360_jpeg_create(r360_jpeg** out) {
    if (!out) return R360_E_INVALID_ARG;
    *out = nullptr;
    r360_jpeg* c = new (std::nothrow) r360_jpeg();
    if (!c) return R360_E_INVALID_ARG;
    // GPU-assisted Huffman decode for large baseline images; falls back inside nvJPEG otherwise.
    // R360_JPEG_BACKEND=hardware asks for the NVJPG engine (baseline, single scan), =hybrid for CPU Huffman.
    nvjpegBackend_t backend = NVJPEG_BACKEND_GPU_HYBRID;
    if (const char* env = std::getenv("R360_JPEG_BACKEND")) {
        if (!std::strcmp(env, "hardware")) backend = NVJPEG_BACKEND_HARDWARE;
        else if (!std::strcmp(env, "hybrid")) backend = NVJPEG_BACKEND_HYBRID;
    }
    static nvjpegDevAllocator_t dev_alloc = {&dev_malloc, &dev_free};
    static nvjpegPinnedAllocator_t pin_alloc = {&pinned_malloc, &pinned_free};
    const bool cached = std::getenv("R360_JPEG_PLAIN_ALLOC") == nullptr;       // experiments: nvJPEG's own allocator
    nvjpegDevAllocator_t* da = cached ? &dev_alloc : nullptr;
    nvjpegPinnedAllocator_t* pa = cached ? &pin_alloc : nullptr;
    nvjpegStatus_t st = nvjpegCreateEx(backend, da, pa, NVJPEG_FLAGS_DEFAULT, &c->handle);
    if (st != NVJPEG_STATUS_SUCCESS && backend != NVJPEG_BACKEND_GPU_HYBRID)
        st = nvjpegCreateEx(NVJPEG_BACKEND_GPU_HYBRID, da, pa, NVJPEG_FLAGS_DEFAULT, &c->handle);
    if (st != NVJPEG_STATUS_SUCCESS) st = nvjpegCreateSimple(&c->handle);
    if (st != NVJPEG_STATUS_SUCCESS) { delete c; return fail("nvjpegCreate", (int)st); }
    if ((st = nvjpegJpegStateCreate(c->handle, &c->decoder)) != NVJPEG_STATUS_SUCCESS ||
        (st = nvjpegEncoderStateCreate(c->handle, &c->encoder, nullptr)) != NVJPEG_STATUS_SUCCESS ||
        (st = nvjpegEncoderParamsCreate(c->handle, &c->params, nullptr)) != NVJPEG_STATUS_SUCCESS) {
        r360_jpeg_destroy(c);
        return fail("nvjpeg state creation", (int)st);
    }
    *out = c;
    return R360_OK;
}

void r360_jpeg_destroy(r360_jpeg* c) {
    if (!c) return;
    if (c->params) nvjpegEncoderParamsDestroy(c->params);
    if (c->encoder) nvjpegEncoderStateDestroy(c->encoder);
    if (c->decoder) nvjpegJpegStateDestroy(c->decoder);
    if (c->handle) nvjpegDestroy(c->handle);
    delete c;
}

int r360_jpeg_info(r360_jpeg* c, const uint8_t* data, size_t size, int32_t* width, int32_t* height,
                   int32_t* channels) {
    if (!c || !data || !size) return R360_E_INVALID_ARG;
    int comps = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    R360_NVJ(nvjpegGetImageInfo(c->handle, data, size, &comps, &sub, w, h));
    if (width) *width = w[0];
    if (height) *height = h[0];
    if (channels) *channels = comps == 1 ? 1 : 3;
    return R360_OK;
}

int r360_jpeg_decode(r360_jpeg* c, const uint8_t* data, size_t size, const r360_images* dst, int32_t index,
                     int32_t channel_order, void* stream) {
    if (!c || !data || !size || !usable(dst, index)) return R360_E_INVALID_ARG;
    if (channel_order != R360_ORDER_BGR && channel_order != R360_ORDER_RGB) return R360_E_INVALID_ARG;
    int32_t w = 0, h = 0, ch = 0;
    const int rc = r360_jpeg_info(c, data, size, &w, &h, &ch);
    if (rc != R360_OK) return rc;
    if (w != dst->width || h != dst->height) return R360_E_INVALID_ARG;
    nvjpegImage_t out;
    std::memset(&out, 0, sizeof(out));
    out.channel[0] = static_cast<unsigned char*>(dst->data) + (int64_t)index * dst->image_stride_bytes;
    out.pitch[0] = (size_t)dst->pitch_bytes;
    const nvjpegOutputFormat_t fmt = dst->channels == 1 ? NVJPEG_OUTPUT_Y
                                    : channel_order == R360_ORDER_BGR ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI;
    R360_NVJ(nvjpegDecode(c->handle, c->decoder, data, size, fmt, &out, static_cast<cudaStream_t>(stream)));
    return R360_OK;
}

int r360_jpeg_encode(r360_jpeg* c, const r360_images* src, int32_t index, int32_t quality, int32_t channel_order,
                     size_t* size, void* stream) {
    if (!c || !size || !usable(src, index)) return R360_E_INVALID_ARG;
    if (channel_order != R360_ORDER_BGR && channel_order != R360_ORDER_RGB) return R360_E_INVALID_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int q = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const bool grey = src->channels == 1;
    if (q != c->quality || grey != c->grey) {
        R360_NVJ(nvjpegEncoderParamsSetQuality(c->params, q, s));
        // optimised Huffman tables (a second pass over the coefficients) unless R360_JPEG_HUFFMAN=default
        static const int optimise = []() { const char* e = std::getenv("R360_JPEG_HUFFMAN"); return (e && !std::strcmp(e, "default")) ? 0 : 1; }();
        R360_NVJ(nvjpegEncoderParamsSetOptimizedHuffman(c->params, optimise, s));
        R360_NVJ(nvjpegEncoderParamsSetSamplingFactors(c->params, grey ? NVJPEG_CSS_GRAY : NVJPEG_CSS_444, s));
        c->quality = q; c->grey = grey;
    }
    nvjpegImage_t in;
    std::memset(&in, 0, sizeof(in));
    in.channel[0] = static_cast<unsigned char*>(src->data) + (int64_t)index * src->image_stride_bytes;
    in.pitch[0] = (size_t)src->pitch_bytes;
    if (grey) {
        R360_NVJ(nvjpegEncodeYUV(c->handle, c->encoder, c->params, &in, NVJPEG_CSS_GRAY, src->width, src->height, s));
    } else {
        R360_NVJ(nvjpegEncodeImage(c->handle, c->encoder, c->params, &in,
                                   channel_order == R360_ORDER_BGR ? NVJPEG_INPUT_BGRI : NVJPEG_INPUT_RGBI,
                                   src->width, src->height, s));
    }
    R360_NVJ(nvjpegEncodeRetrieveBitstream(c->handle, c->encoder, nullptr, size, s));
    return R360_OK;
}

int r360_jpeg_retrieve(r360_jpeg* c, uint8_t* out, size_t capacity, size_t* size, void* stream) {
    if (!c || !out || !size) return R360_E_INVALID_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    size_t need = 0;
    R360_NVJ(nvjpegEncodeRetrieveBitstream(c->handle, c->encoder, nullptr, &need, s));
    if (need > capacity) return R360_E_INVALID_ARG;
    *size = need;
    R360_NVJ(nvjpegEncodeRetrieveBitstream(c->handle, c->encoder, out, size, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail("cudaStreamSynchronize", 0);
    return R360_OK;
}

}  // extern "C"
