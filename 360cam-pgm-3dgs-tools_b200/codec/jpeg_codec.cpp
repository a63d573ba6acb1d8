// libr360codec.so: the C ABI of include/remap360_codec.h over NVIDIA nvJPEG (library code).
// Host-side glue only -- there is no kernel of ours in this file.
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

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
    nvjpegStatus_t st = nvjpegCreateEx(backend, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &c->handle);
    if (st != NVJPEG_STATUS_SUCCESS && backend != NVJPEG_BACKEND_GPU_HYBRID)
        st = nvjpegCreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &c->handle);
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
        R360_NVJ(nvjpegEncoderParamsSetOptimizedHuffman(c->params, 1, s));
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
