/*
 * remap360_codec.h -- JPEG decode / encode next to the remap, on the GPU (nvJPEG).
 *
 * The reference pays for a JPEG decode of the source and a JPEG encode of every view on the CPU:
 * each ffmpeg process decodes the panorama again and writes `-c:v mjpeg -q:v 1 -pix_fmt yuvj444p`
 * (cli_tools/gs360_360PerspCut.py:317-339), and the dual-fisheye tool uses cv2.imread / cv2.imwrite
 * with IMWRITE_JPEG_QUALITY (cli_tools/gs360_DualFisheyeDistortionCalibration.py:735, :1826-1840).
 * With the frame already in HBM these two steps dominate a real run, so they are offered here as a
 * separate small library (libr360codec.so, library code: NVIDIA nvJPEG, linked statically) behind the
 * same plain-C conventions as remap360.h: device pointers + stream in, status codes out, nothing
 * thrown, nothing allocated on behalf of the caller except inside the opaque codec object.
 *
 * A codec object is NOT thread-safe: use one per host thread.  Images are 8-bit, interleaved,
 * 3 channels (B, G, R as cv2 / or R, G, B) or 1 channel; chroma is never subsampled on encode (4:4:4,
 * like the reference's yuvj444p / what cv2 writes at quality >= 90... see DESIGN.md).
 */
#ifndef REMAP360_CODEC_H
#define REMAP360_CODEC_H

#include <stddef.h>
#include <stdint.h>

#include "remap360.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { R360_E_CODEC = -6 };              /* nvJPEG reported an error: r360_codec_last_error() */

typedef struct r360_jpeg r360_jpeg;      /* opaque: nvJPEG handle, decoder / encoder states */

int         r360_jpeg_create(r360_jpeg** out);
void        r360_jpeg_destroy(r360_jpeg* codec);
const char* r360_codec_last_error(void);  /* thread-local text */

/* Size of a JPEG byte stream without decoding it (channels: 1 grey, 3 colour). */
int r360_jpeg_info(r360_jpeg* codec, const uint8_t* data, size_t size,
                   int32_t* width, int32_t* height, int32_t* channels);

/* Decode into image `index` of `dst` (U8, channels 3 or 1, any pitch).  `channel_order` as in
 * r360_apply_lut (R360_ORDER_BGR matches cv2.imread).  Asynchronous on `stream` after the Huffman
 * stage; the byte stream may be released when the call returns. */
int r360_jpeg_decode(r360_jpeg* codec, const uint8_t* data, size_t size,
                     const r360_images* dst, int32_t index, int32_t channel_order, void* stream);

/* Encode image `index` of `src` (device memory) at `quality` 1..100, 4:4:4, optimised Huffman
 * tables.  Returns the byte count through `size`; then r360_jpeg_retrieve copies the stream to
 * host memory (`capacity` >= *size).  Both calls synchronise `stream`. */
int r360_jpeg_encode(r360_jpeg* codec, const r360_images* src, int32_t index, int32_t quality,
                     int32_t channel_order, size_t* size, void* stream);
int r360_jpeg_retrieve(r360_jpeg* codec, uint8_t* out, size_t capacity, size_t* size, void* stream);

#ifdef __cplusplus
}
#endif
#endif
