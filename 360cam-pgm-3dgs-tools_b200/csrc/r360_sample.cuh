// Per-pixel sampling with cv2.remap's arithmetic (the call the reference makes at
// cli_tools/gs360_DualFisheyeDistortionCalibration.py:2001-2008), parameterised on where the
// taps come from (global memory with border rules, or a staged shared-memory patch).
//
//   float32 map value -> s = rint(m * 32); integer part s >> 5 (saturated to int16), fraction s & 31
//   uint8 : 15-bit fixed-point 2-D weights, (sum + 16384) >> 15, saturate
//   others: float32 weights wy[k]*wx[k], products and sums rounded one by one in cv2's order
#pragma once

#include <type_traits>

#include "r360_common.cuh"

namespace r360 {

__device__ __align__(16) WeightTables g_tables;

__device__ __forceinline__ int sat_short(int v) { return min(max(v, -32768), 32767); }

// ---- tap sources -------------------------------------------------------------------------------

// Global memory, panorama border: columns wrap at the seam, rows clamp at the poles.
template <typename TIn>
struct ErpGlobalTaps {
    const unsigned char* img; long long pitch; int w, h, channels;
    static constexpr bool kConstantBorder = false;
    __device__ __forceinline__ bool exists(int, int) const { return true; }
    // the same tap of NF frames `fstride` bytes apart (NF = 1: one frame)
    template <int NF>
    __device__ __forceinline__ void loadn(int x, int y, long long fstride, float (*out)[4]) const {
        if ((unsigned)x >= (unsigned)w) {            // the seam: rare, so the division stays off the common path
            x = x % w;
            if (x < 0) x += w;
        }
        y = min(max(y, 0), h - 1);
        const unsigned char* p = img + (long long)y * pitch + (long long)x * channels * (int)sizeof(TIn);
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < channels) out[f][c] = Elem<TIn>::to_float(__ldg(reinterpret_cast<const TIn*>(p + f * fstride) + c));
    }
};

// Global memory, cv2 BORDER_CONSTANT applied tap by tap.
template <typename TIn>
struct ConstBorderGlobalTaps {
    const unsigned char* img; long long pitch; int w, h, channels; float border;
    static constexpr bool kConstantBorder = true;
    __device__ __forceinline__ bool exists(int x, int y) const {
        return (unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h;
    }
    template <int NF>
    __device__ __forceinline__ void loadn(int x, int y, long long fstride, float (*out)[4]) const {
        if (exists(x, y)) {
            const unsigned char* p = img + (long long)y * pitch + (long long)x * channels * (int)sizeof(TIn);
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < channels) out[f][c] = Elem<TIn>::to_float(__ldg(reinterpret_cast<const TIn*>(p + f * fstride) + c));
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int c = 0; c < 4; ++c) out[f][c] = border;
        }
    }
};

// A staged patch in shared memory that already contains every tap (seam unwrapped, pole rows
// replicated): no border logic at all.  (x, y) are source pixel indices; x may be unwrapped.
template <typename TIn>
struct PatchTaps {
    const unsigned char* base;   // address of byte column xb0 of source row y0
    int pitch;                   // bytes between patch rows
    int xb0;                     // unwrapped source byte column of patch byte 0
    int y0;                      // source row (unclamped index) of patch row 0
    int channels;
    static constexpr bool kConstantBorder = false;
    __device__ __forceinline__ bool exists(int, int) const { return true; }
    template <int NF>
    __device__ __forceinline__ void loadn(int x, int y, long long fstride, float (*out)[4]) const {
        const unsigned char* p = base + (y - y0) * pitch + (x * channels * (int)sizeof(TIn) - xb0);
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < channels) out[f][c] = Elem<TIn>::to_float(reinterpret_cast<const TIn*>(p + f * fstride)[c]);
    }
};

// ---- the sampler ---------------------------------------------------------------------------------

// One output pixel of NF frames that share the map (a batch): the coordinate, the tap positions and the weights are
// the frames' common part; `src_fstride` (bytes) and `dst_fstride` (elements) separate the frames.  The arithmetic per
// frame is exactly the single-frame one, and the loads of the NF frames are independent of one another, which is what
// the latency-bound fallback kernel needs.
template <int INTERP, typename TIn, typename TOut, int NF, typename Taps>
__device__ __forceinline__ void sample_pixel_frames(const Taps& taps, long long src_fstride, int channels, int src_w, int src_h,
                                                    float border, float x32, float y32, TOut* dst, long long dst_fstride) {
    float t[NF][4];
    if constexpr (INTERP == kNearest) {
        taps.template loadn<NF>(sat_short(__float2int_rn(x32)), sat_short(__float2int_rn(y32)), src_fstride, t);
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < channels) dst[f * dst_fstride + c] = Finish<TIn, TOut>::run(t[f][c]);   // exact: every element fits a float
        return;
    } else {
        const int sx = __float2int_rn(x32 * 32.0f);
        const int sy = __float2int_rn(y32 * 32.0f);
        const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
        const int fx = sx & 31, fy = sy & 31;

        if constexpr (std::is_same<TIn, uint8_t>::value) {
            int acc[NF][4];
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[f][c] = 0;
            if constexpr (INTERP == kLinear) {
                // (32-fx)(32-fy) ... in units of 1/1024: cv2's 15-bit table divided by 32, exactly
                const int wx[2] = {32 - fx, fx}, wy[2] = {32 - fy, fy};
#pragma unroll
                for (int ky = 0; ky < 2; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 2; ++kx) {
                        taps.template loadn<NF>(ix + kx, iy + ky, src_fstride, t);
                        const int wgt = wx[kx] * wy[ky];
#pragma unroll
                        for (int f = 0; f < NF; ++f)
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[f][c] += wgt * (int)t[f][c];
                    }
#pragma unroll
                for (int f = 0; f < NF; ++f)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < channels) dst[f * dst_fstride + c] = Finish<uint8_t, TOut>::run((float)((acc[f][c] + 512) >> 10));
            } else {
                constexpr int K = INTERP == kCubic ? 4 : 8, OFF = K / 2 - 1;
                const short* wt = (INTERP == kCubic ? g_tables.cubic_fixed : g_tables.lanczos_fixed) + (fy * 32 + fx) * (K * K);
#pragma unroll(K == 4 ? 4 : 1)
                for (int ky = 0; ky < K; ++ky)
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        taps.template loadn<NF>(ix - OFF + kx, iy - OFF + ky, src_fstride, t);
                        const int wgt = wt[ky * K + kx];
#pragma unroll
                        for (int f = 0; f < NF; ++f)
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[f][c] += wgt * (int)t[f][c];
                    }
#pragma unroll
                for (int f = 0; f < NF; ++f)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < channels)
                            dst[f * dst_fstride + c] = Finish<uint8_t, TOut>::run((float)min(max((acc[f][c] + 16384) >> 15, 0), 255));
            }
        } else {
            float acc[NF][4];
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[f][c] = 0.f;
            if constexpr (INTERP == kLinear) {
                const float tx = (float)fx * (1.0f / 32.0f), ty = (float)fy * (1.0f / 32.0f);
                const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty};
#pragma unroll
                for (int ky = 0; ky < 2; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 2; ++kx) {
                        taps.template loadn<NF>(ix + kx, iy + ky, src_fstride, t);
                        const float wgt = __fmul_rn(wy[ky], wx[kx]);
#pragma unroll
                        for (int f = 0; f < NF; ++f)
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float term = __fmul_rn(t[f][c], wgt);
                                acc[f][c] = (ky == 0 && kx == 0) ? term : __fadd_rn(acc[f][c], term);
                            }
                    }
            } else {
                constexpr int K = INTERP == kCubic ? 4 : 8, OFF = K / 2 - 1;
                const float* tab = INTERP == kCubic ? g_tables.cubic_1d : g_tables.lanczos_1d;
                const float* wx = tab + K * fx;
                const float* wy = tab + K * fy;
                const int x0 = ix - OFF, y0 = iy - OFF;
                bool interior = true;
                if constexpr (Taps::kConstantBorder)
                    interior = x0 >= 0 && x0 < max(src_w - (K - 1), 0) && y0 >= 0 && y0 < max(src_h - (K - 1), 0);
                if (interior) {
                    // each row summed left to right, rows added to a running sum that starts at zero
#pragma unroll(K == 4 ? 4 : 1)
                    for (int ky = 0; ky < K; ++ky) {
                        float row[NF][4];
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            taps.template loadn<NF>(x0 + kx, y0 + ky, src_fstride, t);
                            const float wgt = __fmul_rn(wy[ky], wx[kx]);
#pragma unroll
                            for (int f = 0; f < NF; ++f)
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const float term = __fmul_rn(t[f][c], wgt);
                                    row[f][c] = kx == 0 ? term : __fadd_rn(row[f][c], term);
                                }
                        }
#pragma unroll
                        for (int f = 0; f < NF; ++f)
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[f][c] = __fadd_rn(acc[f][c], row[f][c]);
                    }
                } else {
                    // near the sensor edge cv2 starts from the border value and adds
                    // (tap - border) * w for the taps that exist
#pragma unroll
                    for (int f = 0; f < NF; ++f)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[f][c] = border;
                    for (int ky = 0; ky < K; ++ky)
                        for (int kx = 0; kx < K; ++kx) {
                            if (!taps.exists(x0 + kx, y0 + ky)) continue;
                            taps.template loadn<NF>(x0 + kx, y0 + ky, src_fstride, t);
                            const float wgt = __fmul_rn(wy[ky], wx[kx]);
#pragma unroll
                            for (int f = 0; f < NF; ++f)
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    acc[f][c] = __fadd_rn(acc[f][c], __fmul_rn(__fsub_rn(t[f][c], border), wgt));
                        }
                }
            }
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < channels) dst[f * dst_fstride + c] = Finish<TIn, TOut>::run(acc[f][c]);
        }
    }
}

template <int INTERP, typename TIn, typename TOut, typename Taps>
__device__ __forceinline__ void sample_pixel(const Taps& taps, int channels, int src_w, int src_h, float border,
                                             float x32, float y32, TOut* dst) {
    sample_pixel_frames<INTERP, TIn, TOut, 1>(taps, 0, channels, src_w, src_h, border, x32, y32, dst, 0);
}

}  // namespace r360
