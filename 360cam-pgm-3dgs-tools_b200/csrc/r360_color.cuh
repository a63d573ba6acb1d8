// Input colour pipeline of the dual-fisheye tool, the step right before the remap on every frame
// when a LUT is configured (gs360_DualFisheyeDistortionCalibration.py:684-725):
//   image -> float32 [0, 1] (DF:599-609) -> .cube 3-D LUT, trilinear in RGB (DF:625-681)
//         -> optionally Rec.709 OETF^-1 then sRGB OETF (DF:568-596) -> back to the image dtype
//            with round-half-even (DF:612-622).
// One thread per pixel, float32 throughout with the reference's operation order (every product and
// sum rounded on its own: no fused multiply-add); the LUT is stored as float4 per node so that a
// corner is one 16-byte load (it lives in L2: 33^3 nodes = 575 KB).  HBM-bound: reads and writes
// each pixel once.
#pragma once

#include "r360_common.cuh"

namespace r360 {

struct LutParams {
    ImageSetDev src, dst;
    int channels;            // >= 3; channels beyond the first three are copied (DF:720-724)
    int n_images;
    int size;                // LUT nodes per axis
    int to_srgb;             // 1: rec709_to_srgb after the LUT; 0: clip only ("passthrough")
    int rgb_order;           // 0: memory order is B, G, R (cv2.imread); 1: R, G, B
    float dmin[3], span[3];  // DOMAIN_MIN, DOMAIN_MAX - DOMAIN_MIN (R, G, B)
    const float4* table;     // [b][g][r] nodes, .xyz = output R, G, B
};

template <typename T> struct Unit01;
template <> struct Unit01<uint8_t> {
    static __device__ __forceinline__ float load(uint8_t v) { return __fdiv_rn((float)v, 255.0f); }
    static __device__ __forceinline__ uint8_t store(float v) { return (uint8_t)__float2int_rn(__fmul_rn(v, 255.0f)); }
};
template <> struct Unit01<uint16_t> {
    static __device__ __forceinline__ float load(uint16_t v) { return __fdiv_rn((float)v, 65535.0f); }
    static __device__ __forceinline__ uint16_t store(float v) { return (uint16_t)__float2int_rn(__fmul_rn(v, 65535.0f)); }
};
template <> struct Unit01<float> {
    static __device__ __forceinline__ float load(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
    static __device__ __forceinline__ float store(float v) { return v; }
};

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// DF:568-596 on one value
__device__ __forceinline__ float rec709_to_srgb_1(float v) {
    v = clip01(v);
    const float lin = v < 0.081f ? __fdiv_rn(v, 4.5f)
                                 : powf(__fdiv_rn(__fadd_rn(v, 0.099f), 1.099f), (float)(1.0 / 0.45));
    const float l = clip01(lin);
    const float enc = l <= 0.0031308f ? __fmul_rn(12.92f, l)
                                      : __fsub_rn(__fmul_rn(1.055f, powf(l, (float)(1.0 / 2.4))), 0.055f);
    return clip01(enc);
}

__device__ __forceinline__ float lerp_ref(float a, float b, float f) {       // a + (b - a) * f, three roundings
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), f));
}

template <typename T>
__global__ void __launch_bounds__(256) lut_kernel(const __grid_constant__ LutParams p) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= p.src.width) return;
    const T* s = reinterpret_cast<const T*>(p.src.data + (long long)n * p.src.image_stride + (long long)y * p.src.pitch) +
                 (long long)x * p.channels;
    T* d = reinterpret_cast<T*>(p.dst.data + (long long)n * p.dst.image_stride + (long long)y * p.dst.pitch) +
           (long long)x * p.channels;
    const int ir = p.rgb_order ? 0 : 2, ib = 2 - ir;
    const float in[3] = {Unit01<T>::load(s[ir]), Unit01<T>::load(s[1]), Unit01<T>::load(s[ib])};
    const float top = (float)(p.size - 1);
    int i0[3], i1[3];
    float fr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float coord = clip01(__fdiv_rn(__fsub_rn(in[c], p.dmin[c]), p.span[c]));
        const float pos = __fmul_rn(coord, top);
        const float fl = floorf(pos);
        i0[c] = (int)fl;
        i1[c] = min(i0[c] + 1, p.size - 1);
        fr[c] = __fsub_rn(pos, fl);
    }
    const int n1 = p.size, n2 = p.size * p.size;
    const float4 c000 = __ldg(p.table + i0[2] * n2 + i0[1] * n1 + i0[0]);
    const float4 c100 = __ldg(p.table + i0[2] * n2 + i0[1] * n1 + i1[0]);
    const float4 c010 = __ldg(p.table + i0[2] * n2 + i1[1] * n1 + i0[0]);
    const float4 c110 = __ldg(p.table + i0[2] * n2 + i1[1] * n1 + i1[0]);
    const float4 c001 = __ldg(p.table + i1[2] * n2 + i0[1] * n1 + i0[0]);
    const float4 c101 = __ldg(p.table + i1[2] * n2 + i0[1] * n1 + i1[0]);
    const float4 c011 = __ldg(p.table + i1[2] * n2 + i1[1] * n1 + i0[0]);
    const float4 c111 = __ldg(p.table + i1[2] * n2 + i1[1] * n1 + i1[0]);
    float out[3];
#define R360_TRI(m)                                                                                         \
    lerp_ref(lerp_ref(lerp_ref(c000.m, c100.m, fr[0]), lerp_ref(c010.m, c110.m, fr[0]), fr[1]),             \
             lerp_ref(lerp_ref(c001.m, c101.m, fr[0]), lerp_ref(c011.m, c111.m, fr[0]), fr[1]), fr[2])
    out[0] = R360_TRI(x); out[1] = R360_TRI(y); out[2] = R360_TRI(z);
#undef R360_TRI
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = p.to_srgb ? rec709_to_srgb_1(out[c]) : clip01(out[c]);
    d[ir] = Unit01<T>::store(out[0]);
    d[1] = Unit01<T>::store(out[1]);
    d[ib] = Unit01<T>::store(out[2]);
    for (int c = 3; c < p.channels; ++c) d[c] = s[c];
}

// ---- video colour step ---------------------------------------------------------------------------------
// The cutter's video branch puts `colorspace=iall=bt709:all=smpte170m[:trc=iec61966-2-1]` in front of the
// remap (gs360_360PerspCut.py:299-309; gs360_Video2Frames.py:462-464 after it).  In R'G'B' terms that filter
// is: decode the source transfer curve, change primaries with a 3x3 matrix on linear light, encode with
// the destination curve.  One thread per pixel, float32; the matrix comes from the host (float64
// chromaticity algebra, rounded once).
enum Trc : int { kTrcBt709 = 0, kTrcSrgb = 1, kTrcLinear = 2 };      // bt709 == smpte170m == bt2020 curve

struct ColorConvertParams {
    ImageSetDev src, dst;
    int channels, n_images;
    int in_trc, out_trc;
    int rgb_order;           // 0: memory order B, G, R; 1: R, G, B
    float m[9];              // linear RGB -> linear RGB, row-major
};

__device__ __forceinline__ float trc_decode(int trc, float v) {      // code value -> linear light (odd extension)
    const float a = fabsf(v);
    float l;
    if (trc == kTrcBt709) l = a < 0.081242858f ? a * (1.0f / 4.5f) : powf((a + 0.09929682f) * (1.0f / 1.09929682f), 1.0f / 0.45f);
    else if (trc == kTrcSrgb) l = a <= 0.04045f ? a * (1.0f / 12.92f) : powf((a + 0.055f) * (1.0f / 1.055f), 2.4f);
    else l = a;
    return copysignf(l, v);
}
__device__ __forceinline__ float trc_encode(int trc, float l) {
    const float a = fabsf(l);
    float v;
    if (trc == kTrcBt709) v = a < 0.018053968f ? 4.5f * a : 1.09929682f * powf(a, 0.45f) - 0.09929682f;
    else if (trc == kTrcSrgb) v = a <= 0.0031308f ? 12.92f * a : 1.055f * powf(a, 1.0f / 2.4f) - 0.055f;
    else v = a;
    return copysignf(v, l);
}

template <typename T>
__global__ void __launch_bounds__(256) color_convert_kernel(const __grid_constant__ ColorConvertParams p) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= p.src.width) return;
    const T* s = reinterpret_cast<const T*>(p.src.data + (long long)n * p.src.image_stride + (long long)y * p.src.pitch) +
                 (long long)x * p.channels;
    T* d = reinterpret_cast<T*>(p.dst.data + (long long)n * p.dst.image_stride + (long long)y * p.dst.pitch) +
           (long long)x * p.channels;
    const int ir = p.rgb_order ? 0 : 2, ib = 2 - ir;
    const float r = trc_decode(p.in_trc, Unit01<T>::load(s[ir])), g = trc_decode(p.in_trc, Unit01<T>::load(s[1])),
                b = trc_decode(p.in_trc, Unit01<T>::load(s[ib]));
    const float ro = fmaf(p.m[0], r, fmaf(p.m[1], g, p.m[2] * b));
    const float go = fmaf(p.m[3], r, fmaf(p.m[4], g, p.m[5] * b));
    const float bo = fmaf(p.m[6], r, fmaf(p.m[7], g, p.m[8] * b));
    d[ir] = Unit01<T>::store(clip01(trc_encode(p.out_trc, ro)));
    d[1] = Unit01<T>::store(clip01(trc_encode(p.out_trc, go)));
    d[ib] = Unit01<T>::store(clip01(trc_encode(p.out_trc, bo)));
    for (int c = 3; c < p.channels; ++c) d[c] = s[c];
}

}  // namespace r360
