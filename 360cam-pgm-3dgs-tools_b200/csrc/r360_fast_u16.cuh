// Float-weight sampling of 16-bit 3-channel pixels from a staged shared-memory patch
// (BASELINE config 4: 8K uint16 frames, bicubic, uint16 / fp16 out).
//
// Same arithmetic as the generic sampler (cv2.remap on CV_16UC3: float32 weights wy[k] * wx[k], every
// product and sum rounded on its own, rows summed left to right) -- bit for bit -- but organised for
// instruction count: a row of taps is fetched as aligned 64-bit words instead of one 16-bit load per
// channel, the pixel's byte offset inside the first word (0, 2, 4 or 6) is removed with selects and
// funnel shifts, and uint16 -> float is a byte permute into the mantissa of 2^23 followed by one
// subtraction (exact, and on the FP32 pipe rather than the conversion unit).
#pragma once

#include "r360_fast_u8.cuh"

namespace r360 {

// low / high half of a word as float: (0x4B000000 | h) is 2^23 + h exactly
__device__ __forceinline__ float u16lo_f(uint32_t w) { return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)), 8388608.0f); }
__device__ __forceinline__ float u16hi_f(uint32_t w) { return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)), 8388608.0f); }

// Address bias for (round_bits(x) >> 5, round_bits(y) >> 5) -> byte address of the top-left tap of a
// uint16 x 3 patch (6 bytes per pixel); see patch_bias_u8c3.
__device__ __forceinline__ uint32_t patch_bias_u16c3(uint32_t patch_saddr, int pitch, int xb0, int py0) {
    const uint32_t m = kMagicBits >> 5;
    return patch_saddr - (uint32_t)xb0 - (uint32_t)(py0 * pitch) - 6u * m - (uint32_t)pitch * m;
}

// NW words starting at the 8-byte boundary below `addr`, shifted so that out[0] starts at `addr`
// (addr is even).  NW = 6: four pixels (24 bytes); NW = 3: two pixels (12 bytes).
template <int NW>
__device__ __forceinline__ void load_row_u16(uint32_t addr, uint32_t* out) {
    constexpr int NL = (NW + 2 + 1) / 2;                 // 64-bit loads covering NW + 2 words
    uint32_t w[2 * NL];
    const uint32_t a8 = addr & ~7u;
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        const uint2 v = lds64(a8 + 8 * k);
        w[2 * k] = v.x; w[2 * k + 1] = v.y;
    }
    const bool odd = (addr & 4u) != 0;
    const uint32_t sh = (addr & 2u) << 3;                 // 0 or 16 bits
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint32_t lo = odd ? w[k + 1] : w[k], hi = odd ? w[k + 2] : w[k + 1];
        out[k] = shf_r(lo, hi, sh);
    }
}

// Bicubic, split like the 8-bit samplers: `prep` fetches what depends on the coordinate only (tap address, the two
// rows of the 1-D cubic table [32][4] -- global, L1-resident), `taps` samples one frame of the item.  `bias` must
// address tap (ix - 1, iy - 1): the patch bias minus (6 + pitch).
struct BicubicPrepU16 {
    uint32_t addr;
    float wx[4], wy[4];
};
__device__ __forceinline__ BicubicPrepU16 bicubic_prep_u16c3(uint32_t bias, uint32_t pitch, const float* wtab,
                                                             uint32_t ux, uint32_t uy) {
    BicubicPrepU16 p;
    const uint32_t fx = ux & 31u, fy = uy & 31u;
    p.addr = (ux >> 5) * 6u + (uy >> 5) * pitch + bias;
    const float4 wx = __ldg(reinterpret_cast<const float4*>(wtab) + fx);
    const float4 wy = __ldg(reinterpret_cast<const float4*>(wtab) + fy);
    p.wx[0] = wx.x; p.wx[1] = wx.y; p.wx[2] = wx.z; p.wx[3] = wx.w;
    p.wy[0] = wy.x; p.wy[1] = wy.y; p.wy[2] = wy.z; p.wy[3] = wy.w;
    return p;
}
// Three float accumulators (R/G/B in memory order) for the frame whose patch starts `frame_off` bytes further.
__device__ __forceinline__ void bicubic_taps_u16c3(const BicubicPrepU16& p, uint32_t pitch, uint32_t frame_off, float* acc) {
    uint32_t addr = p.addr + frame_off;
    acc[0] = acc[1] = acc[2] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        uint32_t v[6];                                    // (c0 c1)(c2 c0)(c1 c2)(c0 c1)(c2 c0)(c1 c2)
        load_row_u16<6>(addr, v);
        addr += pitch;
        const float t[12] = {u16lo_f(v[0]), u16hi_f(v[0]), u16lo_f(v[1]), u16hi_f(v[1]), u16lo_f(v[2]), u16hi_f(v[2]),
                             u16lo_f(v[3]), u16hi_f(v[3]), u16lo_f(v[4]), u16hi_f(v[4]), u16lo_f(v[5]), u16hi_f(v[5])};
        float row[3];
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            const float wgt = __fmul_rn(p.wy[ky], p.wx[kx]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float term = __fmul_rn(t[3 * kx + c], wgt);
                row[c] = kx == 0 ? term : __fadd_rn(row[c], term);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], row[c]);
    }
}

// Bilinear (cv2: one left-to-right expression over the four taps).
struct BilinearPrepU16 {
    uint32_t addr;
    float tx, ty;
};
__device__ __forceinline__ BilinearPrepU16 bilinear_prep_u16c3(uint32_t bias, uint32_t pitch, uint32_t ux, uint32_t uy) {
    BilinearPrepU16 p;
    const uint32_t fx = ux & 31u, fy = uy & 31u;
    p.addr = (ux >> 5) * 6u + (uy >> 5) * pitch + bias;
    p.tx = (float)fx * (1.0f / 32.0f); p.ty = (float)fy * (1.0f / 32.0f);
    return p;
}
__device__ __forceinline__ void bilinear_taps_u16c3(const BilinearPrepU16& p, uint32_t pitch, uint32_t frame_off, float* acc) {
    const uint32_t addr = p.addr + frame_off;
    const float wxs[2] = {1.0f - p.tx, p.tx}, wys[2] = {1.0f - p.ty, p.ty};
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        uint32_t v[3];
        load_row_u16<3>(addr + ky * pitch, v);
        const float t[6] = {u16lo_f(v[0]), u16hi_f(v[0]), u16lo_f(v[1]), u16hi_f(v[1]), u16lo_f(v[2]), u16hi_f(v[2])};
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
            const float wgt = __fmul_rn(wys[ky], wxs[kx]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float term = __fmul_rn(t[3 * kx + c], wgt);
                acc[c] = (ky == 0 && kx == 0) ? term : __fadd_rn(acc[c], term);
            }
        }
    }
}

}  // namespace r360
