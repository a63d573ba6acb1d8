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

// ---- lane-per-column samplers (the tiled kernel's full tiles) ----------------------------------------------
// Same arithmetic again, organised for the pipes of the SM (tools/pipe_probe.cu): with one pixel per lane, lanes
// 12 bytes apart at the 2:1 minification of an 8K -> 1600 px view, aligned 32-bit loads are bank-conflict free
// (stride of three words) where the 64-bit loads above take four wavefronts; a 16-bit sample never straddles a
// 32-bit word, so the low / high halves are taken with one byte permute against zero and no funnel shift; and
// uint16 -> float32 costs nothing: the zero-extended sample IS the float32 denormal v * 2^-149, the weights carry
// 2^126 (wy * wx <= 1.2 stays finite), products and sums are then the reference's scaled by 2^-23 bit for bit
// (scaling by powers of two commutes with rounding while nothing underflows: the smallest non-zero product is
// 2^-149 * 2^126 * 2^-15 > 2^-126), and one exact multiplication by 2^23 ends the accumulation.

// The seven aligned words that hold a row of four RGB pixels starting at byte `addr` (even), and the twelve
// samples in memory order as zero-extended words.  `hi_first` = addr & 2: the first sample sits in a high half.
__device__ __forceinline__ void load_row_u16_words(uint32_t addr, bool hi_first, uint32_t* s) {
    const uint32_t a4 = addr & ~3u;
    uint32_t w[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) w[k] = lds32(a4 + 4 * k);
    const uint32_t sel_even = hi_first ? 0x4432u : 0x4410u;      // sample 2k: half of word k chosen by the alignment
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        s[2 * k] = __byte_perm(w[k], 0u, sel_even);
        // sample 2k + 1: high half of word k, or low half of word k + 1
        s[2 * k + 1] = __byte_perm(hi_first ? w[k + 1] : w[k], 0u, hi_first ? 0x4410u : 0x4432u);
    }
}

struct BicubicColPrepU16 {
    uint32_t addr;
    float wgt[16];          // wy[ky] * wx[kx] * 2^126, the reference's float32 products scaled
};
__device__ __forceinline__ BicubicColPrepU16 bicubic_col_prep_u16c3(uint32_t bias, uint32_t pitch, const float* wtab,
                                                                   uint32_t ux, uint32_t uy) {
    BicubicColPrepU16 p;
    const uint32_t fx = ux & 31u, fy = uy & 31u;
    p.addr = (ux >> 5) * 6u + (uy >> 5) * pitch + bias;
    const float4 wx = __ldg(reinterpret_cast<const float4*>(wtab) + fx);
    const float4 wy = __ldg(reinterpret_cast<const float4*>(wtab) + fy);
    const float wxs[4] = {wx.x, wx.y, wx.z, wx.w}, wys[4] = {wy.x, wy.y, wy.z, wy.w};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
            p.wgt[ky * 4 + kx] = __fmul_rn(__fmul_rn(wys[ky], wxs[kx]), 8.507059173023462e37f);      // * 2^126, exact
    return p;
}
__device__ __forceinline__ void bicubic_col_taps_u16c3(const BicubicColPrepU16& p, uint32_t pitch, uint32_t frame_off, float* acc) {
    uint32_t addr = p.addr + frame_off;
    const bool hi_first = (p.addr & 2u) != 0;
    acc[0] = acc[1] = acc[2] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        uint32_t s[12];
        load_row_u16_words(addr, hi_first, s);
        addr += pitch;
        float row[3];
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float term = __fmul_rn(__uint_as_float(s[3 * kx + c]), p.wgt[ky * 4 + kx]);
                row[c] = kx == 0 ? term : __fadd_rn(row[c], term);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], row[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = __fmul_rn(acc[c], 8388608.0f);                               // * 2^23, exact
}

struct BilinearColPrepU16 {
    uint32_t addr;
    float wgt[4];
};
__device__ __forceinline__ BilinearColPrepU16 bilinear_col_prep_u16c3(uint32_t bias, uint32_t pitch, uint32_t ux, uint32_t uy) {
    BilinearColPrepU16 p;
    const uint32_t fx = ux & 31u, fy = uy & 31u;
    p.addr = (ux >> 5) * 6u + (uy >> 5) * pitch + bias;
    const float tx = (float)fx * (1.0f / 32.0f), ty = (float)fy * (1.0f / 32.0f);
    const float wxs[2] = {1.0f - tx, tx}, wys[2] = {1.0f - ty, ty};
#pragma unroll
    for (int ky = 0; ky < 2; ++ky)
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) p.wgt[ky * 2 + kx] = __fmul_rn(__fmul_rn(wys[ky], wxs[kx]), 8.507059173023462e37f);
    return p;
}
__device__ __forceinline__ void bilinear_col_taps_u16c3(const BilinearColPrepU16& p, uint32_t pitch, uint32_t frame_off, float* acc) {
    const uint32_t addr = p.addr + frame_off;
    const bool hi_first = (p.addr & 2u) != 0;
#pragma unroll
    for (int ky = 0; ky < 2; ++ky) {
        const uint32_t a4 = (addr + ky * pitch) & ~3u;
        const uint32_t w0 = lds32(a4), w1 = lds32(a4 + 4), w2 = lds32(a4 + 8), w3 = lds32(a4 + 12);
        const uint32_t sel_even = hi_first ? 0x4432u : 0x4410u, sel_odd = hi_first ? 0x4410u : 0x4432u;
        const uint32_t s[6] = {__byte_perm(w0, 0u, sel_even), __byte_perm(hi_first ? w1 : w0, 0u, sel_odd),
                               __byte_perm(w1, 0u, sel_even), __byte_perm(hi_first ? w2 : w1, 0u, sel_odd),
                               __byte_perm(w2, 0u, sel_even), __byte_perm(hi_first ? w3 : w2, 0u, sel_odd)};
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float term = __fmul_rn(__uint_as_float(s[3 * kx + c]), p.wgt[ky * 2 + kx]);
                acc[c] = (ky == 0 && kx == 0) ? term : __fadd_rn(acc[c], term);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = __fmul_rn(acc[c], 8388608.0f);
}

}  // namespace r360
