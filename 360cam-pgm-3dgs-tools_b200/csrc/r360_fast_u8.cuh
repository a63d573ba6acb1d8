// Packed-integer sampling of 8-bit 3-channel pixels from a staged shared-memory patch.
//
// Same arithmetic as the generic sampler (cv2.remap, DF:2001-2008) -- bit for bit -- but organised
// for instruction count: taps are fetched as aligned 32-bit words and funnel-shifted into place,
// channels are gathered with byte permutes, and the weighted sums run on dp4a / dp2a.
//
//   bilinear: weights (32-fx, fx) x (32-fy, fy) in units of 1/1024 (== cv2's 15-bit table / 32):
//             horizontal pass with dp4a (8-bit weights), vertical pass with two IMAD, (s+512)>>10
//   bicubic : cv2's 15-bit 4x4 table (shared-memory copy), rows of 4 taps with dp2a
//             (16-bit weight pairs x 8-bit pixels), (s+16384)>>15, saturate
#pragma once

#include "r360_common.cuh"

namespace r360 {

// Position of the (fy, fx) entry inside a plane of the shared-memory weight table.
__host__ __device__ __forceinline__ uint32_t table_entry_index(uint32_t fy, uint32_t fx) { return (fy << 5) + fx; }

// Shared-memory loads by 32-bit shared-window address, written as ordinary loads from the dynamic shared array so
// that the compiler may schedule them: ahead of the arithmetic of the previous pixel / frame is where the
// memory-level parallelism of the sampling loops comes from (`asm volatile` loads, shifts and dot products pin the
// program order: load, wait, compute, load, ...).  The mbarrier waits carry "memory" clobbers, so no load moves
// above the wait that guards its patch.
extern __shared__ __align__(128) unsigned char r360_dyn_smem[];
__device__ __forceinline__ const unsigned char* smem_at(uint32_t saddr) {
    return r360_dyn_smem + (saddr - (uint32_t)__cvta_generic_to_shared(r360_dyn_smem));
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) { return *reinterpret_cast<const uint32_t*>(smem_at(saddr)); }
__device__ __forceinline__ uint2 lds64(uint32_t saddr) { return *reinterpret_cast<const uint2*>(smem_at(saddr)); }
__device__ __forceinline__ uint4 lds128(uint32_t saddr) { return *reinterpret_cast<const uint4*>(smem_at(saddr)); }
// funnel shift / dp4a without the `volatile` of the toolkit's intrinsics (pure functions of their operands)
__device__ __forceinline__ uint32_t shf_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    uint32_t d;
    asm("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sh));
    return d;
}
__device__ __forceinline__ uint32_t dp4a_u8(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_s16_u8(uint32_t w_pair, uint32_t bytes, int acc) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w_pair), "r"(bytes), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp2a_hi_s16_u8(uint32_t w_pair, uint32_t bytes, int acc) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w_pair), "r"(bytes), "r"(acc));
    return d;
}

// float (holding 32 * coordinate, |value| < 2^22) -> integer bits: adding 1.5 * 2^23 leaves the
// round-half-even integer in the low mantissa bits (what cvRound does), on the full-rate FP32
// pipe instead of the conversion unit.  Result = 0x4B400000 + rint(v).
constexpr uint32_t kMagicBits = 0x4B400000u;
__device__ __forceinline__ uint32_t round_bits(float v) { return __float_as_uint(__fadd_rn(v, 12582912.0f)); }

// Address bias that turns (round_bits(x) >> 5, round_bits(y) >> 5) into a shared-memory byte
// address of the top-left tap: patch + (iy - py0) * pitch + ix * 3 - xb0, with the magic offsets
// folded in (32-bit wrap-around arithmetic).
__device__ __forceinline__ uint32_t patch_bias_u8c3(uint32_t patch_saddr, int pitch, int xb0, int py0) {
    const uint32_t m = kMagicBits >> 5;
    return patch_saddr - (uint32_t)xb0 - (uint32_t)(py0 * pitch) - 3u * m - (uint32_t)pitch * m;
}

// Nearest neighbour (cv2.remap INTER_NEAREST: the float32 map value rounded half to even).  `ux`, `uy` are
// round_bits() of the coordinate itself (not of 32 x), `bias_px` folds the patch origin and the magic offsets for
// `px_bytes` bytes per pixel.
__device__ __forceinline__ uint32_t patch_bias_nearest_u8(uint32_t patch_saddr, int pitch, int xb0, int py0, uint32_t px_bytes) {
    return patch_saddr - (uint32_t)xb0 - (uint32_t)(py0 * pitch) - px_bytes * kMagicBits - (uint32_t)pitch * kMagicBits;
}
__device__ __forceinline__ uint32_t nearest_u8c3(uint32_t bias_px, uint32_t pitch, uint32_t ux, uint32_t uy) {
    const uint32_t addr = ux * 3u + uy * pitch + bias_px;
    const uint32_t a4 = addr & ~3u;
    return shf_r(lds32(a4), lds32(a4 + 4), (addr & 3u) << 3);      // R | G << 8 | B << 16 | (next byte) << 24
}
__device__ __forceinline__ uint32_t nearest_u8c1(uint32_t bias_px, uint32_t pitch, uint32_t ux, uint32_t uy) {
    const uint32_t addr = ux + uy * pitch + bias_px;
    return (lds32(addr & ~3u) >> ((addr & 3u) << 3)) & 0xffu;
}

// Bilinear, split in two so that a tile's coordinates serve several frames of a batch: `prep` turns the
// quantised coordinate into the tap address and the weight words (frame independent), `taps` samples one frame.
struct BilinearPrep {
    uint32_t a4, sh;        // aligned address of the first tap word, bit shift of the pixel inside it
    uint32_t w_lo;          // bytes (32 - fx, fx, 0, 0)
    uint32_t fy;
};
__device__ __forceinline__ BilinearPrep bilinear_prep_u8c3(uint32_t bias, uint32_t pitch, uint32_t ux, uint32_t uy) {
    BilinearPrep p;
    const uint32_t fx = ux & 31u;
    const uint32_t addr = (ux >> 5) * 3u + (uy >> 5) * pitch + bias;
    p.a4 = addr & ~3u; p.sh = (addr & 3u) << 3;
    p.w_lo = 32u + 255u * fx;
    p.fy = uy & 31u;
    return p;
}
// One bilinear sample of the frame whose patch starts `frame_off` bytes further; returns R | G << 8 | B << 16.
__device__ __forceinline__ uint32_t bilinear_taps_u8c3(const BilinearPrep& p, uint32_t pitch, uint32_t frame_off) {
    const uint32_t sh = p.sh;
    const uint32_t a4 = p.a4 + frame_off;
    const uint32_t t0 = lds32(a4), t1 = lds32(a4 + 4), t2 = lds32(a4 + 8);
    const uint32_t b0 = lds32(a4 + pitch), b1 = lds32(a4 + pitch + 4), b2 = lds32(a4 + pitch + 8);
    const uint32_t lo_t = shf_r(t0, t1, sh), hi_t = shf_r(t1, t2, sh);   // R0 G0 B0 R1 | G1 B1 . .
    const uint32_t lo_b = shf_r(b0, b1, sh), hi_b = shf_r(b1, b2, sh);
    const uint32_t rr = __byte_perm(lo_t, lo_b, 0x7430);      // R00 R01 R10 R11
    const uint32_t gb_t = __byte_perm(lo_t, hi_t, 0x5241);    // G00 G01 B00 B01
    const uint32_t gb_b = __byte_perm(lo_b, hi_b, 0x5241);    // G10 G11 B10 B11
    const uint32_t w_lo = p.w_lo;                             // bytes (32 - fx, fx, 0, 0)
    const uint32_t w_hi = w_lo << 16;                         // bytes (0, 0, 32 - fx, fx)
    const uint32_t fy = p.fy, wy0 = 32u - fy;
    const uint32_t r = (dp4a_u8(rr, w_lo, 0u) * wy0 + dp4a_u8(rr, w_hi, 0u) * fy + 512u) >> 10;
    const uint32_t g = (dp4a_u8(gb_t, w_lo, 0u) * wy0 + dp4a_u8(gb_b, w_lo, 0u) * fy + 512u) >> 10;
    const uint32_t b = (dp4a_u8(gb_t, w_hi, 0u) * wy0 + dp4a_u8(gb_b, w_hi, 0u) * fy + 512u) >> 10;
    return r | (g << 8) | (b << 16);
}
__device__ __forceinline__ uint32_t bilinear_u8c3(uint32_t bias, uint32_t pitch, uint32_t ux, uint32_t uy) {
    return bilinear_taps_u8c3(bilinear_prep_u8c3(bias, pitch, ux, uy), pitch, 0u);
}

// Bicubic.  `table_saddr` is the shared-memory copy of cv2's fixed-point table, split into two planes
// [ky/2][fy][fx][ky%2][kx] int16 (16-byte entries, 16 KB apart).  `bias` must address tap (ix - 1, iy - 1): pass
// the bilinear bias minus (3 + pitch).
struct BicubicPrep {
    uint32_t a4, sh;
    uint32_t w[8];          // tap row ky: (w0 | w1 << 16, w2 | w3 << 16) in w[2 ky], w[2 ky + 1]
};
__device__ __forceinline__ BicubicPrep bicubic_prep_u8c3(uint32_t bias, uint32_t pitch, uint32_t table_saddr,
                                                         uint32_t ux, uint32_t uy) {
    BicubicPrep p;
    const uint32_t fx = ux & 31u, fy = uy & 31u;
    const uint32_t addr = (ux >> 5) * 3u + (uy >> 5) * pitch + bias;
    p.a4 = addr & ~3u; p.sh = (addr & 3u) << 3;
    const uint32_t wt = table_saddr + table_entry_index(fy, fx) * 16u;
    const uint4 wa = lds128(wt), wb = lds128(wt + 16384u);      // plane of rows 0,1 | plane of rows 2,3
    p.w[0] = wa.x; p.w[1] = wa.y; p.w[2] = wa.z; p.w[3] = wa.w;
    p.w[4] = wb.x; p.w[5] = wb.y; p.w[6] = wb.z; p.w[7] = wb.w;
    return p;
}
// One bicubic sample of the frame whose patch starts `frame_off` bytes further; returns R | G << 8 | B << 16.
// Tap loads: a row of taps is 12 bytes at a byte address of any alignment, fetched as four aligned 32-bit words.  At
// the 2:1 minification of an 8K -> 1600 px view a warp's lanes are 6 bytes apart: 2 shared-memory wavefronts per
// load when the lanes sit in one patch row (tools/pipe_probe.cu: 2.0 cycles), 2.4 measured in the kernel where a
// warp's pixels drift over several patch rows.  Three 64-bit loads + selects move the same bytes in 6 cycles instead
// of 8 in one row but in 9 over three rows, and measured slower in the kernel (125 vs 140 Gpix/s): not used.
__device__ __forceinline__ uint32_t bicubic_taps_u8c3(const BicubicPrep& p, uint32_t pitch, uint32_t frame_off) {
    const uint32_t sh = p.sh;
    int r = 16384, g = 16384, b = 16384;
    uint32_t a4 = p.a4 + frame_off;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        const uint32_t q0 = lds32(a4), q1 = lds32(a4 + 4), q2 = lds32(a4 + 8), q3 = lds32(a4 + 12);
        a4 += pitch;
        const uint32_t p0 = shf_r(q0, q1, sh);       // R0 G0 B0 R1
        const uint32_t p1 = shf_r(q1, q2, sh);       // G1 B1 R2 G2
        const uint32_t p2 = shf_r(q2, q3, sh);       // B2 R3 G3 B3
        const uint32_t rr = __byte_perm(__byte_perm(p0, p1, 0x0630), p2, 0x5210);   // R0 R1 R2 R3
        const uint32_t gg = __byte_perm(__byte_perm(p0, p1, 0x0741), p2, 0x6210);   // G0 G1 G2 G3
        const uint32_t bb = __byte_perm(__byte_perm(p0, p1, 0x0052), p2, 0x7410);   // B0 B1 B2 B3
        r = dp2a_hi_s16_u8(p.w[2 * ky + 1], rr, dp2a_lo_s16_u8(p.w[2 * ky], rr, r));
        g = dp2a_hi_s16_u8(p.w[2 * ky + 1], gg, dp2a_lo_s16_u8(p.w[2 * ky], gg, g));
        b = dp2a_hi_s16_u8(p.w[2 * ky + 1], bb, dp2a_lo_s16_u8(p.w[2 * ky], bb, b));
    }
    // saturate to 0..255 and pack in two instructions (cvt.pack.sat: bytes {a, b} on top of the low half of c)
    uint32_t hi, out;                                  // d = (c << 16) | sat(a) << 8 | sat(b)
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(0), "r"(b >> 15), "r"(0));            // . . 0 B
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(g >> 15), "r"(r >> 15), "r"(hi));    // 0 B G R
    return out;
}
__device__ __forceinline__ uint32_t bicubic_u8c3(uint32_t bias, uint32_t pitch, uint32_t table_saddr,
                                                 uint32_t ux, uint32_t uy) {
    return bicubic_taps_u8c3(bicubic_prep_u8c3(bias, pitch, table_saddr, ux, uy), pitch, 0u);
}

// One lanczos4 sample (cv2's 8 x 8 kernel, 15-bit table [fy][fx][ky][kx]: 128 bytes per pixel, read through L1 / L2
// -- it is four times the bicubic table and does not fit beside the patch ring).  `bias` must address tap
// (ix - 3, iy - 3): the bilinear bias minus (9 + 3 * pitch).  Returns R | G << 8 | B << 16.
template <bool SMEM_TABLE>
__device__ __forceinline__ uint32_t lanczos4_u8c3(uint32_t bias, uint32_t pitch, const short* wtab, uint32_t table_saddr,
                                                  uint32_t ux, uint32_t uy) {
    const uint32_t fx = ux & 31u, fy = uy & 31u;
    const uint32_t addr = (ux >> 5) * 3u + (uy >> 5) * pitch + bias;
    uint32_t a4 = addr & ~3u;
    const uint32_t sh = (addr & 3u) << 3;
    const uint32_t entry = (fy << 5) + fx;
    const uint4* w = reinterpret_cast<const uint4*>(wtab + entry * 64u);               // 8 rows x (w0|w1, w2|w3, w4|w5, w6|w7)
    const uint32_t wt = table_saddr + entry * 128u;                                     // shared-memory copy: row ky in slot (ky + entry) mod 8
    int r = 16384, g = 16384, b = 16384;
#pragma unroll 2
    for (int ky = 0; ky < 8; ++ky) {
        uint4 wr;
        if constexpr (SMEM_TABLE) wr = lds128(wt + (((uint32_t)ky + entry) & 7u) * 16u);
        else wr = __ldg(w + ky);
        const uint32_t q0 = lds32(a4), q1 = lds32(a4 + 4), q2 = lds32(a4 + 8), q3 = lds32(a4 + 12), q4 = lds32(a4 + 16),
                       q5 = lds32(a4 + 20), q6 = lds32(a4 + 24);
        a4 += pitch;
        const uint32_t p0 = shf_r(q0, q1, sh), p1 = shf_r(q1, q2, sh), p2 = shf_r(q2, q3, sh),
                       p3 = shf_r(q3, q4, sh), p4 = shf_r(q4, q5, sh), p5 = shf_r(q5, q6, sh);
        // (p0 p1 p2) hold pixels 0..3 and (p3 p4 p5) pixels 4..7, both as R G B R | G B R G | B R G B
        const uint32_t r_lo = __byte_perm(__byte_perm(p0, p1, 0x0630), p2, 0x5210), r_hi = __byte_perm(__byte_perm(p3, p4, 0x0630), p5, 0x5210);
        const uint32_t g_lo = __byte_perm(__byte_perm(p0, p1, 0x0741), p2, 0x6210), g_hi = __byte_perm(__byte_perm(p3, p4, 0x0741), p5, 0x6210);
        const uint32_t b_lo = __byte_perm(__byte_perm(p0, p1, 0x0052), p2, 0x7410), b_hi = __byte_perm(__byte_perm(p3, p4, 0x0052), p5, 0x7410);
        r = dp2a_hi_s16_u8(wr.w, r_hi, dp2a_lo_s16_u8(wr.z, r_hi, dp2a_hi_s16_u8(wr.y, r_lo, dp2a_lo_s16_u8(wr.x, r_lo, r))));
        g = dp2a_hi_s16_u8(wr.w, g_hi, dp2a_lo_s16_u8(wr.z, g_hi, dp2a_hi_s16_u8(wr.y, g_lo, dp2a_lo_s16_u8(wr.x, g_lo, g))));
        b = dp2a_hi_s16_u8(wr.w, b_hi, dp2a_lo_s16_u8(wr.z, b_hi, dp2a_hi_s16_u8(wr.y, b_lo, dp2a_lo_s16_u8(wr.x, b_lo, b))));
    }
    uint32_t hi, out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(0), "r"(b >> 15), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(g >> 15), "r"(r >> 15), "r"(hi));
    return out;
}

// Four RGB pixels (each R | G<<8 | B<<16) -> three packed words = 12 output bytes.
__device__ __forceinline__ void pack4_rgb(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3,
                                          uint32_t& w0, uint32_t& w1, uint32_t& w2) {
    w0 = __byte_perm(p0, p1, 0x4210);     // R0 G0 B0 R1
    w1 = __byte_perm(p1, p2, 0x5421);     // G1 B1 R2 G2
    w2 = __byte_perm(p2, p3, 0x6542);     // B2 R3 G3 B3
}

}  // namespace r360
