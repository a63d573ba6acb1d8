// Direct path: one thread per output pixel, float64 projection, taps gathered through L1/L2.
//
// This is the reference-grade device path: it handles every layout, border and projection the
// ABI accepts, and the tiled fast path falls back to it for tiles it cannot stage (poles,
// sensor edges).  Arithmetic per pixel:
//   ray (gs360_GUI.py:377-392) -> lon/lat -> ERP pixel (gs360_GUI.py:419-424)            [ERP]
//   ray -> equisolid + Brown -> sensor pixel, validity (DF:1794-1821)                     [fisheye]
//   float32 cast -> 1/32-px quantisation -> cv2.remap weights (DF:2001-2008)              [sampling]
#pragma once

#include <type_traits>

#include "r360_common.cuh"

namespace r360 {

__device__ WeightTables g_tables;

// ---- tap addressing ------------------------------------------------------------------------

template <int PROJ> struct Border;

// panorama: columns wrap at the seam, rows clamp at the poles
template <> struct Border<kProjErp> {
    static __device__ __forceinline__ bool resolve(int& x, int& y, int w, int h) {
        x = x % w;
        if (x < 0) x += w;
        y = min(max(y, 0), h - 1);
        return true;
    }
};
// fisheye: cv2 BORDER_CONSTANT, tap by tap
template <> struct Border<kProjFisheye> {
    static __device__ __forceinline__ bool resolve(int& x, int& y, int w, int h) {
        return (unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h;
    }
};

template <int PROJ, typename TIn>
__device__ __forceinline__ void load_tap(const unsigned char* img, long long pitch, int w, int h,
                                         int x, int y, int channels, float border, float* out) {
    if (Border<PROJ>::resolve(x, y, w, h)) {
        const TIn* p = reinterpret_cast<const TIn*>(img + (long long)y * pitch) + (long long)x * channels;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < channels) out[c] = Elem<TIn>::to_float(__ldg(p + c));
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) out[c] = border;
    }
}

__device__ __forceinline__ int sat_short(int v) { return min(max(v, -32768), 32767); }

// ---- per-pixel sampling (cv2.remap arithmetic) -------------------------------------------

template <int PROJ, int INTERP, typename TIn, typename TOut>
__device__ __forceinline__ void sample_pixel(const unsigned char* img, long long pitch, int w, int h,
                                             int channels, float border, float x32, float y32,
                                             TOut* dst) {
    if (INTERP == kNearest) {
        const int ix = sat_short(__float2int_rn(x32));
        const int iy = sat_short(__float2int_rn(y32));
        float v[4];
        load_tap<PROJ, TIn>(img, pitch, w, h, ix, iy, channels, border, v);
        // nearest copies the element; going through float is exact for u8/u16/f16/f32
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < channels) dst[c] = Finish<TIn, TOut>::run(v[c]);
        return;
    }

    const int sx = __float2int_rn(x32 * 32.0f);
    const int sy = __float2int_rn(y32 * 32.0f);
    const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
    const int fx = sx & 31, fy = sy & 31;

    if (std::is_same<TIn, uint8_t>::value) {
        // ---- 8-bit: 15-bit fixed-point weights -------------------------------------------
        int acc[4] = {0, 0, 0, 0};
        float t[4];
        if (INTERP == kLinear) {
            // weights are (32-fx)(32-fy) etc. in units of 1/1024; identical to cv2's table * 32
            const int wx[2] = {32 - fx, fx}, wy[2] = {32 - fy, fy};
#pragma unroll
            for (int ky = 0; ky < 2; ++ky)
#pragma unroll
                for (int kx = 0; kx < 2; ++kx) {
                    load_tap<PROJ, TIn>(img, pitch, w, h, ix + kx, iy + ky, channels, border, t);
                    const int wgt = wx[kx] * wy[ky];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] += wgt * (int)t[c];
                }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < channels) dst[c] = (TOut)(uint8_t)((acc[c] + 512) >> 10);
        } else {
            const short* wt = g_tables.cubic_fixed + (fy * 32 + fx) * 16;
#pragma unroll
            for (int ky = 0; ky < 4; ++ky)
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    load_tap<PROJ, TIn>(img, pitch, w, h, ix - 1 + kx, iy - 1 + ky, channels, border, t);
                    const int wgt = wt[ky * 4 + kx];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[c] += wgt * (int)t[c];
                }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < channels) dst[c] = (TOut)(uint8_t)min(max((acc[c] + 16384) >> 15, 0), 255);
        }
        return;
    }

    // ---- 16-bit / float: float32 weights, cv2's summation order, no FMA contraction -------
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float t[4];
    if (INTERP == kLinear) {
        const float tx = (float)fx * (1.0f / 32.0f), ty = (float)fy * (1.0f / 32.0f);
        const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty};
#pragma unroll
        for (int ky = 0; ky < 2; ++ky)
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) {
                load_tap<PROJ, TIn>(img, pitch, w, h, ix + kx, iy + ky, channels, border, t);
                const float wgt = __fmul_rn(wy[ky], wx[kx]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float term = __fmul_rn(t[c], wgt);
                    acc[c] = (ky == 0 && kx == 0) ? term : __fadd_rn(acc[c], term);
                }
            }
    } else {
        const float* wx = g_tables.cubic_1d + 4 * fx;
        const float* wy = g_tables.cubic_1d + 4 * fy;
        const int x0 = ix - 1, y0 = iy - 1;
        const bool interior = PROJ == kProjErp ||
            (x0 >= 0 && x0 < max(w - 3, 0) && y0 >= 0 && y0 < max(h - 3, 0));
        if (interior) {
            // rows summed left to right, then added to a running sum that starts at zero
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                float row[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    load_tap<PROJ, TIn>(img, pitch, w, h, x0 + kx, y0 + ky, channels, border, t);
                    const float wgt = __fmul_rn(wy[ky], wx[kx]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float term = __fmul_rn(t[c], wgt);
                        row[c] = kx == 0 ? term : __fadd_rn(row[c], term);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[c] = __fadd_rn(acc[c], row[c]);
            }
        } else {
            // near the sensor edge cv2 starts from the border value and adds (tap - border) * w
            // for the taps that exist
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = border;
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    int xx = x0 + kx, yy = y0 + ky;
                    if (!Border<kProjFisheye>::resolve(xx, yy, w, h)) continue;
                    load_tap<kProjFisheye, TIn>(img, pitch, w, h, xx, yy, channels, border, t);
                    const float wgt = __fmul_rn(wy[ky], wx[kx]);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[c] = __fadd_rn(acc[c], __fmul_rn(__fsub_rn(t[c], border), wgt));
                }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < channels) dst[c] = Finish<TIn, TOut>::run(acc[c]);
}

// ---- projection of one output pixel to (x32, y32, valid) ----------------------------------------

template <int PROJ>
__device__ __forceinline__ bool project_pixel(const ViewDev& view, const ErpDev& erp, const LensDev* lens,
                                              int i, int j, double& x, double& y) {
    double dx, dy, dz;
    ray_at(view, i, j, dx, dy, dz);
    if (PROJ == kProjErp) {
        double lon, lat;
        erp_lonlat(dx, dy, dz, lon, lat);
        erp_xy(erp, lon, lat, x, y);
        return true;
    }
    return fisheye_xy(lens[view.slot], dx, dy, dz, x, y);
}

template <int PROJ, int INTERP, typename TIn, typename TOut>
__global__ void __launch_bounds__(256) remap_direct_kernel(const __grid_constant__ LaunchParams p) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= p.dst.width || j >= p.dst.height) return;
    const int v = blockIdx.z % p.n_views;
    const int g = blockIdx.z / p.n_views;
    const ViewDev& view = p.views[v];

    double x, y;
    const bool valid = project_pixel<PROJ>(view, p.erp, p.lens, i, j, x, y);

    const long long dst_img = (long long)g * p.n_views_total + p.view_base + v;
    TOut* dst = reinterpret_cast<TOut*>(p.dst.data + dst_img * p.dst.image_stride + (long long)j * p.dst.pitch)
                + (long long)i * p.channels;
    if (PROJ == kProjFisheye && p.fill_invalid && !valid) {
        for (int c = 0; c < p.channels; ++c) dst[c] = Finish<TIn, TOut>::run(p.border_value);
        return;
    }
    const unsigned char* img = p.src.data + ((long long)g * p.n_lenses + view.slot) * p.src.image_stride;
    sample_pixel<PROJ, INTERP, TIn, TOut>(img, p.src.pitch, p.src.width, p.src.height, p.channels,
                                          p.border_value, (float)x, (float)y, dst);
}

__global__ void __launch_bounds__(256) coords_kernel(const __grid_constant__ CoordParams p, int proj) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= p.out_w || j >= p.out_h) return;
    const int v = blockIdx.z;
    double x, y;
    const bool valid = proj == kProjErp ? project_pixel<kProjErp>(p.views[v], p.erp, p.lens, i, j, x, y)
                                        : project_pixel<kProjFisheye>(p.views[v], p.erp, p.lens, i, j, x, y);
    const long long o = ((p.view_base + v) * p.out_h + j) * (long long)p.out_w + i;
    if (p.x32) p.x32[o] = (float)x;
    if (p.y32) p.y32[o] = (float)y;
    if (p.x64) p.x64[o] = x;
    if (p.y64) p.y64[o] = y;
    if (p.valid) p.valid[o] = valid ? 1 : 0;
}

}  // namespace r360
